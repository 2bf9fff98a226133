/*
 * Plain-C caller of the drop-in boundary (include/pcv_attn.h): no Python, no torch — device buffers from the CUDA
 * runtime, one pcv_attn_fwd call per implementation on a stream of its own, result checked against a double-precision
 * host loop that restates perceiver/model/core/modules.py:146-164 (scores, finite-fill padding mask, right-aligned
 * causal mask, softmax, P.V) for this one small case.  tests/test_gpu_c_abi.py builds and runs it on the GPU box.
 *
 *   gcc -O2 -I include -I /usr/local/cuda/include examples/c_abi_demo.c -o build/c_abi_demo \
 *       -L perceiver_io_b200/lib -lpcv_attn -L /usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/perceiver_io_b200/lib
 */
#include <cuda_runtime_api.h>
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pcv_attn.h"

#define CHECK_CUDA(x)                                                                       \
  do {                                                                                      \
    cudaError_t e_ = (x);                                                                   \
    if (e_ != cudaSuccess) {                                                                \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 2;                                                                             \
    }                                                                                       \
  } while (0)

static uint16_t f32_to_bf16(float f) { /* round to nearest even */
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static float bf16_to_f32(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static float frand(uint32_t* s) { /* xorshift, roughly N(0,1) as a sum of uniforms */
  float acc = 0.f;
  for (int i = 0; i < 4; ++i) {
    *s ^= *s << 13; *s ^= *s >> 17; *s ^= *s << 5;
    acc += (float)(*s & 0xFFFFFF) / 16777216.0f - 0.5f;
  }
  return acc * 1.7320508f;
}

int main(void) {
  enum { B = 2, H = 2, N = 70, M = 333, D = 64 }; /* ragged on purpose: partial query and key tiles */
  const float scale = 1.0f / sqrtf((float)D);
  const size_t nq = (size_t)B * N * H * D, nk = (size_t)B * M * H * D;
  uint16_t *q = malloc(nq * 2), *k = malloc(nk * 2), *v = malloc(nk * 2), *o = malloc(nq * 2);
  uint8_t* pad = calloc((size_t)B * M, 1);
  double* ref = malloc(nq * sizeof(double));
  uint32_t seed = 12345u;
  for (size_t i = 0; i < nq; ++i) q[i] = f32_to_bf16(frand(&seed));
  for (size_t i = 0; i < nk; ++i) k[i] = f32_to_bf16(frand(&seed));
  for (size_t i = 0; i < nk; ++i) v[i] = f32_to_bf16(frand(&seed));
  for (int j = 0; j < 37; ++j) pad[j] = 1;            /* batch row 0: left padding            */
  for (int j = 0; j < M; ++j) pad[M + j] = 1;         /* batch row 1: fully padded -> uniform */

  if (pcv_abi_version() != PCV_ABI_VERSION) {
    fprintf(stderr, "ABI version mismatch\n");
    return 2;
  }
  void *dq, *dk, *dv, *dout, *dpad;
  CHECK_CUDA(cudaMalloc(&dq, nq * 2));
  CHECK_CUDA(cudaMalloc(&dk, nk * 2));
  CHECK_CUDA(cudaMalloc(&dv, nk * 2));
  CHECK_CUDA(cudaMalloc(&dout, nq * 2));
  CHECK_CUDA(cudaMalloc(&dpad, (size_t)B * M));
  CHECK_CUDA(cudaMemcpy(dq, q, nq * 2, cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(dk, k, nk * 2, cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(dv, v, nk * 2, cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(dpad, pad, (size_t)B * M, cudaMemcpyHostToDevice));
  cudaStream_t stream;
  CHECK_CUDA(cudaStreamCreate(&stream));

  int failures = 0;
  for (int causal = 0; causal < 2; ++causal) {
    /* host restatement in double */
    double ref_max = 0.0;
    for (int b = 0; b < B; ++b)
      for (int h = 0; h < H; ++h)
        for (int i = 0; i < N; ++i) {
          static double s[M];
          double mx = -DBL_MAX;
          for (int j = 0; j < M; ++j) {
            double acc = 0.0;
            for (int c = 0; c < D; ++c)
              acc += (double)bf16_to_f32(q[((size_t)(b * N + i) * H + h) * D + c]) *
                     (double)bf16_to_f32(k[((size_t)(b * M + j) * H + h) * D + c]);
            acc *= scale;
            if (pad[b * M + j] || (causal && j > i + (M - N))) acc = -(double)FLT_MAX; /* modules.py:152-158 */
            s[j] = acc;
            if (acc > mx) mx = acc;
          }
          double l = 0.0;
          for (int j = 0; j < M; ++j) {
            s[j] = exp(s[j] - mx);
            l += s[j];
          }
          for (int c = 0; c < D; ++c) {
            double acc = 0.0;
            for (int j = 0; j < M; ++j) acc += s[j] * (double)bf16_to_f32(v[((size_t)(b * M + j) * H + h) * D + c]);
            acc /= l;
            ref[((size_t)(b * N + i) * H + h) * D + c] = acc;
            if (fabs(acc) > ref_max) ref_max = fabs(acc);
          }
        }

    const int impls[2] = {PCV_IMPL_AUTO, PCV_IMPL_SIMT};
    for (int t = 0; t < 2; ++t) {
      pcv_attn_params p;
      memset(&p, 0, sizeof(p));
      p.q = dq; p.k = dk; p.v = dv; p.out = dout;
      p.q_stride_b = (int64_t)N * H * D; p.q_stride_n = H * D; p.q_stride_h = D;
      p.k_stride_b = (int64_t)M * H * D; p.k_stride_m = H * D; p.k_stride_h = D;
      p.v_stride_b = (int64_t)M * H * D; p.v_stride_m = H * D; p.v_stride_h = D;
      p.o_stride_b = (int64_t)N * H * D; p.o_stride_n = H * D; p.o_stride_h = D;
      p.B = B; p.H = H; p.N = N; p.M = M; p.dqk = D; p.dv = D;
      p.scale = scale; p.dtype = PCV_BF16; p.causal = causal; p.m_total = M; p.m_offset = 0;
      p.pad_mask = (const uint8_t*)dpad; p.pad_stride_b = M;
      p.impl = impls[t];
      size_t need = 0;
      if (pcv_attn_workspace_bytes(&p, &need) != PCV_OK) {
        fprintf(stderr, "workspace query failed: %s\n", pcv_last_error());
        return 2;
      }
      void* ws = NULL;
      if (need) CHECK_CUDA(cudaMalloc(&ws, need));
      p.workspace = ws; p.workspace_bytes = need;
      CHECK_CUDA(cudaMemsetAsync(dout, 0xFF, nq * 2, stream));
      if (pcv_attn_fwd(&p, stream) != PCV_OK) {
        fprintf(stderr, "pcv_attn_fwd failed: %s\n", pcv_last_error());
        return 2;
      }
      CHECK_CUDA(cudaStreamSynchronize(stream));
      CHECK_CUDA(cudaMemcpy(o, dout, nq * 2, cudaMemcpyDeviceToHost));
      double err = 0.0;
      for (size_t i = 0; i < nq; ++i) {
        const double d = fabs((double)bf16_to_f32(o[i]) - ref[i]);
        if (!(d <= err)) err = d; /* also catches NaN */
      }
      const double bound = 1.2e-2 * ref_max;
      printf("causal=%d impl=%d workspace=%zu max_err=%.3e bound=%.3e %s\n", causal, impls[t], need, err, bound,
             err <= bound ? "ok" : "FAIL");
      if (!(err <= bound)) ++failures;
      if (ws) CHECK_CUDA(cudaFree(ws));
    }
  }
  /* error convention: a bad argument returns a status and leaves a message */
  pcv_attn_params bad;
  memset(&bad, 0, sizeof(bad));
  if (pcv_attn_fwd(&bad, stream) == PCV_OK || strlen(pcv_last_error()) == 0) {
    printf("bad-argument call did not fail\n");
    ++failures;
  }
  printf(failures ? "C_ABI_DEMO_FAILED\n" : "C_ABI_DEMO_OK launches=%llu\n", (unsigned long long)pcv_launch_count());
  return failures ? 1 : 0;
}
