// pcv_attn_decode.cu — attention for a handful of query rows against a long key/value cache (sm_100a):
// the Perceiver AR decode step (reference modules.py:146-164 with i = 1 query, j = n cached keys; SURVEY.md
// §8(f)3).  With N <= 4 queries the two contractions are matrix-vector products: every K and V byte is used once,
// the tensor cores have nothing to amortise (the 128-row tcgen05 tile would waste 127/128 of its MMA rows) and the
// roofline is HBM: algorithmic bytes = B*M*(Dqk + Dv)*2 per call.  So this is a pure streaming kernel:
//
//   grid = B * splits * H CTAs with the head index fastest (CTAs that run together cover all heads of the same key
//   rows, so DRAM pages are consumed whole); CTA = 4 warps, one (b, h) and a contiguous key range;
//   a key's row is split over LPK lanes x 16 bytes (LPK = max head dim / 8 rounded up to a power of two, <= 32), so a
//   warp covers 32/LPK keys per step with fully coalesced 16-byte loads; kUnroll steps are in flight per warp
//   (all K and V loads of a block are issued before the first is consumed);
//   scores are reduced inside the lane group with shuffles; online softmax per lane group in the log2 domain with
//   one rescale per block of kUnroll keys; probabilities stay fp32 (no bf16 rounding of P: closer to the fp64
//   reference than the tensor-core path);
//   lane groups -> warps -> CTA are merged through shared memory, the split states go to the workspace and the LAST
//   CTA of every (b, h) (atomic ticket) merges them and writes the normalised output or the partial state: one
//   launch, no follow-up merge kernel (CUDA-graph friendly).
// Masks follow include/pcv_attn.h: finite fill for padding / causal keys (a fully masked row is the uniform average).
#include "pcv_common.cuh"

#include <algorithm>

namespace pcv {
namespace {

constexpr int kDecWarps = 4;
constexpr int kDecThreads = kDecWarps * 32;
constexpr int kMaxQ = 4;

struct DecParams {
  pcv_attn_params a;
  int nsplit, keys_per_split;
  float* ws_o;        // [B*H][nsplit][NQ][dv]
  float* ws_m;        // [B*H][nsplit][NQ]
  float* ws_l;        // [B*H][nsplit][NQ]
  unsigned int* tickets;  // [B*H], zero on entry; the last CTA of a (b,h) resets its ticket
};

template <int NQ>
struct Unroll {
  static constexpr int value = 4;  // warp steps per register buffer; two buffers: 16 independent 16-byte loads per lane
};

template <typename T>
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const typename Elem<T>::T2* h = reinterpret_cast<const typename Elem<T>::T2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 x = Elem<T>::to_f2(h[i]);
    f[2 * i] = x.x;
    f[2 * i + 1] = x.y;
  }
}

// LPK lanes share one key; NQ query rows
template <typename T, int LPK, int NQ>
__global__ void __launch_bounds__(kDecThreads) attn_decode_kernel(const DecParams p) {
  constexpr int kUnroll = Unroll<NQ>::value;
  constexpr int KPW = 32 / LPK;           // keys per warp step
  constexpr int KPB = KPW * kUnroll;      // keys per warp block
  const pcv_attn_params& a = p.a;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % LPK;             // 16-byte chunk of the row this lane owns
  const int grp = lane / LPK;             // key of the warp step this lane works on
  // blockIdx.x = (b * nsplit + split) * H + h
  const int h = blockIdx.x % a.H;
  const int split = (blockIdx.x / a.H) % p.nsplit;
  const int b = blockIdx.x / (a.H * p.nsplit);
  const int bh = b * a.H + h;
  const int kb = split * p.keys_per_split;
  const int ke = min(a.M, kb + p.keys_per_split);
  const int c0 = sub * 8;
  const bool kq_live = c0 < a.dqk, v_live = c0 < a.dv;

  const T* qp = reinterpret_cast<const T*>(a.q) + (a.q_stride_b ? (int64_t)b * a.q_stride_b : 0) + (int64_t)h * a.q_stride_h + c0;
  const T* kp = reinterpret_cast<const T*>(a.k) + (int64_t)b * a.k_stride_b + (int64_t)h * a.k_stride_h + c0;
  const T* vp = reinterpret_cast<const T*>(a.v) + (int64_t)b * a.v_stride_b + (int64_t)h * a.v_stride_h + c0;
  const uint8_t* pad = a.pad_mask ? a.pad_mask + (int64_t)b * a.pad_stride_b : nullptr;

  const float scale_log2 = a.scale * kLog2e;
  float q[NQ][8];
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    uint4 u = make_uint4(0, 0, 0, 0);
    if (kq_live && i < a.N) u = *reinterpret_cast<const uint4*>(qp + (int64_t)i * a.q_stride_n);
    unpack8<T>(u, q[i]);
#pragma unroll
    for (int c = 0; c < 8; ++c) q[i][c] *= scale_log2;   // scores come out in the log2 domain
  }
  const int causal_shift = a.m_total - a.N;  // key jg masked for query n iff jg > n + causal_shift

  float m[NQ], l[NQ], acc[NQ][8];
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    m[i] = -INFINITY;
    l[i] = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[i][c] = 0.f;
  }

  // keys of this CTA are dealt to the warps in blocks of KPB keys: warp w takes blocks w, w + kDecWarps, ...
  // Register double buffering: the loads of block i+1 are issued BEFORE block i is consumed, so every lane always
  // has kUnroll K rows + kUnroll V rows (16 bytes each) in flight while it computes.
  uint4 ku[2][kUnroll], vu[2][kUnroll];
  bool masked[2][kUnroll];
  auto load_block = [&](int buf, int j0) {
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int j = j0 + u * KPW + grp;
      ku[buf][u] = make_uint4(0, 0, 0, 0);
      vu[buf][u] = make_uint4(0, 0, 0, 0);
      masked[buf][u] = false;
      if (j < ke) {
        if (kq_live) ku[buf][u] = __ldcs(reinterpret_cast<const uint4*>(kp + (int64_t)j * a.k_stride_m));
        if (v_live) vu[buf][u] = __ldcs(reinterpret_cast<const uint4*>(vp + (int64_t)j * a.v_stride_m));
        masked[buf][u] = pad != nullptr && pad[j] != 0;
      }
    }
  };
  auto consume_block = [&](int buf, int j0) {
    float s[NQ][kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      float kf[8];
      unpack8<T>(ku[buf][u], kf);
      const int j = j0 + u * KPW + grp;
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        float d = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) d = fmaf(q[i][c], kf[c], d);
#pragma unroll
        for (int o = LPK / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        if (masked[buf][u] || (a.causal && a.m_offset + j > i + causal_shift)) d = kMaskedScore;
        if (j >= ke) d = -INFINITY;
        s[i][u] = d;
      }
    }
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      float mb = s[i][0];
#pragma unroll
      for (int u = 1; u < kUnroll; ++u) mb = fmaxf(mb, s[i][u]);
      const float m_new = fmaxf(m[i], mb);
      if (m_new == -INFINITY) continue;  // no live key in this block for this lane group
      const float alpha = exp2f(m[i] - m_new);
      m[i] = m_new;
      l[i] *= alpha;
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[i][c] *= alpha;
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const float pe = exp2f(s[i][u] - m_new);
        l[i] += pe;
        float vf[8];
        unpack8<T>(vu[buf][u], vf);
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[i][c] = fmaf(pe, vf[c], acc[i][c]);
      }
    }
  };
  {
    constexpr int kStride = kDecWarps * KPB;
    int j0 = kb + warp * KPB;
    if (j0 < ke) load_block(0, j0);
    while (j0 < ke) {
      if (j0 + kStride < ke) load_block(1, j0 + kStride);
      consume_block(0, j0);
      j0 += kStride;
      if (j0 >= ke) break;
      if (j0 + kStride < ke) load_block(0, j0 + kStride);
      consume_block(1, j0);
      j0 += kStride;
    }
  }

  // ---- merge: lane groups of a warp -> warps of the CTA (shared memory) -----------------------------------
  __shared__ float sm_m[kDecWarps][NQ], sm_l[kDecWarps][NQ];
  __shared__ float sm_o[kDecWarps][NQ][LPK * 8];
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    // groups: lanes with equal `sub` hold the same channels for different keys
#pragma unroll
    for (int o = LPK; o < 32; o <<= 1) {
      const float m_o = __shfl_xor_sync(0xffffffffu, m[i], o);
      const float l_o = __shfl_xor_sync(0xffffffffu, l[i], o);
      const float m_new = fmaxf(m[i], m_o);
      const float wa = (m[i] == -INFINITY) ? 0.f : exp2f(m[i] - m_new);
      const float wb = (m_o == -INFINITY) ? 0.f : exp2f(m_o - m_new);
      l[i] = l[i] * wa + l_o * wb;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float a_o = __shfl_xor_sync(0xffffffffu, acc[i][c], o);
        acc[i][c] = acc[i][c] * wa + a_o * wb;
      }
      m[i] = m_new;
    }
    if (grp == 0) {
      if (sub == 0) {
        sm_m[warp][i] = m[i];
        sm_l[warp][i] = l[i];
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) sm_o[warp][i][c0 + c] = acc[i][c];
    }
  }
  __syncthreads();

  const int dvp = LPK * 8;
  // CTA state -> workspace: thread t handles (query i, channel c)
  const int64_t wbase = ((int64_t)bh * p.nsplit + split) * NQ;
  for (int idx = threadIdx.x; idx < NQ * dvp; idx += kDecThreads) {
    const int i = idx / dvp, c = idx - i * dvp;
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < kDecWarps; ++w) mm = fmaxf(mm, sm_m[w][i]);
    float o = 0.f, ll = 0.f;
#pragma unroll
    for (int w = 0; w < kDecWarps; ++w) {
      const float wt = (sm_m[w][i] == -INFINITY) ? 0.f : exp2f(sm_m[w][i] - mm);
      o = fmaf(sm_o[w][i][c], wt, o);
      ll = fmaf(sm_l[w][i], wt, ll);
    }
    if (c < a.dv) p.ws_o[(wbase + i) * a.dv + c] = o;
    if (c == 0) {
      p.ws_m[wbase + i] = mm;
      p.ws_l[wbase + i] = ll;
    }
  }

  // ---- the last CTA of this (b, h) merges the splits ---------------------------------------------------------
  __shared__ unsigned int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(p.tickets + bh, 1u);
    s_last = (t == (unsigned int)p.nsplit - 1) ? 1u : 0u;
    if (s_last) p.tickets[bh] = 0u;  // ready for the next launch on this workspace
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int idx = threadIdx.x; idx < NQ * a.dv; idx += kDecThreads) {
    const int i = idx / a.dv, c = idx - i * a.dv;
    if (i >= a.N) continue;
    const int64_t sb = (int64_t)bh * p.nsplit * NQ + i;
    float mm = -INFINITY;
    for (int sp = 0; sp < p.nsplit; ++sp) mm = fmaxf(mm, __ldcg(p.ws_m + sb + (int64_t)sp * NQ));
    float o = 0.f, ll = 0.f;
    for (int sp = 0; sp < p.nsplit; ++sp) {
      const float ms = __ldcg(p.ws_m + sb + (int64_t)sp * NQ);
      const float wt = (ms == -INFINITY) ? 0.f : exp2f(ms - mm);
      o = fmaf(__ldcg(p.ws_o + (sb + (int64_t)sp * NQ) * a.dv + c), wt, o);
      ll = fmaf(__ldcg(p.ws_l + sb + (int64_t)sp * NQ), wt, ll);
    }
    if (!a.write_partial) {
      T* out = reinterpret_cast<T*>(a.out) + (int64_t)b * a.o_stride_b + (int64_t)i * a.o_stride_n + (int64_t)h * a.o_stride_h;
      out[c] = Elem<T>::from_f(o / ll);
    } else {
      const int64_t r = ((int64_t)b * a.H + h) * a.N + i;
      a.part_o[r * a.dv + c] = o;
      if (c == 0) {
        a.part_m[r] = mm;
        a.part_l[r] = ll;
      }
    }
  }
}

int choose_split(const pcv_attn_params& a, int* nsplit, int* keys_per_split) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t bh = (int64_t)a.B * a.H;
  // ~12 CTAs (of 4 warps) per SM overall — several waves, so the tail wave is short — at least 256 keys per CTA,
  // splits on 128-key boundaries
  int64_t want = std::max<int64_t>(1, (12LL * sms + bh - 1) / bh);
  const int64_t max_by_keys = std::max<int64_t>(1, a.M / 256);
  want = std::min<int64_t>(std::min<int64_t>(want, max_by_keys), 256);
  int64_t kps = (a.M + want - 1) / want;
  kps = (kps + 127) / 128 * 128;
  *keys_per_split = (int)kps;
  *nsplit = (int)((a.M + kps - 1) / kps);
  return PCV_OK;
}

size_t align256(size_t x) { return (x + 255) / 256 * 256; }

template <typename T, int LPK>
int launch_nq(const DecParams& p, cudaStream_t stream) {
  dim3 grid((unsigned)((int64_t)p.nsplit * p.a.B * p.a.H));
  if (p.a.N == 1)
    attn_decode_kernel<T, LPK, 1><<<grid, kDecThreads, 0, stream>>>(p);
  else  // 2-4 query rows share the four-row instantiation (rows beyond N are zero queries whose results are dropped)
    attn_decode_kernel<T, LPK, 4><<<grid, kDecThreads, 0, stream>>>(p);
  PCV_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return PCV_OK;
}

template <typename T>
int launch_lpk(const DecParams& p, int lpk, cudaStream_t stream) {
  switch (lpk) {  // rows of up to 32 channels use the 4-lane instantiation with idle lanes
    case 1:
    case 2:
    case 4: return launch_nq<T, 4>(p, stream);
    case 8: return launch_nq<T, 8>(p, stream);
    case 16: return launch_nq<T, 16>(p, stream);
    default: return launch_nq<T, 32>(p, stream);
  }
}

int lanes_per_key(const pcv_attn_params& a) {
  const int chunks = (std::max(a.dqk, a.dv) + 7) / 8;
  int lpk = 1;
  while (lpk < chunks) lpk <<= 1;
  return lpk;
}

}  // namespace

bool attn_decode_supported(const pcv_attn_params& a, const char** why) {
  auto fail = [&](const char* w) {
    *why = w;
    return false;
  };
  if (a.N > kMaxQ) return fail("more than 4 query rows");
  if (a.dqk > 256 || a.dv > 256) return fail("head dim > 256");
  if ((a.dqk % 8) || (a.dv % 8)) return fail("head dims must be multiples of 8");
  auto al16 = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
  if (!al16(a.q) || !al16(a.k) || !al16(a.v)) return fail("q/k/v must be 16-byte aligned");
  if ((a.q_stride_n % 8) || (a.k_stride_m % 8) || (a.v_stride_m % 8) || (a.q_stride_h % 8) || (a.k_stride_h % 8) ||
      (a.v_stride_h % 8) || (a.q_stride_b % 8) || (a.k_stride_b % 8) || (a.v_stride_b % 8))
    return fail("strides must be multiples of 8 elements");
  if (a.M < 1024) return fail("short key axis (the general kernels are as fast)");
  return true;
}

int attn_decode_workspace_bytes(const pcv_attn_params& a, size_t* bytes) {
  int nsplit = 1, kps = a.M;
  choose_split(a, &nsplit, &kps);
  const int nq = a.N <= 1 ? 1 : 4;
  const size_t rows = (size_t)a.B * a.H * nsplit * nq;
  *bytes = align256(rows * a.dv * 4) + 2 * align256(rows * 4) + align256((size_t)a.B * a.H * 4);
  return PCV_OK;
}

int launch_attn_decode(const pcv_attn_params& a, cudaStream_t stream) {
  size_t need = 0;
  attn_decode_workspace_bytes(a, &need);
  PCV_REQUIRE(a.workspace != nullptr && a.workspace_bytes >= need, PCV_ERR_WORKSPACE,
              "decode attention: workspace of %zu bytes required, %zu given", need, a.workspace_bytes);
  DecParams p{};
  p.a = a;
  choose_split(a, &p.nsplit, &p.keys_per_split);
  const int nq = a.N <= 1 ? 1 : 4;
  const size_t rows = (size_t)a.B * a.H * p.nsplit * nq;
  char* ws = reinterpret_cast<char*>(a.workspace);
  p.ws_o = reinterpret_cast<float*>(ws);
  ws += align256(rows * a.dv * 4);
  p.ws_m = reinterpret_cast<float*>(ws);
  ws += align256(rows * 4);
  p.ws_l = reinterpret_cast<float*>(ws);
  ws += align256(rows * 4);
  p.tickets = reinterpret_cast<unsigned int*>(ws);
  // the workspace is caller memory with arbitrary contents: the tickets must start at zero
  PCV_CHECK_CUDA(cudaMemsetAsync(p.tickets, 0, (size_t)a.B * a.H * 4, stream));
  const int lpk = lanes_per_key(a);
  prof_mark_begin(stream);
  const int rc = a.dtype == PCV_BF16 ? launch_lpk<__nv_bfloat16>(p, lpk, stream) : launch_lpk<__half>(p, lpk, stream);
  prof_mark_end(stream);
  return rc;
}

}  // namespace pcv
