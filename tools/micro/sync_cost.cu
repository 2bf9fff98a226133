// Micro-benchmark (GPU box): cost, in SM cycles as seen by the issuing thread, of the synchronisation
// instructions on the MMA issuer's serial path: mbarrier.try_wait on a completed phase (dependent chain and two
// independent probes), tcgen05.commit with nothing outstanding, tcgen05.fence::after_thread_sync, mbarrier.arrive.
// Build + run: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/sync_cost tools/micro/sync_cost.cu && build/sync_cost
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void k(long long* out) {
  __shared__ uint64_t bars[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[i])), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x >= 32) return;
  const int R = 64;
  long long t0, t1;
  uint32_t acc = 0;
  // fresh barrier: waiting for parity 1 succeeds immediately ("previous phase" complete)
  // (a) dependent chain of try_wait on a complete phase
  t0 = clock64();
  for (int i = 0; i < R; ++i) { acc += try_wait(&bars[(i + acc) & 1], 1); }
  t1 = clock64();
  if (threadIdx.x == 0) out[0] = (t1 - t0) / R;
  // (b) two independent probes per step
  t0 = clock64();
  for (int i = 0; i < R; ++i) { bool a = try_wait(&bars[0], 1); bool b = try_wait(&bars[1], 1); acc += (a && b); acc &= 1; }
  t1 = clock64();
  if (threadIdx.x == 0) out[1] = (t1 - t0) / R;
  // (c) three independent probes per step
  t0 = clock64();
  for (int i = 0; i < R; ++i) { bool a = try_wait(&bars[0], 1); bool b = try_wait(&bars[1], 1); bool c = try_wait(&bars[2], 1); acc += (a && b && c); acc &= 1; }
  t1 = clock64();
  if (threadIdx.x == 0) out[2] = (t1 - t0) / R;
  // (d) test_wait dependent chain
  t0 = clock64();
  for (int i = 0; i < R; ++i) { acc += test_wait(&bars[(i + acc) & 1], 1); }
  t1 = clock64();
  if (threadIdx.x == 0) out[3] = (t1 - t0) / R;
  // (e) tcgen05.fence::after_thread_sync
  t0 = clock64();
  for (int i = 0; i < R; ++i) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  t1 = clock64();
  if (threadIdx.x == 0) out[4] = (t1 - t0) / R;
  // (f) tcgen05.commit, nothing outstanding (one lane), each followed by a wait for its arrival (round trip)
  if (threadIdx.x == 0) {
    t0 = clock64();
    for (int i = 0; i < R; ++i) commit(&bars[3 + (i & 3)]);
    t1 = clock64();
    out[5] = (t1 - t0) / R;
    // round trip: commit -> phase completion visible
    uint32_t par = 0;  // bars[7] untouched so far: phase 0 pending
    t0 = clock64();
    for (int i = 0; i < 16; ++i) {
      commit(&bars[7]);
      while (!try_wait(&bars[7], par)) {}
      par ^= 1;
    }
    t1 = clock64();
    out[6] = (t1 - t0) / 16;
    // (g) mbarrier.arrive round trip on own barrier
    t0 = clock64();
    for (int i = 0; i < 16; ++i) {
      asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(&bars[7])) : "memory");
      while (!try_wait(&bars[7], par)) {}
      par ^= 1;
    }
    t1 = clock64();
    out[7] = (t1 - t0) / 16;
    // (h) try_wait on an INCOMPLETE phase (how long one failed probe blocks)
    t0 = clock64();
    for (int i = 0; i < 4; ++i) acc += try_wait(&bars[7], par);
    t1 = clock64();
    out[8] = (t1 - t0) / 4;
    t0 = clock64();
    for (int i = 0; i < 4; ++i) acc += test_wait(&bars[7], par);
    t1 = clock64();
    out[9] = (t1 - t0) / 4;
  }
  if (acc == 12345) out[15] = acc;
}

int main() {
  long long* d;
  cudaMalloc(&d, 16 * sizeof(long long));
  cudaMemset(d, 0, 16 * sizeof(long long));
  k<<<1, 64>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[16];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("status %s\n", cudaGetErrorString(e));
  const char* names[] = {"try_wait complete, dependent chain", "2 independent try_wait", "3 independent try_wait",
                         "test_wait complete, dependent chain", "tcgen05.fence::after_thread_sync",
                         "tcgen05.commit issue (nothing outstanding)", "commit -> completion visible (round trip)",
                         "mbarrier.arrive -> completion visible", "try_wait on incomplete phase (one failed probe)",
                         "test_wait on incomplete phase"};
  for (int i = 0; i < 10; ++i) printf("%-50s %lld cycles\n", names[i], h[i]);
  return 0;
}
