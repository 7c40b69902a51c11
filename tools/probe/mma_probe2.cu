// Microbenchmark 2: does the SS-mode issue floor of tcgen05.mma (M=128 K=16 bf16: ~86 cycles for N <= 128) belong to
// the SM's tensor pipe or to one issuing CTA?  Runs the same issue loop with 1 or 2 co-resident CTAs per SM
// (half the shared memory / TMEM columns each), with M=64 tiles, and with the A operand in TMEM (TS mode).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe2 mma_probe2.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../terrain_diffusion_b200/csrc/tdx_ptx.cuh"
using namespace tdx;

struct Cfg { int m, n, nacc, iters, ts, tmem_cols; };

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(128, 2) probe(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(&slot, c.tmem_cols); tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_bf16(c.m, c.n);
    const uint32_t a0 = smem_u32(smem) + 176, b0 = smem_u32(smem) + 48 * 1024;
    const uint32_t a_tmem = tb + c.nacc * c.n;   // TS mode: A (M x 16 bf16 = 8 columns per K step) behind the accumulators
    long long t0 = clock64();
    for (int i = 0; i < c.iters; ++i) {
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t d = tb + ((i * 4 + j) % c.nacc) * c.n;
          const uint64_t bd = make_smem_desc(b0 + j * (c.n * 32), c.n * 16, 128);
          if (c.ts) umma_bf16_ts(d, a_tmem + j * 8, bd, idesc, 1);
          else umma_bf16(d, make_smem_desc(a0 + j * 5760, 2880, 160), bd, idesc, 1);
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0, 1);
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tb, c.tmem_cols); }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, probe, 128, 100 * 1024);
  printf("SMs %d, probe CTAs/SM by occupancy API: %d\n", nsm, occ);
  struct { const char* name; Cfg c; } tests[] = {
    {"M=128 N=64  SS 4 acc", {128, 64, 4, 256, 0, 256}},
    {"M=128 N=128 SS 2 acc", {128, 128, 2, 256, 0, 256}},
    {"M=128 N=32  SS 4 acc", {128, 32, 4, 256, 0, 256}},
    {"M=128 N=16  SS 4 acc", {128, 16, 4, 256, 0, 256}},
    {"M=128 N=192 SS 1 acc", {128, 192, 1, 256, 0, 256}},
    {"M=128 N=256 SS 1 acc", {128, 256, 1, 256, 0, 256}},
    {"M=64  N=64  SS 4 acc", {64, 64, 4, 256, 0, 256}},
    {"M=64  N=128 SS 2 acc", {64, 128, 2, 256, 0, 256}},
    {"M=64  N=256 SS 1 acc", {64, 256, 1, 256, 0, 256}},
    {"M=128 N=64  TS 2 acc", {128, 64, 2, 256, 1, 256}},
    {"M=128 N=128 TS 1 acc", {128, 128, 1, 256, 1, 256}},
    {"M=128 N=32  TS 4 acc", {128, 32, 4, 256, 1, 256}},
  };
  for (auto& t : tests) {
    for (int per_sm = 1; per_sm <= 2; ++per_sm) {
      // grid of nsm * per_sm CTAs: with 100 KB of shared memory each, two are co-resident on every SM
      probe<<<nsm * per_sm, 128, (per_sm == 1 ? 100 : 100) * 1024>>>(t.c, d);
      long long h = 0; cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf("%s: CUDA error %s\n", t.name, cudaGetErrorString(e)); return 1; }
      printf("%-24s %d CTA/SM: %7.1f cycles/MMA per CTA -> %7.1f cycles/MMA per SM\n", t.name, per_sm,
             (double)h / (t.c.iters * 4), (double)h / (t.c.iters * 4) / per_sm);
    }
  }
  return 0;
}
