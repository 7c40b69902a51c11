// Microbenchmark: issue rate of tcgen05.mma (M=128, K=16, bf16) for different N, smem layouts and accumulator reuse.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../terrain_diffusion_b200/csrc/tdx_ptx.cuh"
using namespace tdx;

struct Cfg { int n; int layout; uint32_t a_lbo, a_sbo, b_lbo, b_sbo; uint32_t a_off; int nacc; int iters; int a_step; };

__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t lbo, uint32_t sbo, int layout) {
  uint64_t d = make_smem_desc(addr, lbo, sbo);
  d |= (uint64_t)layout << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) probe(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_bf16(128, c.n);
    const uint32_t a0 = smem_u32(smem) + c.a_off, b0 = smem_u32(smem) + 96 * 1024;
    long long t0 = clock64();
    for (int i = 0; i < c.iters; ++i) {
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          umma_bf16(tb + ((i * 4 + j) % c.nacc) * c.n, desc(a0 + j * c.a_step, c.a_lbo, c.a_sbo, c.layout),
                    desc(b0 + j * 32, c.b_lbo, c.b_sbo, c.layout), idesc, 1);
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
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct { const char* name; Cfg c; } tests[] = {
    // name, n, layout(0 none,2 sw128), a_lbo,a_sbo,b_lbo,b_sbo,a_off,nacc,iters,a_step
    {"N=64  noswz SBO160 off176 (conv tap)  1 acc", {64, 0, 2880, 160, 1024, 128, 176, 1, 256, 5760}},
    {"N=64  noswz SBO160 off176             4 acc", {64, 0, 2880, 160, 1024, 128, 176, 4, 256, 5760}},
    {"N=64  noswz SBO128 off0 aligned       4 acc", {64, 0, 2048, 128, 1024, 128, 0, 4, 256, 4096}},
    {"N=64  noswz SBO128 off0 aligned       1 acc", {64, 0, 2048, 128, 1024, 128, 0, 1, 256, 4096}},
    {"N=64  sw128 (canonical K-major)       4 acc", {64, 2, 16, 1024, 16, 1024, 0, 4, 256, 32}},
    {"N=64  sw128 (canonical K-major)       1 acc", {64, 2, 16, 1024, 16, 1024, 0, 1, 256, 32}},
    {"N=128 noswz SBO160 off176             2 acc", {128, 0, 2880, 160, 2048, 128, 176, 2, 256, 5760}},
    {"N=128 sw128                           2 acc", {128, 2, 16, 1024, 16, 1024, 0, 2, 256, 32}},
    {"N=256 noswz SBO160 off176             1 acc", {256, 0, 2880, 160, 4096, 128, 176, 1, 256, 5760}},
    {"N=256 sw128                           1 acc", {256, 2, 16, 1024, 16, 1024, 0, 1, 256, 32}},
    {"N=256 sw128                           2 acc", {256, 2, 16, 1024, 16, 1024, 0, 2, 256, 32}},
    {"N=32  noswz SBO160                    4 acc", {32, 0, 2880, 160, 512, 128, 176, 4, 256, 5760}},
    {"N=16  noswz SBO160                    4 acc", {16, 0, 2880, 160, 256, 128, 176, 4, 256, 5760}},
  };
  for (auto& t : tests) {
    probe<<<1, 128, 160 * 1024>>>(t.c, d);
    long long h = 0; cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", t.name, cudaGetErrorString(e)); return 1; }
    printf("%-52s %7.1f cycles/MMA (1 CTA)\n", t.name, (double)h / (t.c.iters * 4));
    probe<<<148, 128, 160 * 1024>>>(t.c, d);
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-52s %7.1f cycles/MMA (148 CTAs)\n", "", (double)h / (t.c.iters * 4));
  }
  return 0;
}
