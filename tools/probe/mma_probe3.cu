// Microbenchmark 3: is the ~86-cycle SS-mode issue floor of tcgen05.mma (M=128 K=16) a limit of ONE issuing thread?
// W issuer warps of one CTA (one elected lane each) issue MMAs concurrently, into separate accumulators or into the
// SAME accumulator; the accumulator is read back and checked against the exact expected sum (all operands are
// powers of two), so lost updates from racing issuers would show.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe3 mma_probe3.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../terrain_diffusion_b200/csrc/tdx_ptx.cuh"
using namespace tdx;

struct Cfg { int n, warps, same_acc, iters; };

__global__ void __launch_bounds__(192, 1) probe(Cfg c, long long* out, float* val) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  __shared__ long long t_start[4], t_end[4];
  const int warp = threadIdx.x >> 5;
  // bf16 0x3C00 = 2^-7: every product is 2^-14, a K=16 MMA adds 2^-10 to every accumulator element
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
  if (warp == 4) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  if (warp == 5) {   // zero the accumulators: one MMA with accumulate = 0 per accumulator, then wait
    const uint32_t idesc = make_idesc_bf16(128, c.n);
    if (elect_one()) {
      for (int w = 0; w < 4 && (w + 1) * c.n <= 512; ++w)
        umma_bf16(tb + w * c.n, make_smem_desc(smem_u32(smem), 2880, 160),
                  make_smem_desc(smem_u32(smem) + 96 * 1024, c.n * 16, 128), idesc, 0);
      umma_commit(&bar[3]);
    }
    __syncwarp();
    mbar_wait(&bar[3], 0, 9);
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp < c.warps) {
    const uint32_t idesc = make_idesc_bf16(128, c.n);
    const uint32_t d = tb + (c.same_acc ? 0 : warp * c.n);
    uint64_t ad[4], bd[4];
    for (int j = 0; j < 4; ++j) {
      ad[j] = make_smem_desc(smem_u32(smem) + 176 + warp * 23040 + j * 5760, 2880, 160);
      bd[j] = make_smem_desc(smem_u32(smem) + 96 * 1024 + j * (c.n * 32), c.n * 16, 128);
    }
    long long t0 = clock64();
    for (int i = 0; i < c.iters; ++i) {
      if (elect_one()) {
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int j = 0; j < 4; ++j) umma_bf16(d, ad[j], bd[j], idesc, 1);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar[warp]);
    __syncwarp();
    mbar_wait(&bar[warp], 0, warp);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) { t_start[warp] = t0; t_end[warp] = t1; }
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp == 0) {
    uint32_t r[16];
    float tot = 0.f;
    for (int w = 0; w < (c.same_acc ? 1 : c.warps); ++w) {
      tmem_ld16(tb + w * c.n, r);
      tmem_ld_wait();
      tot += __uint_as_float(r[3]);
    }
    if (threadIdx.x == 5 && blockIdx.x == 0) val[0] = tot;
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    long long a = t_start[0], b = t_end[0];
    for (int w = 1; w < c.warps; ++w) { a = min(a, t_start[w]); b = max(b, t_end[w]); }
    out[0] = b - a;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  float* v; cudaMalloc(&v, 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int ns[] = {64, 128, 256};
  for (int n : ns) {
    for (int same = 0; same <= 1; ++same) {
      for (int w = 1; w <= 4; ++w) {
        if (!same && w * n > 512) continue;
        Cfg c = {n, w, same, 128};
        for (int grid : {1, 148}) {
          probe<<<grid, 192, 180 * 1024>>>(c, d, v);
          long long h = 0; float hv = 0;
          cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
          cudaMemcpy(&hv, v, 4, cudaMemcpyDeviceToHost);
          if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
          const double total = (double)c.iters * 8 * w;
          const double expect = total / 1024.0 + (same ? 1 : w) * (1.0 / 1024.0);   // + the zeroing MMAs
          printf("N=%3d %s acc, %d issuer warps, grid %3d: %7.1f cycles per MMA (SM rate), acc sum %.6f expect %.6f %s\n",
                 n, same ? "same" : "sep ", w, grid, (double)h / total, hv, expect,
                 fabs(hv - expect) < 1e-6 ? "OK" : "MISMATCH");
        }
      }
    }
  }
  return 0;
}
