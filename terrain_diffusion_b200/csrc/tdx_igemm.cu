// Persistent, warp-specialised implicit-GEMM convolution for sm_100a.
//
//   D[128 pixels x Cout] (fp32, TMEM) += A[128 x 16] (bf16, smem halo patch) * B[Cout x 16] (bf16, smem weights)
//
// * M tile = 16 rows x 8 columns of output pixels.  The A operand of every filter tap is the SAME (18 x 10)-pixel halo
//   patch in shared memory: activations are stored NC8HW8, so TMA drops the patch as [kc][18][10][8ch] and the
//   tcgen05 K-major/no-swizzle descriptor (8-row core matrices 128 B contiguous, SBO = 10 px * 16 B between pixel
//   rows, LBO = one 8-channel plane) addresses tap (r,c) by just adding (r*10+c)*16 B to the start address.  The patch
//   is fetched once per 64-channel chunk and used by all 9 taps (1.4x halo overhead instead of 9x re-fetch), and conv
//   zero padding is TMA out-of-bounds fill.
// * B (weights) are pre-packed on the host in exactly the shared-memory image order, one stage per (chunk, tap), and
//   streamed with 1-D bulk copies through their own ring.
// * Warp roles: w0 A-producer (TMA tiled), w1 B-producer (bulk copy), w2 MMA issuer (one thread), w3 TMEM allocator,
//   w4..w7 epilogue (one TMEM lane quadrant each).  Two TMEM accumulators (2 x 256 columns) let the epilogue of tile i
//   overlap the MMAs of tile i+1.
// * Epilogue: each thread owns one pixel and ALL Cout channels, so the per-pixel channel reductions of the EDM2 block
//   (pixel-norm) are thread-local; it applies emb-scale+mp_silu / residual mp_sum + clip / pixel-norm and writes up to
//   three bf16 NC8HW8 outputs (raw, activated, activated+resampled) for the consumers.
//
// Reference math being replaced: models/mp_layers.py:201-221 (MPConv), models/unet_block.py:116-156 (UNetBlock).
#include "tdx_common.h"
#include "tdx_ptx.cuh"

namespace tdx {

constexpr int kTileH = 16, kTileW = 8;
constexpr int kPatchH = kTileH + 2, kPatchW = kTileW + 2;
constexpr int kKcBytes = kPatchH * kPatchW * 16;  // one 8-channel plane of the halo patch: 2880 B
constexpr int kAStageBytes = 8 * kKcBytes;        // 64 channels: 23040 B
constexpr int kSA = 3;                            // A ring depth
constexpr int kMaxSB = 8;                         // B ring depth (max)
constexpr int kThreads = 256;
constexpr int kSmemBudget = 227 * 1024;

struct IgemmParams {
  int nseg;
  int seg_chunks[3];
  int seg_taps[3];
  const __nv_bfloat16* B;
  int cout;
  int H, W, nimg, tiles_x, tiles_y, num_tiles;
  int SB;
  int epi;
  const float* cvec;
  const uint4* resid;
  int resid_spatial, resid_pnorm;
  float resid_scale, clip;
  TdxOutSpec out[3];
};

__device__ __forceinline__ void decode_tile(const IgemmParams& p, int tile, int& img, int& Y0, int& X0) {
  int tx = tile % p.tiles_x;
  int t = tile / p.tiles_x;
  int ty = t % p.tiles_y;
  img = t / p.tiles_y;
  Y0 = ty * kTileH;
  X0 = tx * kTileW;
}

__global__ void __launch_bounds__(kThreads, 1)
igemm_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
             const __grid_constant__ CUtensorMap tm2, const __grid_constant__ IgemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + kSA * kAStageBytes;
  const int b_stage_bytes = p.cout * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_ring + p.SB * b_stage_bytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kSA;
  uint64_t* b_full = a_empty + kSA;
  uint64_t* b_empty = b_full + kMaxSB;
  uint64_t* t_full = b_empty + kMaxSB;
  uint64_t* t_empty = t_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm0);
    if (p.nseg > 1) tma_prefetch_desc(&tm1);
    if (p.nseg > 2) tma_prefetch_desc(&tm2);
  }
  if (warp == 2 && lane == 0) {
    for (int i = 0; i < kSA; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < p.SB; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 3) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ A producer (halo patches via tiled TMA)
    if (lane == 0) {
      int sa = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int img, Y0, X0;
        decode_tile(p, tile, img, Y0, X0);
        for (int seg = 0; seg < p.nseg; ++seg) {
          const CUtensorMap* tm = seg == 0 ? &tm0 : (seg == 1 ? &tm1 : &tm2);
          for (int ch = 0; ch < p.seg_chunks[seg]; ++ch) {
            mbar_wait(&a_empty[sa], ph ^ 1, 100 + sa);
            mbar_expect_tx(&a_full[sa], kAStageBytes);
            tma_load_4d(tm, &a_full[sa], a_ring + sa * kAStageBytes, (X0 - 1) * 8, Y0 - 1, ch * 8, img);
            if (++sa == kSA) { sa = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ B producer (pre-packed weight stages)
    if (lane == 0) {
      int sb = 0;
      uint32_t ph = 0;
      int stages_per_tile = 0;
      for (int seg = 0; seg < p.nseg; ++seg) stages_per_tile += p.seg_chunks[seg] * p.seg_taps[seg];
      const uint8_t* bsrc = reinterpret_cast<const uint8_t*>(p.B);
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int ks = 0; ks < stages_per_tile; ++ks) {
          mbar_wait(&b_empty[sb], ph ^ 1, 200 + sb);
          mbar_expect_tx(&b_full[sb], b_stage_bytes);
          bulk_load_1d(bsrc + (size_t)ks * b_stage_bytes, &b_full[sb], b_ring + sb * b_stage_bytes, b_stage_bytes);
          if (++sb == p.SB) { sb = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, p.cout);
      const uint32_t b_lbo = p.cout * 16;
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t accph = (it >> 1) & 1;
        mbar_wait(&t_empty[acc], accph ^ 1, 300 + acc);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        uint32_t accumulate = 0;
        for (int seg = 0; seg < p.nseg; ++seg) {
          const int taps = p.seg_taps[seg];
          for (int ch = 0; ch < p.seg_chunks[seg]; ++ch) {
            mbar_wait(&a_full[sa], pha, 400 + sa);
            tc_fence_after();
            const uint32_t a_base = smem_u32(a_ring + sa * kAStageBytes);
            for (int tap = 0; tap < taps; ++tap) {
              const int r = taps == 9 ? tap / 3 : 1;
              const int c = taps == 9 ? tap % 3 : 1;
              mbar_wait(&b_full[sb], phb, 500 + sb);
              tc_fence_after();
              const uint32_t b_base = smem_u32(b_ring + sb * b_stage_bytes);
              const uint32_t a_tap = a_base + (r * kPatchW + c) * 16;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint64_t adesc = make_smem_desc(a_tap + 2 * j * kKcBytes, kKcBytes, kPatchW * 16);
                const uint64_t bdesc = make_smem_desc(b_base + 2 * j * b_lbo, b_lbo, 128);
                umma_bf16(d_tmem, adesc, bdesc, idesc, accumulate);
                accumulate = 1;
              }
              umma_commit(&b_empty[sb]);
              if (++sb == p.SB) { sb = 0; phb ^= 1; }
            }
            umma_commit(&a_empty[sa]);
            if (++sa == kSA) { sa = 0; pha ^= 1; }
          }
        }
        umma_commit(&t_full[acc]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (TMEM -> registers -> global)
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int y = m >> 3, x = m & 7;
    const int C8 = p.cout >> 3;
    const bool need_norm = (p.epi & TDX_EPI_PNORM) || p.out[0].kind == TDX_OUT_PNORM_SILU ||
                           p.out[1].kind == TDX_OUT_PNORM_SILU || p.out[2].kind == TDX_OUT_PNORM_SILU;
    const int npass = need_norm ? 2 : 1;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t accph = (it >> 1) & 1;
      int img, Y0, X0;
      decode_tile(p, tile, img, Y0, X0);
      const int Y = Y0 + y, X = X0 + x;
      const bool valid = (Y < p.H) && (X < p.W);

      // residual addressing (+ optional pixel-norm of the residual vector)
      const uint4* rptr = nullptr;
      size_t rplane = 0;
      float rscale = p.resid_scale;
      if ((p.epi & TDX_EPI_RESID) && valid) {
        int Hr = p.H, Wr = p.W, Yr = Y, Xr = X;
        if (p.resid_spatial == TDX_SP_UP2) { Hr = p.H >> 1; Wr = p.W >> 1; Yr = Y >> 1; Xr = X >> 1; }
        else if (p.resid_spatial == TDX_SP_DOWN2) { Hr = p.H << 1; Wr = p.W << 1; Yr = Y << 1; Xr = X << 1; }
        rplane = (size_t)Hr * Wr;
        rptr = p.resid + ((size_t)img * C8) * rplane + (size_t)Yr * Wr + Xr;
        if (p.resid_pnorm) {
          float ss = 0.f;
          for (int g = 0; g < C8; ++g) {
            uint4 u = __ldg(rptr + g * rplane);
            float a, b;
            unpack_bf16x2(u.x, a, b); ss += a * a + b * b;
            unpack_bf16x2(u.y, a, b); ss += a * a + b * b;
            unpack_bf16x2(u.z, a, b); ss += a * a + b * b;
            unpack_bf16x2(u.w, a, b); ss += a * a + b * b;
          }
          rscale = p.resid_scale / (1e-4f + sqrtf(ss / (float)p.cout));
        }
      }

      mbar_wait(&t_full[acc], accph, 600 + acc);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * 256 + ((uint32_t)(q * 32) << 16);
      const float* cv = p.cvec ? p.cvec + (size_t)img * p.cout : nullptr;

      float sumsq = 0.f, inv = 1.f;
      for (int pass = 0; pass < npass; ++pass) {
        const bool last = (pass == npass - 1);
        if (need_norm && last) inv = 1.0f / (1e-4f + sqrtf(sumsq / (float)p.cout));
        for (int c0 = 0; c0 < p.cout; c0 += 32) {
          uint32_t r[32];
          __syncwarp();
          tmem_ld32(taddr + c0, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          if (p.epi & TDX_EPI_EMB_SILU) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              float4 c4 = __ldg(reinterpret_cast<const float4*>(cv + c0 + i));
              v[i + 0] = mp_silu_f(v[i + 0] * c4.x);
              v[i + 1] = mp_silu_f(v[i + 1] * c4.y);
              v[i + 2] = mp_silu_f(v[i + 2] * c4.z);
              v[i + 3] = mp_silu_f(v[i + 3] * c4.w);
            }
          }
          if (p.epi & TDX_EPI_RESID) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 u = valid ? __ldg(rptr + (size_t)((c0 >> 3) + g) * rplane) : make_uint4(0, 0, 0, 0);
              float rr[8];
              unpack_bf16x2(u.x, rr[0], rr[1]);
              unpack_bf16x2(u.y, rr[2], rr[3]);
              unpack_bf16x2(u.z, rr[4], rr[5]);
              unpack_bf16x2(u.w, rr[6], rr[7]);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[g * 8 + i] = fmaf(rscale, rr[i], v[g * 8 + i]);
            }
          }
          if (p.clip > 0.f) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fminf(fmaxf(v[i], -p.clip), p.clip);
          }
          if (need_norm && !last) {
#pragma unroll
            for (int i = 0; i < 32; ++i) sumsq = fmaf(v[i], v[i], sumsq);
            continue;
          }
          if (p.epi & TDX_EPI_PNORM) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= inv;
          }
          if (!valid) continue;
#pragma unroll
          for (int o = 0; o < 3; ++o) {
            const TdxOutSpec& os = p.out[o];
            if (os.kind == TDX_OUT_NONE) continue;
            float sc = os.scale;
            if (os.kind == TDX_OUT_PNORM_SILU) sc = (p.epi & TDX_EPI_PNORM) ? 1.0f : inv;
            int Ho = p.H, Wo = p.W, Yo = Y, Xo = X;
            if (os.spatial == TDX_SP_DOWN2) {
              if ((Y | X) & 1) continue;
              Ho >>= 1; Wo >>= 1; Yo >>= 1; Xo >>= 1;
            } else if (os.spatial == TDX_SP_UP2) {
              Ho <<= 1; Wo <<= 1; Yo <<= 1; Xo <<= 1;
            }
            const size_t oplane = (size_t)Ho * Wo;
            uint4* optr = reinterpret_cast<uint4*>(os.ptr) + ((size_t)img * C8 + (c0 >> 3)) * oplane +
                          (size_t)Yo * Wo + Xo;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float w[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float t = v[g * 8 + i];
                w[i] = (os.kind == TDX_OUT_RAW) ? t : mp_silu_f(t * sc);
              }
              uint4 u;
              u.x = pack_bf16x2(w[0], w[1]);
              u.y = pack_bf16x2(w[2], w[3]);
              u.z = pack_bf16x2(w[4], w[5]);
              u.w = pack_bf16x2(w[6], w[7]);
              uint4* dst = optr + (size_t)g * oplane;
              dst[0] = u;
              if (os.spatial == TDX_SP_UP2) {
                dst[1] = u;
                dst[Wo] = u;
                dst[Wo + 1] = u;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------- host side
static int smem_layout(int cout, int* SB_out) {
  const int fixed = kSA * kAStageBytes + 1024;  // rings + barriers/tmem slot
  int SB = (kSmemBudget - fixed) / (cout * 128);
  if (SB > kMaxSB) SB = kMaxSB;
  if (SB < 2) return -1;
  *SB_out = SB;
  int bytes = kSA * kAStageBytes + SB * cout * 128 + 1024;
  if (bytes < 120 * 1024) bytes = 120 * 1024;  // one CTA per SM: each CTA allocates all 512 TMEM columns
  return bytes;
}

int igemm_prepare() {
  static bool attr_set = false;
  if (!attr_set) {
    TDX_CHECK_CUDA(cudaFuncSetAttribute(igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    attr_set = true;
  }
  return TDX_OK;
}

int igemm_launch(const TdxIgemmDesc& d, const CUtensorMap* tms, cudaStream_t stream) {
  IgemmParams p;
  memset(&p, 0, sizeof(p));
  p.nseg = d.n_seg;
  for (int s = 0; s < d.n_seg; ++s) {
    p.seg_chunks[s] = d.a_channels[s] / 64;
    p.seg_taps[s] = d.a_taps[s];
  }
  p.B = reinterpret_cast<const __nv_bfloat16*>(d.b_packed);
  p.cout = d.c_out;
  p.H = d.height;
  p.W = d.width;
  p.nimg = d.n_img;
  p.tiles_x = (d.width + kTileW - 1) / kTileW;
  p.tiles_y = (d.height + kTileH - 1) / kTileH;
  p.num_tiles = p.tiles_x * p.tiles_y * d.n_img;
  int SB = 0;
  const int smem = smem_layout(d.c_out, &SB);
  TDX_REQUIRE(smem > 0, "igemm: c_out=%d does not fit the shared-memory plan", d.c_out);
  p.SB = SB;
  p.epi = d.epi_flags;
  p.cvec = d.cvec;
  p.resid = reinterpret_cast<const uint4*>(d.resid);
  p.resid_spatial = d.resid_spatial;
  p.resid_pnorm = d.resid_pnorm;
  p.resid_scale = d.resid_scale;
  p.clip = d.clip;
  for (int o = 0; o < 3; ++o) p.out[o] = d.out[o];

  int rc_prep = igemm_prepare();
  if (rc_prep != TDX_OK) return rc_prep;
  int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  const CUtensorMap& t0 = tms[0];
  const CUtensorMap& t1 = tms[d.n_seg > 1 ? 1 : 0];
  const CUtensorMap& t2 = tms[d.n_seg > 2 ? 2 : 0];
  igemm_kernel<<<grid, kThreads, smem, stream>>>(t0, t1, t2, p);
  TDX_CHECK_CUDA(cudaGetLastError());
  return TDX_OK;
}

int igemm_validate(const TdxIgemmDesc& d) {
  TDX_REQUIRE(d.n_seg >= 1 && d.n_seg <= 3, "igemm: n_seg=%d not in 1..3", d.n_seg);
  for (int s = 0; s < d.n_seg; ++s) {
    TDX_REQUIRE(d.a_ptr[s] != nullptr, "igemm: a_ptr[%d] is null", s);
    TDX_REQUIRE(d.a_channels[s] > 0 && d.a_channels[s] % 64 == 0, "igemm: a_channels[%d]=%d not a multiple of 64", s,
                d.a_channels[s]);
    TDX_REQUIRE(d.a_taps[s] == 9 || d.a_taps[s] == 1, "igemm: a_taps[%d]=%d not 9 or 1", s, d.a_taps[s]);
  }
  TDX_REQUIRE(d.b_packed != nullptr, "igemm: b_packed is null");
  TDX_REQUIRE(d.c_out >= 32 && d.c_out <= 256 && d.c_out % 32 == 0, "igemm: c_out=%d must be a multiple of 32 <= 256",
              d.c_out);
  TDX_REQUIRE(d.n_img >= 1 && d.height >= 8 && d.width >= 8 && d.height % 8 == 0 && d.width % 8 == 0,
              "igemm: bad shape n=%d h=%d w=%d (h, w multiples of 8)", d.n_img, d.height, d.width);
  if (d.epi_flags & TDX_EPI_EMB_SILU) TDX_REQUIRE(d.cvec != nullptr, "igemm: EMB_SILU needs cvec");
  if (d.epi_flags & TDX_EPI_RESID) {
    TDX_REQUIRE(d.resid != nullptr, "igemm: RESID needs resid");
    if (d.resid_spatial == TDX_SP_UP2)
      TDX_REQUIRE(d.height % 2 == 0 && d.width % 2 == 0, "igemm: UP2 residual needs even h, w");
  }
  for (int o = 0; o < 3; ++o) {
    if (d.out[o].kind == TDX_OUT_NONE) continue;
    TDX_REQUIRE(d.out[o].ptr != nullptr, "igemm: out[%d].ptr is null", o);
    TDX_REQUIRE(d.out[o].kind >= 1 && d.out[o].kind <= 3, "igemm: out[%d].kind=%d", o, d.out[o].kind);
    if (d.out[o].spatial == TDX_SP_DOWN2)
      TDX_REQUIRE(d.height % 2 == 0 && d.width % 2 == 0, "igemm: DOWN2 output needs even h, w");
  }
  return TDX_OK;
}

}  // namespace tdx

extern "C" int64_t tdx_igemm_packed_weight_elems(const int32_t* a_channels, const int32_t* a_taps, int32_t n_seg,
                                                 int32_t c_out) {
  int64_t n = 0;
  for (int s = 0; s < n_seg; ++s) n += (int64_t)a_channels[s] * a_taps[s] * c_out;
  return n;
}

extern "C" int tdx_igemm_run(const TdxIgemmDesc* desc, void* stream) {
  if (!desc) {
    tdx::set_error("igemm: null descriptor");
    return TDX_E_INVALID;
  }
  int rc = tdx::igemm_validate(*desc);
  if (rc != TDX_OK) return rc;
  CUtensorMap tms[3];
  for (int s = 0; s < desc->n_seg; ++s) {
    rc = tdx::make_act_tensor_map(&tms[s], desc->a_ptr[s], desc->n_img, desc->a_channels[s], desc->height,
                                  desc->width);
    if (rc != TDX_OK) return rc;
  }
  return tdx::igemm_launch(*desc, tms, reinterpret_cast<cudaStream_t>(stream));
}
