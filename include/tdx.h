/*
 * tdx.h -- C ABI of the B200-native InfiniteDiffusion sampling hot path (libtdx.so).
 *
 * Plain C: pointers, sizes, POD structs.  No torch / C++ types cross this boundary.  Every device pointer is owned
 * by the caller unless a handle says otherwise; every launch goes to the `stream` argument (a CUstream/cudaStream_t
 * passed as void*).  All functions return 0 on success or a negative TDX_E_* code; tdx_last_error() returns a
 * human-readable message for the calling thread.  There is no CPU fallback anywhere behind this header.
 *
 * Each entry point cites the reference interface (xandergos/terrain-diffusion @ 82a0431) it replaces.
 * The reference has no FFI of its own (pure Python); INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Activation layout ("NC8HW8"): a [N, C, H, W] activation is stored as bf16 [N][C/8][H][W][8] -- channel groups of 8
 * (16 bytes) are the innermost unit, so a TMA box of (W-run x rows x groups) lands in shared memory exactly in the
 * tcgen05 K-major no-swizzle core-matrix order, and one thread's 8-channel store is a 16-byte vector.
 */
#ifndef TDX_H_
#define TDX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDX_OK 0
#define TDX_E_INVALID -1   /* bad argument / unsupported shape */
#define TDX_E_CUDA -2      /* CUDA runtime / driver error */
#define TDX_E_UNSUPPORTED -3

const char* tdx_last_error(void);
/* Library + device probe: fills sm count, compute capability; fails if the device is not sm_100. */
int tdx_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* sizeof() of the public structs (0: TdxOutSpec, 1: TdxIgemmDesc, ...) so bindings can verify their layout. */
int tdx_abi_sizeof(int which);

/* ------------------------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution on tcgen05 (the operator under MPConv.forward, models/mp_layers.py:201-221, with the
 * surrounding UNetBlock elementwise math, models/unet_block.py:116-156, fused into its epilogue).
 *
 * out[m, n] = sum over segments s, taps (r,c), channels k of  A_s[pixel(m)+(r-1,c-1), k] * B[n, s, k, r, c]
 *   m: output pixel inside a 16x8 tile (M = 128), n: output channel (N = Cout, whole), K = sum_s taps_s * C_s.
 * Up to 3 K-segments (e.g. the two halves of an mp_concat, or a 3x3 residual conv + a 1x1 skip conv fused as extra K).
 * ------------------------------------------------------------------------------------------------------------------ */
enum { TDX_OUT_NONE = 0, TDX_OUT_RAW = 1, TDX_OUT_SILU = 2, TDX_OUT_PNORM_SILU = 3 };
enum { TDX_SP_SAME = 0, TDX_SP_DOWN2 = 1, TDX_SP_UP2 = 2 };
enum { TDX_EPI_EMB_SILU = 1, TDX_EPI_RESID = 2, TDX_EPI_PNORM = 4 };

typedef struct TdxOutSpec {
  void* ptr;      /* bf16 NC8HW8, Cout channels; spatial size per `spatial` */
  int32_t kind;   /* TDX_OUT_*: raw v | mp_silu(scale*v) | mp_silu(pixelnorm(v)) */
  int32_t spatial;/* TDX_SP_*: same HxW | every 2nd pixel into H/2 x W/2 | nearest x2 into 2H x 2W */
  float scale;
  int32_t _pad;
} TdxOutSpec;

typedef struct TdxIgemmDesc {
  /* A operand: up to 3 activation tensors (bf16 NC8HW8, all n_img x C_s x H x W) */
  const void* a_ptr[3];
  int32_t a_channels[3];   /* multiple of 64 */
  int32_t a_taps[3];       /* 9 (3x3, pad 1) or 1 (1x1) */
  int32_t n_seg;
  /* B operand: packed bf16 weights, stage order (segment, 64-channel chunk, tap), each stage [8][Cout][8] */
  const void* b_packed;
  int32_t c_out;           /* multiple of 32, <= 256 */
  int32_t n_img, height, width;   /* output == input spatial size; multiples of 8 */
  /* epilogue */
  int32_t epi_flags;       /* TDX_EPI_* */
  const float* cvec;       /* [n_img][c_out] fp32 embedding scale (TDX_EPI_EMB_SILU): v = mp_silu(v * c)          */
  const void* resid;       /* bf16 NC8HW8 residual (TDX_EPI_RESID): v = clip(v + resid_scale * r', +-clip)       */
  int32_t resid_spatial;   /* TDX_SP_SAME | TDX_SP_UP2 (r is H/2 x W/2) | TDX_SP_DOWN2 (r is 2H x 2W)            */
  int32_t resid_pnorm;     /* r' = pixelnorm(r) over channels (unet_block.py:121) when non-zero                  */
  float resid_scale;
  float clip;
  TdxOutSpec out[3];
} TdxIgemmDesc;

/* Bytes of packed B for a descriptor's segments. */
int64_t tdx_igemm_packed_weight_elems(const int32_t* a_channels, const int32_t* a_taps, int32_t n_seg, int32_t c_out);
/* One launch of the persistent tcgen05 kernel. */
int tdx_igemm_run(const TdxIgemmDesc* desc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TDX_H_ */
