// Elevation read-out ("post step"): the reference does this on the CPU with torchvision
// on every WorldPipeline.get() (world_pipeline.py:1277-1313, data/laplacian_encoder.py:6-137).  Here each primitive is
// one small fp32 kernel, thread per output element, so the read-out stays on the device next to the canvases:
//   normalise-on-read, linear-extrapolation padding, torch's anti-aliased separable bilinear resize (one axis per
//   launch), torchvision's reflect-padded Gaussian blur, and add + crop + signed square (+ int16 pack, api.py:73-77).
// All arithmetic is written with explicitly rounded operations in the order of the CPU restatement (oracle/postproc.py,
// pinned against the reference), so results agree to float32 round-off.  HBM-bound and tiny (a few hundred KB per call).
#include "tdx_common.h"
#include "tdx_ptx.cuh"

namespace tdx {

__global__ void post_normalize_kernel(const float* __restrict__ num, const float* __restrict__ den, long pitch,
                                      float* __restrict__ out, int h, int w, float scale, float offset) {
  pdl_launch_dependents();
  pdl_wait();
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w || y >= h) return;
  const float q = __fdiv_rn(num[(long)y * pitch + x], den[(long)y * pitch + x]);
  out[(long)y * w + x] = __fadd_rn(__fmul_rn(q, scale), offset);
}

// laplacian_encoder.py:6-40: rows are extrapolated first, then the columns of the row-padded tensor.
__device__ __forceinline__ float row_padded(const float* x, int h, int w, int r, int c) {   // r in [-1, h], c in [0, w)
  if (r >= 0 && r < h) return x[(long)r * w + c];
  if (h == 1) return x[c];
  if (r < 0) return __fsub_rn(__fmul_rn(2.0f, x[c]), x[(long)w + c]);
  return __fsub_rn(__fmul_rn(2.0f, x[(long)(h - 1) * w + c]), x[(long)(h - 2) * w + c]);
}
__global__ void post_pad_extrapolate_kernel(const float* __restrict__ x, int h, int w, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y;
  if (ox >= w + 2 || oy >= h + 2) return;
  const int r = oy - 1, c = ox - 1;
  float v;
  if (c >= 0 && c < w) v = row_padded(x, h, w, r, c);
  else if (w == 1) v = row_padded(x, h, w, r, 0);
  else if (c < 0) v = __fsub_rn(__fmul_rn(2.0f, row_padded(x, h, w, r, 0)), row_padded(x, h, w, r, 1));
  else v = __fsub_rn(__fmul_rn(2.0f, row_padded(x, h, w, r, w - 1)), row_padded(x, h, w, r, w - 2));
  out[(long)oy * (w + 2) + ox] = v;
}

// One axis of torch's interpolate(mode="bilinear", antialias=True) (ATen UpSampleKernel.cpp,
// _compute_indices_min_size_weights_aa with the triangle filter): weights are normalised first, then accumulated in
// tap order.  axis 1: out[y][i] over rows of x;  axis 0: out[i][x] over columns.
__global__ void resize_aa_axis_kernel(const float* __restrict__ x, int h, int w, float* __restrict__ out, int out_size,
                                      int axis) {
  pdl_launch_dependents();
  pdl_wait();
  const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y;
  const int oh = axis == 0 ? out_size : h, ow = axis == 1 ? out_size : w;
  if (ox >= ow || oy >= oh) return;
  const int in_size = axis == 1 ? w : h;
  const int i = axis == 1 ? ox : oy;
  const float scale = __fdiv_rn((float)in_size, (float)out_size);
  const float support = scale >= 1.0f ? scale : 1.0f;
  const float invscale = scale >= 1.0f ? __fdiv_rn(1.0f, scale) : 1.0f;
  const float center = __fmul_rn(scale, (float)i + 0.5f);
  int lo = (int)__fadd_rn(__fsub_rn(center, support), 0.5f);
  lo = lo > 0 ? lo : 0;
  int hi = (int)__fadd_rn(__fadd_rn(center, support), 0.5f);
  hi = hi < in_size ? hi : in_size;
  const int n = hi - lo;
  auto weight = [&](int j) {
    float t = __fmul_rn(__fadd_rn(__fsub_rn((float)(j + lo), center), 0.5f), invscale);
    t = t < 0.f ? -t : t;
    return t < 1.0f ? __fsub_rn(1.0f, t) : 0.0f;
  };
  float total = 0.f;
  for (int j = 0; j < n; ++j) total = __fadd_rn(total, weight(j));
  const long stride = axis == 1 ? 1 : w;
  const float* src = axis == 1 ? x + (long)oy * w + lo : x + (long)lo * w + ox;
  float acc = 0.f;
  for (int j = 0; j < n; ++j) {
    float wj = weight(j);
    if (total != 0.f) wj = __fdiv_rn(wj, total);
    const float term = __fmul_rn(src[(long)j * stride], wj);
    acc = j == 0 ? term : __fadd_rn(acc, term);
  }
  out[(long)oy * ow + ox] = acc;
}

struct BlurParams {
  float k[32];
  int ksize;
};
// torchvision gaussian_blur: reflect padding ksize/2, 2-D kernel = outer(k, k), taps accumulated row-major.
__global__ void gaussian_blur_kernel(const float* __restrict__ x, int h, int w, float* __restrict__ out,
                                     const BlurParams bp) {
  pdl_launch_dependents();
  pdl_wait();
  const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y;
  if (ox >= w || oy >= h) return;
  const int p = bp.ksize / 2;
  float acc = 0.f;
  for (int dy = 0; dy < bp.ksize; ++dy) {
    int yy = oy + dy - p;
    yy = yy < 0 ? -yy : (yy >= h ? 2 * h - 2 - yy : yy);
    for (int dx = 0; dx < bp.ksize; ++dx) {
      int xx = ox + dx - p;
      xx = xx < 0 ? -xx : (xx >= w ? 2 * w - 2 - xx : xx);
      acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(bp.k[dy], bp.k[dx]), x[(long)yy * w + xx]));
    }
  }
  out[(long)oy * w + ox] = acc;
}

__global__ void post_combine_kernel(const float* __restrict__ a, long a_pitch, const float* __restrict__ b, long b_pitch,
                                    float* __restrict__ out, int16_t* __restrict__ out_i16, int h, int w,
                                    int signed_square) {
  pdl_launch_dependents();
  pdl_wait();
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w || y >= h) return;
  float v = __fadd_rn(a[(long)y * a_pitch + x], b[(long)y * b_pitch + x]);
  if (signed_square) v = v == 0.f ? 0.f : copysignf(__fmul_rn(v, v), v);   // sign(v) * v^2 (world_pipeline.py:1312)
  if (out) out[(long)y * w + x] = v;
  if (out_i16) {
    float f = floorf(v);                                                   // api.py:73-77: floor, clip, '<i2'
    f = fminf(fmaxf(f, -32768.f), 32767.f);
    out_i16[(long)y * w + x] = (int16_t)f;
  }
}

static int launch2d(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, int h, int w, cudaStream_t stream) {
  fill_launch_config(cfg, attr, dim3((w + 127) / 128, h), dim3(128), 0, stream);
  return TDX_OK;
}

}  // namespace tdx

using namespace tdx;

extern "C" int tdx_post_normalize(const float* num, const float* den, int64_t pitch, float* out, int32_t h, int32_t w,
                                  float scale, float offset, void* stream) {
  TDX_REQUIRE(num && den && out, "post_normalize: null pointer");
  TDX_REQUIRE(h >= 1 && w >= 1 && h <= 65535 && pitch >= w, "post_normalize: bad shape %d x %d (pitch %lld)", h, w,
              (long long)pitch);
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  launch2d(&cfg, attr, h, w, reinterpret_cast<cudaStream_t>(stream));
  TDX_CHECK_CUDA(cudaLaunchKernelEx(&cfg, post_normalize_kernel, num, den, (long)pitch, out, (int)h, (int)w, scale,
                                    offset));
  return TDX_OK;
}

extern "C" int tdx_post_pad_extrapolate(const float* x, int32_t h, int32_t w, float* out, void* stream) {
  TDX_REQUIRE(x && out, "post_pad_extrapolate: null pointer");
  TDX_REQUIRE(h >= 1 && w >= 1 && h + 2 <= 65535, "post_pad_extrapolate: bad shape %d x %d", h, w);
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  launch2d(&cfg, attr, h + 2, w + 2, reinterpret_cast<cudaStream_t>(stream));
  TDX_CHECK_CUDA(cudaLaunchKernelEx(&cfg, post_pad_extrapolate_kernel, x, (int)h, (int)w, out));
  return TDX_OK;
}

extern "C" int tdx_resize_aa_axis(const float* x, int32_t h, int32_t w, float* out, int32_t out_size, int32_t axis,
                                  void* stream) {
  TDX_REQUIRE(x && out, "resize_aa_axis: null pointer");
  TDX_REQUIRE(axis == 0 || axis == 1, "resize_aa_axis: axis=%d (0 = height, 1 = width)", axis);
  TDX_REQUIRE(h >= 1 && w >= 1 && out_size >= 1, "resize_aa_axis: bad shape %d x %d -> %d", h, w, out_size);
  const int oh = axis == 0 ? out_size : h, ow = axis == 1 ? out_size : w;
  TDX_REQUIRE(oh <= 65535, "resize_aa_axis: more than 65535 output rows");
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  launch2d(&cfg, attr, oh, ow, reinterpret_cast<cudaStream_t>(stream));
  TDX_CHECK_CUDA(cudaLaunchKernelEx(&cfg, resize_aa_axis_kernel, x, (int)h, (int)w, out, (int)out_size, (int)axis));
  return TDX_OK;
}

extern "C" int tdx_gaussian_blur(const float* x, int32_t h, int32_t w, float* out, int32_t ksize, float sigma,
                                 void* stream) {
  TDX_REQUIRE(x && out && x != out, "gaussian_blur: null or aliased pointers");
  TDX_REQUIRE(ksize >= 1 && ksize <= 31 && (ksize & 1), "gaussian_blur: kernel size %d (odd, <= 31)", ksize);
  TDX_REQUIRE(sigma > 0.f, "gaussian_blur: sigma must be positive");
  TDX_REQUIRE(h > ksize / 2 && w > ksize / 2 && h <= 65535,
              "gaussian_blur: reflect padding of %d needs both dimensions larger than that (got %d x %d)", ksize / 2, h, w);
  BlurParams bp;
  bp.ksize = ksize;
  // torchvision _get_gaussian_kernel1d in float32: taps on linspace(-(k-1)/2, (k-1)/2, k), normalised by their sum
  const float half = (ksize - 1) * 0.5f;
  float sum = 0.f;
  for (int i = 0; i < ksize; ++i) {
    const float t = (-half + (float)i) / sigma;
    bp.k[i] = expf(-0.5f * (t * t));
    sum += bp.k[i];
  }
  for (int i = 0; i < ksize; ++i) bp.k[i] /= sum;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  launch2d(&cfg, attr, h, w, reinterpret_cast<cudaStream_t>(stream));
  TDX_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gaussian_blur_kernel, x, (int)h, (int)w, out, bp));
  return TDX_OK;
}

extern "C" int tdx_post_combine(const float* a, int64_t a_pitch, const float* b, int64_t b_pitch, float* out,
                                int16_t* out_i16, int32_t h, int32_t w, int32_t signed_square, void* stream) {
  TDX_REQUIRE(a && b && (out || out_i16), "post_combine: null pointer");
  TDX_REQUIRE(h >= 1 && w >= 1 && h <= 65535 && a_pitch >= w && b_pitch >= w, "post_combine: bad shape %d x %d", h, w);
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  launch2d(&cfg, attr, h, w, reinterpret_cast<cudaStream_t>(stream));
  TDX_CHECK_CUDA(cudaLaunchKernelEx(&cfg, post_combine_kernel, a, (long)a_pitch, b, (long)b_pitch, out, out_i16, (int)h,
                                    (int)w, (int)signed_square));
  return TDX_OK;
}
