/* ORACLE (test infrastructure, not product): plain-C restatement of the reference's portable noise generator.
 *
 * Reference (xandergos/terrain-diffusion @ 82a0431):
 *   _pcg64_next / next_seed                 terrain_diffusion/inference/portable_rng.py:24-42
 *   _fill_standard_normal_impl              terrain_diffusion/inference/portable_rng.py:55-74   (numba, fp64 math)
 *   _tile_seed / gaussian_noise_patch       terrain_diffusion/inference/world_pipeline.py:58-115
 *
 * PCG-XSH-RR 64/32 (64-bit LCG state, 32-bit output) + Marsaglia polar; U in (0,1], V = 2U-1, accept 0 < S < 1,
 * X = V*sqrt(-2 ln S / S); both values of an accepted pair are emitted (the second only if there is room).
 * Pinned by tests/test_oracle_golden.py against SURVEY Appendix E vectors and vectors generated from the reference's
 * numba implementation (tests/golden/make_golden.py).
 *
 * Build: make -C oracle   (gcc -O2 -shared -fPIC; -ffp-contract=off so no FMA contraction changes fp64 results)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PCG_MULT 6364136223846793005ULL
#define PCG_INC 1442695040888963407ULL

static inline uint32_t pcg_next(uint64_t* state) {
  *state = *state * PCG_MULT + PCG_INC;
  uint64_t s = *state;
  uint32_t x = (uint32_t)(((s >> 18) ^ s) >> 27);
  uint32_t rot = (uint32_t)(s >> 59);
  return (x >> rot) | (x << ((32 - rot) & 31));
}

uint64_t oracle_next_seed(uint64_t seed) {
  uint64_t st = seed;
  uint64_t lo = pcg_next(&st);
  uint64_t hi = pcg_next(&st);
  return (hi << 32) | lo;
}

uint64_t oracle_tile_seed(uint64_t base_seed, int64_t ty, int64_t tx) {
  uint64_t h = base_seed * 0x9E3779B9ULL;
  h = h + ((uint64_t)ty & 0xFFFFFFFFULL);
  h = h * 0x9E3779B9ULL + ((uint64_t)tx & 0xFFFFFFFFULL);
  return h;
}

/* Fill out[0..n) with standard normals; as_f32 != 0 stores float, else double. */
void oracle_fill_standard_normal(uint64_t seed, void* out, int64_t n, int as_f32) {
  uint64_t state = seed;
  const double inv_2p32 = 1.0 / 4294967296.0;
  int64_t i = 0;
  float* of = (float*)out;
  double* od = (double*)out;
  while (i < n) {
    uint32_t u1 = pcg_next(&state);
    uint32_t u2 = pcg_next(&state);
    double v1 = 2.0 * ((double)u1 + 1.0) * inv_2p32 - 1.0;
    double v2 = 2.0 * ((double)u2 + 1.0) * inv_2p32 - 1.0;
    double s = v1 * v1 + v2 * v2;
    if (s > 0.0 && s < 1.0) {
      double f = sqrt(-2.0 * log(s) / s);
      if (as_f32) of[i] = (float)(v1 * f); else od[i] = v1 * f;
      i++;
      if (i < n) {
        if (as_f32) of[i] = (float)(v2 * f); else od[i] = v2 * f;
        i++;
      }
    }
  }
}

static int64_t floordiv(int64_t a, int64_t b) {
  int64_t q = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) q--;
  return q;
}

/* (C, h, w) fp32 patch at integer origin (y0, x0) of the infinite tile-seeded field; every overlapped
 * (tile_h x tile_w x C) tile is regenerated in full, exactly as the reference does. */
int oracle_gaussian_noise_patch(uint64_t base_seed, int64_t y0, int64_t x0, int64_t h, int64_t w, int64_t channels,
                                int64_t tile_h, int64_t tile_w, float* out) {
  float* tile = (float*)malloc(sizeof(float) * (size_t)(channels * tile_h * tile_w));
  if (!tile) return -1;
  int64_t ty0 = floordiv(y0, tile_h), ty1 = floordiv(y0 + h - 1, tile_h);
  int64_t tx0 = floordiv(x0, tile_w), tx1 = floordiv(x0 + w - 1, tile_w);
  for (int64_t ty = ty0; ty <= ty1; ++ty) {
    for (int64_t tx = tx0; tx <= tx1; ++tx) {
      int64_t tyo = ty * tile_h, txo = tx * tile_w;
      int64_t oy0 = y0 > tyo ? y0 : tyo, oy1 = (y0 + h) < (tyo + tile_h) ? (y0 + h) : (tyo + tile_h);
      int64_t ox0 = x0 > txo ? x0 : txo, ox1 = (x0 + w) < (txo + tile_w) ? (x0 + w) : (txo + tile_w);
      oracle_fill_standard_normal(oracle_tile_seed(base_seed, ty, tx), tile, channels * tile_h * tile_w, 1);
      for (int64_t c = 0; c < channels; ++c)
        for (int64_t y = oy0; y < oy1; ++y)
          memcpy(out + (c * h + (y - y0)) * w + (ox0 - x0), tile + (c * tile_h + (y - tyo)) * tile_w + (ox0 - txo),
                 sizeof(float) * (size_t)(ox1 - ox0));
    }
  }
  free(tile);
  return 0;
}
