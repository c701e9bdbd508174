// 2-D output tiling shared by the per-pixel kernels: a (TILE_H x TILE_W) tile of full-resolution
// pixels per CTA, the ~half-resolution operands staged in shared memory, and the deterministic
// in-tile adjoint of the bilinear upsample.
#pragma once
#include "common.cuh"

namespace ged {

constexpr int TILE_W = 256;   // output pixels per tile row  (64 threads x float4)
constexpr int TILE_H = 4;     // output rows per tile        (4 thread rows)
constexpr int TX = TILE_W / 4;
constexpr int NSLOPE = 11;    // slope bins -5..5 degrees (encoder_decoder.py:68)
// half-resolution staging tile: scale ~ 1/2 plus one halo tap on each side, with slack
constexpr int ST_W = TILE_W / 2 + 6;
constexpr int ST_H = TILE_H / 2 + 4;

struct SrcWindow {
  int y0, x0, h, w;   // origin and extent of the staged source tile
};

__device__ __forceinline__ SrcWindow src_window(int oy0, int ox0, int H, int W, int h2, int w2,
                                                float sy, float sx, bool align) {
  SrcWindow s;
  int oy1 = min(oy0 + TILE_H, H) - 1, ox1 = min(ox0 + TILE_W, W) - 1;
  s.y0 = tap(oy0, sy, align, h2).i0;
  s.x0 = tap(ox0, sx, align, w2).i0;
  s.h = tap(oy1, sy, align, h2).i1 - s.y0 + 1;
  s.w = tap(ox1, sx, align, w2).i1 - s.x0 + 1;
  return s;
}

// Adjoint of the bilinear upsample restricted to one tile: s_g holds per-pixel gradients of the
// tile's outputs (C channels); every half-resolution pixel touched by the tile gathers its share
// deterministically and adds it to global memory (interior pixels get exactly one add).
template <int C>
__device__ __forceinline__ void tile_adjoint_upsample(
    const float (*s_g)[TILE_H][TILE_W + 1], float* __restrict__ g_half, int64_t chan_stride, int oy0,
    int ox0, int H, int W, int h2, int w2, float sy, float sx, bool align, const SrcWindow& sw) {
  const int tid = threadIdx.y * TX + threadIdx.x;
  const int th = min(TILE_H, H - oy0), tw = min(TILE_W, W - ox0);
  for (int i = tid; i < sw.h * sw.w; i += TX * TILE_H) {
    const int r = i / sw.w, c = i - r * sw.w;
    const int j = sw.y0 + r, k = sw.x0 + c;
    // weights of source row j in each of the tile's (<= TILE_H) output rows: evaluated once, not per column
    float wyv[TILE_H];
    bool any_y = false;
#pragma unroll
    for (int yy = 0; yy < TILE_H; ++yy) {
      wyv[yy] = 0.f;
      if (yy < th) {
        const Tap ty = tap(oy0 + yy, sy, align, h2);
        wyv[yy] = (ty.i0 == j ? ty.l0 : 0.f) + (ty.i1 == j ? ty.l1 : 0.f);
      }
      any_y |= wyv[yy] != 0.f;
    }
    if (!any_y) continue;
    int xlo, xhi;
    adjoint_range(k, sx, align, w2, W, xlo, xhi);
    xlo = max(xlo, ox0) - ox0; xhi = min(xhi, ox0 + tw - 1) - ox0;
    float acc[C];
#pragma unroll
    for (int ch = 0; ch < C; ++ch) acc[ch] = 0.f;
    bool any = false;
    for (int xx = xlo; xx <= xhi; ++xx) {
      const Tap tx = tap(ox0 + xx, sx, align, w2);
      const float wx = (tx.i0 == k ? tx.l0 : 0.f) + (tx.i1 == k ? tx.l1 : 0.f);
      if (wx == 0.f) continue;
      any = true;
#pragma unroll
      for (int yy = 0; yy < TILE_H; ++yy) {
        const float w = wyv[yy] * wx;
        if (w != 0.f) {
#pragma unroll
          for (int ch = 0; ch < C; ++ch) acc[ch] += w * s_g[ch][yy][xx];
        }
      }
    }
    if (any) {
#pragma unroll
      for (int ch = 0; ch < C; ++ch) atomicAdd(g_half + ch * chan_stride + (int64_t)j * w2 + k, acc[ch]);
    }
  }
}


// a15 per-pixel evaluation (encoder_decoder.py:84-100): softmax over the 11 slope bins, expected slope, tan,
// inverse-depth shift, range mask.
struct SlopeEval {
  float p[NSLOPE];
  float theta, k, den, off, m;
};

__device__ __forceinline__ void slope_eval(const float* L, float pe, float h, float depth_scale,
                                           SlopeEval& e) {
  float mx = L[0];
#pragma unroll
  for (int c = 1; c < NSLOPE; ++c) mx = fmaxf(mx, L[c]);
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < NSLOPE; ++c) { e.p[c] = __expf(L[c] - mx); sum += e.p[c]; }
  const float inv = 1.f / sum;
  float th = 0.f;
#pragma unroll
  for (int c = 0; c < NSLOPE; ++c) { e.p[c] *= inv; th += e.p[c] * (float)(c - 5); }
  e.theta = th;
  // |theta| <= 5 degrees (a convex combination of the bins): tan by its series, exact to fp32 for |x| <= 0.0873
  const float xr = th * 0.017453292519943295f, x2 = xr * xr;
  e.k = xr * (1.f + x2 * (0.33333333333f + x2 * (0.13333333333f + x2 * 0.05396825397f)));
  const float a = -h / (pe + 1e-8f);
  e.den = (a - e.k) + 1e-8f;
  e.off = -h / e.den;
  // in-place thresholds of encoder_decoder.py:97-100: <0 -> 0, >depth_scale -> 0, >0 -> 1
  // (a NaN offset survives all three and poisons the pixel exactly as in the reference)
  float mm = e.off;
  if (mm < 0.f) mm = 0.f;
  if (mm > depth_scale) mm = 0.f;
  if (mm > 0.f) mm = 1.f;
  e.m = mm;
}


}  // namespace ged
