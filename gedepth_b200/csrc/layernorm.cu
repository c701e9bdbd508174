// LayerNorm over the channel dimension of a token matrix (rows x C), sm_100a.
// a5/a6/a9/a10: nn.LayerNorm(eps=1e-5) sites at embed.py:299-300, depthformer_swin.py:118,463,469,1178.
// One warp per row, two-pass statistics (mean, then centred variance - same numerics class as ATen's
// Welford), float4 accesses; mean / rstd are kept for the backward.  HBM-bound: 8 B/element fwd.
#include "common.cuh"

namespace ged {

__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x,
                                                             const float* __restrict__ w,
                                                             const float* __restrict__ b,
                                                             float* __restrict__ y, float* __restrict__ mean,
                                                             float* __restrict__ rstd, int64_t rows, int C,
                                                             float eps) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = (const float4*)(x + row * C);
  const int C4 = C >> 2;
  float s = 0.f;
  for (int i = lane; i < C4; i += 32) { float4 v = __ldg(xr + i); s += (v.x + v.y) + (v.z + v.w); }
  const float mu = warp_sum(s) / (float)C;
  float q = 0.f;
  for (int i = lane; i < C4; i += 32) {
    float4 v = __ldg(xr + i);
    float a = v.x - mu, b2 = v.y - mu, c = v.z - mu, d = v.w - mu;
    q += (a * a + b2 * b2) + (c * c + d * d);
  }
  const float rs = rsqrtf(warp_sum(q) / (float)C + eps);
  if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
  float4* yr = (float4*)(y + row * C);
  for (int i = lane; i < C4; i += 32) {
    float4 v = __ldg(xr + i), g = __ldg((const float4*)w + i), bb = __ldg((const float4*)b + i);
    float4 o;
    o.x = (v.x - mu) * rs * g.x + bb.x; o.y = (v.y - mu) * rs * g.y + bb.y;
    o.z = (v.z - mu) * rs * g.z + bb.z; o.w = (v.w - mu) * rs * g.w + bb.w;
    yr[i] = o;
  }
}

// dx = rstd * (g*w - mean_c(g*w) - xhat * mean_c(g*w*xhat))
__global__ void __launch_bounds__(256) layernorm_bwd_dx_kernel(
    const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ w,
    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ g_add,
    float* __restrict__ dx, int64_t rows, int C) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = (const float4*)(x + row * C);
  const float4* gr = (const float4*)(g + row * C);
  const int C4 = C >> 2;
  const float mu = mean[row], rs = rstd[row];
  float s1 = 0.f, s2 = 0.f;
  for (int i = lane; i < C4; i += 32) {
    float4 v = __ldg(xr + i), gg = __ldg(gr + i), ww = __ldg((const float4*)w + i);
    float a = gg.x * ww.x, b2 = gg.y * ww.y, c = gg.z * ww.z, d = gg.w * ww.w;
    s1 += (a + b2) + (c + d);
    s2 += (a * (v.x - mu) + b2 * (v.y - mu)) + (c * (v.z - mu) + d * (v.w - mu));
  }
  s1 = warp_sum(s1) / (float)C;
  s2 = warp_sum(s2) * rs / (float)C;      // mean(g*w*xhat)
  float4* dr = (float4*)(dx + row * C);
  for (int i = lane; i < C4; i += 32) {
    float4 v = __ldg(xr + i), gg = __ldg(gr + i), ww = __ldg((const float4*)w + i);
    float4 o;
    o.x = rs * (gg.x * ww.x - s1 - (v.x - mu) * rs * s2);
    o.y = rs * (gg.y * ww.y - s1 - (v.y - mu) * rs * s2);
    o.z = rs * (gg.z * ww.z - s1 - (v.z - mu) * rs * s2);
    o.w = rs * (gg.w * ww.w - s1 - (v.w - mu) * rs * s2);
    if (g_add) {     // gradient arriving at x through the residual path (x is also the block's identity)
      const float4 r = __ldg((const float4*)(g_add + row * C) + i);
      o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
    }
    dr[i] = o;
  }
}

// dw[c] += sum_rows g*xhat ; db[c] += sum_rows g.  grid (C/32, row chunks), block (32, 8).
__global__ void __launch_bounds__(256) layernorm_bwd_wb_kernel(
    const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ mean,
    const float* __restrict__ rstd, float* __restrict__ dw, float* __restrict__ db, int64_t rows, int C,
    int rows_per_block) {
  __shared__ float s_w[8][33], s_b[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  float aw = 0.f, ab = 0.f;
  if (c < C) {
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
      const float gg = __ldg(g + r * C + c);
      aw += gg * (__ldg(x + r * C + c) - __ldg(mean + r)) * __ldg(rstd + r);
      ab += gg;
    }
  }
  s_w[threadIdx.y][threadIdx.x] = aw; s_b[threadIdx.y][threadIdx.x] = ab;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    for (int i = 1; i < 8; ++i) { aw += s_w[i][threadIdx.x]; ab += s_b[i][threadIdx.x]; }
    atomicAdd(dw + c, aw);
    atomicAdd(db + c, ab);
  }
}

// ---- rows held in registers (C <= 768) ------------------------------------------------------------------------------------
// The kernels above read a row three times (mean, centred variance, output - the later passes from L1) as three dependent
// round trips, and the backward needs a second kernel that re-reads g and x for the weight / bias gradients.  For
// C <= 32 * 4 * NV a lane keeps its NV float4 of the row in registers: ONE global read per operand, all of a row's loads in
// flight at once, the same summation order (bit-identical y / mean / rstd / dx); the backward accumulates dw / db per lane
// over the rows its warp walks and reduces them once per CTA.  (Measured and rejected for C = 768: re-reading w and g_add late
// instead of holding them - 171 -> 127 registers, two CTAs per SM - 4.95 -> 5.4 ms per step.)
template <int NV>
__global__ void __launch_bounds__(256) layernorm_fwd_reg_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                 const float* __restrict__ b, float* __restrict__ y,
                                                                 float* __restrict__ mean, float* __restrict__ rstd,
                                                                 int64_t rows, int C, float eps) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = (const float4*)(x + row * C);
  const int C4 = C >> 2;
  float4 v[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = lane + 32 * k;
    v[k] = i < C4 ? __ldg(xr + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) if (lane + 32 * k < C4) s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
  const float mu = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    if (lane + 32 * k < C4) {
      const float a = v[k].x - mu, b2 = v[k].y - mu, c = v[k].z - mu, d = v[k].w - mu;
      q += (a * a + b2 * b2) + (c * c + d * d);
    }
  }
  const float rs = rsqrtf(warp_sum(q) / (float)C + eps);
  if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
  float4* yr = (float4*)(y + row * C);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = lane + 32 * k;
    if (i < C4) {
      const float4 g = __ldg((const float4*)w + i), bb = __ldg((const float4*)b + i);
      float4 o;
      o.x = (v[k].x - mu) * rs * g.x + bb.x; o.y = (v[k].y - mu) * rs * g.y + bb.y;
      o.z = (v[k].z - mu) * rs * g.z + bb.z; o.w = (v[k].w - mu) * rs * g.w + bb.w;
      yr[i] = o;
    }
  }
}

// grid-stride over rows (warp granularity); dw / db (may be NULL together) accumulated per lane, reduced over the CTA's 8
// warps in shared memory, one atomicAdd per column per CTA.
template <int NV>
__global__ void __launch_bounds__(256) layernorm_bwd_reg_kernel(
    const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ w,
    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ g_add,
    float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db, int64_t rows, int C) {
  __shared__ float4 s_red[8][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C4 = C >> 2;
  float4 ww[NV], aw[NV], ab[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = lane + 32 * k;
    ww[k] = i < C4 ? __ldg((const float4*)w + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    aw[k] = ab[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < rows; row += (int64_t)gridDim.x * 8) {
    const float4* xr = (const float4*)(x + row * C);
    const float4* gr = (const float4*)(g + row * C);
    float4 v[NV], gg[NV], ga[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int i = lane + 32 * k;
      const bool ok = i < C4;
      v[k] = ok ? __ldg(xr + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      gg[k] = ok ? __ldg(gr + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      ga[k] = (ok && g_add) ? __ldg((const float4*)(g_add + row * C) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      if (lane + 32 * k < C4) {
        const float a = gg[k].x * ww[k].x, b2 = gg[k].y * ww[k].y, c = gg[k].z * ww[k].z, d = gg[k].w * ww[k].w;
        s1 += (a + b2) + (c + d);
        s2 += (a * (v[k].x - mu) + b2 * (v[k].y - mu)) + (c * (v[k].z - mu) + d * (v[k].w - mu));
      }
    }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) * rs / (float)C;      // mean(g*w*xhat)
    float4* dr = (float4*)(dx + row * C);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int i = lane + 32 * k;
      if (i < C4) {
        const float hx = (v[k].x - mu) * rs, hy = (v[k].y - mu) * rs, hz = (v[k].z - mu) * rs, hw = (v[k].w - mu) * rs;
        float4 o;
        o.x = rs * (gg[k].x * ww[k].x - s1 - (v[k].x - mu) * rs * s2);
        o.y = rs * (gg[k].y * ww[k].y - s1 - (v[k].y - mu) * rs * s2);
        o.z = rs * (gg[k].z * ww[k].z - s1 - (v[k].z - mu) * rs * s2);
        o.w = rs * (gg[k].w * ww[k].w - s1 - (v[k].w - mu) * rs * s2);
        if (g_add) { o.x += ga[k].x; o.y += ga[k].y; o.z += ga[k].z; o.w += ga[k].w; }
        dr[i] = o;
        aw[k].x += gg[k].x * hx; aw[k].y += gg[k].y * hy; aw[k].z += gg[k].z * hz; aw[k].w += gg[k].w * hw;
        ab[k].x += gg[k].x; ab[k].y += gg[k].y; ab[k].z += gg[k].z; ab[k].w += gg[k].w;
      }
    }
  }
  if (dw == nullptr) return;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = lane + 32 * k;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      __syncthreads();
      s_red[warp][lane] = which ? ab[k] : aw[k];
      __syncthreads();
      if (warp == 0 && i < C4) {
        float4 t = s_red[0][lane];
#pragma unroll
        for (int j = 1; j < 8; ++j) { const float4 u = s_red[j][lane]; t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
        float* dst = (which ? db : dw) + 4 * i;
        atomicAdd(dst, t.x); atomicAdd(dst + 1, t.y); atomicAdd(dst + 2, t.z); atomicAdd(dst + 3, t.w);
      }
    }
  }
}

}  // namespace ged
using namespace ged;

static int g_ln_reg = 1;     // 0: the three-pass / two-kernel forms (A/B, tests)
GED_API int ged_set_layernorm_reg(int on) { const int prev = g_ln_reg; g_ln_reg = on ? 1 : 0; return prev; }

GED_API int ged_layernorm_fwd(const float* x, const float* w, const float* b, float* y, float* mean,
                              float* rstd, int64_t rows, int C, float eps, cudaStream_t stream) {
  if (!x || !w || !b || !y || !mean || !rstd || rows <= 0) return GED_ERR_ARG;
  if (C % 4 != 0) return GED_ERR_SHAPE;
  if (!aligned16(x) || !aligned16(y) || !aligned16(w) || !aligned16(b)) return GED_ERR_ALIGN;
  const unsigned blocks = (unsigned)((rows + 7) / 8);
  const int nv = (C / 4 + 31) / 32;
  if (!g_ln_reg || nv > 6) layernorm_fwd_kernel<<<blocks, 256, 0, stream>>>(x, w, b, y, mean, rstd, rows, C, eps);
  else if (nv == 1) layernorm_fwd_reg_kernel<1><<<blocks, 256, 0, stream>>>(x, w, b, y, mean, rstd, rows, C, eps);
  else if (nv == 2) layernorm_fwd_reg_kernel<2><<<blocks, 256, 0, stream>>>(x, w, b, y, mean, rstd, rows, C, eps);
  else if (nv == 3) layernorm_fwd_reg_kernel<3><<<blocks, 256, 0, stream>>>(x, w, b, y, mean, rstd, rows, C, eps);
  else if (nv == 4) layernorm_fwd_reg_kernel<4><<<blocks, 256, 0, stream>>>(x, w, b, y, mean, rstd, rows, C, eps);
  else layernorm_fwd_reg_kernel<6><<<blocks, 256, 0, stream>>>(x, w, b, y, mean, rstd, rows, C, eps);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

// dw / db are ACCUMULATED into (caller zeroes or passes the running .grad buffers).
// g_add (optional, [rows][C]): added to dx - the gradient that reaches x through the residual branch when x is both
// the LayerNorm input and the block's identity (depthformer_swin.py:461-472), saving autograd's separate sum pass.
GED_API int ged_layernorm_bwd(const float* g, const float* x, const float* w, const float* mean,
                              const float* rstd, const float* g_add, float* dx, float* dw, float* db, int64_t rows,
                              int C, cudaStream_t stream) {
  if (!g || !x || !w || !mean || !rstd || !dx || rows <= 0) return GED_ERR_ARG;
  if (C % 4 != 0) return GED_ERR_SHAPE;
  if (!aligned16(x) || !aligned16(g) || !aligned16(dx) || !aligned16(w) || (g_add && !aligned16(g_add))) return GED_ERR_ALIGN;
  const int nv = (C / 4 + 31) / 32;
  if (g_ln_reg && nv <= 6 && ((dw && db) || (!dw && !db))) {
    // a warp walks several rows so that the per-lane dw / db partials amortise their reduction; the cap is two waves of
    // what the register footprint lets an SM hold (NV = 6: 171 registers, one CTA per SM)
    const int64_t cap = 148 * (nv <= 2 ? 6 : (nv <= 4 ? 4 : 2));
    const unsigned blocks = (unsigned)((rows + 7) / 8 < cap ? (rows + 7) / 8 : cap);
#define LN_BWD(NV) layernorm_bwd_reg_kernel<NV><<<blocks, 256, 0, stream>>>(g, x, w, mean, rstd, g_add, dx, dw, db, rows, C)
    if (nv == 1) LN_BWD(1); else if (nv == 2) LN_BWD(2); else if (nv == 3) LN_BWD(3); else if (nv == 4) LN_BWD(4); else LN_BWD(6);
#undef LN_BWD
    GED_CHECK_LAUNCH();
    return GED_OK;
  }
  layernorm_bwd_dx_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(g, x, w, mean, rstd, g_add, dx, rows, C);
  if (dw && db) {
    const int rpb = 256;
    dim3 grid(cdiv(C, 32), (unsigned)((rows + rpb - 1) / rpb));
    layernorm_bwd_wb_kernel<<<grid, dim3(32, 8), 0, stream>>>(g, x, mean, rstd, dw, db, rows, C, rpb);
  }
  GED_CHECK_LAUNCH();
  return GED_OK;
}
