// Data-movement kernels around the tensor-core convs, NHWC fp32, sm_100a.  All HBM-bound, 128-bit accesses.
//
//   ged_prep_conv_input   ONE pass that builds the zero-bordered input of a 3x3 conv from up to two sources:
//                         [ bilinear(src0 -> HxW, align_corners=True) | src1 ] along channels.  Replaces
//                         F.interpolate + torch.cat + zero padding (three full-tensor passes) of
//                         densedepth_head.py:24-27 (UpSample) and the cat of hahi.py:329-353.
//   ged_upsample_nhwc_bwd adjoint of that bilinear resize for the first C0 channels of dX.
//   ged_act_bwd           gz = g * act'(.) * row_scale  and  db += column sums of gz, one pass
//                         (ReLU / LeakyReLU / sigmoid from the output, exact GELU from the pre-activation).
//   ged_resize_add_nhwc   acc += bilinear(t -> HxW, align_corners=True)   (pemask_neck.py:52-63)
#include "common.cuh"

namespace ged {

__global__ void __launch_bounds__(256) prep_conv_input_kernel(
    const float* __restrict__ src0, int C0, int h0, int w0, const float* __restrict__ src1, int C1,
    float* __restrict__ dst, int B, int H, int W, float sy, float sx, int64_t bs0, int64_t bs1) {
  const int C = C0 + C1, C4 = C >> 2;
  const int64_t total = (int64_t)B * (H + 2) * (W + 2) * C4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    int64_t p = i / C4;
    const int xp = (int)(p % (W + 2)); p /= (W + 2);
    const int yp = (int)(p % (H + 2));
    const int b = (int)(p / (H + 2));
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (xp > 0 && xp <= W && yp > 0 && yp <= H) {
      const int x = xp - 1, y = yp - 1;
      if (c < C0) {
        if (h0 == H && w0 == W) {
          v = __ldg((const float4*)(src0 + b * bs0 + ((int64_t)y * W + x) * C0 + c));
        } else {
          const Tap ty = tap(y, sy, true, h0), tx = tap(x, sx, true, w0);
          const float* base = src0 + b * bs0 + c;
          const float4 v00 = __ldg((const float4*)(base + ((int64_t)ty.i0 * w0 + tx.i0) * C0));
          const float4 v01 = __ldg((const float4*)(base + ((int64_t)ty.i0 * w0 + tx.i1) * C0));
          const float4 v10 = __ldg((const float4*)(base + ((int64_t)ty.i1 * w0 + tx.i0) * C0));
          const float4 v11 = __ldg((const float4*)(base + ((int64_t)ty.i1 * w0 + tx.i1) * C0));
          v.x = ty.l0 * (tx.l0 * v00.x + tx.l1 * v01.x) + ty.l1 * (tx.l0 * v10.x + tx.l1 * v11.x);
          v.y = ty.l0 * (tx.l0 * v00.y + tx.l1 * v01.y) + ty.l1 * (tx.l0 * v10.y + tx.l1 * v11.y);
          v.z = ty.l0 * (tx.l0 * v00.z + tx.l1 * v01.z) + ty.l1 * (tx.l0 * v10.z + tx.l1 * v11.z);
          v.w = ty.l0 * (tx.l0 * v00.w + tx.l1 * v01.w) + ty.l1 * (tx.l0 * v10.w + tx.l1 * v11.w);
        }
      } else {
        v = __ldg((const float4*)(src1 + b * bs1 + ((int64_t)y * W + x) * C1 + (c - C0)));
      }
    }
    *((float4*)dst + i) = v;
  }
}

// g: (B,H,W,ldg) - uses channels [0,C0); out (B,h0,w0,C0) = resize^T(g[..., :C0])
__global__ void __launch_bounds__(256) upsample_nhwc_bwd_kernel(
    const float* __restrict__ g, int ldg, float* __restrict__ out, int C0, int B, int H, int W, int h0, int w0,
    float sy, float sx) {
  const int C4 = C0 >> 2;
  const int64_t total = (int64_t)B * h0 * w0 * C4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    int64_t p = i / C4;
    const int k = (int)(p % w0); p /= w0;
    const int j = (int)(p % h0);
    const int b = (int)(p / h0);
    int ylo, yhi, xlo, xhi;
    adjoint_range(j, sy, true, h0, H, ylo, yhi);
    adjoint_range(k, sx, true, w0, W, xlo, xhi);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int y = ylo; y <= yhi; ++y) {
      const Tap ty = tap(y, sy, true, h0);
      const float wy = (ty.i0 == j ? ty.l0 : 0.f) + (ty.i1 == j ? ty.l1 : 0.f);
      if (wy == 0.f) continue;
      for (int x = xlo; x <= xhi; ++x) {
        const Tap tx = tap(x, sx, true, w0);
        const float wgt = wy * ((tx.i0 == k ? tx.l0 : 0.f) + (tx.i1 == k ? tx.l1 : 0.f));
        if (wgt == 0.f) continue;
        const float4 v = __ldg((const float4*)(g + (((int64_t)b * H + y) * W + x) * ldg + c));
        acc.x += wgt * v.x; acc.y += wgt * v.y; acc.z += wgt * v.z; acc.w += wgt * v.w;
      }
    }
    *((float4*)out + i) = acc;
  }
}

// Row-structured forms of the two kernels above (default where they apply).  The flat grid-stride kernels decompose a 64-bit
// element index with five divisions per float4 and - in the adjoint - re-derive the bilinear weights of up to 7 x 7 candidate
// pixels for EVERY channel quad (~700 instructions per 16 bytes: instruction-bound at 0.6 of the HBM peak and below).  Here a
// CTA owns one output row of one sample: the row's y taps and, per column, the x taps are evaluated ONCE into shared memory
// and re-used across all channels; the per-element work is one 32-bit division, the loads and the FMAs, in the same
// summation order (bit-identical results).
constexpr int RW_MAXW = 1024;     // widest row the forward's tap table holds (wider maps take the flat kernels)
constexpr int RW_SMEM0 = 40 * 1024; // the adjoint's x-tap tables (w0 x taps x 6 bytes + 4 w0; larger: flat kernel)
constexpr int RW_T = 24;          // most non-zero adjoint taps per axis (2 / scale + 2: ratios up to x10)

__global__ void __launch_bounds__(256) prep_conv_input_rows_kernel(
    const float* __restrict__ src0, int C0, int h0, int w0, const float* __restrict__ src1, int C1,
    float* __restrict__ dst, int B, int H, int W, float sy, float sx, int64_t bs0, int64_t bs1) {
  __shared__ int s_i0[RW_MAXW], s_i1[RW_MAXW];
  __shared__ float s_l1[RW_MAXW];
  const int yp = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int C = C0 + C1, C4 = C >> 2, Wp = W + 2;
  float4* drow = (float4*)(dst + (((int64_t)b * (H + 2) + yp) * Wp) * C);
  const int n = Wp * C4;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (yp == 0 || yp == H + 1) {
    for (int i = tid; i < n; i += 256) drow[i] = zero4;
    return;
  }
  const int y = yp - 1;
  const bool resize = !(h0 == H && w0 == W);
  if (resize) {
    for (int x = tid; x < W; x += 256) { const Tap tx = tap(x, sx, true, w0); s_i0[x] = tx.i0; s_i1[x] = tx.i1; s_l1[x] = tx.l1; }
    __syncthreads();
  }
  const Tap ty = tap(y, sy, true, h0);
  const float* r0 = src0 + b * bs0 + (int64_t)ty.i0 * w0 * C0;
  const float* r1 = src0 + b * bs0 + (int64_t)ty.i1 * w0 * C0;
  const float* s0row = src0 + b * bs0 + (int64_t)y * W * C0;                 // same-size case
  const float* s1row = src1 ? src1 + b * bs1 + (int64_t)y * W * C1 : nullptr;
  for (int i = tid; i < n; i += 256) {
    const int xp = i / C4, c = (i - xp * C4) * 4;
    float4 v = zero4;
    if (xp > 0 && xp <= W) {
      const int x = xp - 1;
      if (c < C0) {
        if (!resize) {
          v = __ldg((const float4*)(s0row + (int64_t)x * C0 + c));
        } else {
          const int i0 = s_i0[x], i1 = s_i1[x];
          const float l1 = s_l1[x], l0 = 1.f - l1;
          const float4 v00 = __ldg((const float4*)(r0 + (int64_t)i0 * C0 + c));
          const float4 v01 = __ldg((const float4*)(r0 + (int64_t)i1 * C0 + c));
          const float4 v10 = __ldg((const float4*)(r1 + (int64_t)i0 * C0 + c));
          const float4 v11 = __ldg((const float4*)(r1 + (int64_t)i1 * C0 + c));
          v.x = ty.l0 * (l0 * v00.x + l1 * v01.x) + ty.l1 * (l0 * v10.x + l1 * v11.x);
          v.y = ty.l0 * (l0 * v00.y + l1 * v01.y) + ty.l1 * (l0 * v10.y + l1 * v11.y);
          v.z = ty.l0 * (l0 * v00.z + l1 * v01.z) + ty.l1 * (l0 * v10.z + l1 * v11.z);
          v.w = ty.l0 * (l0 * v00.w + l1 * v01.w) + ty.l1 * (l0 * v10.w + l1 * v11.w);
        }
      } else {
        v = __ldg((const float4*)(s1row + (int64_t)x * C1 + (c - C0)));
      }
    }
    drow[i] = v;
  }
}

template <int CG>
__global__ void __launch_bounds__(256) upsample_nhwc_bwd_rows_kernel(
    const float* __restrict__ g, int ldg, float* __restrict__ out, int C0, int B, int H, int W, int h0, int w0,
    float sy, float sx, int T) {
  // dynamic shared memory: x-tap weights [w0][T] (float), x-tap columns [w0][T] (short), tap counts [w0] (int)
  extern __shared__ __align__(16) unsigned char rw_smem[];
  float* s_xw = (float*)rw_smem;
  int* s_xn = (int*)(s_xw + (size_t)w0 * T);
  short* s_xi = (short*)(s_xn + w0);
  __shared__ float s_yw[RW_T];
  __shared__ int s_yi[RW_T], s_yn;
  const int j = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int C4 = C0 >> 2;
  if (tid == 0) {
    int ylo, yhi, n = 0;
    adjoint_range(j, sy, true, h0, H, ylo, yhi);
    for (int y = ylo; y <= yhi && n < RW_T; ++y) {
      const Tap ty = tap(y, sy, true, h0);
      const float wy = (ty.i0 == j ? ty.l0 : 0.f) + (ty.i1 == j ? ty.l1 : 0.f);
      if (wy != 0.f) { s_yw[n] = wy; s_yi[n] = y; ++n; }
    }
    s_yn = n;
  }
  for (int k = tid; k < w0; k += 256) {
    int xlo, xhi, n = 0;
    adjoint_range(k, sx, true, w0, W, xlo, xhi);
    for (int x = xlo; x <= xhi && n < T; ++x) {
      const Tap tx = tap(x, sx, true, w0);
      const float wx = (tx.i0 == k ? tx.l0 : 0.f) + (tx.i1 == k ? tx.l1 : 0.f);
      if (wx != 0.f) { s_xw[k * T + n] = wx; s_xi[k * T + n] = (short)x; ++n; }
    }
    s_xn[k] = n;
  }
  __syncthreads();
  // A thread produces CG channel quads (C4 / CG apart) of one source column: the tap weights / columns / addresses are formed
  // once per CG float4 loads (they were ~12 instructions per 16-byte load).
  const int ny = s_yn, Cg = C4 / CG, n = w0 * Cg;
  float4* orow = (float4*)(out + (((int64_t)b * h0 + j) * w0) * C0);
  for (int i = tid; i < n; i += 256) {
    const int k = i / Cg, cq = i - k * Cg, c = cq * 4;      // quads cq, cq + Cg, ..: every load instruction stays fully coalesced
    const int nx = s_xn[k];
    float4 acc[CG];
#pragma unroll
    for (int q = 0; q < CG; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    // the (<= 5, rarely more) x taps of one y tap are requested together, then accumulated in order: a load consumed right
    // after its issue inside a loop of dynamic length leaves ONE load in flight per thread
    constexpr int XC = 5;
    for (int a = 0; a < ny; ++a) {
      const float wy = s_yw[a];
      const float* grow = g + ((int64_t)b * H + s_yi[a]) * W * ldg + c;
      for (int t0 = 0; t0 < nx; t0 += XC) {
        float4 v[XC][CG];
        float wgt[XC];
#pragma unroll
        for (int u = 0; u < XC; ++u) {
          const bool on = t0 + u < nx;
          wgt[u] = on ? wy * s_xw[k * T + t0 + u] : 0.f;
          const float4* src = (const float4*)(grow + (int64_t)(on ? s_xi[k * T + t0 + u] : 0) * ldg);
#pragma unroll
          for (int q = 0; q < CG; ++q) {
            v[u][q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (on && wgt[u] != 0.f) v[u][q] = __ldg(src + q * Cg);
          }
        }
#pragma unroll
        for (int u = 0; u < XC; ++u) {
          if (wgt[u] != 0.f) {
#pragma unroll
            for (int q = 0; q < CG; ++q) {
              acc[q].x += wgt[u] * v[u][q].x; acc[q].y += wgt[u] * v[u][q].y; acc[q].z += wgt[u] * v[u][q].z; acc[q].w += wgt[u] * v[u][q].w;
            }
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < CG; ++q) orow[(size_t)k * C4 + cq + q * Cg] = acc[q];
  }
}

// out (B,H,W,C) = base + bilinear(t (B,h0,w0,C) -> HxW, align_corners=True); out may alias base (in place)
template <typename IT>      // unsigned when the element count fits 32 bits: five divisions per float4
__global__ void __launch_bounds__(256) resize_add_nhwc_kernel(
    const float* __restrict__ t, const float* base_in, float* acc, int C, int B, int H, int W, int h0, int w0, float sy,
    float sx) {
  const int C4 = C >> 2;
  const IT total = (IT)B * (IT)H * (IT)W * (IT)C4;
  for (IT i = (IT)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (IT)gridDim.x * blockDim.x) {
    IT p = i / (IT)C4;
    const int c = (int)(i - p * (IT)C4) * 4;
    const IT p2 = p / (IT)W;
    const int x = (int)(p - p2 * (IT)W);
    const int b = (int)(p2 / (IT)H);
    const int y = (int)(p2 - (IT)b * (IT)H);
    const Tap ty = tap(y, sy, true, h0), tx = tap(x, sx, true, w0);
    const float* base = t + (int64_t)b * h0 * w0 * C + c;
    const float4 v00 = __ldg((const float4*)(base + ((int64_t)ty.i0 * w0 + tx.i0) * C));
    const float4 v01 = __ldg((const float4*)(base + ((int64_t)ty.i0 * w0 + tx.i1) * C));
    const float4 v10 = __ldg((const float4*)(base + ((int64_t)ty.i1 * w0 + tx.i0) * C));
    const float4 v11 = __ldg((const float4*)(base + ((int64_t)ty.i1 * w0 + tx.i1) * C));
    float4 a = *((const float4*)base_in + i);
    a.x += ty.l0 * (tx.l0 * v00.x + tx.l1 * v01.x) + ty.l1 * (tx.l0 * v10.x + tx.l1 * v11.x);
    a.y += ty.l0 * (tx.l0 * v00.y + tx.l1 * v01.y) + ty.l1 * (tx.l0 * v10.y + tx.l1 * v11.y);
    a.z += ty.l0 * (tx.l0 * v00.z + tx.l1 * v01.z) + ty.l1 * (tx.l0 * v10.z + tx.l1 * v11.z);
    a.w += ty.l0 * (tx.l0 * v00.w + tx.l1 * v01.w) + ty.l1 * (tx.l0 * v10.w + tx.l1 * v11.w);
    *((float4*)acc + i) = a;
  }
}

// d/dx [x Phi(x)] = Phi(x) + x phi(x) for the exact (erf) GELU.  Phi through Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7 on
// erf, i.e. 7.5e-8 on Phi - below fp32 round-off of the sum), whose exp(-u^2) with u = |x| / sqrt(2) IS the exp(-x^2 / 2) of
// the density term: one ex2, one rcp and a degree-5 polynomial instead of erff + expf (the kernel is issue-bound on the
// GELU layers: ~170 instructions per float4 of a 48-byte-per-float4 stream).
__device__ __forceinline__ float gelu_grad(float x) {
  const float E = __expf(-0.5f * x * x);
  const float u = fabsf(x) * 0.70710678118654752f;
  const float t = __frcp_rn(1.f + 0.3275911f * u);
  const float poly = t * (0.254829592f + t * (-0.284496736f + t * (1.421413741f + t * (-1.453152027f + t * 1.061405429f))));
  const float erf_abs = 1.f - poly * E;                       // erf(|x| / sqrt 2)
  const float Phi = 0.5f * (1.f + copysignf(erf_abs, x));
  return Phi + x * 0.3989422804014327f * E;
}

// gz[r, c] = g[r, c] * act'(ref[r, c]) * row_scale[r / rows_per_batch];  db[c] += sum_r gz[r, c]
// act: 1 relu, 2 leaky (ref = output), 3 gelu (ref = pre-activation), 4 sigmoid (ref = output), 0 none.
// grid (ceil(N/128), row chunks), block (32, 8): a thread owns 4 consecutive columns.
// GELU = true: act == 3 (the ~100-instruction derivative; one row per iteration keeps 40 registers and 6 CTAs per SM - requesting
// the next row first cost occupancy: 13.3 -> 14.8 ms per step).  Also measured and rejected: 16- / 8-lane thread rows for matrices of
// <= 64 / 32 columns, so that no lane idles (12.5 -> 14.2 ms: two rows per warp halve the bytes per load instruction).  GELU = false: every other case is a pure stream (column sums
// only, DropPath row scale, ReLU-class masks, dropout): two rows in flight per thread.
template <bool GELU>
__global__ void __launch_bounds__(256) act_bwd_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ ref,
                                                       float* __restrict__ gz, float* __restrict__ db,
                                                       const float* __restrict__ row_scale, int rows_per_batch,
                                                       int64_t rows, int N, int act, float slope,
                                                       int rows_per_block, float drop_p, uint32_t drop_seed,
                                                       const int* drop_step) {
  __shared__ float4 s_red[8][32];
  const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool drop = drop_p > 0.f;
  const uint32_t dseed = drop ? drop_seed_eff(drop_seed, drop_step) : 0u, dthresh = drop_threshold(drop_p);
  const float dinv = drop ? drop_scale(dthresh) : 1.f;
  const bool small_rows = rows < (1ll << 31);
  if (c < N) {
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    int64_t r = r0 + threadIdx.y;
    float4 vn = zero4, yn = zero4;
    if (!GELU && r < r1) {
      vn = __ldg((const float4*)(g + r * ldg + c));
      if (act) yn = __ldg((const float4*)(ref + r * N + c));
    }
    for (; r < r1; r += 8) {
      float4 v, y;
      if (GELU) {
        v = __ldg((const float4*)(g + r * ldg + c));
        y = __ldg((const float4*)(ref + r * N + c));
      } else {
        v = vn; y = yn;
        if (r + 8 < r1) {
          vn = __ldg((const float4*)(g + (r + 8) * ldg + c));
          if (act) yn = __ldg((const float4*)(ref + (r + 8) * N + c));
        }
      }
      if (drop) {      // the mask the GEMM epilogue drew for this element
        const uint32_t keep = drop_keep4(dseed, (uint32_t)(r * N + c), dthresh);
        v.x = (keep & 1u) ? v.x * dinv : 0.f;
        v.y = (keep & 2u) ? v.y * dinv : 0.f;
        v.z = (keep & 4u) ? v.z * dinv : 0.f;
        v.w = (keep & 8u) ? v.w * dinv : 0.f;
      }
      if (GELU) {
        v.x *= gelu_grad(y.x); v.y *= gelu_grad(y.y); v.z *= gelu_grad(y.z); v.w *= gelu_grad(y.w);
      } else if (act) {
        float d[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float t = d[e];
          if (act == 1) d[e] = t > 0.f ? 1.f : 0.f;
          else if (act == 2) d[e] = t > 0.f ? 1.f : slope;
          else d[e] = t * (1.f - t);
        }
        v.x *= d[0]; v.y *= d[1]; v.z *= d[2]; v.w *= d[3];
      }
      if (row_scale) {
        const int64_t sb = small_rows ? (int64_t)((unsigned)r / (unsigned)rows_per_batch) : r / rows_per_batch;   // 32-bit division when it fits
        const float s = __ldg(row_scale + sb);
        v.x *= s; v.y *= s; v.z *= s; v.w *= s;
      }
      if (gz) *(float4*)(gz + r * N + c) = v;
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  if (db == nullptr) return;
  s_red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < N) {
    for (int i = 1; i < 8; ++i) {
      const float4 t = s_red[i][threadIdx.x];
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    atomicAdd(db + c, acc.x); atomicAdd(db + c + 1, acc.y); atomicAdd(db + c + 2, acc.z); atomicAdd(db + c + 3, acc.w);
  }
}

// img (B, Cimg, H, W) NCHW, channels [0,Cin) -> tokens (B, DH*DW, Cin*P*P) in Conv2d weight order (c, ky, kx);
// bottom/right zero padding to a multiple of P (embed.py:286-294).  The 4x4/s4 patch-embedding conv is then a GEMM.
__global__ void __launch_bounds__(256) patchify_kernel(const float* __restrict__ img, int64_t bstride,
                                                        float* __restrict__ tok, int Cin, int H, int W, int P,
                                                        int DH, int DW, int64_t total) {
  const int K = Cin * P * P;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    int64_t t = i / K;
    const int px = (int)(t % DW); t /= DW;
    const int py = (int)(t % DH);
    const int b = (int)(t / DH);
    const int c = k / (P * P), r = k - c * P * P, ky = r / P, kx = r - ky * P;
    const int y = py * P + ky, x = px * P + kx;
    tok[i] = (y < H && x < W) ? __ldg(img + b * bstride + ((int64_t)c * H + y) * W + x) : 0.f;
  }
}

// General im2col for a strided / padded conv on NCHW planes: tok[(b,oy,ox)][k], k = (c*kh + ky)*kw + kx < Cin*kh*kw in
// Conv2d weight order, zero for taps outside the image and for the padding columns k >= Cin*kh*kw (row pitch Kp, a
// multiple of 4 so the rows are TMA-addressable).  Turns the 7x7/s2/p3 stem conv (depthformer_swin.py:1032-1039,
// K = 147 -> Kp = 148) into a tcgen05 GEMM.
__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ img, int64_t bstride,
                                                      float* __restrict__ tok, int Cin, int H, int W, int kh, int kw,
                                                      int stride, int pad, int Ho, int Wo, int Kp, int64_t total) {
  const int K = Cin * kh * kw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kp);
    int64_t t = i / Kp;
    const int ox = (int)(t % Wo); t /= Wo;
    const int oy = (int)(t % Ho);
    const int b = (int)(t / Ho);
    float v = 0.f;
    if (k < K) {
      const int c = k / (kh * kw), r = k - c * kh * kw, ky = r / kw, kx = r - ky * kw;
      const int y = oy * stride - pad + ky, x = ox * stride - pad + kx;
      if (y >= 0 && y < H && x >= 0 && x < W) v = __ldg(img + b * bstride + ((int64_t)c * H + y) * W + x);
    }
    tok[i] = v;
  }
}

// Row form (default): one CTA per output row (b, oy); the feature index k -> (channel plane offset, ky, kx) table is built
// once in shared memory, a thread emits four consecutive features as one 128-bit store, all index arithmetic is 32-bit
// (the flat kernel above spends ~300 instructions on 64-bit divisions per 4-byte element).
constexpr int IM_MAXK = 1024;
__global__ void __launch_bounds__(256) im2col_rows_kernel(const float* __restrict__ img, int64_t bstride,
                                                           float* __restrict__ tok, int Cin, int H, int W, int kh, int kw,
                                                           int stride, int pad, int Ho, int Wo, int Kp) {
  __shared__ int s_off[IM_MAXK];
  __shared__ short s_ky[IM_MAXK], s_kx[IM_MAXK];
  const int oy = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int K = Cin * kh * kw;
  for (int k = tid; k < Kp; k += 256) {
    const int c = k / (kh * kw), r = k - c * kh * kw, ky = r / kw, kx = r - ky * kw;
    s_off[k] = k < K ? (c * H + ky) * W + kx : -1;
    s_ky[k] = (short)ky; s_kx[k] = (short)kx;
  }
  __syncthreads();
  const int Kq = Kp >> 2, n4 = Wo * Kq;
  const int y0 = oy * stride - pad;
  const float* ib = img + b * bstride;
  float4* trow = (float4*)(tok + ((int64_t)b * Ho + oy) * Wo * Kp);
  for (int i = tid; i < n4; i += 256) {
    const int ox = i / Kq, k0 = (i - ox * Kq) * 4;
    const int x0 = ox * stride - pad;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = k0 + e, off = s_off[k];
      const int y = y0 + s_ky[k], x = x0 + s_kx[k];
      v[e] = (off >= 0 && y >= 0 && y < H && x >= 0 && x < W) ? __ldg(ib + off + y0 * W + x0) : 0.f;
    }
    trow[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// merge_patches with one CTA per MERGED row (b, y2): a thread owns one (x2, c) and moves the four pixels of its 2 x 2 patch as
// ONE 128-bit access on the merged side (features c*4 .. c*4+3 = (ky, kx) = (0,0) (0,1) (1,0) (1,1)) and four coalesced 4-byte
// accesses on the unmerged side; 32-bit index arithmetic.  (One thread per unmerged element wrote the merged side as 4-byte
// pieces 16 bytes apart: 34 % DRAM utilisation under ncu.)  Pixels beyond an odd H / W read as zero / are not written.
__global__ void __launch_bounds__(256) merge_patches_rows_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                                  int H, int W, int C, int H2, int W2, int dir) {
  const int y2 = blockIdx.x, b = blockIdx.y;
  const int y0 = 2 * y2, y1 = 2 * y2 + 1;
  const bool has_y1 = y1 < H;
  const int64_t urow0 = ((int64_t)b * H + y0) * W * C, urow1 = ((int64_t)b * H + y1) * W * C;   // unmerged row starts
  const int64_t mrow = ((int64_t)b * H2 + y2) * W2 * (4 * (int64_t)C);
  const int n = W2 * C;
  for (int i = threadIdx.x; i < n; i += 256) {
    const int x2 = i / C, c = i - x2 * C;
    const int xa = 2 * x2, xb = 2 * x2 + 1;
    const bool has_xb = xb < W;
    const int64_t m = mrow + (int64_t)x2 * (4 * C) + c * 4;
    if (dir == 0) {
      float4 v;
      v.x = __ldg(src + urow0 + (int64_t)xa * C + c);
      v.y = has_xb ? __ldg(src + urow0 + (int64_t)xb * C + c) : 0.f;
      v.z = has_y1 ? __ldg(src + urow1 + (int64_t)xa * C + c) : 0.f;
      v.w = (has_y1 && has_xb) ? __ldg(src + urow1 + (int64_t)xb * C + c) : 0.f;
      *(float4*)(dst + m) = v;
    } else {
      const float4 v = __ldg((const float4*)(src + m));
      dst[urow0 + (int64_t)xa * C + c] = v.x;
      if (has_xb) dst[urow0 + (int64_t)xb * C + c] = v.y;
      if (has_y1) dst[urow1 + (int64_t)xa * C + c] = v.z;
      if (has_y1 && has_xb) dst[urow1 + (int64_t)xb * C + c] = v.w;
    }
  }
}

// x (B, H, W, C) tokens -> (B, ceil(H/2)*ceil(W/2), 4C) in nn.Unfold(2,2) order: feature = c*4 + ky*2 + kx
// (depthformer_swin.py:86,115; zero padding bottom/right for odd sizes :110-111).  dir=1: backward (scatter = gather).
__global__ void __launch_bounds__(256) merge_patches_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                             int H, int W, int C, int H2, int W2, int64_t total, int dir) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    // i indexes the UNMERGED tensor (b, y, x, c): coalesced on that side; the merged side is a 16-byte-strided gather
    const int c = (int)(i % C);
    int64_t t = i / C;
    const int x = (int)(t % W); t /= W;
    const int y = (int)(t % H);
    const int b = (int)(t / H);
    const int64_t m = (((int64_t)b * H2 + (y >> 1)) * W2 + (x >> 1)) * (4 * C) + c * 4 + (y & 1) * 2 + (x & 1);
    if (dir == 0) dst[m] = __ldg(src + i); else dst[i] = __ldg(src + m);
  }
}

// out (B,1,H,W) = bilinear(clamp(x (B,1,h0,w0), lo, hi) -> HxW, align_corners=True)  (encoder_decoder.py:132-138)
__global__ void __launch_bounds__(256) clamp_resize_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                            int h0, int w0, int H, int W, float sy, float sx, float lo,
                                                            float hi, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xx = (int)(i % W);
    int64_t t = i / W;
    const int yy = (int)(t % H);
    const int b = (int)(t / H);
    const Tap ty = tap(yy, sy, true, h0), tx = tap(xx, sx, true, w0);
    const float* base = x + (int64_t)b * h0 * w0;
    const float v00 = fminf(fmaxf(__ldg(base + ty.i0 * w0 + tx.i0), lo), hi), v01 = fminf(fmaxf(__ldg(base + ty.i0 * w0 + tx.i1), lo), hi);
    const float v10 = fminf(fmaxf(__ldg(base + ty.i1 * w0 + tx.i0), lo), hi), v11 = fminf(fmaxf(__ldg(base + ty.i1 * w0 + tx.i1), lo), hi);
    out[i] = ty.l0 * (tx.l0 * v00 + tx.l1 * v01) + ty.l1 * (tx.l0 * v10 + tx.l1 * v11);
  }
}


}  // namespace ged
using namespace ged;

static int g_layout_rows = 1;     // 0: the flat grid-stride forms of prep_conv_input / upsample_nhwc_bwd (A/B, tests)
GED_API int ged_set_layout_rows(int on) { const int prev = g_layout_rows; g_layout_rows = on ? 1 : 0; return prev; }

static inline unsigned grid_for(int64_t total) {
  const int64_t b = (total + 255) / 256;
  return (unsigned)(b < 148 * 16 ? b : 148 * 16);
}

// dst [B,H+2,W+2,C0+C1] <- zero border | [resize(src0 (B,h0,w0,C0)) , src1 (B,H,W,C1)].  src1 may be NULL (C1=0).
GED_API int ged_prep_conv_input(const float* src0, int C0, int h0, int w0, const float* src1, int C1, float* dst,
                                int B, int H, int W, int64_t src0_bstride, int64_t src1_bstride, cudaStream_t stream) {
  if (!src0 || !dst || B <= 0 || H <= 0 || W <= 0 || C0 <= 0 || (C1 > 0 && !src1)) return GED_ERR_ARG;
  if ((C0 % 4) || (C1 % 4) || (src0_bstride % 4) || (src1_bstride % 4)) return GED_ERR_SHAPE;
  if (!aligned16(src0) || !aligned16(dst) || (src1 && !aligned16(src1))) return GED_ERR_ALIGN;
  // batch strides in floats (0 = dense): a source may be one level's slice of a (B, S, C) token tensor
  const int64_t bs0 = src0_bstride ? src0_bstride : (int64_t)h0 * w0 * C0, bs1 = src1_bstride ? src1_bstride : (int64_t)H * W * C1;
  const int64_t total = (int64_t)B * (H + 2) * (W + 2) * ((C0 + C1) / 4);
  if (g_layout_rows && W <= RW_MAXW && B <= 65535 && (int64_t)(W + 2) * ((C0 + C1) / 4) < (1 << 30))
    prep_conv_input_rows_kernel<<<dim3(H + 2, B), 256, 0, stream>>>(src0, C0, h0, w0, src1, C1, dst, B, H, W,
                                                                   resize_scale(h0, H, true), resize_scale(w0, W, true), bs0, bs1);
  else
    prep_conv_input_kernel<<<grid_for(total), 256, 0, stream>>>(src0, C0, h0, w0, src1, C1, dst, B, H, W,
                                                               resize_scale(h0, H, true), resize_scale(w0, W, true), bs0, bs1);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

GED_API int ged_upsample_nhwc_bwd(const float* g, int ldg, float* out, int C0, int B, int H, int W, int h0, int w0,
                                  cudaStream_t stream) {
  if (!g || !out || B <= 0 || C0 <= 0) return GED_ERR_ARG;
  if ((C0 % 4) || (ldg % 4)) return GED_ERR_SHAPE;
  if (!aligned16(g) || !aligned16(out)) return GED_ERR_ALIGN;
  const int64_t total = (int64_t)B * h0 * w0 * (C0 / 4);
  const float sy = resize_scale(h0, H, true), sx = resize_scale(w0, W, true);
  const int Tx = sx > 0.f ? (int)ceilf(2.f / sx) + 2 : RW_T + 1, Ty = sy > 0.f ? (int)ceilf(2.f / sy) + 2 : RW_T + 1;
  const size_t rsmem = (size_t)w0 * Tx * 6 + (size_t)w0 * 4 + 16;
  if (g_layout_rows && Tx <= RW_T && Ty <= RW_T && rsmem <= RW_SMEM0 && W <= 32767 && B <= 65535 && (int64_t)w0 * (C0 / 4) < (1 << 30))
  {
    if ((C0 / 4) % 2 == 0) upsample_nhwc_bwd_rows_kernel<2><<<dim3(h0, B), 256, rsmem, stream>>>(g, ldg, out, C0, B, H, W, h0, w0, sy, sx, Tx);
    else upsample_nhwc_bwd_rows_kernel<1><<<dim3(h0, B), 256, rsmem, stream>>>(g, ldg, out, C0, B, H, W, h0, w0, sy, sx, Tx);
  }
  else
    upsample_nhwc_bwd_kernel<<<grid_for(total), 256, 0, stream>>>(g, ldg, out, C0, B, H, W, h0, w0, sy, sx);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

GED_API int ged_resize_add_nhwc(const float* t, const float* base, float* out, int C, int B, int H, int W, int h0,
                                int w0, cudaStream_t stream) {
  if (!t || !base || !out || B <= 0 || C <= 0) return GED_ERR_ARG;
  if (C % 4) return GED_ERR_SHAPE;
  if (!aligned16(t) || !aligned16(base) || !aligned16(out)) return GED_ERR_ALIGN;
  const int64_t total = (int64_t)B * H * W * (C / 4);
  if (total + (int64_t)grid_for(total) * 256 < (1ll << 32))
    resize_add_nhwc_kernel<unsigned><<<grid_for(total), 256, 0, stream>>>(t, base, out, C, B, H, W, h0, w0, resize_scale(h0, H, true),
                                                                         resize_scale(w0, W, true));
  else
    resize_add_nhwc_kernel<int64_t><<<grid_for(total), 256, 0, stream>>>(t, base, out, C, B, H, W, h0, w0, resize_scale(h0, H, true),
                                                                        resize_scale(w0, W, true));
  GED_CHECK_LAUNCH();
  return GED_OK;
}

// db (N floats) is ACCUMULATED into when non-NULL.  ref: output (relu/leaky/sigmoid) or pre-activation (gelu).
GED_API int ged_act_bwd(const float* g, int64_t ldg, const float* ref, float* gz, float* db, const float* row_scale,
                        int rows_per_batch, int64_t rows, int N, int act, float slope, cudaStream_t stream) {
  if (!g || (!gz && !db) || rows <= 0 || N <= 0 || (act && !ref) || ldg < N) return GED_ERR_ARG;
  if ((N % 4) || (ldg % 4)) return GED_ERR_SHAPE;
  if (!aligned16(g) || (gz && !aligned16(gz)) || (ref && !aligned16(ref))) return GED_ERR_ALIGN;
  // rows per CTA: 128, or 512 where that still leaves >= 8 CTAs per SM (4x fewer column-sum atomics, 64 instead of 16 row
  // iterations per thread to amortise the fill and the drain of its two-rows-in-flight load pipeline)
  const int rpb = (rows / 512) * cdiv(N, 128) >= 148 * 8 ? 512 : 128;
  dim3 grid(cdiv(N, 128), (unsigned)((rows + rpb - 1) / rpb));
  if (act == 3)
    act_bwd_kernel<true><<<grid, dim3(32, 8), 0, stream>>>(g, ldg, ref, gz, db, row_scale, rows_per_batch > 0 ? rows_per_batch : 1,
                                                          rows, N, act, slope, rpb, 0.f, 0u, nullptr);
  else
    act_bwd_kernel<false><<<grid, dim3(32, 8), 0, stream>>>(g, ldg, ref, gz, db, row_scale, rows_per_batch > 0 ? rows_per_batch : 1,
                                                           rows, N, act, slope, rpb, 0.f, 0u, nullptr);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

// Backward of the dropout drawn in the GEMM epilogue (same p / seed / step counter): gz = g * keep / (1-p) and
// db += column sums of gz (db may be NULL).  No activation on this path.
GED_API int ged_dropout_bwd(const float* g, int64_t ldg, float* gz, float* db, int64_t rows, int N, float drop_p,
                            unsigned drop_seed, const int* drop_step, cudaStream_t stream) {
  if (!g || !gz || rows <= 0 || N <= 0 || ldg < N) return GED_ERR_ARG;
  if ((N % 4) || (ldg % 4) || drop_p <= 0.f || drop_p >= 1.f || rows * N > 0xFFFFFFFFll) return GED_ERR_SHAPE;
  if (!aligned16(g) || !aligned16(gz)) return GED_ERR_ALIGN;
  const int rpb = 128;
  dim3 grid(cdiv(N, 128), (unsigned)((rows + rpb - 1) / rpb));
  act_bwd_kernel<false><<<grid, dim3(32, 8), 0, stream>>>(g, ldg, nullptr, gz, db, nullptr, 1, rows, N, 0, 0.f, rpb, drop_p, drop_seed, drop_step);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

// tokens (B, DH*DW, Cin*P*P) from channels [0,Cin) of an NCHW image batch (batch stride in floats): embed.py:282-297.
GED_API int ged_patchify(const float* img, int64_t batch_stride, float* tok, int B, int Cin, int H, int W, int P,
                         cudaStream_t stream) {
  if (!img || !tok || B <= 0 || Cin <= 0 || P <= 0) return GED_ERR_ARG;
  const int DH = cdiv(H, P), DW = cdiv(W, P);
  const int64_t total = (int64_t)B * DH * DW * Cin * P * P;
  patchify_kernel<<<grid_for(total), 256, 0, stream>>>(img, batch_stride, tok, Cin, H, W, P, DH, DW, total);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

// tok (B*Ho*Wo, Kp) from channels [0,Cin) of an NCHW batch (batch stride in floats); Ho = (H+2p-kh)/s+1.
GED_API int ged_im2col(const float* img, int64_t batch_stride, float* tok, int B, int Cin, int H, int W, int kh, int kw,
                       int stride, int pad, int Kp, cudaStream_t stream) {
  if (!img || !tok || B <= 0 || Cin <= 0 || kh <= 0 || kw <= 0 || stride <= 0 || pad < 0) return GED_ERR_ARG;
  if (Kp < Cin * kh * kw || (Kp & 3)) return GED_ERR_SHAPE;
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  if (Ho <= 0 || Wo <= 0) return GED_ERR_SHAPE;
  const int64_t total = (int64_t)B * Ho * Wo * Kp;
  if (g_layout_rows && Kp <= IM_MAXK && B <= 65535 && (int64_t)Wo * Kp < (1 << 30) && (int64_t)Cin * H * W < (1 << 30) && aligned16(tok))
    im2col_rows_kernel<<<dim3(Ho, B), 256, 0, stream>>>(img, batch_stride, tok, Cin, H, W, kh, kw, stride, pad, Ho, Wo, Kp);
  else
    im2col_kernel<<<grid_for(total), 256, 0, stream>>>(img, batch_stride, tok, Cin, H, W, kh, kw, stride, pad, Ho, Wo, Kp, total);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

// backward=0: x (B,H,W,C) -> merged (B,H2*W2,4C) (zero-filled first when H or W is odd); backward=1: g_merged -> g_x.
GED_API int ged_merge_patches(const float* src, float* dst, int B, int H, int W, int C, int backward,
                              cudaStream_t stream) {
  if (!src || !dst || B <= 0 || C <= 0) return GED_ERR_ARG;
  const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
  if (!backward && ((H | W) & 1))
    if (cudaMemsetAsync(dst, 0, sizeof(float) * (size_t)B * H2 * W2 * 4 * C, stream) != cudaSuccess) return GED_ERR_LAUNCH;
  const int64_t total = (int64_t)B * H * W * C;
  if (g_layout_rows && B <= 65535 && (int64_t)W * C < (1 << 30) && aligned16(backward ? src : dst))
    merge_patches_rows_kernel<<<dim3(H2, B), 256, 0, stream>>>(src, dst, H, W, C, H2, W2, backward);
  else
    merge_patches_kernel<<<grid_for(total), 256, 0, stream>>>(src, dst, H, W, C, H2, W2, total, backward);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

GED_API int ged_clamp_resize(const float* x, float* out, int B, int h0, int w0, int H, int W, float lo, float hi,
                             cudaStream_t stream) {
  if (!x || !out || B <= 0) return GED_ERR_ARG;
  const int64_t total = (int64_t)B * H * W;
  clamp_resize_kernel<<<grid_for(total), 256, 0, stream>>>(x, out, h0, w0, H, W, resize_scale(h0, H, true),
                                                          resize_scale(w0, W, true), lo, hi, total);
  GED_CHECK_LAUNCH();
  return GED_OK;
}
