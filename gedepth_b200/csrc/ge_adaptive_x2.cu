// a15 (Adaptive ground embedding, depth/models/depther/encoder_decoder.py:79-102) when the full-resolution map is EXACTLY
// twice the half-resolution one - every GE config (352 x 1120 over 176 x 560, 384 x 640 over 192 x 320).
//
// align_corners=False x2 upsampling has the closed form
//     out[2j] = 0.25 in[j-1] + 0.75 in[j],  out[2j+1] = 0.75 in[j] + 0.25 in[j+1]     (edges: out[0] = in[0], out[2n-1] = in[n-1])
// so a thread that owns a 4 x 4 block of full-resolution pixels needs a 4 x 4 block of half-resolution ones, per channel:
// 12 shared-memory loads, 32 horizontal and 16 packed (fma.rn.f32x2) vertical operations for 16 pixels instead of 4 loads
// and 7 operations per pixel and channel.  The generic kernel (ground_embed.cu) was bound by instruction issue at ~250
// instructions per pixel; this one spends ~120.
//
//   forward   channels stream through registers: the softmax is stabilised with an UPPER BOUND of the per-pixel maximum -
//             the bilinear interpolation of the per-half-resolution-pixel channel maxima (formed while staging) - so the 11
//             logits of a pixel are never held at once; exp2 / sums / expected slope run two pixels per instruction.  A
//             pixel whose bound is so loose that the sum underflows is re-evaluated with the exact maximum (slope_eval).
//   backward  the adjoint of the x2 upsample is a 4 x 4 GATHER per half-resolution pixel with weights (.25 .75 .75 .25)
//             per axis: no candidate search, no atomics, no zero-fill.  Phase 1 evaluates the 12 per-pixel gradients of one
//             full-resolution row segment and contracts them along x (neighbour columns by warp shuffle) into shared
//             memory; phase 2 contracts along y.  A CTA is one warp wide (32 four-pixel slots, the outer two are halo
//             slots that are evaluated but not emitted), so no contraction crosses a warp.
#include "common.cuh"
#include "tile.cuh"

namespace ged {

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float rcp_fast(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ex2_fast(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// Ampere-style asynchronous copies (LDGSTS): the staged operands never pass through registers, so every load of a tile is
// in flight at once.  src_bytes == 0 zero-fills the destination (out-of-range halo).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kDeg = 0.017453292519943295f;

// tan(theta degrees) for |theta| <= 5: series, exact to fp32 for |x| <= 0.0873 (same polynomial as slope_eval)
__device__ __forceinline__ float tan_deg_small(float th) {
  const float xr = th * kDeg, x2 = xr * xr;
  return xr * (1.f + x2 * (0.33333333333f + x2 * (0.13333333333f + x2 * 0.05396825397f)));
}
// inverse-depth shift and range mask (encoder_decoder.py:92-100); returns off * m, den through the reference
__device__ __forceinline__ float shift_and_mask(float k, float pe, float h, float depth_scale, float& den, float& m) {
  const float a = -h * rcp_fast(pe + 1e-8f);
  den = (a - k) + 1e-8f;
  const float off = -h * rcp_fast(den);
  float mm = off;
  if (mm < 0.f) mm = 0.f;
  if (mm > depth_scale) mm = 0.f;
  if (mm > 0.f) mm = 1.f;
  m = mm;
  return off * mm;
}

// x2 adjoint weights of half-resolution index j (n of them) over full-resolution indices 2j-1 .. 2j+2
__device__ __forceinline__ void x2w(int j, int n, float (&w)[4]) {
  w[0] = j >= 1 ? 0.25f : 0.f;
  w[1] = j >= 1 ? 0.75f : 1.f;
  w[2] = j <= n - 2 ? 0.75f : 1.f;
  w[3] = j <= n - 2 ? 0.25f : 0.f;
}

// ---------------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------------
constexpr int AX_HW = 128;               // half-resolution columns per CTA (256 full-resolution columns)
constexpr int AX_HH = 8;                 // half-resolution rows per CTA (16 full-resolution rows)
constexpr int AX_PITCH = AX_HW + 4;      // staged columns k0-2 .. k0+129 (pairs stay 8-byte aligned)
constexpr int AX_ROWS = AX_HH + 2;       // staged rows jy0-1 .. jy0+8
constexpr int AX_CH = NSLOPE + 2;        // 11 logits, y, -log2(e) * channel maximum
constexpr int AX_TILE = AX_CH * AX_ROWS * AX_PITCH;          // floats
constexpr int AX_SMEM = (AX_TILE + 2 * AX_HH * 2 * AX_HW) * 4;   // + the camera-plane tile (16 x 256 full-resolution pixels)

// off * m of one pixel with the exact softmax maximum, from the staged tile (rare path of the forward kernel)
__device__ __noinline__ float exact_pixel(const float* smem, int oy, int ox, int jy0, int k0, int h2, int w2, float pe, float h,
                                          float depth_scale) {
  const Tap tyy = tap(oy, 0.5f, false, h2), txx = tap(ox, 0.5f, false, w2);
  const int r0 = tyy.i0 - (jy0 - 1), r1 = tyy.i1 - (jy0 - 1), c0 = txx.i0 - (k0 - 2), c1 = txx.i1 - (k0 - 2);
  float Lx[NSLOPE];
#pragma unroll
  for (int ch = 0; ch < NSLOPE; ++ch) {
    const float* sp = smem + ch * (AX_ROWS * AX_PITCH);
    Lx[ch] = tyy.l0 * (txx.l0 * sp[r0 * AX_PITCH + c0] + txx.l1 * sp[r0 * AX_PITCH + c1]) +
             tyy.l1 * (txx.l0 * sp[r1 * AX_PITCH + c0] + txx.l1 * sp[r1 * AX_PITCH + c1]);
  }
  SlopeEval e;
  slope_eval(Lx, pe, h, depth_scale, e);
  return e.off * e.m;
}

template <bool LOGITS>
__global__ void __launch_bounds__(256, 2) ge_adaptive_fwd_x2_kernel(
    const float* __restrict__ pe_raw, int64_t pe_bstride, const float* __restrict__ y_half,
    const float* __restrict__ logits_half, const float* __restrict__ height, float height_scalar, float depth_scale,
    float* __restrict__ y, float* __restrict__ pe_mask, float* __restrict__ logits_full, int H, int W, int h2, int w2) {
  extern __shared__ __align__(16) float smem[];
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 64 + tx;
  const int b = blockIdx.z, jy0 = blockIdx.y * AX_HH, k0 = blockIdx.x * AX_HW;
  const int64_t hw2 = (int64_t)h2 * w2, HW = (int64_t)H * W;

  // stage the half-resolution tile (+1 row / +2 columns of halo) pair by pair and the camera-plane tile, asynchronously;
  // out-of-range pairs are zero-filled and only ever meet a zero weight.
  float* s_pe = smem + AX_TILE;                              // [2 AX_HH][2 AX_HW]
  {
    const uint32_t s0 = smem_u32(smem);
    const float* lsrc = logits_half + (int64_t)b * NSLOPE * hw2;
    const float* ysrc = y_half + b * hw2;
    for (int i = tid; i < AX_ROWS * (AX_PITCH / 2); i += 256) {
      const int r = i / (AX_PITCH / 2), pr = i - r * (AX_PITCH / 2);
      const int j = jy0 - 1 + r, k = k0 - 2 + 2 * pr;
      const bool ok = j >= 0 && j < h2 && k >= 0 && k < w2;
      const int so = ok ? j * w2 + k : 0;
      const int nb = ok ? 8 : 0;
      uint32_t dst = s0 + (r * AX_PITCH + 2 * pr) * 4;
      const float* src = lsrc + so;
#pragma unroll
      for (int ch = 0; ch < NSLOPE; ++ch) {
        cp_async8(dst, src, nb);
        dst += AX_ROWS * AX_PITCH * 4; src += hw2;
      }
      cp_async8(dst, ysrc + so, nb);
    }
    const float* psrc = pe_raw + (int64_t)b * pe_bstride;
    for (int i = tid; i < 2 * AX_HH * (2 * AX_HW / 4); i += 256) {
      const int r = i >> 6, c4 = i & 63;
      const int oy = 2 * jy0 + r, ox = 2 * k0 + 4 * c4;
      const bool ok = oy < H && ox < W;
      cp_async16(smem_u32(s_pe + r * (2 * AX_HW) + 4 * c4), psrc + (ok ? (int64_t)oy * W + ox : 0), ok ? 16 : 0);
    }
    cp_async_commit();
    cp_async_wait<0>();
  }
  __syncthreads();
  // the per-pixel channel maximum becomes a 13th channel (scaled by -log2 e)
  for (int i = tid; i < AX_ROWS * (AX_PITCH / 2); i += 256) {
    const int r = i / (AX_PITCH / 2), pr = i - r * (AX_PITCH / 2);
    const float* sp = smem + r * AX_PITCH + 2 * pr;
    float2 mx = *(const float2*)sp;
#pragma unroll
    for (int ch = 1; ch < NSLOPE; ++ch) {
      const float2 v = *(const float2*)(sp + ch * (AX_ROWS * AX_PITCH));
      mx.x = fmaxf(mx.x, v.x); mx.y = fmaxf(mx.y, v.y);
    }
    *(float2*)(smem + (NSLOPE + 1) * (AX_ROWS * AX_PITCH) + r * AX_PITCH + 2 * pr) = make_float2(-kLog2e * mx.x, -kLog2e * mx.y);
  }
  __syncthreads();

  const int ja = jy0 + 2 * ty;                         // half-resolution rows ja, ja+1 -> full-resolution rows 2ja .. 2ja+3
  const int t = blockIdx.x * 64 + tx;                  // half-resolution columns 2t, 2t+1 -> full-resolution columns 4t .. 4t+3
  if (ja >= h2 || 4 * t >= W) return;
  const bool two = ja + 1 < h2;                        // rows 2, 3 exist
  // vertical weights of output row r over staged rows (r+1)/2 and (r+1)/2 + 1, as packed pairs
  u64 wv[4][2];
  {
    const bool top = ja > 0, mid = ja < h2 - 1, bot = ja + 1 < h2 - 1;
    const float w00 = top ? 0.25f : 0.f, w01 = top ? 0.75f : 1.f;
    const float w10 = mid ? 0.75f : 1.f, w11 = mid ? 0.25f : 0.f;
    const float w30 = bot ? 0.75f : 1.f, w31 = bot ? 0.25f : 0.f;
    wv[0][0] = pk2(w00, w00); wv[0][1] = pk2(w01, w01);
    wv[1][0] = pk2(w10, w10); wv[1][1] = pk2(w11, w11);
    wv[2][0] = pk2(0.25f, 0.25f); wv[2][1] = pk2(0.75f, 0.75f);
    wv[3][0] = pk2(w30, w30); wv[3][1] = pk2(w31, w31);
  }
  const float a0 = t > 0 ? 0.25f : 0.f, a1 = t > 0 ? 0.75f : 1.f;                  // column 4t   over (2t-1, 2t)
  const float b0 = 2 * t + 1 < w2 - 1 ? 0.75f : 1.f, b1 = 2 * t + 1 < w2 - 1 ? 0.25f : 0.f;   // column 4t+3 over (2t+1, 2t+2)
  const float* sbase = smem + (2 * ty) * AX_PITCH + 2 * tx + 2;

  // o[r][q]: output row r, column pair q of one channel
  auto interp = [&](int ch, u64 (&o)[4][2]) {
    const float* sp = sbase + ch * (AX_ROWS * AX_PITCH);
    u64 hz[4][2];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float2 v = *(const float2*)(sp + r * AX_PITCH);
      const float vl = sp[r * AX_PITCH - 1], vr = sp[r * AX_PITCH + 2];
      hz[r][0] = pk2(a0 * vl + a1 * v.x, 0.75f * v.x + 0.25f * v.y);
      hz[r][1] = pk2(0.25f * v.x + 0.75f * v.y, b0 * v.y + b1 * vr);
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      o[0][q] = fma2(wv[0][1], hz[1][q], mul2(wv[0][0], hz[0][q]));
      o[1][q] = fma2(wv[1][1], hz[2][q], mul2(wv[1][0], hz[1][q]));
      o[2][q] = fma2(wv[2][1], hz[2][q], mul2(wv[2][0], hz[1][q]));
      o[3][q] = fma2(wv[3][1], hz[3][q], mul2(wv[3][0], hz[2][q]));
    }
  };

  u64 negm[4][2], s[4][2], tt[4][2];
  interp(NSLOPE + 1, negm);
#pragma unroll
  for (int r = 0; r < 4; ++r) { s[r][0] = s[r][1] = tt[r][0] = tt[r][1] = 0ull; }
  const u64 l2e = pk2(kLog2e, kLog2e);
  const int64_t pix0 = (int64_t)(2 * ja) * W + 4 * t;        // first pixel of the block inside one image plane
#pragma unroll
  for (int ch = 0; ch < NSLOPE; ++ch) {
    u64 L[4][2];
    interp(ch, L);
    if (LOGITS) {
      float* lp = logits_full + ((int64_t)b * NSLOPE + ch) * HW + pix0;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (r < 2 || two) {
          float4 v;
          upk2(L[r][0], v.x, v.y); upk2(L[r][1], v.z, v.w);
          stg_stream((float4*)(lp + (int64_t)r * W), v);
        }
      }
    }
    const float cw = (float)(ch - 5);
    const u64 cw2 = pk2(cw, cw);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float x0, x1;
        upk2(fma2(L[r][q], l2e, negm[r][q]), x0, x1);
        const u64 e = pk2(ex2_fast(x0), ex2_fast(x1));
        s[r][q] = add2(s[r][q], e);
        tt[r][q] = fma2(e, cw2, tt[r][q]);
      }
    }
  }
  u64 yv2[4][2];
  interp(NSLOPE, yv2);
  const float h = height ? __ldg(height + b) : height_scalar;
  const float* pp = s_pe + (4 * ty) * (2 * AX_HW) + 4 * tx;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    if (r >= 2 && !two) break;
    const float4 pe4 = *(const float4*)(pp + r * (2 * AX_HW));
    const float pe[4] = {pe4.x, pe4.y, pe4.z, pe4.w};
    float sv[4], tv[4], yv[4], pm[4];
    upk2(s[r][0], sv[0], sv[1]); upk2(s[r][1], sv[2], sv[3]);
    upk2(tt[r][0], tv[0], tv[1]); upk2(tt[r][1], tv[2], tv[3]);
    upk2(yv2[r][0], yv[0], yv[1]); upk2(yv2[r][1], yv[2], yv[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float den, m;
      if (sv[i] > 1e-30f) {
        const float th = tv[i] * rcp_fast(sv[i]);
        pm[i] = shift_and_mask(tan_deg_small(th), pe[i], h, depth_scale, den, m) * yv[i];
      } else {
        // the bound on the maximum was too loose (or a logit is not finite): exact evaluation from the staged tile
        pm[i] = exact_pixel(smem, 2 * ja + r, 4 * t + i, jy0, k0, h2, w2, pe[i], h, depth_scale) * yv[i];
      }
    }
    const int64_t o = (int64_t)b * HW + pix0 + (int64_t)r * W;
    stg_stream((float4*)(y + o), make_float4(yv[0], yv[1], yv[2], yv[3]));
    stg_stream((float4*)(pe_mask + o), make_float4(pm[0], pm[1], pm[2], pm[3]));
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------------------
constexpr int BX_HH = 7;                   // half-resolution rows emitted per CTA
constexpr int BX_FR = 2 * BX_HH + 2;       // full-resolution rows evaluated: 2 jy0 - 1 .. 2 jy0 + 14
constexpr int BX_EMIT = 30;                // four-pixel slots emitted per CTA (of 32 evaluated)
constexpr int BX_TROWS = BX_HH + 2;        // staged half-resolution rows jy0-1 .. jy0+7
constexpr int BX_TP = 68;                  // staged half-resolution columns 60 bx - 4 .. 60 bx + 63
constexpr int BX_CH = NSLOPE + 1;          // 11 logits + y
constexpr int BX_RING = 6;                 // g_logits_full channels in flight per warp
constexpr int BX_TILE = BX_CH * BX_TROWS * BX_TP;                       // floats
constexpr int BX_ST = BX_CH * BX_FR * 32 * 2;                           // floats (float2 per slot)
constexpr int BX_SMEM = (BX_TILE + BX_ST + 8 * BX_RING * 32 * 4) * 4;

__global__ void __launch_bounds__(256, 2) ge_adaptive_bwd_x2_kernel(
    const float* __restrict__ pe_raw, int64_t pe_bstride, const float* __restrict__ y_half,
    const float* __restrict__ logits_half, const float* __restrict__ height, float height_scalar, float depth_scale,
    const float* __restrict__ g_y, const float* __restrict__ g_pe_mask, const float* __restrict__ g_logits_full,
    float* __restrict__ g_y_half, float* __restrict__ g_logits_half, int H, int W, int h2, int w2) {
  extern __shared__ __align__(16) float smem[];
  float* s_tile = smem;                                                  // [BX_CH][BX_TROWS][BX_TP]
  float2* s_t = (float2*)(smem + BX_TILE);                               // [BX_CH][BX_FR][32]
  const int lane = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + lane;
  float4* s_ring = (float4*)(smem + BX_TILE + BX_ST) + ty * (BX_RING * 32) + lane;     // this lane's slots, stride 32
  const int b = blockIdx.z, jy0 = blockIdx.y * BX_HH, kbase = blockIdx.x * (2 * BX_EMIT) - 4;
  const int hw2 = h2 * w2;
  const int64_t HW = (int64_t)H * W;

  {
    const uint32_t s0 = smem_u32(s_tile);
    const float* lsrc = logits_half + (int64_t)b * NSLOPE * hw2;
    const float* ysrc = y_half + (int64_t)b * hw2;
    for (int i = tid; i < BX_TROWS * (BX_TP / 2); i += 256) {
      const int r = i / (BX_TP / 2), pr = i - r * (BX_TP / 2);
      const int j = jy0 - 1 + r, k = kbase + 2 * pr;
      const bool ok = j >= 0 && j < h2 && k >= 0 && k < w2;
      const int so = ok ? j * w2 + k : 0;
      const int nb = ok ? 8 : 0;
      uint32_t dst = s0 + (r * BX_TP + 2 * pr) * 4;
      const float* src = lsrc + so;
#pragma unroll
      for (int ch = 0; ch < NSLOPE; ++ch) {
        cp_async8(dst, src, nb);
        dst += BX_TROWS * BX_TP * 4; src += hw2;
      }
      cp_async8(dst, ysrc + so, nb);
    }
    cp_async_commit();
    cp_async_wait<0>();
  }
  __syncthreads();

  const int t = blockIdx.x * BX_EMIT - 1 + lane;         // four-pixel slot: full-resolution columns 4t .. 4t+3
  const int c0 = 4 * t;
  const bool col_ok = t >= 0 && c0 < W;
  const float h = height ? __ldg(height + b) : height_scalar;
  const float a0 = t > 0 ? 0.25f : 0.f, a1 = t > 0 ? 0.75f : 1.f;
  const float b0 = 2 * t + 1 < w2 - 1 ? 0.75f : 1.f, b1 = 2 * t + 1 < w2 - 1 ? 0.25f : 0.f;
  float wxa[4], wxb[4];
  x2w(2 * t, w2, wxa);
  x2w(2 * t + 1, w2, wxb);
  const u64 l2e = pk2(kLog2e, kLog2e);
  const uint32_t ring0 = smem_u32(s_ring);

  // phase 1: per-pixel gradients of one full-resolution row segment, contracted along x.  G[ch][q]: pixel pair q of channel ch.
#pragma unroll 1
  for (int r = ty; r < BX_FR; r += 8) {
    const int oy = 2 * jy0 - 1 + r;
    const bool row_ok = oy >= 0 && oy < H;               // warp-uniform
    const bool act = row_ok && col_ok;
    const int64_t po = act ? (int64_t)oy * W + c0 : 0;
    // the 11 rows of g_logits_full stream through a per-warp ring of asynchronous copies: BX_RING channels are in flight
    // while the softmax of this row segment is evaluated, one more is issued per channel consumed
    const float* glsrc = g_logits_full ? g_logits_full + (int64_t)b * NSLOPE * HW + po : nullptr;
    if (g_logits_full) {
#pragma unroll
      for (int c = 0; c < BX_RING; ++c) {
        cp_async16(ring0 + c * 512, glsrc + (int64_t)c * HW, act ? 16 : 0);
        cp_async_commit();
      }
    }
    u64 G[BX_CH][2];
    float4 pe4 = make_float4(1.f, 1.f, 1.f, 1.f), gy4 = make_float4(0.f, 0.f, 0.f, 0.f), gm4 = gy4;
    if (act) {
      pe4 = ldg_stream((const float4*)(pe_raw + (int64_t)b * pe_bstride + po));
      if (g_y) gy4 = ldg_stream((const float4*)(g_y + b * HW + po));
      if (g_pe_mask) gm4 = ldg_stream((const float4*)(g_pe_mask + b * HW + po));
    }
    float Gs[4], nth[4];
    {
      // output row oy over staged rows (r >> 1), (r >> 1) + 1: vertical contraction first (two columns per instruction)
      float wy0, wy1;
      if (r & 1) { const int j = oy >> 1; wy0 = j > 0 ? 0.25f : 0.f; wy1 = j > 0 ? 0.75f : 1.f; }
      else { const int j = (oy - 1) >> 1; wy0 = j < h2 - 1 ? 0.75f : 1.f; wy1 = j < h2 - 1 ? 0.25f : 0.f; }
      const u64 wy0p = pk2(wy0, wy0), wy1p = pk2(wy1, wy1);
      const float* sp0 = s_tile + (r >> 1) * BX_TP + 2 * lane + 2;
#pragma unroll
      for (int ch = 0; ch < BX_CH; ++ch) {
        const float* sp = sp0 + ch * (BX_TROWS * BX_TP);
        const float2 u = *(const float2*)sp, v = *(const float2*)(sp + BX_TP);
        float wl, wr, w0, w1;
        upk2(fma2(wy1p, pk2(v.x, v.y), mul2(wy0p, pk2(u.x, u.y))), w0, w1);
        upk2(fma2(wy1p, pk2(sp[BX_TP - 1], sp[BX_TP + 2]), mul2(wy0p, pk2(sp[-1], sp[2]))), wl, wr);
        G[ch][0] = pk2(a0 * wl + a1 * w0, 0.75f * w0 + 0.25f * w1);
        G[ch][1] = pk2(0.25f * w0 + 0.75f * w1, b0 * w1 + b1 * wr);
      }
      // softmax (exact maximum) and expected slope, two pixels per instruction; G[ch] <- exp2 terms
      u64 nm[2], sum[2], tsum[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float m0, m1;
        upk2(G[0][q], m0, m1);
#pragma unroll
        for (int ch = 1; ch < NSLOPE; ++ch) {
          float x0, x1;
          upk2(G[ch][q], x0, x1);
          m0 = fmaxf(m0, x0); m1 = fmaxf(m1, x1);
        }
        nm[q] = pk2(-kLog2e * m0, -kLog2e * m1);
        sum[q] = tsum[q] = 0ull;
      }
#pragma unroll
      for (int ch = 0; ch < NSLOPE; ++ch) {
        const float cw = (float)(ch - 5);
        const u64 cw2 = pk2(cw, cw);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          float x0, x1;
          upk2(fma2(G[ch][q], l2e, nm[q]), x0, x1);
          const u64 e = pk2(ex2_fast(x0), ex2_fast(x1));
          G[ch][q] = e;
          sum[q] = add2(sum[q], e);
          tsum[q] = fma2(e, cw2, tsum[q]);
        }
      }
      float sv[4], tv[4], yv[4];
      upk2(sum[0], sv[0], sv[1]); upk2(sum[1], sv[2], sv[3]);
      upk2(tsum[0], tv[0], tv[1]); upk2(tsum[1], tv[2], tv[3]);
      upk2(G[NSLOPE][0], yv[0], yv[1]); upk2(G[NSLOPE][1], yv[2], yv[3]);
      const float pe[4] = {pe4.x, pe4.y, pe4.z, pe4.w}, gyv[4] = {gy4.x, gy4.y, gy4.z, gy4.w}, gmv[4] = {gm4.x, gm4.y, gm4.z, gm4.w};
      float gyo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float inv = rcp_fast(sv[i]);
        const float th = tv[i] * inv;
        const float k = tan_deg_small(th);
        float den, m;
        const float offm = shift_and_mask(k, pe[i], h, depth_scale, den, m);
        // d pe_mask / d y = off*m ; d pe_mask / d off = m*y ; d off / d k = -h/den^2 ;
        // d k / d theta = (pi/180)(1+k^2) ; d theta / d L_c = p_c (c-5 - theta).  m is a constant.
        gyo[i] = gyv[i] + gmv[i] * offm;
        const float rd = rcp_fast(den);
        float g = gmv[i] * m * yv[i] * (-h * rd * rd) * (kDeg * (1.f + k * k));
        if (m == 0.f) g = 0.f;                   // 0 * inf guards: the reference multiplies by an exact-zero mask
        Gs[i] = act ? g * inv : 0.f;             // folds the softmax normalisation; inactive lanes contribute nothing
        nth[i] = -th;
      }
      G[NSLOPE][0] = act ? pk2(gyo[0], gyo[1]) : 0ull;
      G[NSLOPE][1] = act ? pk2(gyo[2], gyo[3]) : 0ull;
    }
    const u64 Gs2[2] = {pk2(Gs[0], Gs[1]), pk2(Gs[2], Gs[3])}, nth2[2] = {pk2(nth[0], nth[1]), pk2(nth[2], nth[3])};
    float2* st = s_t + r * 32 + lane;
#pragma unroll
    for (int ch = 0; ch < BX_CH; ++ch) {
      float g0, g1, g2, g3;
      if (ch < NSLOPE) {
        const float cw = (float)(ch - 5);
        const u64 cw2 = pk2(cw, cw);
        u64 gl0 = 0ull, gl1 = 0ull;
        if (g_logits_full) {
          cp_async_wait<BX_RING - 1>();
          const float4 gl = s_ring[(ch % BX_RING) * 32];
          gl0 = pk2(gl.x, gl.y); gl1 = pk2(gl.z, gl.w);
          if (ch + BX_RING < NSLOPE) cp_async16(ring0 + (ch % BX_RING) * 512, glsrc + (int64_t)(ch + BX_RING) * HW, act ? 16 : 0);
          cp_async_commit();
        }
        upk2(fma2(mul2(Gs2[0], G[ch][0]), add2(cw2, nth2[0]), gl0), g0, g1);
        upk2(fma2(mul2(Gs2[1], G[ch][1]), add2(cw2, nth2[1]), gl1), g2, g3);
      } else {
        upk2(G[ch][0], g0, g1); upk2(G[ch][1], g2, g3);
      }
      const float left = __shfl_up_sync(0xffffffffu, g3, 1), right = __shfl_down_sync(0xffffffffu, g0, 1);
      float2 acc;
      acc.x = wxa[0] * left + wxa[1] * g0 + wxa[2] * g1 + wxa[3] * g2;
      acc.y = wxb[0] * g1 + wxb[1] * g2 + wxb[2] * g3 + wxb[3] * right;
      st[ch * (BX_FR * 32)] = acc;
    }
  }
  cp_async_wait<0>();
  __syncthreads();

  // phase 2: contraction along y; thread row = half-resolution row, lanes 1 .. 30 own the emitted slots
  const int jy = jy0 + ty;
  if (ty >= BX_HH || jy >= h2 || lane < 1 || lane > BX_EMIT || 2 * t >= w2) return;
  float wy[4];
  x2w(jy, h2, wy);
  const float2* sp = s_t + (2 * ty) * 32 + lane;
  float* dst = g_logits_half + (int64_t)b * NSLOPE * hw2 + jy * w2 + 2 * t;
#pragma unroll
  for (int ch = 0; ch < BX_CH; ++ch) {
    float2 o = make_float2(0.f, 0.f);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const float2 v = sp[ch * (BX_FR * 32) + a * 32];
      o.x = fmaf(wy[a], v.x, o.x); o.y = fmaf(wy[a], v.y, o.y);
    }
    if (ch == NSLOPE) dst = g_y_half + (int64_t)b * hw2 + jy * w2 + 2 * t;
    *(float2*)dst = o;
    dst += hw2;
  }
}

// host side: 0 = launched, 1 = shape / alignment not eligible (caller uses the generic kernels), < 0 = error
int launch_ge_adaptive_fwd_x2(const float* pe_raw, int64_t pe_bstride, const float* y_half, const float* logits_half,
                              const float* height, float height_scalar, float depth_scale, float* y, float* pe_mask,
                              float* logits_full, int B, int H, int W, int h2, int w2, cudaStream_t stream) {
  if (H != 2 * h2 || W != 2 * w2 || (W % 4) || (pe_bstride % 4) || !aligned16(pe_raw) || !aligned16(y) || !aligned16(pe_mask) ||
      (logits_full && !aligned16(logits_full)) || ((uintptr_t)y_half & 7) || ((uintptr_t)logits_half & 7)) return 1;
  if (cudaFuncSetAttribute(ge_adaptive_fwd_x2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AX_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(ge_adaptive_fwd_x2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AX_SMEM) != cudaSuccess)
    return GED_ERR_LAUNCH;
  dim3 block(64, 4), grid(cdiv(w2, AX_HW), cdiv(h2, AX_HH), B);
  if (logits_full)
    ge_adaptive_fwd_x2_kernel<true><<<grid, block, AX_SMEM, stream>>>(pe_raw, pe_bstride, y_half, logits_half, height, height_scalar, depth_scale, y, pe_mask, logits_full, H, W, h2, w2);
  else
    ge_adaptive_fwd_x2_kernel<false><<<grid, block, AX_SMEM, stream>>>(pe_raw, pe_bstride, y_half, logits_half, height, height_scalar, depth_scale, y, pe_mask, logits_full, H, W, h2, w2);
  return cudaGetLastError() == cudaSuccess ? 0 : GED_ERR_LAUNCH;
}

int launch_ge_adaptive_bwd_x2(const float* pe_raw, int64_t pe_bstride, const float* y_half, const float* logits_half,
                              const float* height, float height_scalar, float depth_scale, const float* g_y,
                              const float* g_pe_mask, const float* g_logits_full, float* g_y_half, float* g_logits_half,
                              int B, int H, int W, int h2, int w2, cudaStream_t stream) {
  if (H != 2 * h2 || W != 2 * w2 || (W % 4) || (pe_bstride % 4) || !aligned16(pe_raw) || (g_y && !aligned16(g_y)) ||
      (g_pe_mask && !aligned16(g_pe_mask)) || (g_logits_full && !aligned16(g_logits_full)) || ((uintptr_t)y_half & 7) ||
      ((uintptr_t)logits_half & 7) || ((uintptr_t)g_y_half & 7) || ((uintptr_t)g_logits_half & 7)) return 1;
  if (cudaFuncSetAttribute(ge_adaptive_bwd_x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BX_SMEM) != cudaSuccess)
    return GED_ERR_LAUNCH;
  dim3 block(32, 8), grid(cdiv(cdiv(w2, 2), BX_EMIT), cdiv(h2, BX_HH), B);
  ge_adaptive_bwd_x2_kernel<<<grid, block, BX_SMEM, stream>>>(pe_raw, pe_bstride, y_half, logits_half, height, height_scalar, depth_scale, g_y, g_pe_mask, g_logits_full, g_y_half, g_logits_half, H, W, h2, w2);
  return cudaGetLastError() == cudaSuccess ? 0 : GED_ERR_LAUNCH;
}

}  // namespace ged
