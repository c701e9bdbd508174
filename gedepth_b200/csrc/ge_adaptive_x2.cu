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
#include <cuda.h>
#include <cudaTypedefs.h>
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
constexpr int AX_HH = 4;                 // half-resolution rows per CTA (8 full-resolution rows): 4 CTAs of 128 threads per SM, so
                                         // one tile is always in flight while others compute
constexpr int AX_THREADS = 32 * AX_HH;   // 64 x AX_HH/2 threads, 16 pixels each
constexpr int AX_PITCH = AX_HW + 8;      // staged columns k0-4 .. k0+131 (rows are whole 16-byte units: one TMA box)
constexpr int AX_ROWS = AX_HH + 2;       // staged rows jy0-1 .. jy0+AX_HH
constexpr int AX_CHF = AX_ROWS * AX_PITCH;                   // floats per staged channel
// shared-memory map (bytes; every TMA destination 128-byte aligned)
constexpr int AX_OFF_Y = ((NSLOPE * AX_CHF * 4 + 127) / 128) * 128;
constexpr int AX_OFF_M = AX_OFF_Y + ((AX_CHF * 4 + 127) / 128) * 128;
constexpr int AX_OFF_PE = AX_OFF_M + ((AX_CHF * 4 + 127) / 128) * 128;
constexpr int AX_OFF_BAR = AX_OFF_PE + 2 * AX_HH * 2 * AX_HW * 4;
constexpr int AX_SMEM = AX_OFF_BAR + 16;

struct AxMaps { CUtensorMap logits, yh, pe; };     // {w2, h2, B*11} / {w2, h2, B} / {W, H, B}

__device__ __forceinline__ void ax_tma3(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               :: "r"(dst), "l"((uint64_t)map), "r"(bar), "r"(x), "r"(y), "r"(z) : "memory");
}

// off * m of one pixel with the exact softmax maximum, from the staged tile (rare path of the forward kernel)
__device__ __noinline__ float exact_pixel(const float* smem, int oy, int ox, int jy0, int k0, int h2, int w2, float pe, float h,
                                          float depth_scale) {
  const Tap tyy = tap(oy, 0.5f, false, h2), txx = tap(ox, 0.5f, false, w2);
  const int r0 = tyy.i0 - (jy0 - 1), r1 = tyy.i1 - (jy0 - 1), c0 = txx.i0 - (k0 - 4), c1 = txx.i1 - (k0 - 4);
  float Lx[NSLOPE];
#pragma unroll
  for (int ch = 0; ch < NSLOPE; ++ch) {
    const float* sp = smem + ch * AX_CHF;
    Lx[ch] = tyy.l0 * (txx.l0 * sp[r0 * AX_PITCH + c0] + txx.l1 * sp[r0 * AX_PITCH + c1]) +
             tyy.l1 * (txx.l0 * sp[r1 * AX_PITCH + c0] + txx.l1 * sp[r1 * AX_PITCH + c1]);
  }
  SlopeEval e;
  slope_eval(Lx, pe, h, depth_scale, e);
  return e.off * e.m;
}

// TMA: the three tiles (11 logit planes, y, camera plane) arrive as three bulk tensor copies issued by one thread (rows of
// the half-resolution maps must be 16-byte multiples: W % 8 == 0); otherwise 8 / 16-byte asynchronous copies by all threads.
// NAT: pixel pairs in memory order (0,1)(2,3) - needed where the interpolated logits are stored; otherwise the pairs are
// (0,3)(1,2), which lets the horizontal pass run two columns per instruction as well.
template <bool LOGITS, bool TMA>
__global__ void __launch_bounds__(AX_THREADS, 512 / AX_THREADS) ge_adaptive_fwd_x2_kernel(
    const __grid_constant__ AxMaps maps, const float* __restrict__ pe_raw, int64_t pe_bstride, const float* __restrict__ y_half,
    const float* __restrict__ logits_half, const float* __restrict__ height, float height_scalar, float depth_scale,
    float* __restrict__ y, float* __restrict__ pe_mask, float* __restrict__ logits_full, int H, int W, int h2, int w2) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* smem = (float*)smem_raw;
  float* s_y = (float*)(smem_raw + AX_OFF_Y);
  float* s_m = (float*)(smem_raw + AX_OFF_M);
  float* s_pe = (float*)(smem_raw + AX_OFF_PE);                // [2 AX_HH][2 AX_HW]
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 64 + tx;
  const int b = blockIdx.z, jy0 = blockIdx.y * AX_HH, k0 = blockIdx.x * AX_HW;
  const int hw2 = h2 * w2;
  const int64_t HW = (int64_t)H * W;

  if (TMA) {
    const uint32_t bar = smem_u32(smem_raw + AX_OFF_BAR);
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                   :: "r"(bar), "r"((NSLOPE + 1) * AX_CHF * 4 + 2 * AX_HH * 2 * AX_HW * 4) : "memory");
      ax_tma3(smem_u32(smem), &maps.logits, bar, k0 - 4, jy0 - 1, b * NSLOPE);
      ax_tma3(smem_u32(s_y), &maps.yh, bar, k0 - 4, jy0 - 1, b);
      ax_tma3(smem_u32(s_pe), &maps.pe, bar, 2 * k0, 2 * jy0, b);
    }
    __syncthreads();                                           // the barrier is initialised before anyone polls it
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "AXW_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
        "@p bra AXD_%=;\n\t"
        "bra AXW_%=;\n\t"
        "AXD_%=:\n\t}" :: "r"(bar) : "memory");
  } else {
    // out-of-range pairs are zero-filled and only ever meet a zero weight
    const uint32_t s0 = smem_u32(smem);
    const float* lsrc = logits_half + (int64_t)b * NSLOPE * hw2;
    const float* ysrc = y_half + (int64_t)b * hw2;
    for (int i = tid; i < AX_ROWS * (AX_PITCH / 2); i += AX_THREADS) {
      const int r = i / (AX_PITCH / 2), pr = i - r * (AX_PITCH / 2);
      const int j = jy0 - 1 + r, k = k0 - 4 + 2 * pr;
      const bool ok = j >= 0 && j < h2 && k >= 0 && k < w2;
      const int so = ok ? j * w2 + k : 0;
      const int nb = ok ? 8 : 0;
      uint32_t dst = s0 + (r * AX_PITCH + 2 * pr) * 4;
      const float* src = lsrc + so;
#pragma unroll
      for (int ch = 0; ch < NSLOPE; ++ch) {
        cp_async8(dst, src, nb);
        dst += AX_CHF * 4; src += hw2;
      }
      cp_async8(smem_u32(s_y) + (r * AX_PITCH + 2 * pr) * 4, ysrc + so, nb);
    }
    const float* psrc = pe_raw + (int64_t)b * pe_bstride;
    for (int i = tid; i < 2 * AX_HH * (2 * AX_HW / 4); i += AX_THREADS) {
      const int r = i >> 6, c4 = i & 63;
      const int oy = 2 * jy0 + r, ox = 2 * k0 + 4 * c4;
      const bool ok = oy < H && ox < W;
      cp_async16(smem_u32(s_pe + r * (2 * AX_HW) + 4 * c4), psrc + (ok ? (int64_t)oy * W + ox : 0), ok ? 16 : 0);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
  }
  // per half-resolution pixel: -log2(e) * (maximum over the 11 channels); +inf outside the map (never the minimum below)
  for (int i = tid; i < AX_ROWS * (AX_PITCH / 4); i += AX_THREADS) {
    const int r = i / (AX_PITCH / 4), q = i - r * (AX_PITCH / 4);
    const int j = jy0 - 1 + r, k = k0 - 4 + 4 * q;
    const float* sp = smem + r * AX_PITCH + 4 * q;
    float4 mx = *(const float4*)sp;
#pragma unroll
    for (int ch = 1; ch < NSLOPE; ++ch) {
      const float4 v = *(const float4*)(sp + ch * AX_CHF);
      mx.x = fmaxf(mx.x, v.x); mx.y = fmaxf(mx.y, v.y); mx.z = fmaxf(mx.z, v.z); mx.w = fmaxf(mx.w, v.w);
    }
    const bool rok = j >= 0 && j < h2;
    float4 o;
    o.x = rok && k >= 0 && k < w2 ? -kLog2e * mx.x : INFINITY;
    o.y = rok && k + 1 >= 0 && k + 1 < w2 ? -kLog2e * mx.y : INFINITY;
    o.z = rok && k + 2 >= 0 && k + 2 < w2 ? -kLog2e * mx.z : INFINITY;
    o.w = rok && k + 3 >= 0 && k + 3 < w2 ? -kLog2e * mx.w : INFINITY;
    *(float4*)(s_m + r * AX_PITCH + 4 * q) = o;
  }
  __syncthreads();

  const int ja = jy0 + 2 * ty;                         // half-resolution rows ja, ja+1 -> full-resolution rows 2ja .. 2ja+3
  const int t = blockIdx.x * 64 + tx;                  // half-resolution columns 2t, 2t+1 -> full-resolution columns 4t .. 4t+3
  if (ja >= h2 || 4 * t >= W) return;
  const bool two = ja + 1 < h2;                        // rows 2, 3 exist
  // vertical weights of output row r over staged rows (r+1)/2 and (r+1)/2 + 1 (the packed operations broadcast them)
  float wv[4][2];
  {
    const bool top = ja > 0, mid = ja < h2 - 1, bot = ja + 1 < h2 - 1;
    wv[0][0] = top ? 0.25f : 0.f; wv[0][1] = top ? 0.75f : 1.f;
    wv[1][0] = mid ? 0.75f : 1.f; wv[1][1] = mid ? 0.25f : 0.f;
    wv[2][0] = 0.25f; wv[2][1] = 0.75f;
    wv[3][0] = bot ? 0.75f : 1.f; wv[3][1] = bot ? 0.25f : 0.f;
  }
  const float a0 = t > 0 ? 0.25f : 0.f, a1 = t > 0 ? 0.75f : 1.f;                  // column 4t   over (2t-1, 2t)
  const float b0 = 2 * t + 1 < w2 - 1 ? 0.75f : 1.f, b1 = 2 * t + 1 < w2 - 1 ? 0.25f : 0.f;   // column 4t+3 over (2t+1, 2t+2)
  const u64 c7525 = pk2(0.75f, 0.25f), c2575 = pk2(0.25f, 0.75f), wo = pk2(a0, b1), wi = pk2(a1, b0);
  const int soff = (2 * ty) * AX_PITCH + 2 * tx + 4;

  // o[r][q]: output row r, pixel pair q of one channel; pairs are (0,1)(2,3) when LOGITS, else (0,3)(1,2)
  auto interp = [&](const float* chan, u64 (&o)[4][2]) {
    const float* sp = chan + soff;
    u64 hz[4][2];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float2 v = *(const float2*)(sp + r * AX_PITCH);
      const float vl = sp[r * AX_PITCH - 1], vr = sp[r * AX_PITCH + 2];
      if (LOGITS) {
        hz[r][0] = pk2(a0 * vl + a1 * v.x, 0.75f * v.x + 0.25f * v.y);
        hz[r][1] = pk2(0.25f * v.x + 0.75f * v.y, b0 * v.y + b1 * vr);
      } else {
        hz[r][0] = fma2(wi, pk2(v.x, v.y), mul2(wo, pk2(vl, vr)));                       // columns 0, 3
        hz[r][1] = fma2(pk2(v.y, v.y), c2575, mul2(pk2(v.x, v.x), c7525));               // columns 1, 2
      }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      o[0][q] = fma2(pk2(wv[0][1], wv[0][1]), hz[1][q], mul2(pk2(wv[0][0], wv[0][0]), hz[0][q]));
      o[1][q] = fma2(pk2(wv[1][1], wv[1][1]), hz[2][q], mul2(pk2(wv[1][0], wv[1][0]), hz[1][q]));
      o[2][q] = fma2(pk2(wv[2][1], wv[2][1]), hz[2][q], mul2(pk2(wv[2][0], wv[2][0]), hz[1][q]));
      o[3][q] = fma2(pk2(wv[3][1], wv[3][1]), hz[3][q], mul2(pk2(wv[3][0], wv[3][0]), hz[2][q]));
    }
  };

  // bound on -log2(e) * max_c L_c over the thread's 16 pixels: the minimum over its 4 x 4 half-resolution neighbourhood
  // (an interpolated logit never exceeds the largest of its taps)
  float negm;
  {
    const float* sp = s_m + soff;
    negm = INFINITY;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float2 v = *(const float2*)(sp + r * AX_PITCH);
      negm = fminf(negm, fminf(fminf(v.x, v.y), fminf(sp[r * AX_PITCH - 1], sp[r * AX_PITCH + 2])));
    }
  }
  const u64 negm2 = pk2(negm, negm);
  u64 s[4][2], tt[4][2];
#pragma unroll
  for (int r = 0; r < 4; ++r) { s[r][0] = s[r][1] = tt[r][0] = tt[r][1] = 0ull; }
  const u64 l2e = pk2(kLog2e, kLog2e);
  const int64_t pix0 = (int64_t)(2 * ja) * W + 4 * t;        // first pixel of the block inside one image plane
#pragma unroll
  for (int ch = 0; ch < NSLOPE; ++ch) {
    u64 L[4][2];
    interp(smem + ch * AX_CHF, L);
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
        upk2(fma2(L[r][q], l2e, negm2), x0, x1);
        const u64 e = pk2(ex2_fast(x0), ex2_fast(x1));
        s[r][q] = add2(s[r][q], e);
        tt[r][q] = fma2(e, cw2, tt[r][q]);
      }
    }
  }
  u64 yv2[4][2];
  interp(s_y, yv2);
  const float h = height ? __ldg(height + b) : height_scalar;
  const u64 nh2 = pk2(-h, -h), eps2 = pk2(1e-8f, 1e-8f), deg2 = pk2(kDeg, kDeg);
  const float* pp = s_pe + (4 * ty) * (2 * AX_HW) + 4 * tx;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    if (r >= 2 && !two) break;
    const float4 pe4 = *(const float4*)(pp + r * (2 * AX_HW));
    // pixel order of pair q: LOGITS (0,1)(2,3), else (0,3)(1,2)
    const int px[2][2] = {{0, LOGITS ? 1 : 3}, {LOGITS ? 2 : 1, LOGITS ? 3 : 2}};
    const float pe[4] = {pe4.x, pe4.y, pe4.z, pe4.w};
    float yo[4], pm[4];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      float s0, s1;
      upk2(s[r][q], s0, s1);
      // expected slope, tan, inverse-depth shift: two pixels per instruction (shift_and_mask / tan_deg_small, packed)
      const u64 th = mul2(tt[r][q], pk2(rcp_fast(s0), rcp_fast(s1)));
      const u64 xr = mul2(th, deg2), x2 = mul2(xr, xr);
      u64 pl = fma2(x2, pk2(0.05396825397f, 0.05396825397f), pk2(0.13333333333f, 0.13333333333f));
      pl = fma2(x2, pl, pk2(0.33333333333f, 0.33333333333f));
      pl = fma2(x2, pl, pk2(1.f, 1.f));
      const u64 k2 = mul2(xr, pl);
      float d0, d1;
      upk2(add2(pk2(pe[px[q][0]], pe[px[q][1]]), eps2), d0, d1);
      const u64 a2 = mul2(nh2, pk2(rcp_fast(d0), rcp_fast(d1)));
      upk2(add2(fma2(k2, pk2(-1.f, -1.f), a2), eps2), d0, d1);
      float off[2], yy[2];
      upk2(mul2(nh2, pk2(rcp_fast(d0), rcp_fast(d1))), off[0], off[1]);
      upk2(yv2[r][q], yy[0], yy[1]);
      const float ss[2] = {s0, s1};
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float mm = off[i];
        if (mm < 0.f) mm = 0.f;
        if (mm > depth_scale) mm = 0.f;
        if (mm > 0.f) mm = 1.f;
        float v = (off[i] * mm) * yy[i];
        if (!(ss[i] > 1e-30f))   // the bound on the maximum was too loose (or a logit is not finite): exact evaluation
          v = exact_pixel(smem, 2 * ja + r, 4 * t + px[q][i], jy0, k0, h2, w2, pe[px[q][i]], h, depth_scale) * yy[i];
        pm[px[q][i]] = v;
        yo[px[q][i]] = yy[i];
      }
    }
    const int64_t o = (int64_t)b * HW + pix0 + (int64_t)r * W;
    stg_stream((float4*)(y + o), make_float4(yo[0], yo[1], yo[2], yo[3]));
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
constexpr int BX_RING = 8;                 // g_logits_full channels in flight per warp
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
    // NOT awaited here: the first row segment's own loads (ring, pe, g_y, g_pe_mask) are issued first, so that the staging
    // latency and theirs overlap instead of adding up (a CTA lives for two row segments per warp: one exposed latency less
    // out of three)
  }

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
  const u64 c7525 = pk2(0.75f, 0.25f), c2575 = pk2(0.25f, 0.75f), wo = pk2(a0, b1), wi = pk2(a1, b0);
  // x contraction of a slot's four gradients g0..g3 and its neighbours' (left, right), as pairs (acc.x, acc.y):
  // (wxa0, wxb3) (left, right) + (wxa1, wxb2) (g0, g3) + (wxa2, wxb0) g1 + (wxa3, wxb1) g2
  const u64 xw_out = pk2(wxa[0], wxb[3]), xw_03 = pk2(wxa[1], wxb[2]), xw_1 = pk2(wxa[2], wxb[0]), xw_2 = pk2(wxa[3], wxb[1]);
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
    {
      // L2 prefetch of everything the NEXT row segment of this thread streams (and of this segment's channels beyond the ring)
      const int oyn = oy + 8;
      const bool nxt = r + 8 < BX_FR && oyn < H && col_ok;
      const int64_t pn = (int64_t)oyn * W + c0;
      if (g_logits_full) {
        const float* q = g_logits_full + (int64_t)b * NSLOPE * HW;
#pragma unroll
        for (int c = 0; c < NSLOPE; ++c) {
          if (nxt) asm volatile("prefetch.global.L2 [%0];" :: "l"(q + pn));
          if (c >= BX_RING && act) asm volatile("prefetch.global.L2 [%0];" :: "l"(q + po));
          q += HW;
        }
      }
      if (nxt) {
        asm volatile("prefetch.global.L2 [%0];" :: "l"(pe_raw + (int64_t)b * pe_bstride + pn));
        if (g_y) asm volatile("prefetch.global.L2 [%0];" :: "l"(g_y + b * HW + pn));
        if (g_pe_mask) asm volatile("prefetch.global.L2 [%0];" :: "l"(g_pe_mask + b * HW + pn));
      }
    }
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
    if (r == ty) {                 // first segment of every warp (ty < BX_FR): the staged tile must have landed
      if (g_logits_full) cp_async_wait<BX_RING>(); else cp_async_wait<0>();
      __syncthreads();
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
        float w0, w1;
        const u64 wmid = fma2(wy1p, pk2(v.x, v.y), mul2(wy0p, pk2(u.x, u.y)));
        const u64 wout = fma2(wy1p, pk2(sp[BX_TP - 1], sp[BX_TP + 2]), mul2(wy0p, pk2(sp[-1], sp[2])));
        upk2(wmid, w0, w1);
        G[ch][0] = fma2(wi, wmid, mul2(wo, wout));                                   // pixels 0, 3
        G[ch][1] = fma2(pk2(w1, w1), c2575, mul2(pk2(w0, w0), c7525));               // pixels 1, 2
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
      float sv[4], tv[4], yv[4];                 // pixel order from here on
      upk2(sum[0], sv[0], sv[3]); upk2(sum[1], sv[1], sv[2]);
      upk2(tsum[0], tv[0], tv[3]); upk2(tsum[1], tv[1], tv[2]);
      upk2(G[NSLOPE][0], yv[0], yv[3]); upk2(G[NSLOPE][1], yv[1], yv[2]);
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
      G[NSLOPE][0] = act ? pk2(gyo[0], gyo[3]) : 0ull;
      G[NSLOPE][1] = act ? pk2(gyo[1], gyo[2]) : 0ull;
    }
    const u64 Gs2[2] = {pk2(Gs[0], Gs[3]), pk2(Gs[1], Gs[2])}, nth2[2] = {pk2(nth[0], nth[3]), pk2(nth[1], nth[2])};
    const float* glnext = glsrc ? glsrc + (int64_t)BX_RING * HW : nullptr;
    float2* st = s_t + r * 32 + lane;
#pragma unroll
    for (int ch = 0; ch < BX_CH; ++ch) {
      u64 g03, g12;
      if (ch < NSLOPE) {
        const float cw = (float)(ch - 5);
        const u64 cw2 = pk2(cw, cw);
        u64 gl03 = 0ull, gl12 = 0ull;
        if (g_logits_full) {
          cp_async_wait<BX_RING - 1>();
          const float4 gl = s_ring[(ch % BX_RING) * 32];
          gl03 = pk2(gl.x, gl.w); gl12 = pk2(gl.y, gl.z);
          if (ch + BX_RING < NSLOPE) { cp_async16(ring0 + (ch % BX_RING) * 512, glnext, act ? 16 : 0); glnext += HW; }
          cp_async_commit();
        }
        g03 = fma2(mul2(Gs2[0], G[ch][0]), add2(cw2, nth2[0]), gl03);
        g12 = fma2(mul2(Gs2[1], G[ch][1]), add2(cw2, nth2[1]), gl12);
      } else {
        g03 = G[ch][0]; g12 = G[ch][1];
      }
      float g0, g1, g2, g3;
      upk2(g03, g0, g3); upk2(g12, g1, g2);
      const float left = __shfl_up_sync(0xffffffffu, g3, 1), right = __shfl_down_sync(0xffffffffu, g0, 1);
      u64 acc = fma2(xw_03, g03, mul2(xw_out, pk2(left, right)));
      acc = fma2(xw_1, pk2(g1, g1), acc);
      acc = fma2(xw_2, pk2(g2, g2), acc);
      float2 a2;
      upk2(acc, a2.x, a2.y);
      st[ch * (BX_FR * 32)] = a2;
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
    u64 o2 = 0ull;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const float2 v = sp[ch * (BX_FR * 32) + a * 32];
      o2 = fma2(pk2(wy[a], wy[a]), pk2(v.x, v.y), o2);
    }
    float2 o;
    upk2(o2, o.x, o.y);
    if (ch == NSLOPE) dst = g_y_half + (int64_t)b * hw2 + jy * w2 + 2 * t;
    *(float2*)dst = o;
    dst += hw2;
  }
}

static PFN_cuTensorMapEncodeTiled_v12000 ax_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
  if (!enc) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      enc = (PFN_cuTensorMapEncodeTiled_v12000)ptr;
  }
  return enc;
}
// 3-D fp32 map {w, h, n} with plane stride `plane` elements, box {bw, bh, bd}; out-of-range elements read 0
static bool ax_map3(CUtensorMap* m, const float* base, int w, int h, int n, int64_t plane, int bw, int bh, int bd) {
  auto enc = ax_encode();
  if (!enc) return false;
  cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[2] = {(cuuint64_t)w * 4, (cuuint64_t)plane * 4};
  cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bd};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <bool LOGITS, bool TMA>
static int ax_launch(const AxMaps& maps, const float* pe_raw, int64_t pe_bstride, const float* y_half, const float* logits_half,
                     const float* height, float height_scalar, float depth_scale, float* y, float* pe_mask, float* logits_full,
                     int B, int H, int W, int h2, int w2, cudaStream_t stream) {
  if (cudaFuncSetAttribute(ge_adaptive_fwd_x2_kernel<LOGITS, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, AX_SMEM) != cudaSuccess)
    return GED_ERR_LAUNCH;
  dim3 block(64, AX_HH / 2), grid(cdiv(w2, AX_HW), cdiv(h2, AX_HH), B);
  ge_adaptive_fwd_x2_kernel<LOGITS, TMA><<<grid, block, AX_SMEM, stream>>>(maps, pe_raw, pe_bstride, y_half, logits_half, height,
                                                                         height_scalar, depth_scale, y, pe_mask, logits_full, H, W, h2, w2);
  return cudaGetLastError() == cudaSuccess ? 0 : GED_ERR_LAUNCH;
}

static int g_ax_tma = 1;     // 0: asynchronous-copy staging even where TMA is possible (A/B, tests)
void set_ge_x2_tma(int on) { g_ax_tma = on; }

// host side: 0 = launched, 1 = shape / alignment not eligible (caller uses the generic kernels), < 0 = error
int launch_ge_adaptive_fwd_x2(const float* pe_raw, int64_t pe_bstride, const float* y_half, const float* logits_half,
                              const float* height, float height_scalar, float depth_scale, float* y, float* pe_mask,
                              float* logits_full, int B, int H, int W, int h2, int w2, cudaStream_t stream) {
  if (H != 2 * h2 || W != 2 * w2 || (W % 4) || (pe_bstride % 4) || !aligned16(pe_raw) || !aligned16(y) || !aligned16(pe_mask) ||
      (logits_full && !aligned16(logits_full)) || ((uintptr_t)y_half & 7) || ((uintptr_t)logits_half & 7)) return 1;
  AxMaps maps;
  bool tma = g_ax_tma && (w2 % 4 == 0) && aligned16(y_half) && aligned16(logits_half) && (int64_t)B * NSLOPE < (1ll << 31);
  if (tma)
    tma = ax_map3(&maps.logits, logits_half, w2, h2, B * NSLOPE, (int64_t)h2 * w2, AX_PITCH, AX_ROWS, NSLOPE) &&
          ax_map3(&maps.yh, y_half, w2, h2, B, (int64_t)h2 * w2, AX_PITCH, AX_ROWS, 1) &&
          ax_map3(&maps.pe, pe_raw, W, H, B, pe_bstride, 2 * AX_HW, 2 * AX_HH, 1);
#define AX_GO(L, T) ax_launch<L, T>(maps, pe_raw, pe_bstride, y_half, logits_half, height, height_scalar, depth_scale, y, pe_mask, \
                                    logits_full, B, H, W, h2, w2, stream)
  if (logits_full) return tma ? AX_GO(true, true) : AX_GO(true, false);
  return tma ? AX_GO(false, true) : AX_GO(false, false);
#undef AX_GO
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
