// Ground-embedding kernels (the path BASELINE.json names), sm_100a.  HBM-bound per-pixel work:
// float4-vectorised, coalesced, the half-resolution operands (attention map, slope logits) staged
// through shared memory once per tile.
//
//   ged_ground_plane      a1   tools/preprocess_data_kitti.py:47-53 (+ loading.py:388-403, transforms.py:40-48)
//   ged_pixel_grid        a1   the int64 (u,v) meshgrid of preprocess_data_kitti.py:52
//   ged_ge_vanilla_*      a14  depth/models/depther/encoder_decoder.py:112-123
//   ged_ge_adaptive_*     a15  depth/models/depther/encoder_decoder.py:79-102
//   ged_fuse_head_*       a17  depth/models/decode_heads/decode_head.py:489-508
//   ged_find_k            (f)1 tools/preprocess_data_kitti.py:59-63,86-89 / preprocess_data_ddad.py:47-51,77-82
//
// Algorithmic bytes per full-resolution pixel (fp32): ground_plane 8; vanilla fwd 13;
// vanilla bwd 13; adaptive fwd 24 (+44 when the full-resolution logits are written); fuse_head ~9.
#include "common.cuh"
#include "tile.cuh"

namespace ged {

// ---------------------------------------------------------------------------------------------
// a1: ground-plane generator.  fp64 on the integer grid, separate roundings (no FMA contraction)
// so that the result equals numpy's `num / (cu*u + cv*v + c1)` bit for bit before the fp32 cast.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ground_plane_kernel(
    float* __restrict__ ch3, float* __restrict__ ch4, int64_t batch_stride3, int64_t batch_stride4,
    int B, int H, int W, double num, double cu, double cv, double c1, double u0, double v0, double su,
    double sv, float depth_scale, float clamp_max) {
  int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  int y = blockIdx.y;
  if (x4 >= W) return;
  double v = __dadd_rn(__dmul_rn((double)y, sv), v0);
  double row = __dmul_rn(cv, v);
  float raw[4], nrm[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double u = __dadd_rn(__dmul_rn((double)(x4 + i), su), u0);
    double den = __dadd_rn(__dadd_rn(__dmul_rn(cu, u), row), c1);
    float pe = (float)__ddiv_rn(num, den);
    raw[i] = pe;
    float c = pe;
    if (c > clamp_max) c = 0.f;      // loading.py:400-401 (NaN compares false and stays, like numpy)
    if (c < 0.f) c = 0.f;
    if (c > 0.f) c = c / depth_scale;  // transforms.py:44
    nrm[i] = c;
  }
  for (int b = 0; b < B; ++b) {
    float* p3 = ch3 + b * batch_stride3 + (int64_t)y * W + x4;
    float* p4 = ch4 + b * batch_stride4 + (int64_t)y * W + x4;
    if (x4 + 3 < W && aligned16(p3) && aligned16(p4)) {
      stg_stream((float4*)p3, make_float4(nrm[0], nrm[1], nrm[2], nrm[3]));
      stg_stream((float4*)p4, make_float4(raw[0], raw[1], raw[2], raw[3]));
    } else {
      for (int i = 0; i < 4 && x4 + i < W; ++i) { p3[i] = nrm[i]; p4[i] = raw[i]; }
    }
  }
}

__global__ void pixel_grid_kernel(long long* __restrict__ u, long long* __restrict__ v, int H, int W,
                                  int u0, int v0) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  u[(int64_t)y * W + x] = (long long)(u0 + x);
  v[(int64_t)y * W + x] = (long long)(v0 + y);
}

// ---------------------------------------------------------------------------------------------
// a14: Vanilla.  y = up(y_half) (align_corners=False); pe_mask = pe_norm * y * 200.
// ---------------------------------------------------------------------------------------------
// Tile: 256 columns x VT_H rows per CTA; a thread owns 4 consecutive columns and VT_H/4 rows, so the four
// column taps are computed once and re-used down the rows (the kernel was issue-bound, not HBM-bound, when
// every pixel recomputed both taps: ncu sm__throughput 83 % at 45 % DRAM).
constexpr int VT_H = 16;
constexpr int VS_H = VT_H / 2 + 4;
template <bool VEC>
__global__ void __launch_bounds__(TX * 4) ge_vanilla_fwd_kernel(
    const float* __restrict__ pe_norm, int64_t pe_bstride, const float* __restrict__ y_half,
    float* __restrict__ y, float* __restrict__ pe_mask, int H, int W, int h2, int w2, float sy,
    float sx) {
  __shared__ float s_y[VS_H][ST_W];
  const int b = blockIdx.z, oy0 = blockIdx.y * VT_H, ox0 = blockIdx.x * TILE_W;
  const int oy_last = min(oy0 + VT_H, H) - 1, ox_last = min(ox0 + TILE_W, W) - 1;
  const int sy0 = tap(oy0, sy, false, h2).i0, sx0 = tap(ox0, sx, false, w2).i0;
  const int sh = tap(oy_last, sy, false, h2).i1 - sy0 + 1, sw_ = tap(ox_last, sx, false, w2).i1 - sx0 + 1;
  const int ox = ox0 + threadIdx.x * 4;
  // all of the thread's pe rows are requested before the first one is used: the streaming loads / stores are volatile
  // asm, so a load written inside the row loop could not be moved above the previous row's stores and every row paid a
  // full memory latency (SASS: LDG, LDS.., STG, STG, LDG, ..).  Issued ahead of the staging so that they overlap it too.
  float4 pe4[VT_H / 4];
  if (VEC && ox < W) {
#pragma unroll
    for (int k = 0; k < VT_H / 4; ++k) {
      const int oy = oy0 + threadIdx.y + 4 * k;
      pe4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (oy < H) pe4[k] = ldg_stream((const float4*)(pe_norm + (int64_t)b * pe_bstride + (int64_t)oy * W + ox));
    }
  }
  const float* yh = y_half + (int64_t)b * h2 * w2;
  for (int r = threadIdx.y; r < sh; r += 4)
    for (int c = threadIdx.x; c < sw_; c += TX) s_y[r][c] = __ldg(yh + (int64_t)(sy0 + r) * w2 + sx0 + c);
  __syncthreads();
  if (ox >= W) return;
  const int n = min(4, W - ox);
  int c0[4], c1[4];
  float a0[4], a1[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const Tap tx = tap(min(ox + i, W - 1), sx, false, w2);
    c0[i] = tx.i0 - sx0; c1[i] = tx.i1 - sx0; a0[i] = tx.l0; a1[i] = tx.l1;
  }
#pragma unroll
  for (int k = 0; k < VT_H / 4; ++k) {
    const int oy = oy0 + threadIdx.y + 4 * k;
    if (oy >= H) break;
    const Tap ty = tap(oy, sy, false, h2);
    const float* r0 = s_y[ty.i0 - sy0];
    const float* r1 = s_y[ty.i1 - sy0];
    const int64_t o = ((int64_t)b * H + oy) * W + ox;
    const float* pp = pe_norm + (int64_t)b * pe_bstride + (int64_t)oy * W + ox;
    float pe[4], yo[4], mo[4];
    if (VEC) { pe[0] = pe4[k].x; pe[1] = pe4[k].y; pe[2] = pe4[k].z; pe[3] = pe4[k].w; }
    else { for (int i = 0; i < n; ++i) pe[i] = __ldg(pp + i); }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float v = ty.l0 * (a0[i] * r0[c0[i]] + a1[i] * r0[c1[i]]) + ty.l1 * (a0[i] * r1[c0[i]] + a1[i] * r1[c1[i]]);
      yo[i] = v;
      mo[i] = pe[i] * v * 200.f;    // literal 200, not depth_scale (encoder_decoder.py:122)
    }
    if (VEC) {
      stg_stream((float4*)(y + o), make_float4(yo[0], yo[1], yo[2], yo[3]));
      stg_stream((float4*)(pe_mask + o), make_float4(mo[0], mo[1], mo[2], mo[3]));
    } else {
      for (int i = 0; i < n; ++i) { y[o + i] = yo[i]; pe_mask[o + i] = mo[i]; }
    }
  }
}

// a14 forward when the full-resolution map is exactly twice the half-resolution one (every GE config), streaming form: a
// WARP owns a strip of 128 full-resolution columns and walks VF_FR rows down it.  The x-interpolated half-resolution row
// (out[2k] = 1/4 in[k-1] + 3/4 in[k], out[2k+1] = 3/4 in[k] + 1/4 in[k+1], clamped at the map's edges) is formed once per
// half-resolution row from one float2 per lane + two shuffles and kept in registers; every full-resolution row is a
// vertical blend of two such rows.  The pe rows and the half-resolution row of the NEXT iteration are requested before the
// current one is blended and stored; no shared memory, no barrier.
struct VfHalf { float2 v; float e; };
template <int VF_FR>
__global__ void __launch_bounds__(128) ge_vanilla_fwd_x2s_kernel(
    const float* __restrict__ pe_norm, int64_t pe_bstride, const float* __restrict__ y_half,
    float* __restrict__ y, float* __restrict__ pe_mask, int H, int W, int h2, int w2) {
  const int lane = threadIdx.x, strip = blockIdx.x * 4 + threadIdx.y;
  if (strip * 128 >= W) return;                          // warp-uniform
  const int t = strip * 32 + lane, c0 = 4 * t, b = blockIdx.z;
  const bool col_ok = c0 < W;
  const float* pp0 = pe_norm + (int64_t)b * pe_bstride;
  const float* yh = y_half + (int64_t)b * h2 * w2;
  // half-resolution column just outside the lane's pair: lane 0 needs 2t-1, lane 31 needs 2t+2 (the others: shuffle)
  const int ce = lane == 0 ? 2 * t - 1 : 2 * t + 2;
  const bool edge_ok = (lane == 0 || lane == 31) && ce >= 0 && ce < w2 && col_ok;
  const float a0 = t > 0 ? 0.25f : 0.f, a1 = t > 0 ? 0.75f : 1.f;
  const float b0 = 2 * t + 2 < w2 ? 0.75f : 1.f, b1 = 2 * t + 2 < w2 ? 0.25f : 0.f;
  auto pload = [&](int oy) {
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (oy >= 0 && oy < H && col_ok) p = ldg_stream((const float4*)(pp0 + (int64_t)oy * W + c0));
    return p;
  };
  auto hload = [&](int j) {
    const int jc = min(max(j, 0), h2 - 1);
    VfHalf q;
    q.v = make_float2(0.f, 0.f); q.e = 0.f;
    if (col_ok) {
      q.v = __ldg((const float2*)(yh + (int64_t)jc * w2 + 2 * t));
      if (edge_ok) q.e = __ldg(yh + (int64_t)jc * w2 + ce);
    }
    return q;
  };
  auto hrow = [&](const VfHalf& q) {
    float left = __shfl_up_sync(0xffffffffu, q.v.y, 1), right = __shfl_down_sync(0xffffffffu, q.v.x, 1);
    if (lane == 0) left = q.e;
    if (lane == 31) right = q.e;
    return make_float4(a0 * left + a1 * q.v.x, 0.75f * q.v.x + 0.25f * q.v.y, 0.25f * q.v.x + 0.75f * q.v.y,
                       b0 * q.v.y + b1 * right);
  };
  // iteration k emits rows 2k+1 and 2k+2 from half-resolution rows k, k+1; the chunk owns rows [r0, r1)
  const int r0 = blockIdx.y * VF_FR, r1 = min(r0 + VF_FR, H);
  const int k_first = r0 == 0 ? -1 : (r0 - 1) >> 1;     // r0 even: rows r0-1 (not ours), r0
  const int k_last = (r1 - 2) >> 1;                      // last row r1-1 = 2 k_last + 1 or + 2
  VfHalf hn = hload(k_first + 1);
  float4 hk = hrow(hload(k_first));
  float4 pA = pload(2 * k_first + 1), pB = pload(2 * k_first + 2);
  for (int k = k_first; k <= k_last; ++k) {
    VfHalf hnn = hn;
    float4 pAn = pA, pBn = pB;
    if (k < k_last) { hnn = hload(k + 2); pAn = pload(2 * k + 3); pBn = pload(2 * k + 4); }
    const float4 hk1 = hrow(hn);
    const int oyA = 2 * k + 1, oyB = 2 * k + 2;
    if (oyA >= r0 && oyA < r1 && col_ok) {
      const float wl = k < h2 - 1 ? 0.75f : 1.f, wh = k < h2 - 1 ? 0.25f : 0.f;
      const float4 yo = make_float4(wl * hk.x + wh * hk1.x, wl * hk.y + wh * hk1.y, wl * hk.z + wh * hk1.z, wl * hk.w + wh * hk1.w);
      const int64_t o = ((int64_t)b * H + oyA) * W + c0;
      stg_stream((float4*)(y + o), yo);
      stg_stream((float4*)(pe_mask + o), make_float4(pA.x * yo.x * 200.f, pA.y * yo.y * 200.f, pA.z * yo.z * 200.f, pA.w * yo.w * 200.f));
    }
    if (oyB >= r0 && oyB < r1 && col_ok) {
      const float wl = k + 1 > 0 ? 0.25f : 0.f, wh = k + 1 > 0 ? 0.75f : 1.f;
      const float4 yo = make_float4(wl * hk.x + wh * hk1.x, wl * hk.y + wh * hk1.y, wl * hk.z + wh * hk1.z, wl * hk.w + wh * hk1.w);
      const int64_t o = ((int64_t)b * H + oyB) * W + c0;
      stg_stream((float4*)(y + o), yo);
      stg_stream((float4*)(pe_mask + o), make_float4(pB.x * yo.x * 200.f, pB.y * yo.y * 200.f, pB.z * yo.z * 200.f, pB.w * yo.w * 200.f));
    }
    hk = hk1; hn = hnn; pA = pAn; pB = pBn;
  }
}

// a14 backward: g_y_half += up^T( g_y + g_pe_mask * pe_norm * 200 ).  g_y_half must be zeroed.
__global__ void __launch_bounds__(TX * TILE_H) ge_vanilla_bwd_kernel(
    const float* __restrict__ pe_norm, int64_t pe_bstride, const float* __restrict__ g_y,
    const float* __restrict__ g_pe_mask, float* __restrict__ g_y_half, int H, int W, int h2, int w2,
    float sy, float sx) {
  __shared__ float s_g[1][TILE_H][TILE_W + 1];
  const int b = blockIdx.z, oy0 = blockIdx.y * TILE_H, ox0 = blockIdx.x * TILE_W;
  const SrcWindow sw = src_window(oy0, ox0, H, W, h2, w2, sy, sx, false);
  const int oy = oy0 + threadIdx.y;
  if (oy < H) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int lx = i * TX + threadIdx.x, ox = ox0 + lx;   // lane-contiguous: coalesced, conflict-free
      float g = 0.f;
      if (ox < W) {
        const int64_t o = ((int64_t)b * H + oy) * W + ox;
        const float pe = __ldg(pe_norm + (int64_t)b * pe_bstride + (int64_t)oy * W + ox);
        g = (g_y ? __ldg(g_y + o) : 0.f) + (g_pe_mask ? __ldg(g_pe_mask + o) * pe * 200.f : 0.f);
      }
      s_g[0][threadIdx.y][lx] = g;
    }
  }
  __syncthreads();
  tile_adjoint_upsample<1>(s_g, g_y_half + (int64_t)b * h2 * w2, 0, oy0, ox0, H, W, h2, w2, sy, sx,
                           false, sw);
}

// a14 backward when the full-resolution map is EXACTLY twice the half-resolution one (every GE config: 352 x 1120 over
// 176 x 560, 384 x 640 over 192 x 320).  align_corners=False x2 upsampling has the closed form
//     out[2j] = 0.25 in[j-1] + 0.75 in[j],  out[2j+1] = 0.75 in[j] + 0.25 in[j+1]   (edges: out[0] = in[0], out[2n-1] = in[n-1])
// so its adjoint is a 4 x 4 GATHER per half-resolution pixel with weights (0.25, 0.75, 0.75, 0.25) per axis - no
// search over candidate taps, no atomics, no zero-fill.  A thread produces two neighbouring half-resolution pixels of
// one row from 4 rows x (one aligned float4 + the two columns beside it) of each operand.
__device__ __forceinline__ void x2_weights(int j, int n, float (&w)[4]) {
  w[0] = j >= 1 ? 0.25f : 0.f;
  w[1] = j >= 1 ? 0.75f : 1.f;
  w[2] = j <= n - 2 ? 0.75f : 1.f;
  w[3] = j <= n - 2 ? 0.25f : 0.f;
}
// (Tried: issuing the loads of all 4-5 rows of a thread before the first use - 96 registers, 2 CTAs per SM: 0.565 -> 0.38 of
// the HBM peak; capped at 80 registers with spills: 0.50.  The 48-register loop below with 5 CTAs per SM stays.)
// Kept as the A/B partner of the streaming kernel below (ged_set_ge_x2(2)).
// Separable: a CTA owns X2_H half-resolution rows x 128 columns.  Pass 1: every full-resolution row of the footprint is
// read ONCE with aligned float4 loads (a lane's left / right neighbour columns come from its neighbours by warp
// shuffle), G = g_y + g_pe_mask * pe * 200 is formed and contracted along x into shared memory; pass 2 contracts along y.
constexpr int X2_H = 8;
__global__ void __launch_bounds__(256, 4) ge_vanilla_bwd_x2_kernel(
    const float* __restrict__ pe_norm, int64_t pe_bstride, const float* __restrict__ g_y,
    const float* __restrict__ g_pe_mask, float* __restrict__ g_y_half, int H, int W, int h2, int w2) {
  __shared__ float2 s_t[2 * X2_H + 2][64];
  const int tx = threadIdx.x, ty = threadIdx.y, lane = tx & 31;
  const int t = blockIdx.x * 64 + tx;                    // pair of half-resolution columns 2t, 2t+1
  const int jy0 = blockIdx.y * X2_H, b = blockIdx.z;
  const int64_t HW = (int64_t)H * W;
  const int c0 = 4 * t;
  const bool col_ok = c0 < W;
  float wxa[4], wxb[4];
  x2_weights(2 * t, w2, wxa);
  x2_weights(2 * t + 1, w2, wxb);
  // One row segment per iteration.  The loads of the NEXT iteration (and the warp's two outer neighbour columns, which only
  // lanes 0 / 31 need) are issued before the current row is consumed: two rows in flight per warp, no dependent second
  // round trip for the edge lanes.
  struct RowLoads { float4 p4, y4, m4; float ep, ey, em; bool ok; };
  const int ce = lane == 0 ? c0 - 1 : c0 + 4;
  const bool edge_ok = (lane == 0 || lane == 31) && ce >= 0 && ce < W;
  auto issue = [&](int r) {
    RowLoads q;
    const int oy = 2 * jy0 - 1 + r;
    q.ok = r < 2 * X2_H + 2 && oy >= 0 && oy < H && col_ok;
    q.p4 = q.y4 = q.m4 = make_float4(0.f, 0.f, 0.f, 0.f);
    q.ep = q.ey = q.em = 0.f;
    if (q.ok) {
      const int64_t row = (int64_t)oy * W;
      const float* pp = pe_norm + (int64_t)b * pe_bstride + row;
      const float* gy = g_y ? g_y + b * HW + row : nullptr;
      const float* gm = g_pe_mask ? g_pe_mask + b * HW + row : nullptr;
      q.p4 = __ldg((const float4*)(pp + c0));
      if (gy) q.y4 = ldg_stream((const float4*)(gy + c0));
      if (gm) q.m4 = ldg_stream((const float4*)(gm + c0));
      if (edge_ok) {      // raw operands only; combined when the row is consumed (no wait on the loads just issued)
        q.ep = __ldg(pp + ce);
        if (gy) q.ey = __ldg(gy + ce);
        if (gm) q.em = __ldg(gm + ce);
      }
    }
    return q;
  };
  RowLoads cur = issue(ty);
  for (int r = ty; r < 2 * X2_H + 2; r += 4) {
    const RowLoads nxt = issue(r + 4);
    const float4 G = make_float4(cur.y4.x + cur.m4.x * cur.p4.x * 200.f, cur.y4.y + cur.m4.y * cur.p4.y * 200.f,
                                 cur.y4.z + cur.m4.z * cur.p4.z * 200.f, cur.y4.w + cur.m4.w * cur.p4.w * 200.f);
    float left = __shfl_up_sync(0xffffffffu, G.w, 1), right = __shfl_down_sync(0xffffffffu, G.x, 1);
    float2 acc = make_float2(0.f, 0.f);
    if (cur.ok) {
      const float Ge = cur.ey + cur.em * cur.ep * 200.f;
      if (lane == 0) left = Ge;
      if (lane == 31) right = Ge;
      if (c0 + 4 >= W) right = 0.f;
      acc.x = wxa[0] * left + wxa[1] * G.x + wxa[2] * G.y + wxa[3] * G.z;
      acc.y = wxb[0] * G.y + wxb[1] * G.z + wxb[2] * G.w + wxb[3] * right;
    }
    s_t[r][tx] = acc;
    cur = nxt;
  }
  __syncthreads();
  if (2 * t >= w2) return;
  const bool has_b = 2 * t + 1 < w2;
#pragma unroll
  for (int k = 0; k < X2_H / 4; ++k) {
    const int jl = ty + 4 * k, jy = jy0 + jl;
    if (jy >= h2) break;
    float wy[4];
    x2_weights(jy, h2, wy);
    float2 o = make_float2(0.f, 0.f);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const float2 v = s_t[2 * jl + a][tx];
      o.x += wy[a] * v.x; o.y += wy[a] * v.y;
    }
    float* dst = g_y_half + ((int64_t)b * h2 + jy) * w2 + 2 * t;
    if (has_b && ((((uintptr_t)dst) & 7) == 0)) *(float2*)dst = o;
    else { dst[0] = o.x; if (has_b) dst[1] = o.y; }
  }
}

// Streaming form of the same adjoint (default).  A WARP owns a strip of 128 full-resolution columns and walks VB_CH
// half-resolution rows down it: the x-contracted values of the two previous full-resolution rows stay in registers, the two
// new rows of iteration j+1 are in flight while iteration j is contracted along y and stored (4 rows x 3 operands x 512 B =
// 6 KB per warp in flight), no shared memory, no barrier; the halo is 2 rows per 2 VB_CH.
// What held BOTH forms at 0.52-0.57 of the HBM peak was not the tiling but the strip's outer columns: lanes 0 / 31 loaded
// their neighbour column's three operands and combined them at once (LDG, LDG, LDG, FMUL in the SASS), so the whole warp
// waited a full memory latency right after issuing every row.  With the raw operands carried to the point of use:
// tiled 0.575 -> 0.756 / 0.80, streaming 0.52 -> 0.844 / 0.91 at 32 / 64 x 1024 x 2048 (tools/ab_ge_vbwd.py).
struct VbRow { float4 p4, y4, m4; float ep, ey, em; };
template <int VB_CH, int MINB, int UNR>
__global__ void __launch_bounds__(128, MINB) ge_vanilla_bwd_x2s_kernel(
    const float* __restrict__ pe_norm, int64_t pe_bstride, const float* __restrict__ g_y,
    const float* __restrict__ g_pe_mask, float* __restrict__ g_y_half, int H, int W, int h2, int w2) {
  const int lane = threadIdx.x, strip = blockIdx.x * 4 + threadIdx.y;
  if (strip * 128 >= W) return;                          // warp-uniform
  const int t = strip * 32 + lane;                       // half-resolution columns 2t, 2t+1 <- full-resolution 4t-1 .. 4t+4
  const int c0 = 4 * t, b = blockIdx.z;
  const int jy0 = blockIdx.y * VB_CH, jy1 = min(jy0 + VB_CH, h2);
  const bool col_ok = c0 < W;
  const int64_t HW = (int64_t)H * W;
  const float* pp0 = pe_norm + (int64_t)b * pe_bstride;
  const float* gy0 = g_y ? g_y + b * HW : nullptr;
  const float* gm0 = g_pe_mask ? g_pe_mask + b * HW : nullptr;
  float wxa[4], wxb[4];
  x2_weights(2 * t, w2, wxa);
  x2_weights(2 * t + 1, w2, wxb);
  // the column just outside the warp's strip: lane 0 needs 4t-1, lane 31 needs 4t+4 (the others get theirs by shuffle).
  // The three raw operands are only LOADED here and combined in contract(), one iteration later: combining them at once
  // made the whole warp wait a full memory latency on every row (the first version of this kernel, and the tiled one).
  const int ce = lane == 0 ? c0 - 1 : c0 + 4;
  const bool edge_ok = (lane == 0 || lane == 31) && ce >= 0 && ce < W && col_ok;
  auto issue = [&](int oy) {
    VbRow q;
    q.p4 = q.y4 = q.m4 = make_float4(0.f, 0.f, 0.f, 0.f);
    q.ep = q.ey = q.em = 0.f;
    if (oy >= 0 && oy < H && col_ok) {
      const int64_t row = (int64_t)oy * W;
      const float* pp = pp0 + row;
      q.p4 = ldg_stream((const float4*)(pp + c0));
      if (gy0) q.y4 = ldg_stream((const float4*)(gy0 + row + c0));
      if (gm0) q.m4 = ldg_stream((const float4*)(gm0 + row + c0));
      if (edge_ok) {
        q.ep = __ldg(pp + ce);
        if (gy0) q.ey = __ldg(gy0 + row + ce);
        if (gm0) q.em = __ldg(gm0 + row + ce);
      }
    }
    return q;
  };
  // rows outside the map and slots right of it were never loaded: their G is 0 and so is everything contracted from it
  auto contract = [&](const VbRow& q) {
    const float4 G = make_float4(q.y4.x + q.m4.x * q.p4.x * 200.f, q.y4.y + q.m4.y * q.p4.y * 200.f,
                                 q.y4.z + q.m4.z * q.p4.z * 200.f, q.y4.w + q.m4.w * q.p4.w * 200.f);
    const float Ge = q.ey + q.em * q.ep * 200.f;
    float left = __shfl_up_sync(0xffffffffu, G.w, 1), right = __shfl_down_sync(0xffffffffu, G.x, 1);
    if (lane == 0) left = Ge;
    if (lane == 31) right = Ge;
    return make_float2(wxa[0] * left + wxa[1] * G.x + wxa[2] * G.y + wxa[3] * G.z,
                       wxb[0] * G.y + wxb[1] * G.z + wxb[2] * G.w + wxb[3] * right);
  };
  VbRow a0 = issue(2 * jy0 - 1), a1 = issue(2 * jy0);
  VbRow n0 = issue(2 * jy0 + 1), n1 = issue(2 * jy0 + 2);
  float2 tm1 = contract(a0), t0 = contract(a1);
  const bool emit = 2 * t < w2;
  const bool has_b = 2 * t + 1 < w2;
  float* dst = g_y_half + ((int64_t)b * h2 + jy0) * w2 + 2 * t;
  const bool dst8 = has_b && ((((uintptr_t)dst) & 7) == 0) && ((w2 & 1) == 0);
#pragma unroll UNR
  for (int jy = jy0; jy < jy1; ++jy) {
    const VbRow c0r = n0, c1r = n1;
    if (jy + 1 < jy1) { n0 = issue(2 * jy + 3); n1 = issue(2 * jy + 4); }
    const float2 t1 = contract(c0r), t2 = contract(c1r);
    float wy[4];
    x2_weights(jy, h2, wy);
    float2 o = make_float2(0.f, 0.f);
    o.x += wy[0] * tm1.x; o.y += wy[0] * tm1.y;
    o.x += wy[1] * t0.x; o.y += wy[1] * t0.y;
    o.x += wy[2] * t1.x; o.y += wy[2] * t1.y;
    o.x += wy[3] * t2.x; o.y += wy[3] * t2.y;
    if (emit) {
      if (dst8) *(float2*)dst = o;
      else { dst[0] = o.x; if (has_b) dst[1] = o.y; }
    }
    dst += w2;
    tm1 = t1; t0 = t2;
  }
}

// ---------------------------------------------------------------------------------------------
// a15: Adaptive.  Per pixel: L = up(logits_half) (11 bins), theta = sum softmax(L) * (c-5),
// k = tan(theta deg), a = -h/(pe+1e-8), off = -h/((a-k)+1e-8), m = [0 < off <= depth_scale],
// pe_mask = off*m*y.
// ---------------------------------------------------------------------------------------------
// SlopeEval / slope_eval: tile.cuh (shared with ge_adaptive_x2.cu)

// LOGITS: also write the 11 full-resolution logits (training: CE loss input).  A thread takes 4 pixels of one row
// spaced TX apart, so a warp covers 32 CONSECUTIVE pixels: every global access is one coalesced 128-byte segment
// and the half-resolution taps of neighbouring lanes coincide pairwise (bank-conflict-free shared-memory reads).
template <bool LOGITS>
__global__ void __launch_bounds__(TX * TILE_H) ge_adaptive_fwd_kernel(
    const float* __restrict__ pe_raw, int64_t pe_bstride, const float* __restrict__ y_half,
    const float* __restrict__ logits_half, const float* __restrict__ height, float height_scalar,
    float depth_scale, float* __restrict__ y, float* __restrict__ pe_mask,
    float* __restrict__ logits_full, int H, int W, int h2, int w2, float sy, float sx) {
  __shared__ float s_y[ST_H][ST_W];
  __shared__ float s_l[NSLOPE][ST_H][ST_W];
  const int b = blockIdx.z, oy0 = blockIdx.y * TILE_H, ox0 = blockIdx.x * TILE_W;
  const SrcWindow sw = src_window(oy0, ox0, H, W, h2, w2, sy, sx, false);
  const int64_t hw2 = (int64_t)h2 * w2;
  for (int i = threadIdx.y * TX + threadIdx.x; i < sw.h * sw.w; i += TX * TILE_H) {
    int r = i / sw.w, c = i - r * sw.w;
    const int64_t so = (int64_t)(sw.y0 + r) * w2 + sw.x0 + c;
    s_y[r][c] = __ldg(y_half + b * hw2 + so);
#pragma unroll
    for (int ch = 0; ch < NSLOPE; ++ch) s_l[ch][r][c] = __ldg(logits_half + ((int64_t)b * NSLOPE + ch) * hw2 + so);
  }
  __syncthreads();
  const int oy = oy0 + threadIdx.y;
  if (oy >= H) return;
  const float h = height ? __ldg(height + b) : height_scalar;
  const Tap ty = tap(oy, sy, false, h2);
  const int r0 = ty.i0 - sw.y0, r1 = ty.i1 - sw.y0;
  const int64_t HW = (int64_t)H * W;
  const int64_t row = ((int64_t)b * H + oy) * W;
  const float* pp = pe_raw + (int64_t)b * pe_bstride + (int64_t)oy * W;
#pragma unroll 2
  for (int i = 0; i < 4; ++i) {
    const int ox = ox0 + i * TX + threadIdx.x;
    if (ox >= W) break;
    const Tap tx = tap(ox, sx, false, w2);
    const int c0 = tx.i0 - sw.x0, c1 = tx.i1 - sw.x0;
    float L[NSLOPE];
#pragma unroll
    for (int ch = 0; ch < NSLOPE; ++ch)
      L[ch] = ty.l0 * (tx.l0 * s_l[ch][r0][c0] + tx.l1 * s_l[ch][r0][c1]) +
              ty.l1 * (tx.l0 * s_l[ch][r1][c0] + tx.l1 * s_l[ch][r1][c1]);
    const float yv = ty.l0 * (tx.l0 * s_y[r0][c0] + tx.l1 * s_y[r0][c1]) +
                     ty.l1 * (tx.l0 * s_y[r1][c0] + tx.l1 * s_y[r1][c1]);
    if (LOGITS) {
#pragma unroll
      for (int ch = 0; ch < NSLOPE; ++ch)
        __stcs(logits_full + ((int64_t)b * NSLOPE + ch) * HW + (int64_t)oy * W + ox, L[ch]);
    }
    SlopeEval e;
    slope_eval(L, __ldcs(pp + ox), h, depth_scale, e);
    __stcs(y + row + ox, yv);
    __stcs(pe_mask + row + ox, (e.off * e.m) * yv);
  }
}

// a15 backward.  Recomputes L / softmax / offset from the half-resolution operands (11 B/px read
// instead of 44), forms the 12 per-pixel gradients (11 logits + y) in shared memory and applies
// the tile adjoint.  g_y_half and g_logits_half must be zeroed.
__global__ void __launch_bounds__(TX * TILE_H) ge_adaptive_bwd_kernel(
    const float* __restrict__ pe_raw, int64_t pe_bstride, const float* __restrict__ y_half,
    const float* __restrict__ logits_half, const float* __restrict__ height, float height_scalar,
    float depth_scale, const float* __restrict__ g_y, const float* __restrict__ g_pe_mask,
    const float* __restrict__ g_logits_full, float* __restrict__ g_y_half,
    float* __restrict__ g_logits_half, int H, int W, int h2, int w2, float sy, float sx) {
  extern __shared__ float smem[];
  float (*s_g)[TILE_H][TILE_W + 1] = (float (*)[TILE_H][TILE_W + 1])smem;            // 12 channels
  float (*s_y)[ST_W] = (float (*)[ST_W])(smem + (NSLOPE + 1) * TILE_H * (TILE_W + 1));
  float (*s_l)[ST_H][ST_W] = (float (*)[ST_H][ST_W])(smem + (NSLOPE + 1) * TILE_H * (TILE_W + 1) + ST_H * ST_W);
  const int b = blockIdx.z, oy0 = blockIdx.y * TILE_H, ox0 = blockIdx.x * TILE_W;
  const SrcWindow sw = src_window(oy0, ox0, H, W, h2, w2, sy, sx, false);
  const int64_t hw2 = (int64_t)h2 * w2, HW = (int64_t)H * W;
  for (int i = threadIdx.y * TX + threadIdx.x; i < sw.h * sw.w; i += TX * TILE_H) {
    int r = i / sw.w, c = i - r * sw.w;
    const int64_t so = (int64_t)(sw.y0 + r) * w2 + sw.x0 + c;
    s_y[r][c] = __ldg(y_half + b * hw2 + so);
#pragma unroll
    for (int ch = 0; ch < NSLOPE; ++ch) s_l[ch][r][c] = __ldg(logits_half + ((int64_t)b * NSLOPE + ch) * hw2 + so);
  }
  __syncthreads();
  const int oy = oy0 + threadIdx.y;
  const float h = height ? __ldg(height + b) : height_scalar;
  if (oy < H) {
    const Tap ty = tap(oy, sy, false, h2);
    const int r0 = ty.i0 - sw.y0, r1 = ty.i1 - sw.y0;
    for (int i = 0; i < 4; ++i) {
      const int lx = i * TX + threadIdx.x, ox = ox0 + lx;   // lane-contiguous: coalesced, conflict-free
      if (ox >= W) {
#pragma unroll
        for (int ch = 0; ch <= NSLOPE; ++ch) s_g[ch][threadIdx.y][lx] = 0.f;
        continue;
      }
      const Tap tx = tap(ox, sx, false, w2);
      const int c0 = tx.i0 - sw.x0, c1 = tx.i1 - sw.x0;
      float L[NSLOPE];
#pragma unroll
      for (int ch = 0; ch < NSLOPE; ++ch)
        L[ch] = ty.l0 * (tx.l0 * s_l[ch][r0][c0] + tx.l1 * s_l[ch][r0][c1]) +
                ty.l1 * (tx.l0 * s_l[ch][r1][c0] + tx.l1 * s_l[ch][r1][c1]);
      const float yv = ty.l0 * (tx.l0 * s_y[r0][c0] + tx.l1 * s_y[r0][c1]) +
                       ty.l1 * (tx.l0 * s_y[r1][c0] + tx.l1 * s_y[r1][c1]);
      const int64_t po = (int64_t)oy * W + ox;
      const float pe = __ldg(pe_raw + (int64_t)b * pe_bstride + po);
      SlopeEval e;
      slope_eval(L, pe, h, depth_scale, e);
      const float gpm = g_pe_mask ? __ldg(g_pe_mask + b * HW + po) : 0.f;
      // d pe_mask / d y = off*m ; d pe_mask / d off = m*y ; d off / d k = -h/den^2 ;
      // d k / d theta = (pi/180)(1+k^2) ; d theta / d L_c = p_c (c-5 - theta).  m is a constant.
      const float gy = (g_y ? __ldg(g_y + b * HW + po) : 0.f) + gpm * (e.off * e.m);
      float G = gpm * e.m * yv * (-h / (e.den * e.den)) * (0.017453292519943295f * (1.f + e.k * e.k));
      if (e.m == 0.f) G = 0.f;   // 0 * inf guards: the reference multiplies by an exact-zero mask
      s_g[NSLOPE][threadIdx.y][lx] = gy;
#pragma unroll
      for (int ch = 0; ch < NSLOPE; ++ch) {
        float gl = G * e.p[ch] * ((float)(ch - 5) - e.theta);
        if (g_logits_full) gl += __ldg(g_logits_full + ((int64_t)b * NSLOPE + ch) * HW + po);
        s_g[ch][threadIdx.y][lx] = gl;
      }
    }
  }
  __syncthreads();
  // channels 0..10 -> g_logits_half, channel 11 -> g_y_half
  const int tid = threadIdx.y * TX + threadIdx.x;
  const int th = min(TILE_H, H - oy0), tw = min(TILE_W, W - ox0);
  for (int i = tid; i < sw.h * sw.w; i += TX * TILE_H) {
    const int r = i / sw.w, c = i - r * sw.w;
    const int j = sw.y0 + r, k = sw.x0 + c;
    float wyv[TILE_H];
    bool any_y = false;
#pragma unroll
    for (int yy = 0; yy < TILE_H; ++yy) {
      wyv[yy] = 0.f;
      if (yy < th) {
        const Tap ty = tap(oy0 + yy, sy, false, h2);
        wyv[yy] = (ty.i0 == j ? ty.l0 : 0.f) + (ty.i1 == j ? ty.l1 : 0.f);
      }
      any_y |= wyv[yy] != 0.f;
    }
    if (!any_y) continue;
    int xlo, xhi;
    adjoint_range(k, sx, false, w2, W, xlo, xhi);
    xlo = max(xlo, ox0) - ox0; xhi = min(xhi, ox0 + tw - 1) - ox0;
    float acc[NSLOPE + 1];
#pragma unroll
    for (int ch = 0; ch <= NSLOPE; ++ch) acc[ch] = 0.f;
    bool any = false;
    for (int xx = xlo; xx <= xhi; ++xx) {
      const Tap tx = tap(ox0 + xx, sx, false, w2);
      const float wx = (tx.i0 == k ? tx.l0 : 0.f) + (tx.i1 == k ? tx.l1 : 0.f);
      if (wx == 0.f) continue;
      any = true;
#pragma unroll
      for (int yy = 0; yy < TILE_H; ++yy) {
        const float w = wyv[yy] * wx;
        if (w != 0.f) {
#pragma unroll
          for (int ch = 0; ch <= NSLOPE; ++ch) acc[ch] += w * s_g[ch][yy][xx];
        }
      }
    }
    if (any) {
      const int64_t so = (int64_t)j * w2 + k;
#pragma unroll
      for (int ch = 0; ch < NSLOPE; ++ch) atomicAdd(g_logits_half + ((int64_t)b * NSLOPE + ch) * hw2 + so, acc[ch]);
      atomicAdd(g_y_half + b * hw2 + so, acc[NSLOPE]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// a17: head fusion at half resolution.  pe_h, y_h = down(pe_mask), down(y) (align_corners=True);
// out = d*(1-y_h) + pe_h + min_depth.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fuse_head_fwd_kernel(
    const float* __restrict__ d, const float* __restrict__ pe_mask, const float* __restrict__ y,
    float* __restrict__ out, float* __restrict__ y_h, int H, int W, int h2, int w2, float sy, float sx,
    float min_depth) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, yy = blockIdx.y, b = blockIdx.z;
  if (x >= w2) return;
  const Tap ty = tap(yy, sy, true, H), tx = tap(x, sx, true, W);
  const int64_t base = (int64_t)b * H * W;
  const int64_t o00 = base + (int64_t)ty.i0 * W + tx.i0, o01 = base + (int64_t)ty.i0 * W + tx.i1;
  const int64_t o10 = base + (int64_t)ty.i1 * W + tx.i0, o11 = base + (int64_t)ty.i1 * W + tx.i1;
  const float pe_h = ty.l0 * (tx.l0 * __ldg(pe_mask + o00) + tx.l1 * __ldg(pe_mask + o01)) +
                     ty.l1 * (tx.l0 * __ldg(pe_mask + o10) + tx.l1 * __ldg(pe_mask + o11));
  const float yh = ty.l0 * (tx.l0 * __ldg(y + o00) + tx.l1 * __ldg(y + o01)) +
                   ty.l1 * (tx.l0 * __ldg(y + o10) + tx.l1 * __ldg(y + o11));
  const int64_t o = ((int64_t)b * h2 + yy) * w2 + x;
  out[o] = ((__ldg(d + o) * (1.f - yh)) + pe_h) + min_depth;
  y_h[o] = yh;
}

// backward at half resolution: g_d = g_out*(1-y_h); the full-resolution gradients are gathered
// (deterministically) from the <=3x3 half-resolution pixels whose taps touch each pixel.
__global__ void __launch_bounds__(256) fuse_head_bwd_half_kernel(
    const float* __restrict__ g_out, const float* __restrict__ y_h, float* __restrict__ g_d, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) g_d[i] = __ldg(g_out + i) * (1.f - __ldg(y_h + i));
}

__global__ void __launch_bounds__(256) fuse_head_bwd_full_kernel(
    const float* __restrict__ g_out, const float* __restrict__ d, const float* __restrict__ g_yh_extra,
    float* __restrict__ g_pe_mask, float* __restrict__ g_y, int H, int W, int h2, int w2, float sy,
    float sx) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, yy = blockIdx.y, b = blockIdx.z;
  if (x >= W) return;
  int jlo, jhi, klo, khi;
  adjoint_range(yy, sy, true, H, h2, jlo, jhi);
  adjoint_range(x, sx, true, W, w2, klo, khi);
  float gp = 0.f, gy = 0.f;
  for (int j = jlo; j <= jhi; ++j) {
    const Tap ty = tap(j, sy, true, H);
    const float wy = (ty.i0 == yy ? ty.l0 : 0.f) + (ty.i1 == yy ? ty.l1 : 0.f);
    if (wy == 0.f) continue;
    for (int k = klo; k <= khi; ++k) {
      const Tap tx = tap(k, sx, true, W);
      const float wx = (tx.i0 == x ? tx.l0 : 0.f) + (tx.i1 == x ? tx.l1 : 0.f);
      if (wx == 0.f) continue;
      const int64_t o = ((int64_t)b * h2 + j) * w2 + k;
      const float go = __ldg(g_out + o);
      gp += wy * wx * go;
      gy += wy * wx * (-go * __ldg(d + o) + (g_yh_extra ? __ldg(g_yh_extra + o) : 0.f));
    }
  }
  const int64_t o = ((int64_t)b * H + yy) * W + x;
  g_pe_mask[o] = gp;
  g_y[o] = gy;
}

// ---------------------------------------------------------------------------------------------
// (f)1: slope labels.  k = deg(atan(h/gt - h/pe)); KITTI rounds half-to-even (np.around),
// DDAD truncates (astype(int64)); clip to +-5; 255 where gt == 0.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) find_k_kernel(const float* __restrict__ gt,
                                                      const float* __restrict__ pe, int64_t pe_bstride,
                                                      float* __restrict__ k_out, int64_t HW, double h,
                                                      int truncate) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (i >= HW) return;
  const double g = (double)__ldg(gt + b * HW + i);
  const float pf = __ldg(pe + b * pe_bstride + i);
  // the arithmetic types of the reference scripts: KITTI divides -1.65 by the float32 plane IN float32 and adds it to
  // the float64 1.65 / gt (preprocess_data_kitti.py:59-63,81); DDAD keeps everything in float64 (preprocess_data_ddad.py:47-51)
  const double a = truncate ? (-h) / (double)pf : (double)__fdiv_rn(-(float)h, pf);
  double k = h / g + a;
  k = atan(k) * 57.29577951308232;
  double r;
  if (truncate) r = isfinite(k) ? trunc(k) : 0.0; else r = rint(k);
  if (r > 5.0) r = 5.0;
  if (r < -5.0) r = -5.0;
  if (g == 0.0) r = 255.0;
  k_out[b * HW + i] = (float)r;
}

}  // namespace ged

namespace ged {
// ge_adaptive_x2.cu: exact x2 kernels; 0 = launched, 1 = not eligible (generic kernels below), < 0 = error
int launch_ge_adaptive_fwd_x2(const float*, int64_t, const float*, const float*, const float*, float, float, float*, float*,
                              float*, int, int, int, int, int, cudaStream_t);
int launch_ge_adaptive_bwd_x2(const float*, int64_t, const float*, const float*, const float*, float, float, const float*,
                              const float*, const float*, float*, float*, int, int, int, int, int, cudaStream_t);
void set_ge_x2_tma(int on);
}  // namespace ged

using namespace ged;

// 1 = closed-form x2 kernels whenever the shape allows, staged by TMA where the rows allow (default); 2 = x2 kernels staged
// by per-thread asynchronous copies only; 0 = generic bilinear kernels only (A/B, tests)
static int g_ge_x2 = 1;
GED_API int ged_set_ge_x2(int on) {
  const int prev = g_ge_x2;
  g_ge_x2 = on == 2 ? 2 : (on ? 1 : 0);
  set_ge_x2_tma(g_ge_x2 == 1);
  return prev;
}

// ============================================================================================
// C-ABI (include/gedepth.h)
// ============================================================================================
GED_API int ged_ground_plane(float* ch3, float* ch4, int64_t batch_stride3, int64_t batch_stride4,
                             int B, int H, int W, const double* coef4, double u0, double v0,
                             double su, double sv, float depth_scale, float clamp_max,
                             cudaStream_t stream) {
  if (!ch3 || !ch4 || !coef4 || B <= 0 || H <= 0 || W <= 0) return GED_ERR_ARG;
  dim3 block(256), grid(cdiv(cdiv(W, 4), 256), H);
  ground_plane_kernel<<<grid, block, 0, stream>>>(ch3, ch4, batch_stride3, batch_stride4, B, H, W,
                                                 coef4[0], coef4[1], coef4[2], coef4[3], u0, v0, su, sv,
                                                 depth_scale, clamp_max);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

GED_API int ged_pixel_grid(long long* u, long long* v, int H, int W, int u0, int v0,
                           cudaStream_t stream) {
  if (!u || !v || H <= 0 || W <= 0) return GED_ERR_ARG;
  pixel_grid_kernel<<<dim3(cdiv(W, 256), H), 256, 0, stream>>>(u, v, H, W, u0, v0);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

static inline bool half_shape_ok(int H, int W, int h2, int w2) {
  // the staged source window of one tile must fit the shared-memory tile (half map ~ H/2 x W/2)
  if (h2 <= 0 || w2 <= 0 || H <= 0 || W <= 0) return false;
  const float sy = resize_scale(h2, H, false), sx = resize_scale(w2, W, false);
  const int th = H < TILE_H ? H : TILE_H, tw = W < TILE_W ? W : TILE_W;
  return (float)th * sy + 3.f <= (float)ST_H && (float)tw * sx + 3.f <= (float)ST_W;
}

GED_API int ged_ge_vanilla_fwd(const float* pe_norm, int64_t pe_batch_stride, const float* y_half,
                               float* y, float* pe_mask, int B, int H, int W, int h2, int w2,
                               cudaStream_t stream) {
  if (!pe_norm || !y_half || !y || !pe_mask || B <= 0) return GED_ERR_ARG;
  if (!half_shape_ok(H, W, h2, w2)) return GED_ERR_SHAPE;
  const float sy = resize_scale(h2, H, false), sx = resize_scale(w2, W, false);
  dim3 block(TX, 4), grid(cdiv(W, TILE_W), cdiv(H, VT_H), B);
  if ((float)(H < VT_H ? H : VT_H) * sy + 3.f > (float)VS_H) return GED_ERR_SHAPE;
  const bool vec = (W % 4 == 0) && aligned16(pe_norm) && aligned16(y) && aligned16(pe_mask) &&
                   (pe_batch_stride % 4 == 0);
  if (g_ge_x2 == 1 && vec && H == 2 * h2 && W == 2 * w2 && ((((uintptr_t)y_half) & 7) == 0)) {
    // exact x2 (every GE config): streaming kernel; ged_set_ge_x2(2 / 0) keeps the tiled one for A/B
    // rows per warp: 16, or 8 when that leaves less than one wave of warps (a warp's rows are a serial chain of loads).
    // Measured (fraction of the HBM peak at 32 / 16 / 8 rows per warp): 64x352x1120 0.73 / 0.78 / 0.78,
    // 64x384x640 0.60 / 0.74 / 0.71, 32x1024x2048 0.85 / 0.89 / 0.88, 64x1024x2048 0.92 / 0.93 / 0.91, 8x352x1120 0.28 / 0.36 / 0.44
    const int64_t strips_b = (int64_t)B * cdiv(W, 128);
    const int fr = strips_b * cdiv(H, 16) >= 148 * 36 ? 16 : 8;
#define VF_LAUNCH(FR) ge_vanilla_fwd_x2s_kernel<FR><<<dim3(cdiv(W, 512), cdiv(H, FR), B), dim3(32, 4), 0, stream>>>(pe_norm, pe_batch_stride, \
        y_half, y, pe_mask, H, W, h2, w2)
    if (fr == 16) VF_LAUNCH(16); else VF_LAUNCH(8);
#undef VF_LAUNCH
    GED_CHECK_LAUNCH();
    return GED_OK;
  }
  if (vec) ge_vanilla_fwd_kernel<true><<<grid, block, 0, stream>>>(pe_norm, pe_batch_stride, y_half, y, pe_mask, H, W, h2, w2, sy, sx);
  else ge_vanilla_fwd_kernel<false><<<grid, block, 0, stream>>>(pe_norm, pe_batch_stride, y_half, y, pe_mask, H, W, h2, w2, sy, sx);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

GED_API int ged_ge_vanilla_bwd(const float* pe_norm, int64_t pe_batch_stride, const float* g_y,
                               const float* g_pe_mask, float* g_y_half, int B, int H, int W, int h2,
                               int w2, cudaStream_t stream) {
  if (!pe_norm || !g_y_half || B <= 0) return GED_ERR_ARG;
  if (!half_shape_ok(H, W, h2, w2)) return GED_ERR_SHAPE;
  const float sy = resize_scale(h2, H, false), sx = resize_scale(w2, W, false);
  if (g_ge_x2 && H == 2 * h2 && W == 2 * w2 && (W % 4 == 0) && (pe_batch_stride % 4 == 0) && aligned16(pe_norm) &&
      (!g_y || aligned16(g_y)) && (!g_pe_mask || aligned16(g_pe_mask))) {
    // exact x2: closed-form gather (every GE config); ged_set_ge_x2(2) keeps the tiled round-2a kernel for A/B
    if (g_ge_x2 == 1) {
      // 6 CTAs of 4 warps per SM (measured against 4 / 5 CTAs)
      dim3 block(32, 4);
      // half-resolution rows per warp: 16 with >= 3 waves of warps, else 8, else 4 (measured at 16 / 8 / 4 rows: 64x1024x2048
      // 0.92 / 0.90 / 0.84, 32x1024x2048 0.85 / 0.84 / 0.80, 64x352x1120 0.66 / 0.70 / 0.69, 64x384x640 0.61 / 0.66 / 0.63)
      const int64_t strips_b = (int64_t)B * cdiv(W, 128);
      const int ch = strips_b * cdiv(h2, 16) >= 148 * 24 * 3 ? 16 : (strips_b * cdiv(h2, 8) >= 148 * 24 ? 8 : 4);
#define VB_LAUNCH(CH) ge_vanilla_bwd_x2s_kernel<CH, 6, 1><<<dim3(cdiv(W, 512), cdiv(h2, CH), B), block, 0, stream>>>(pe_norm, pe_batch_stride, \
          g_y, g_pe_mask, g_y_half, H, W, h2, w2)
      if (ch == 16) VB_LAUNCH(16); else if (ch == 8) VB_LAUNCH(8); else VB_LAUNCH(4);
#undef VB_LAUNCH
    } else {
      dim3 block(64, 4), grid(cdiv(cdiv(w2, 2), 64), cdiv(h2, X2_H), B);
      ge_vanilla_bwd_x2_kernel<<<grid, block, 0, stream>>>(pe_norm, pe_batch_stride, g_y, g_pe_mask, g_y_half, H, W, h2, w2);
    }
    GED_CHECK_LAUNCH();
    return GED_OK;
  }
  if (cudaMemsetAsync(g_y_half, 0, sizeof(float) * (size_t)B * h2 * w2, stream) != cudaSuccess) return GED_ERR_LAUNCH;
  dim3 block(TX, TILE_H), grid(cdiv(W, TILE_W), cdiv(H, TILE_H), B);
  ge_vanilla_bwd_kernel<<<grid, block, 0, stream>>>(pe_norm, pe_batch_stride, g_y, g_pe_mask, g_y_half, H, W, h2, w2, sy, sx);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

GED_API int ged_ge_adaptive_fwd(const float* pe_raw, int64_t pe_batch_stride, const float* y_half,
                                const float* logits_half, const float* height, float height_scalar,
                                float depth_scale, float* y, float* pe_mask, float* logits_full, int B,
                                int H, int W, int h2, int w2, cudaStream_t stream) {
  if (!pe_raw || !y_half || !logits_half || !y || !pe_mask || B <= 0) return GED_ERR_ARG;
  if (!half_shape_ok(H, W, h2, w2)) return GED_ERR_SHAPE;
  if ((W % 4 == 0) && !(aligned16(y) && aligned16(pe_mask) && (!logits_full || aligned16(logits_full)))) return GED_ERR_ALIGN;
  if (g_ge_x2) {
    const int rc = launch_ge_adaptive_fwd_x2(pe_raw, pe_batch_stride, y_half, logits_half, height, height_scalar, depth_scale, y,
                                             pe_mask, logits_full, B, H, W, h2, w2, stream);
    if (rc <= 0) return rc;
  }
  const float sy = resize_scale(h2, H, false), sx = resize_scale(w2, W, false);
  dim3 block(TX, TILE_H), grid(cdiv(W, TILE_W), cdiv(H, TILE_H), B);
  if (logits_full)
    ge_adaptive_fwd_kernel<true><<<grid, block, 0, stream>>>(pe_raw, pe_batch_stride, y_half, logits_half, height, height_scalar, depth_scale, y, pe_mask, logits_full, H, W, h2, w2, sy, sx);
  else
    ge_adaptive_fwd_kernel<false><<<grid, block, 0, stream>>>(pe_raw, pe_batch_stride, y_half, logits_half, height, height_scalar, depth_scale, y, pe_mask, logits_full, H, W, h2, w2, sy, sx);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

GED_API int ged_ge_adaptive_bwd(const float* pe_raw, int64_t pe_batch_stride, const float* y_half,
                                const float* logits_half, const float* height, float height_scalar,
                                float depth_scale, const float* g_y, const float* g_pe_mask,
                                const float* g_logits_full, float* g_y_half, float* g_logits_half,
                                int B, int H, int W, int h2, int w2, cudaStream_t stream) {
  if (!pe_raw || !y_half || !logits_half || !g_y_half || !g_logits_half || B <= 0) return GED_ERR_ARG;
  if (!half_shape_ok(H, W, h2, w2)) return GED_ERR_SHAPE;
  const float sy = resize_scale(h2, H, false), sx = resize_scale(w2, W, false);
  if (g_ge_x2) {
    const int rc = launch_ge_adaptive_bwd_x2(pe_raw, pe_batch_stride, y_half, logits_half, height, height_scalar, depth_scale, g_y,
                                             g_pe_mask, g_logits_full, g_y_half, g_logits_half, B, H, W, h2, w2, stream);
    if (rc <= 0) return rc;
  }
  const size_t smem = sizeof(float) * ((NSLOPE + 1) * TILE_H * (TILE_W + 1) + ST_H * ST_W + NSLOPE * ST_H * ST_W);
  // the attribute is per device: set it on every call (cheap), not once per process
  if (cudaFuncSetAttribute(ge_adaptive_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return GED_ERR_LAUNCH;
  if (cudaMemsetAsync(g_y_half, 0, sizeof(float) * (size_t)B * h2 * w2, stream) != cudaSuccess) return GED_ERR_LAUNCH;
  if (cudaMemsetAsync(g_logits_half, 0, sizeof(float) * (size_t)B * NSLOPE * h2 * w2, stream) != cudaSuccess) return GED_ERR_LAUNCH;
  dim3 block(TX, TILE_H), grid(cdiv(W, TILE_W), cdiv(H, TILE_H), B);
  ge_adaptive_bwd_kernel<<<grid, block, smem, stream>>>(pe_raw, pe_batch_stride, y_half, logits_half, height, height_scalar, depth_scale, g_y, g_pe_mask, g_logits_full, g_y_half, g_logits_half, H, W, h2, w2, sy, sx);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

GED_API int ged_fuse_head_fwd(const float* d, const float* pe_mask, const float* y, float* out,
                              float* y_h, float min_depth, int B, int H, int W, int h2, int w2,
                              cudaStream_t stream) {
  if (!d || !pe_mask || !y || !out || !y_h || B <= 0) return GED_ERR_ARG;
  const float sy = resize_scale(H, h2, true), sx = resize_scale(W, w2, true);
  fuse_head_fwd_kernel<<<dim3(cdiv(w2, 256), h2, B), 256, 0, stream>>>(d, pe_mask, y, out, y_h, H, W, h2, w2, sy, sx, min_depth);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

GED_API int ged_fuse_head_bwd(const float* g_out, const float* g_yh_extra, const float* d,
                              const float* y_h, float* g_d, float* g_pe_mask, float* g_y, int B, int H,
                              int W, int h2, int w2, cudaStream_t stream) {
  if (!g_out || !d || !y_h || !g_d || !g_pe_mask || !g_y || B <= 0) return GED_ERR_ARG;
  const float sy = resize_scale(H, h2, true), sx = resize_scale(W, w2, true);
  const int64_t n = (int64_t)B * h2 * w2;
  fuse_head_bwd_half_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(g_out, y_h, g_d, n);
  fuse_head_bwd_full_kernel<<<dim3(cdiv(W, 256), H, B), 256, 0, stream>>>(g_out, d, g_yh_extra, g_pe_mask, g_y, H, W, h2, w2, sy, sx);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

GED_API int ged_find_k(const float* gt, const float* pe, int64_t pe_batch_stride, float* k_out, int B,
                       int H, int W, double cam_height, int truncate, cudaStream_t stream) {
  if (!gt || !pe || !k_out || B <= 0) return GED_ERR_ARG;
  const int64_t HW = (int64_t)H * W;
  find_k_kernel<<<dim3((unsigned)((HW + 255) / 256), B), 256, 0, stream>>>(gt, pe, pe_batch_stride, k_out, HW, cam_height, truncate);
  GED_CHECK_LAUNCH();
  return GED_OK;
}
