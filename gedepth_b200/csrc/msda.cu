// Multi-scale deformable attention sampling, sm_100a (a11: depth/models/necks/hahi.py:280-289,
// 316-325 call mmcv.ops.MultiScaleDeformableAttention [external, mmcv-full 1.3.13]; semantics
// restated from multi_scale_deformable_attn_pytorch: softmax over the 32 (level, point) weights,
// loc = ref + offset / (W_l, H_l), bilinear sample with zero padding at pixel = loc*size - 0.5).
//
// One warp per (batch, query, head); a lane owns 2 of the 64 head channels, so every corner fetch is
// one coalesced 256-byte segment value[b, pos, head, 0:64].  A CTA takes 8 CONSECUTIVE queries of one
// head: neighbouring stem pixels sample neighbouring positions, so corner segments are re-used from
// L1.  Softmax, location arithmetic and the gather are one kernel (the reference runs 6).
// Gather-bound (L1/L2), not a tensor-core op.
#include "common.cuh"

namespace ged {

constexpr int MS_L = 4, MS_P = 8, MS_HD = 64, MS_WARPS = 8;

struct MsdaShapes {
  int h[MS_L], w[MS_L], start[MS_L];
};

struct PointGeom {
  int64_t o00, o01, o10, o11;   // element offsets of the four corners (pos * nH*HD), -1 when outside
  float w00, w01, w10, w11;     // bilinear weights
  float lx, ly;                 // fractional parts (for the location gradient)
};

__device__ __forceinline__ PointGeom point_geom(float x, float y, int H, int W, int start, int rowpitch) {
  PointGeom g;
  const float xf = floorf(x), yf = floorf(y);
  const int x0 = (int)xf, y0 = (int)yf;
  g.lx = x - xf; g.ly = y - yf;
  const bool vx0 = x0 >= 0 && x0 < W, vx1 = x0 + 1 >= 0 && x0 + 1 < W;
  const bool vy0 = y0 >= 0 && y0 < H, vy1 = y0 + 1 >= 0 && y0 + 1 < H;
  g.w00 = (1.f - g.ly) * (1.f - g.lx); g.w01 = (1.f - g.ly) * g.lx;
  g.w10 = g.ly * (1.f - g.lx);         g.w11 = g.ly * g.lx;
  const int64_t base = (int64_t)start + (int64_t)y0 * W + x0;
  g.o00 = (vy0 && vx0) ? base * rowpitch : -1;
  g.o01 = (vy0 && vx1) ? (base + 1) * rowpitch : -1;
  g.o10 = (vy1 && vx0) ? (base + W) * rowpitch : -1;
  g.o11 = (vy1 && vx1) ? (base + W + 1) * rowpitch : -1;
  return g;
}

// lane = (level, point): load the lane's logit/offset, softmax across the warp, pixel coordinates
__device__ __forceinline__ void lane_point(const float* __restrict__ off, const float* __restrict__ logit,
                                           float rx, float ry, const MsdaShapes& sh, int lane, float& aw,
                                           float& px, float& py) {
  const float lg = __ldg(logit + lane);
  const float mx = warp_max(lg);
  const float e = __expf(lg - mx);
  aw = e / warp_sum(e);
  const float2 o = __ldg((const float2*)off + lane);
  const int l = lane >> 3;
  const float Wl = (float)sh.w[l], Hl = (float)sh.h[l];
  px = (rx + o.x / Wl) * Wl - 0.5f;
  py = (ry + o.y / Hl) * Hl - 0.5f;
}

__global__ void __launch_bounds__(MS_WARPS * 32) msda_fwd_kernel(
    const float* __restrict__ value, const float* __restrict__ ref, const float* __restrict__ off,
    const float* __restrict__ logit, float* __restrict__ out, MsdaShapes sh, int B, int S, int Q, int nH,
    int ref_bstride) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * MS_WARPS + warp, h = blockIdx.y, b = blockIdx.z;
  if (q >= Q) return;
  const int64_t bq = (int64_t)b * Q + q;
  const float rx = __ldg(ref + (int64_t)b * ref_bstride + q * 2), ry = __ldg(ref + (int64_t)b * ref_bstride + q * 2 + 1);
  float aw, px, py;
  lane_point(off + (bq * nH + h) * (MS_L * MS_P * 2), logit + (bq * nH + h) * (MS_L * MS_P), rx, ry, sh, lane, aw, px, py);
  const int rowpitch = nH * MS_HD;
  const float* vb = value + (int64_t)b * S * rowpitch + h * MS_HD + lane * 2;
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll 4
  for (int j = 0; j < MS_L * MS_P; ++j) {
    const float x = __shfl_sync(0xffffffffu, px, j), y = __shfl_sync(0xffffffffu, py, j);
    const float a = __shfl_sync(0xffffffffu, aw, j);
    const int l = j >> 3;
    const PointGeom g = point_geom(x, y, sh.h[l], sh.w[l], sh.start[l], rowpitch);
    float2 v00 = make_float2(0.f, 0.f), v01 = v00, v10 = v00, v11 = v00;
    if (g.o00 >= 0) v00 = __ldg((const float2*)(vb + g.o00));
    if (g.o01 >= 0) v01 = __ldg((const float2*)(vb + g.o01));
    if (g.o10 >= 0) v10 = __ldg((const float2*)(vb + g.o10));
    if (g.o11 >= 0) v11 = __ldg((const float2*)(vb + g.o11));
    acc.x += a * (g.w00 * v00.x + g.w01 * v01.x + g.w10 * v10.x + g.w11 * v11.x);
    acc.y += a * (g.w00 * v00.y + g.w01 * v01.y + g.w10 * v10.y + g.w11 * v11.y);
  }
  *(float2*)(out + bq * rowpitch + h * MS_HD + lane * 2) = acc;
}

// butterfly reduce-scatter: every lane enters with N partial sums, leaves with N/2
template <int N>
__device__ __forceinline__ void halve(const float* in, float* outv, int mask, bool upper) {
#pragma unroll
  for (int i = 0; i < N / 2; ++i) {
    const float send = upper ? in[i] : in[i + N / 2];
    const float keep = upper ? in[i + N / 2] : in[i];
    outv[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
  }
}

__global__ void __launch_bounds__(MS_WARPS * 32) msda_bwd_kernel(
    const float* __restrict__ value, const float* __restrict__ ref, const float* __restrict__ off,
    const float* __restrict__ logit, const float* __restrict__ g_out, float* __restrict__ g_value,
    float* __restrict__ g_ref, float* __restrict__ g_off, float* __restrict__ g_logit, MsdaShapes sh,
    int B, int S, int Q, int nH, int ref_bstride) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * MS_WARPS + warp, h = blockIdx.y, b = blockIdx.z;
  if (q >= Q) return;
  const int64_t bq = (int64_t)b * Q + q;
  const float rx = __ldg(ref + (int64_t)b * ref_bstride + q * 2), ry = __ldg(ref + (int64_t)b * ref_bstride + q * 2 + 1);
  float aw, px, py;
  lane_point(off + (bq * nH + h) * (MS_L * MS_P * 2), logit + (bq * nH + h) * (MS_L * MS_P), rx, ry, sh, lane, aw, px, py);
  const int rowpitch = nH * MS_HD;
  const int64_t voff = (int64_t)b * S * rowpitch + h * MS_HD + lane * 2;
  const float* vb = value + voff;
  float* gvb = g_value + voff;
  const float2 go = __ldg((const float2*)(g_out + bq * rowpitch + h * MS_HD + lane * 2));
  float part[96];   // [point][gw, gx, gy] partial sums over this lane's two channels
#pragma unroll
  for (int j = 0; j < MS_L * MS_P; ++j) {
    const float x = __shfl_sync(0xffffffffu, px, j), y = __shfl_sync(0xffffffffu, py, j);
    const float a = __shfl_sync(0xffffffffu, aw, j);
    const int l = j >> 3;
    const PointGeom g = point_geom(x, y, sh.h[l], sh.w[l], sh.start[l], rowpitch);
    float2 v00 = make_float2(0.f, 0.f), v01 = v00, v10 = v00, v11 = v00;
    if (g.o00 >= 0) { v00 = __ldg((const float2*)(vb + g.o00)); atomicAdd((float2*)(gvb + g.o00), make_float2(go.x * a * g.w00, go.y * a * g.w00)); }
    if (g.o01 >= 0) { v01 = __ldg((const float2*)(vb + g.o01)); atomicAdd((float2*)(gvb + g.o01), make_float2(go.x * a * g.w01, go.y * a * g.w01)); }
    if (g.o10 >= 0) { v10 = __ldg((const float2*)(vb + g.o10)); atomicAdd((float2*)(gvb + g.o10), make_float2(go.x * a * g.w10, go.y * a * g.w10)); }
    if (g.o11 >= 0) { v11 = __ldg((const float2*)(vb + g.o11)); atomicAdd((float2*)(gvb + g.o11), make_float2(go.x * a * g.w11, go.y * a * g.w11)); }
    // sample and its derivatives w.r.t. the pixel coordinates, dotted with g_out over channels
    const float sx_ = g.w00 * v00.x + g.w01 * v01.x + g.w10 * v10.x + g.w11 * v11.x;
    const float sy_ = g.w00 * v00.y + g.w01 * v01.y + g.w10 * v10.y + g.w11 * v11.y;
    const float dxx = (1.f - g.ly) * (v01.x - v00.x) + g.ly * (v11.x - v10.x);
    const float dxy = (1.f - g.ly) * (v01.y - v00.y) + g.ly * (v11.y - v10.y);
    const float dyx = (1.f - g.lx) * (v10.x - v00.x) + g.lx * (v11.x - v01.x);
    const float dyy = (1.f - g.lx) * (v10.y - v00.y) + g.lx * (v11.y - v01.y);
    part[j * 3 + 0] = go.x * sx_ + go.y * sy_;
    part[j * 3 + 1] = a * (go.x * dxx + go.y * dxy);
    part[j * 3 + 2] = a * (go.x * dyx + go.y * dyy);
  }
  float p48[48], p24[24], p12[12], p6[6], p3[3];
  halve<96>(part, p48, 16, (lane & 16) != 0);
  halve<48>(p48, p24, 8, (lane & 8) != 0);
  halve<24>(p24, p12, 4, (lane & 4) != 0);
  halve<12>(p12, p6, 2, (lane & 2) != 0);
  halve<6>(p6, p3, 1, (lane & 1) != 0);
  // lane now holds (gw, g_xpix, g_ypix) of point `lane`
  const float gw = p3[0], gx = p3[1], gy = p3[2];
  const float dot = warp_sum(aw * gw);
  g_logit[(bq * nH + h) * (MS_L * MS_P) + lane] = aw * (gw - dot);
  // x_pix = (ref + off/W)*W - 0.5  ->  d/d off = 1, d/d ref = W_l
  *((float2*)g_off + (bq * nH + h) * (MS_L * MS_P) + lane) = make_float2(gx, gy);
  if (g_ref) {
    const int l = lane >> 3;
    const float grx = warp_sum(gx * (float)sh.w[l]), gry = warp_sum(gy * (float)sh.h[l]);
    if (lane == 0) { atomicAdd(g_ref + bq * 2, grx); atomicAdd(g_ref + bq * 2 + 1, gry); }
  }
}

}  // namespace ged
using namespace ged;

static int fill_shapes(const int* hw, int L, int S, MsdaShapes& sh) {
  if (L != MS_L) return GED_ERR_SHAPE;
  int start = 0;
  for (int l = 0; l < MS_L; ++l) {
    sh.h[l] = hw[2 * l]; sh.w[l] = hw[2 * l + 1]; sh.start[l] = start;
    if (sh.h[l] <= 0 || sh.w[l] <= 0) return GED_ERR_SHAPE;
    start += sh.h[l] * sh.w[l];
  }
  return start == S ? GED_OK : GED_ERR_SHAPE;
}

// value (B,S,nH,64); ref (Bref,Q,2) Bref in {1,B}; off (B,Q,nH,4,8,2); logit (B,Q,nH,32); out (B,Q,nH*64)
GED_API int ged_msda_fwd(const float* value, const float* ref, int ref_batch, const float* off,
                         const float* logit, float* out, const int* level_hw, int num_levels, int B,
                         int S, int Q, int nH, int head_dim, int num_points, cudaStream_t stream) {
  if (!value || !ref || !off || !logit || !out || !level_hw) return GED_ERR_ARG;
  if (head_dim != MS_HD || num_points != MS_P || (ref_batch != 1 && ref_batch != B)) return GED_ERR_SHAPE;
  MsdaShapes sh;
  if (int e = fill_shapes(level_hw, num_levels, S, sh)) return e;
  dim3 grid(cdiv(Q, MS_WARPS), nH, B);
  msda_fwd_kernel<<<grid, MS_WARPS * 32, 0, stream>>>(value, ref, off, logit, out, sh, B, S, Q, nH, ref_batch == 1 ? 0 : Q * 2);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

// g_value must be zeroed by the caller (it may accumulate over several calls); g_ref may be NULL
// (constant reference points) and must be zeroed otherwise.
GED_API int ged_msda_bwd(const float* value, const float* ref, int ref_batch, const float* off,
                         const float* logit, const float* g_out, float* g_value, float* g_ref,
                         float* g_off, float* g_logit, const int* level_hw, int num_levels, int B, int S,
                         int Q, int nH, int head_dim, int num_points, cudaStream_t stream) {
  if (!value || !ref || !off || !logit || !g_out || !g_value || !g_off || !g_logit || !level_hw) return GED_ERR_ARG;
  if (head_dim != MS_HD || num_points != MS_P || (ref_batch != 1 && ref_batch != B)) return GED_ERR_SHAPE;
  if (g_ref && ref_batch != B) return GED_ERR_SHAPE;
  MsdaShapes sh;
  if (int e = fill_shapes(level_hw, num_levels, S, sh)) return e;
  dim3 grid(cdiv(Q, MS_WARPS), nH, B);
  msda_bwd_kernel<<<grid, MS_WARPS * 32, 0, stream>>>(value, ref, off, logit, g_out, g_value, g_ref, g_off, g_logit, sh, B, S, Q, nH, ref_batch == 1 ? 0 : Q * 2);
  GED_CHECK_LAUNCH();
  return GED_OK;
}
