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

// Per-point geometry, computed ONCE by the lane that owns the point and shared through smem:
// four corner element offsets (0 with the valid bit cleared when outside), fractional parts,
// softmax weight and the valid mask.
struct __align__(16) PointRec {
  int o00, o01, o10, o11;
  float lx, ly, a;
  int valid;            // bit0..3 = corner 00,01,10,11 inside the map
};

__device__ __forceinline__ PointRec make_point(float x, float y, float a, int H, int W, int start, int rowpitch) {
  PointRec g;
  const float xf = floorf(x), yf = floorf(y);
  const int x0 = (int)xf, y0 = (int)yf;
  g.lx = x - xf; g.ly = y - yf; g.a = a;
  const bool vx0 = x0 >= 0 && x0 < W, vx1 = x0 + 1 >= 0 && x0 + 1 < W;
  const bool vy0 = y0 >= 0 && y0 < H, vy1 = y0 + 1 >= 0 && y0 + 1 < H;
  const int base = start + y0 * W + x0;
  g.valid = (vy0 && vx0 ? 1 : 0) | (vy0 && vx1 ? 2 : 0) | (vy1 && vx0 ? 4 : 0) | (vy1 && vx1 ? 8 : 0);
  g.o00 = (g.valid & 1) ? base * rowpitch : 0;
  g.o01 = (g.valid & 2) ? (base + 1) * rowpitch : 0;
  g.o10 = (g.valid & 4) ? (base + W) * rowpitch : 0;
  g.o11 = (g.valid & 8) ? (base + W + 1) * rowpitch : 0;
  return g;
}

// lane = (level, point): softmax over the warp's 32 logits, pixel coordinates, record into smem
__device__ __forceinline__ float lane_point(const float* __restrict__ off, const float* __restrict__ logit,
                                            float rx, float ry, const MsdaShapes& sh, int lane, int rowpitch,
                                            PointRec* rec) {
  // the offsets / logits are a 4.8 GB stream per step that is touched once: kept out of L1, which the value rows need
  // (measured: within noise, 28.9 vs 28.3-30.2 ms per step - the gather is bound by L2 -> SM bandwidth, not by L1 capacity)
  float lg;
  float2 o;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(lg) : "l"(logit + lane));
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(o.x), "=f"(o.y) : "l"((const float2*)off + lane));
  const float mx = warp_max(lg);
  const float e = __expf(lg - mx);
  const float aw = e / warp_sum(e);
  const int l = lane >> 3;
  const float Wl = (float)sh.w[l], Hl = (float)sh.h[l];
  const float px = (rx + o.x / Wl) * Wl - 0.5f, py = (ry + o.y / Hl) * Hl - 0.5f;
  rec[lane] = make_point(px, py, aw, sh.h[l], sh.w[l], sh.start[l], rowpitch);
  return aw;
}

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg((const float4*)p); }

// Work mapping of the 8 warps of a CTA.
//   MAP 0: 8 consecutive queries of ONE head (grid Q/8 x nH x B): neighbouring queries re-use corner rows from L1,
//          but every CTA in flight hits the same head's value / g_value map.
//   MAP 1: ONE query, 8 heads (grid Q x 1 x B; nH == 8): a CTA reads its offsets/logits as one 3 KB block and the
//          gathers / atomics of concurrent CTAs are spread over all heads' rows (8x fewer same-address collisions
//          in the L2 atomic units on the coarse levels).
template <int MAP>
__device__ __forceinline__ void map_work(int warp, int& q, int& h) {
  if (MAP == 0) { q = blockIdx.x * MS_WARPS + warp; h = blockIdx.y; }
  else { q = blockIdx.x; h = warp; }
}

// One warp per (batch, query, head).  Half-warps take alternate points; a lane owns 4 of the 64 channels.
template <int MAP>
__global__ void __launch_bounds__(MS_WARPS * 32) msda_fwd_kernel(
    const float* __restrict__ value, const float* __restrict__ ref, const float* __restrict__ off,
    const float* __restrict__ logit, float* __restrict__ out, MsdaShapes sh, int B, int S, int Q, int nH,
    int ref_bstride) {
  __shared__ PointRec s_rec[MS_WARPS][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int q, h;
  map_work<MAP>(warp, q, h);
  const int b = blockIdx.z;
  if (q >= Q) return;
  const int64_t bq = (int64_t)b * Q + q;
  const float rx = __ldg(ref + (int64_t)b * ref_bstride + q * 2), ry = __ldg(ref + (int64_t)b * ref_bstride + q * 2 + 1);
  const int rowpitch = nH * MS_HD;
  lane_point(off + (bq * nH + h) * (MS_L * MS_P * 2), logit + (bq * nH + h) * (MS_L * MS_P), rx, ry, sh, lane, rowpitch, s_rec[warp]);
  __syncwarp();
  const int half = lane >> 4, cl = (lane & 15) * 4;
  const float* vb = value + (int64_t)b * S * rowpitch + h * MS_HD + cl;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int j = 0; j < MS_L * MS_P / 2; ++j) {
    const PointRec g = s_rec[warp][2 * j + half];
    const float w00 = g.a * (1.f - g.ly) * (1.f - g.lx) * (float)(g.valid & 1);
    const float w01 = g.a * (1.f - g.ly) * g.lx * (float)((g.valid >> 1) & 1);
    const float w10 = g.a * g.ly * (1.f - g.lx) * (float)((g.valid >> 2) & 1);
    const float w11 = g.a * g.ly * g.lx * (float)((g.valid >> 3) & 1);
    const float4 v00 = ld4(vb + g.o00), v01 = ld4(vb + g.o01), v10 = ld4(vb + g.o10), v11 = ld4(vb + g.o11);
    acc.x += w00 * v00.x + w01 * v01.x + w10 * v10.x + w11 * v11.x;
    acc.y += w00 * v00.y + w01 * v01.y + w10 * v10.y + w11 * v11.y;
    acc.z += w00 * v00.z + w01 * v01.z + w10 * v10.z + w11 * v11.z;
    acc.w += w00 * v00.w + w01 * v01.w + w10 * v10.w + w11 * v11.w;
  }
  acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16);
  acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
  acc.z += __shfl_xor_sync(0xffffffffu, acc.z, 16);
  acc.w += __shfl_xor_sync(0xffffffffu, acc.w, 16);
  if (half == 0) stg_stream((float4*)(out + bq * rowpitch + h * MS_HD + cl), acc);
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

// g_value scatter alone: no value loads, no per-point partial sums - a small register footprint, so many more
// warps keep atomics in flight than in the fused kernel.
template <int MAP>
__global__ void __launch_bounds__(MS_WARPS * 32) msda_bwd_scatter_kernel(
    const float* __restrict__ ref, const float* __restrict__ off, const float* __restrict__ logit,
    const float* __restrict__ g_out, float* __restrict__ g_value, MsdaShapes sh, int B, int S, int Q, int nH,
    int ref_bstride) {
  __shared__ PointRec s_rec[MS_WARPS][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int q, h;
  map_work<MAP>(warp, q, h);
  const int b = blockIdx.z;
  if (q >= Q) return;
  const int64_t bq = (int64_t)b * Q + q;
  const float rx = __ldg(ref + (int64_t)b * ref_bstride + q * 2), ry = __ldg(ref + (int64_t)b * ref_bstride + q * 2 + 1);
  const int rowpitch = nH * MS_HD;
  lane_point(off + (bq * nH + h) * (MS_L * MS_P * 2), logit + (bq * nH + h) * (MS_L * MS_P), rx, ry, sh, lane, rowpitch, s_rec[warp]);
  __syncwarp();
  const int half = lane >> 4, cl = (lane & 15) * 4;
  float* gvb = g_value + (int64_t)b * S * rowpitch + h * MS_HD + cl;
  const float4 go = ld4(g_out + bq * rowpitch + h * MS_HD + cl);
#pragma unroll 4
  for (int j = 0; j < MS_L * MS_P / 2; ++j) {
    const PointRec g = s_rec[warp][2 * j + half];
    const float a0 = g.a * (1.f - g.ly), a1 = g.a * g.ly;
    if (g.valid & 1) { const float w = a0 * (1.f - g.lx); atomicAdd((float4*)(gvb + g.o00), make_float4(go.x * w, go.y * w, go.z * w, go.w * w)); }
    if (g.valid & 2) { const float w = a0 * g.lx; atomicAdd((float4*)(gvb + g.o01), make_float4(go.x * w, go.y * w, go.z * w, go.w * w)); }
    if (g.valid & 4) { const float w = a1 * (1.f - g.lx); atomicAdd((float4*)(gvb + g.o10), make_float4(go.x * w, go.y * w, go.z * w, go.w * w)); }
    if (g.valid & 8) { const float w = a1 * g.lx; atomicAdd((float4*)(gvb + g.o11), make_float4(go.x * w, go.y * w, go.z * w, go.w * w)); }
  }
}

// SCATTER = true: fused (g_value atomics + offset / weight gradients); false: offset / weight gradients only
template <int MAP, bool SCATTER>
__global__ void __launch_bounds__(MS_WARPS * 32) msda_bwd_kernel(
    const float* __restrict__ value, const float* __restrict__ ref, const float* __restrict__ off,
    const float* __restrict__ logit, const float* __restrict__ g_out, float* __restrict__ g_value,
    float* __restrict__ g_ref, float* __restrict__ g_off, float* __restrict__ g_logit, MsdaShapes sh,
    int B, int S, int Q, int nH, int ref_bstride) {
  __shared__ PointRec s_rec[MS_WARPS][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int q, h;
  map_work<MAP>(warp, q, h);
  const int b = blockIdx.z;
  if (q >= Q) return;
  const int64_t bq = (int64_t)b * Q + q;
  const float rx = __ldg(ref + (int64_t)b * ref_bstride + q * 2), ry = __ldg(ref + (int64_t)b * ref_bstride + q * 2 + 1);
  const int rowpitch = nH * MS_HD;
  const float aw = lane_point(off + (bq * nH + h) * (MS_L * MS_P * 2), logit + (bq * nH + h) * (MS_L * MS_P), rx, ry, sh, lane, rowpitch, s_rec[warp]);
  __syncwarp();
  const int half = lane >> 4, cl = (lane & 15) * 4;
  const int64_t voff = (int64_t)b * S * rowpitch + h * MS_HD + cl;
  const float* vb = value + voff;
  float* gvb = g_value + voff;
  const float4 go = ld4(g_out + bq * rowpitch + h * MS_HD + cl);
  float part[48];   // [j][gw, gx, gy] partial sums over this lane's four channels, point 2j+half
#pragma unroll
  for (int j = 0; j < MS_L * MS_P / 2; ++j) {
    const PointRec g = s_rec[warp][2 * j + half];
    const float u00 = (1.f - g.ly) * (1.f - g.lx), u01 = (1.f - g.ly) * g.lx, u10 = g.ly * (1.f - g.lx), u11 = g.ly * g.lx;
    float4 v00 = make_float4(0.f, 0.f, 0.f, 0.f), v01 = v00, v10 = v00, v11 = v00;
    if (g.valid & 1) { v00 = ld4(vb + g.o00); if (SCATTER) { const float w = g.a * u00; atomicAdd((float4*)(gvb + g.o00), make_float4(go.x * w, go.y * w, go.z * w, go.w * w)); } }
    if (g.valid & 2) { v01 = ld4(vb + g.o01); if (SCATTER) { const float w = g.a * u01; atomicAdd((float4*)(gvb + g.o01), make_float4(go.x * w, go.y * w, go.z * w, go.w * w)); } }
    if (g.valid & 4) { v10 = ld4(vb + g.o10); if (SCATTER) { const float w = g.a * u10; atomicAdd((float4*)(gvb + g.o10), make_float4(go.x * w, go.y * w, go.z * w, go.w * w)); } }
    if (g.valid & 8) { v11 = ld4(vb + g.o11); if (SCATTER) { const float w = g.a * u11; atomicAdd((float4*)(gvb + g.o11), make_float4(go.x * w, go.y * w, go.z * w, go.w * w)); } }
    // <g_out, corner> over this lane's channels
    const float d00 = go.x * v00.x + go.y * v00.y + go.z * v00.z + go.w * v00.w;
    const float d01 = go.x * v01.x + go.y * v01.y + go.z * v01.z + go.w * v01.w;
    const float d10 = go.x * v10.x + go.y * v10.y + go.z * v10.z + go.w * v10.w;
    const float d11 = go.x * v11.x + go.y * v11.y + go.z * v11.z + go.w * v11.w;
    part[j * 3 + 0] = u00 * d00 + u01 * d01 + u10 * d10 + u11 * d11;                          // d/d a
    part[j * 3 + 1] = g.a * ((1.f - g.ly) * (d01 - d00) + g.ly * (d11 - d10));               // d/d x_pix
    part[j * 3 + 2] = g.a * ((1.f - g.lx) * (d10 - d00) + g.lx * (d11 - d01));               // d/d y_pix
  }
  float p24[24], p12[12], p6[6], p3[3];
  halve<48>(part, p24, 8, (lane & 8) != 0);
  halve<24>(p24, p12, 4, (lane & 4) != 0);
  halve<12>(p12, p6, 2, (lane & 2) != 0);
  halve<6>(p6, p3, 1, (lane & 1) != 0);
  // lane (half, i) now holds (gw, g_xpix, g_ypix) of local point j = i, i.e. global point 2*i + half;
  // move them to the lane that owns that point (lane index == point index)
  const int src = ((lane & 1) << 4) | (lane >> 1);
  const float gw = __shfl_sync(0xffffffffu, p3[0], src);
  const float gx = __shfl_sync(0xffffffffu, p3[1], src);
  const float gy = __shfl_sync(0xffffffffu, p3[2], src);
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

// Roofline probe for the backward: nothing but the scatter pattern of msda_bwd - every half-warp adds one 256-byte
// row (16 lanes x red.global.add.v4.f32) at a pseudo-random position of a value-map-sized buffer, full occupancy, no
// gathers, no arithmetic.  Its payload rate is the L2 atomic throughput the backward can at best reach.
__global__ void __launch_bounds__(256) atomic_probe_kernel(float* __restrict__ buf, int rows, int rowpitch, int heads,
                                                            int iters) {
  const int lane = threadIdx.x & 31, half = lane >> 4, cl = (lane & 15) * 4;
  const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int h = wid % heads;
  uint32_t state = wid * 2654435761u + 12345u;
  const float4 v = make_float4(1e-9f, 1e-9f, 1e-9f, 1e-9f);
  for (int i = 0; i < iters; ++i) {
    state = state * 1664525u + 1013904223u;
    const uint32_t r = ((state >> 8) + (uint32_t)half * 7919u) % (uint32_t)rows;
    atomicAdd((float4*)(buf + (int64_t)r * rowpitch + h * MS_HD + cl), v);
  }
}

}  // namespace ged
using namespace ged;

// Measures nothing itself: enqueues `iters` row-atomics per half-warp (2 * 256 bytes per warp and iteration) on a
// (rows, heads*64) buffer; bench.py times it with CUDA events.  Returns the number of warps launched.
GED_API int ged_msda_atomic_probe(float* buf, int rows, int heads, int iters, cudaStream_t stream) {
  if (!buf || rows <= 0 || heads <= 0 || iters <= 0) return GED_ERR_ARG;
  const int blocks = 148 * 8;
  atomic_probe_kernel<<<blocks, 256, 0, stream>>>(buf, rows, heads * MS_HD, heads, iters);
  if (cudaGetLastError() != cudaSuccess) return GED_ERR_LAUNCH;
  return blocks * 8;
}

// bit0: MAP (1 = one query x 8 heads per CTA), bit1: split backward (scatter + gather kernels), -1 = auto: MAP 1 for
// cross-attention-sized query sets (Q >= 2 S; measured 4-5 % faster there, slightly slower for Q = S)
static int g_msda_variant = -1;

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
  const int rbs = ref_batch == 1 ? 0 : Q * 2;
  const int variant = g_msda_variant < 0 ? (Q >= 2 * S ? 1 : 0) : g_msda_variant;
  if ((variant & 1) && nH == MS_WARPS) {
    msda_fwd_kernel<1><<<dim3(Q, 1, B), MS_WARPS * 32, 0, stream>>>(value, ref, off, logit, out, sh, B, S, Q, nH, rbs);
  } else {
    msda_fwd_kernel<0><<<dim3(cdiv(Q, MS_WARPS), nH, B), MS_WARPS * 32, 0, stream>>>(value, ref, off, logit, out, sh, B, S, Q, nH, rbs);
  }
  GED_CHECK_LAUNCH();
  return GED_OK;
}

// bit0: CTA = one query x 8 heads instead of 8 queries x one head; bit1: backward as two kernels (g_value scatter,
// then offset / weight gradients).  Returns the previous value.
GED_API int ged_set_msda_variant(int v) {
  const int prev = g_msda_variant;
  if (v >= -1 && v <= 3) g_msda_variant = v;
  return prev;
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
  const int rbs = ref_batch == 1 ? 0 : Q * 2;
  const int variant = g_msda_variant < 0 ? (Q >= 2 * S ? 1 : 0) : g_msda_variant;
  const bool map1 = (variant & 1) && nH == MS_WARPS, split = (variant & 2) != 0;
  const dim3 grid = map1 ? dim3(Q, 1, B) : dim3(cdiv(Q, MS_WARPS), nH, B);
  const int T = MS_WARPS * 32;
  if (split) {
    if (map1) {
      msda_bwd_scatter_kernel<1><<<grid, T, 0, stream>>>(ref, off, logit, g_out, g_value, sh, B, S, Q, nH, rbs);
      msda_bwd_kernel<1, false><<<grid, T, 0, stream>>>(value, ref, off, logit, g_out, g_value, g_ref, g_off, g_logit, sh, B, S, Q, nH, rbs);
    } else {
      msda_bwd_scatter_kernel<0><<<grid, T, 0, stream>>>(ref, off, logit, g_out, g_value, sh, B, S, Q, nH, rbs);
      msda_bwd_kernel<0, false><<<grid, T, 0, stream>>>(value, ref, off, logit, g_out, g_value, g_ref, g_off, g_logit, sh, B, S, Q, nH, rbs);
    }
  } else if (map1) {
    msda_bwd_kernel<1, true><<<grid, T, 0, stream>>>(value, ref, off, logit, g_out, g_value, g_ref, g_off, g_logit, sh, B, S, Q, nH, rbs);
  } else {
    msda_bwd_kernel<0, true><<<grid, T, 0, stream>>>(value, ref, off, logit, g_out, g_value, g_ref, g_off, g_logit, sh, B, S, Q, nH, rbs);
  }
  GED_CHECK_LAUNCH();
  return GED_OK;
}
