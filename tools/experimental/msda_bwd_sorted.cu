// DRAFT - NOT part of libgedepth_sm100.so, NOT yet run on hardware (written at the end of round 1 after the GPU budget
// was spent; compile-checked only).  The design DESIGN.md §9 costs for the MSDA value-gradient scatter:
//
//   * the caller passes `order`, a permutation of the queries sorted by reference point (static per step, shared by
//     batch and heads; any permutation gives the same result, the sort only buys locality);
//   * a CTA takes T = 32 consecutive sorted queries of one (batch, head).  tools/msda_window_stats.py: their bilinear
//     corners fall into a ~10 x 10 box per level; a 12 x 12 window per level holds >= 98.3 % of the 4096 records;
//   * records are counting-sorted into the 4 x 144 window cells with integer shared-memory atomics only, every
//     non-empty cell is reduced in registers by one warp (lane = 2 of the 64 head channels, the tile's g_out rows sit
//     in 8 KB of shared memory) and leaves as ONE global red; the <= 1.7 % of records outside the windows take the
//     round-1 path (one red.v4 row each).  ~30x fewer global atomics than msda_bwd_kernel.
//
// Offset / attention-weight gradients stay with msda_bwd_kernel<MAP, false> (gather + dot, no atomics).
// To try it: add this file to gedepth_b200/csrc/, call ged_msda_bwd_scatter_sorted instead of the scatter half of
// ged_msda_bwd (variant bit 1), compare g_value with tests/test_ops_gpu.py::test_msda_fwd_bwd.
#include "../../gedepth_b200/csrc/common.cuh"

namespace ged {

constexpr int SL = 4, SP = 8, SHD = 64;
constexpr int ST = 32;                     // queries per CTA
constexpr int SWIN = 12;                   // window edge per level
constexpr int SCELLS = SL * SWIN * SWIN;   // 576
constexpr int SREC = ST * SL * SP * 4;     // 4096 corner records per tile
constexpr int STHREADS = 256;

struct SortedShapes {
  int h[SL], w[SL], start[SL];
};

struct SortedSmem {
  float go[ST][SHD];                 // g_out rows of the tile
  float lx[ST][SL * SP], ly[ST][SL * SP], a[ST][SL * SP];
  short x0[ST][SL * SP], y0[ST][SL * SP];
  unsigned short sorted[SREC];       // record ids grouped by cell
  unsigned short outside[SREC];      // record ids that missed their window
  int count[SCELLS], offset[SCELLS + 1], fill[SCELLS];
  int sum_x[SL], sum_y[SL], n_pts[SL], wx0[SL], wy0[SL];
  int n_outside;
  int qidx[ST];                      // original query index of each tile slot (-1 past the end)
};

// record id = (slot << 7) | (point << 2) | corner ; point = level * 8 + p
__device__ __forceinline__ float corner_weight(const SortedSmem& s, int slot, int pt, int corner) {
  const float lx = s.lx[slot][pt], ly = s.ly[slot][pt];
  const float wx = (corner & 1) ? lx : 1.f - lx, wy = (corner & 2) ? ly : 1.f - ly;
  return s.a[slot][pt] * wx * wy;
}

__global__ void __launch_bounds__(STHREADS) msda_bwd_scatter_sorted_kernel(
    const int* __restrict__ order, const float* __restrict__ ref, const float* __restrict__ off,
    const float* __restrict__ logit, const float* __restrict__ g_out, float* __restrict__ g_value, SortedShapes sh,
    int B, int S, int Q, int nH, int ref_bstride) {
  extern __shared__ unsigned char smem_raw[];
  SortedSmem& s = *reinterpret_cast<SortedSmem*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z, t0 = blockIdx.x * ST;
  const int rowpitch = nH * SHD;

  for (int i = threadIdx.x; i < SCELLS; i += STHREADS) { s.count[i] = 0; s.fill[i] = 0; }
  if (threadIdx.x < SL) { s.sum_x[threadIdx.x] = 0; s.sum_y[threadIdx.x] = 0; s.n_pts[threadIdx.x] = 0; }
  if (threadIdx.x == 0) s.n_outside = 0;
  __syncthreads();

  // ---- A: geometry of the tile's points (lane = point), g_out rows -------------------------------------------
  for (int slot = warp; slot < ST; slot += STHREADS / 32) {
    const int qs = t0 + slot;
    const int q = qs < Q ? __ldg(order + qs) : -1;
    if (lane == 0) s.qidx[slot] = q;
    if (q < 0) {
      s.a[slot][lane] = 0.f; s.lx[slot][lane] = 0.f; s.ly[slot][lane] = 0.f; s.x0[slot][lane] = -30000; s.y0[slot][lane] = -30000;
      s.go[slot][lane] = 0.f; s.go[slot][lane + 32] = 0.f;
      continue;
    }
    const int64_t bq = (int64_t)b * Q + q;
    const float rx = __ldg(ref + (int64_t)b * ref_bstride + q * 2), ry = __ldg(ref + (int64_t)b * ref_bstride + q * 2 + 1);
    const float lg = __ldg(logit + (bq * nH + h) * (SL * SP) + lane);
    const float mx = warp_max(lg);
    const float e = __expf(lg - mx);
    const float aw = e / warp_sum(e);
    const float2 o = __ldg((const float2*)(off + (bq * nH + h) * (SL * SP * 2)) + lane);
    const int l = lane >> 3;
    const float Wl = (float)sh.w[l], Hl = (float)sh.h[l];
    const float px = (rx + o.x / Wl) * Wl - 0.5f, py = (ry + o.y / Hl) * Hl - 0.5f;
    const float xf = floorf(px), yf = floorf(py);
    const int x0 = max(-30000, min(30000, (int)xf)), y0 = max(-30000, min(30000, (int)yf));
    s.lx[slot][lane] = px - xf; s.ly[slot][lane] = py - yf; s.a[slot][lane] = aw;
    s.x0[slot][lane] = (short)x0; s.y0[slot][lane] = (short)y0;
    if (x0 >= -1 && x0 < sh.w[l] && y0 >= -1 && y0 < sh.h[l]) {      // at least one corner can be inside the map
      atomicAdd(&s.sum_x[l], x0); atomicAdd(&s.sum_y[l], y0); atomicAdd(&s.n_pts[l], 1);
    }
    const float2 g2 = __ldg((const float2*)(g_out + bq * rowpitch + h * SHD) + lane);
    s.go[slot][2 * lane] = g2.x; s.go[slot][2 * lane + 1] = g2.y;
  }
  __syncthreads();
  // ---- B: window origin per level = mean corner position - WIN/2 ---------------------------------------------
  if (threadIdx.x < SL) {
    const int l = threadIdx.x, n = max(s.n_pts[l], 1);
    const int cx = (int)floorf((float)s.sum_x[l] / (float)n + 0.5f), cy = (int)floorf((float)s.sum_y[l] / (float)n + 0.5f);
    s.wx0[l] = cx - SWIN / 2 + 1; s.wy0[l] = cy - SWIN / 2 + 1;
  }
  __syncthreads();
  // ---- C1: count records per cell; list the ones outside their window ----------------------------------------
  auto cell_of = [&](int rec, bool& valid) -> int {
    const int slot = rec >> 7, pt = (rec >> 2) & 31, corner = rec & 3, l = pt >> 3;
    const int x = s.x0[slot][pt] + (corner & 1), y = s.y0[slot][pt] + (corner >> 1);
    valid = s.qidx[slot] >= 0 && x >= 0 && x < sh.w[l] && y >= 0 && y < sh.h[l];
    const int cx = x - s.wx0[l], cy = y - s.wy0[l];
    if (!valid || cx < 0 || cx >= SWIN || cy < 0 || cy >= SWIN) return -1;
    return (l * SWIN + cy) * SWIN + cx;
  };
  for (int rec = threadIdx.x; rec < SREC; rec += STHREADS) {
    bool valid;
    const int c = cell_of(rec, valid);
    if (c >= 0) atomicAdd(&s.count[c], 1);
    else if (valid) s.outside[atomicAdd(&s.n_outside, 1)] = (unsigned short)rec;
  }
  __syncthreads();
  // ---- C2: exclusive prefix sum over the 576 cells (one warp, 18 cells per lane) ------------------------------
  if (warp == 0) {
    constexpr int PER = SCELLS / 32;
    int local[PER], tot = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) { local[i] = tot; tot += s.count[lane * PER + i]; }
    int incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    const int base = incl - tot;
#pragma unroll
    for (int i = 0; i < PER; ++i) s.offset[lane * PER + i] = base + local[i];
    if (lane == 31) s.offset[SCELLS] = incl;
  }
  __syncthreads();
  // ---- C3: scatter record ids into their cells ----------------------------------------------------------------
  for (int rec = threadIdx.x; rec < SREC; rec += STHREADS) {
    bool valid;
    const int c = cell_of(rec, valid);
    if (c >= 0) s.sorted[s.offset[c] + atomicAdd(&s.fill[c], 1)] = (unsigned short)rec;
  }
  __syncthreads();
  // ---- D: one warp per non-empty cell: register reduction over its records, one global red --------------------
  float* gvb = g_value + (int64_t)b * S * rowpitch + h * SHD + 2 * lane;
  for (int c = warp; c < SCELLS; c += STHREADS / 32) {
    const int beg = s.offset[c], end = s.offset[c + 1];
    if (beg == end) continue;
    float2 acc = make_float2(0.f, 0.f);
    for (int i = beg; i < end; ++i) {
      const int rec = s.sorted[i], slot = rec >> 7;
      const float wgt = corner_weight(s, slot, (rec >> 2) & 31, rec & 3);
      acc.x += wgt * s.go[slot][2 * lane]; acc.y += wgt * s.go[slot][2 * lane + 1];
    }
    const int l = c / (SWIN * SWIN), r = c - l * SWIN * SWIN, cy = r / SWIN, cx = r - cy * SWIN;
    const int64_t pos = sh.start[l] + (int64_t)(s.wy0[l] + cy) * sh.w[l] + (s.wx0[l] + cx);
    atomicAdd((float2*)(gvb + pos * rowpitch), acc);
  }
  // ---- E: the few records outside their window: one 256-byte red each ------------------------------------------
  const int n_out = s.n_outside;
  for (int i = warp; i < n_out; i += STHREADS / 32) {
    const int rec = s.outside[i], slot = rec >> 7, pt = (rec >> 2) & 31, corner = rec & 3, l = pt >> 3;
    const float wgt = corner_weight(s, slot, pt, corner);
    const int x = s.x0[slot][pt] + (corner & 1), y = s.y0[slot][pt] + (corner >> 1);
    const int64_t pos = sh.start[l] + (int64_t)y * sh.w[l] + x;
    atomicAdd((float2*)(gvb + pos * rowpitch), make_float2(wgt * s.go[slot][2 * lane], wgt * s.go[slot][2 * lane + 1]));
  }
}

}  // namespace ged
using namespace ged;

// order (Q) int32: permutation of the queries (sorted by reference point for locality; identity is valid too).
// g_value accumulated.  Same tensor layouts as ged_msda_bwd.
GED_API int ged_msda_bwd_scatter_sorted(const int* order, const float* ref, int ref_batch, const float* off,
                                        const float* logit, const float* g_out, float* g_value, const int* level_hw,
                                        int num_levels, int B, int S, int Q, int nH, int head_dim, int num_points,
                                        cudaStream_t stream) {
  if (!order || !ref || !off || !logit || !g_out || !g_value || !level_hw) return GED_ERR_ARG;
  if (num_levels != SL || head_dim != SHD || num_points != SP || (ref_batch != 1 && ref_batch != B)) return GED_ERR_SHAPE;
  SortedShapes sh;
  int start = 0;
  for (int l = 0; l < SL; ++l) {
    sh.h[l] = level_hw[2 * l]; sh.w[l] = level_hw[2 * l + 1]; sh.start[l] = start;
    if (sh.h[l] <= 0 || sh.w[l] <= 0 || sh.h[l] > 20000 || sh.w[l] > 20000) return GED_ERR_SHAPE;
    start += sh.h[l] * sh.w[l];
  }
  if (start != S) return GED_ERR_SHAPE;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(msda_bwd_scatter_sorted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SortedSmem)) != cudaSuccess) return GED_ERR_LAUNCH;
    attr_set = true;
  }
  dim3 grid(cdiv(Q, ST), nH, B);
  msda_bwd_scatter_sorted_kernel<<<grid, STHREADS, sizeof(SortedSmem), stream>>>(order, ref, off, logit, g_out, g_value, sh, B, S, Q, nH, ref_batch == 1 ? 0 : Q * 2);
  GED_CHECK_LAUNCH();
  return GED_OK;
}
