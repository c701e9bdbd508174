// Locality-aware multi-scale deformable attention, sm_100a (a11: depth/models/necks/hahi.py:280-289,316-325 call
// mmcv.ops.MultiScaleDeformableAttention [external, mmcv-full 1.3.13]; same semantics as msda.cu).
//
// The round-1 kernels (msda.cu) fetch / atomically add one 256-byte row per bilinear corner straight from / to L2:
// 207 GB of gathers and 207 GB of red.global.add.v4.f32 per step at B = 8.  Here the queries are first SORTED by
// their reference point (ged_msda_sort_queries: band / x-bucket counting sort; any permutation gives the same
// numbers, the sort only buys locality).  A CTA owns a TILE of 32 consecutive sorted queries of one (batch, head);
// their 4096 corners fall into a ~10 x 10 box per level, so
//   * forward and the offset / weight gradients stage a 12 x 12 window of value rows per level in shared memory
//     (36 KB, coalesced) and gather from there - L2 reads drop ~7x;
//   * the value gradient counting-sorts the corner records into the 4 x 144 window cells with integer shared-memory
//     atomics, reduces every cell in registers and leaves ONE red.global.add.v4.f32 row per non-empty cell - ~13x
//     fewer L2 atomics.
// Corners outside their window (~1 %) take the direct path.  fp32 throughout.
#include "common.cuh"

namespace ged {

constexpr int TL = 4, TP = 8, THD = 64;
constexpr int TWARPS = 8, TTHREADS = TWARPS * 32;
constexpr int TQ = 32;                 // queries per tile
constexpr int TQW = TQ / TWARPS;       // query slots per warp
constexpr int TWIN = 12, TCELLS = TWIN * TWIN;

struct TileShapes {
  int h[TL], w[TL], start[TL];
};
// kernel-parameter arrays indexed by a run-time level: selects instead of a local-memory copy of the struct
__device__ __forceinline__ int pick(const int (&a)[TL], int l) { return l == 0 ? a[0] : (l == 1 ? a[1] : (l == 2 ? a[2] : a[3])); }

// lane = (level, point) of the warp's TQW query slots
struct TileGeom {
  float px[TQW], py[TQW], aw[TQW];
  int q[TQW];                          // original query index (warp-uniform), -1 past the end
};

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void fma4(float4& acc, float w, const float4& v) {
  acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// Softmax over the 32 logits, pixel coordinates of the 32 points (same arithmetic as msda.cu: loc = ref + off / (W,H),
// pixel = loc * (W,H) - 0.5), and the per-level sums the window origin is derived from.
__device__ __forceinline__ void tile_geometry(TileGeom& g, const int* __restrict__ order, const float* __restrict__ ref,
                                              const float* __restrict__ off, const float* __restrict__ logit,
                                              const TileShapes& sh, int b, int h, int Q, int nH, int ref_bstride, int t0,
                                              int warp, int lane, int* s_sum) {
  const int l = lane >> 3;
  const int Wi = pick(sh.w, l), Hi = pick(sh.h, l);
  const float Wl = (float)Wi, Hl = (float)Hi;
  int sx = 0, sy = 0, sn = 0;
#pragma unroll
  for (int i = 0; i < TQW; ++i) {
    const int qs = t0 + warp * TQW + i;
    const int q = qs < Q ? __ldg(order + qs) : -1;
    g.q[i] = q;
    if (q < 0) { g.px[i] = -30000.f; g.py[i] = -30000.f; g.aw[i] = 0.f; continue; }
    const int64_t bq = (int64_t)b * Q + q;
    const float rx = __ldg(ref + (int64_t)b * ref_bstride + q * 2), ry = __ldg(ref + (int64_t)b * ref_bstride + q * 2 + 1);
    const float lg = __ldg(logit + (bq * nH + h) * (TL * TP) + lane);
    const float mx = warp_max(lg);
    const float e = __expf(lg - mx);
    g.aw[i] = e / warp_sum(e);
    const float2 o = __ldg((const float2*)(off + (bq * nH + h) * (TL * TP * 2)) + lane);
    // NaN / huge coordinates: clamped far outside every map (no corner is valid there)
    const float px = fminf(fmaxf((rx + o.x / Wl) * Wl - 0.5f, -30000.f), 30000.f);
    const float py = fminf(fmaxf((ry + o.y / Hl) * Hl - 0.5f, -30000.f), 30000.f);
    g.px[i] = px; g.py[i] = py;
    const int x0 = (int)floorf(px), y0 = (int)floorf(py);
    if (x0 >= -1 && x0 < Wi && y0 >= -1 && y0 < Hi) { sx += x0; sy += y0; sn += 1; }
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); sn += __shfl_xor_sync(0xffffffffu, sn, o);
  }
  if ((lane & 7) == 0 && sn > 0) { atomicAdd(&s_sum[l * 3], sx); atomicAdd(&s_sum[l * 3 + 1], sy); atomicAdd(&s_sum[l * 3 + 2], sn); }
}

// Window origin per level: centred on the mean corner block of the tile, clamped into the map.
__device__ __forceinline__ void tile_origin(const int* s_sum, int* s_org, const TileShapes& sh, int l) {
  const int n = max(s_sum[l * 3 + 2], 1);
  const int mx = (int)floorf((float)s_sum[l * 3] / (float)n + 0.5f), my = (int)floorf((float)s_sum[l * 3 + 1] / (float)n + 0.5f);
  s_org[l * 2] = max(0, min(mx - (TWIN / 2 - 1), pick(sh.w, l) - TWIN));
  s_org[l * 2 + 1] = max(0, min(my - (TWIN / 2 - 1), pick(sh.h, l) - TWIN));
}

// Stage the TWIN x TWIN window of value rows of one (batch, head, level); cells outside the map are zero.
__device__ __forceinline__ void stage_window(float* s_win, const float* __restrict__ vb, const TileShapes& sh, int l, int wx0,
                                             int wy0, int rowpitch, int tid) {
  const int W = pick(sh.w, l), H = pick(sh.h, l), start = pick(sh.start, l);
  for (int idx = tid; idx < TCELLS * (THD / 4); idx += TTHREADS) {
    const int cell = idx >> 4, part = idx & 15;
    const int cy = cell / TWIN, cx = cell - cy * TWIN;
    const int x = wx0 + cx, y = wy0 + cy;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (x < W && y < H) v = ldg4(vb + (int64_t)(start + y * W + x) * rowpitch + part * 4);
    *reinterpret_cast<float4*>(s_win + cell * THD + part * 4) = v;
  }
}

// One corner on the general path: zero outside the map, shared-memory window when inside it, L2 otherwise.
__device__ __forceinline__ float4 corner_value(const float* s_win, const float* __restrict__ vb, int x, int y, int W, int H,
                                               int start, int wx0, int wy0, int rowpitch, int cl) {
  if (x < 0 || x >= W || y < 0 || y >= H) return make_float4(0.f, 0.f, 0.f, 0.f);
  const int cx = x - wx0, cy = y - wy0;
  if ((unsigned)cx < (unsigned)TWIN && (unsigned)cy < (unsigned)TWIN) return lds4(s_win + (cy * TWIN + cx) * THD + cl);
  return ldg4(vb + (int64_t)(start + y * W + x) * rowpitch + cl);
}

struct Corners {
  float4 v00, v01, v10, v11;
};

__device__ __forceinline__ Corners load_corners(const float* s_win, const float* __restrict__ vb, int x0, int y0, int W, int H,
                                                int start, int wx0, int wy0, int rowpitch, int cl) {
  Corners c;
  const int cx = x0 - wx0, cy = y0 - wy0;
  if (cx >= 0 && cx < TWIN - 1 && cy >= 0 && cy < TWIN - 1) {       // whole 2 x 2 block inside the window
    const float* p = s_win + (cy * TWIN + cx) * THD + cl;
    c.v00 = lds4(p); c.v01 = lds4(p + THD); c.v10 = lds4(p + TWIN * THD); c.v11 = lds4(p + TWIN * THD + THD);
  } else {
    c.v00 = corner_value(s_win, vb, x0, y0, W, H, start, wx0, wy0, rowpitch, cl);
    c.v01 = corner_value(s_win, vb, x0 + 1, y0, W, H, start, wx0, wy0, rowpitch, cl);
    c.v10 = corner_value(s_win, vb, x0, y0 + 1, W, H, start, wx0, wy0, rowpitch, cl);
    c.v11 = corner_value(s_win, vb, x0 + 1, y0 + 1, W, H, start, wx0, wy0, rowpitch, cl);
  }
  return c;
}

// ---------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TTHREADS, 4) msda_tile_fwd_kernel(
    const float* __restrict__ value, const float* __restrict__ ref, const float* __restrict__ off,
    const float* __restrict__ logit, const int* __restrict__ order, float* __restrict__ out, TileShapes sh, int B, int S,
    int Q, int nH, int ref_bstride) {
  __shared__ __align__(16) float s_win[TCELLS * THD];
  __shared__ int s_sum[TL * 3], s_org[TL * 2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t0 = blockIdx.x * TQ, h = blockIdx.y, b = blockIdx.z;
  if (tid < TL * 3) s_sum[tid] = 0;
  __syncthreads();
  TileGeom g;
  tile_geometry(g, order, ref, off, logit, sh, b, h, Q, nH, ref_bstride, t0, warp, lane, s_sum);
  __syncthreads();
  if (tid < TL) tile_origin(s_sum, s_org, sh, tid);
  __syncthreads();
  const int half = lane >> 4, cl = (lane & 15) * 4;
  const int rowpitch = nH * THD;
  const float* vb = value + (int64_t)b * S * rowpitch + h * THD;
  float4 acc[TQW];
#pragma unroll
  for (int i = 0; i < TQW; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int l = 0; l < TL; ++l) {
    const int W = pick(sh.w, l), H = pick(sh.h, l), start = pick(sh.start, l), wx0 = s_org[l * 2], wy0 = s_org[l * 2 + 1];
    stage_window(s_win, vb, sh, l, wx0, wy0, rowpitch, tid);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < TQW; ++i) {
#pragma unroll
      for (int j = 0; j < TP / 2; ++j) {
        const int src = l * TP + 2 * j + half;
        const float x = __shfl_sync(0xffffffffu, g.px[i], src), y = __shfl_sync(0xffffffffu, g.py[i], src);
        const float a = __shfl_sync(0xffffffffu, g.aw[i], src);
        const float xf = floorf(x), yf = floorf(y);
        const float lx = x - xf, ly = y - yf;
        const Corners c = load_corners(s_win, vb, (int)xf, (int)yf, W, H, start, wx0, wy0, rowpitch, cl);
        const float a0 = a * (1.f - ly), a1 = a * ly;
        fma4(acc[i], a0 * (1.f - lx), c.v00); fma4(acc[i], a0 * lx, c.v01);
        fma4(acc[i], a1 * (1.f - lx), c.v10); fma4(acc[i], a1 * lx, c.v11);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TQW; ++i) {
    acc[i].x += __shfl_xor_sync(0xffffffffu, acc[i].x, 16); acc[i].y += __shfl_xor_sync(0xffffffffu, acc[i].y, 16);
    acc[i].z += __shfl_xor_sync(0xffffffffu, acc[i].z, 16); acc[i].w += __shfl_xor_sync(0xffffffffu, acc[i].w, 16);
    if (half == 0 && g.q[i] >= 0) *reinterpret_cast<float4*>(out + ((int64_t)b * Q + g.q[i]) * rowpitch + h * THD + cl) = acc[i];
  }
}

// butterfly reduce-scatter: every lane enters with N partial sums, leaves with N/2
template <int N>
__device__ __forceinline__ void halve_t(const float* in, float* outv, int mask, bool upper) {
#pragma unroll
  for (int i = 0; i < N / 2; ++i) {
    const float send = upper ? in[i] : in[i + N / 2];
    const float keep = upper ? in[i + N / 2] : in[i];
    outv[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward, part 1: offset / attention-weight / reference-point gradients (gather + dot, no atomics on g_value)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TTHREADS, 2) msda_tile_bwd_dot_kernel(
    const float* __restrict__ value, const float* __restrict__ ref, const float* __restrict__ off,
    const float* __restrict__ logit, const int* __restrict__ order, const float* __restrict__ g_out,
    float* __restrict__ g_ref, float* __restrict__ g_off, float* __restrict__ g_logit, TileShapes sh, int B, int S, int Q,
    int nH, int ref_bstride) {
  __shared__ __align__(16) float s_win[TCELLS * THD];
  __shared__ __align__(16) float s_go[TQ * THD];
  __shared__ int s_sum[TL * 3], s_org[TL * 2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t0 = blockIdx.x * TQ, h = blockIdx.y, b = blockIdx.z;
  if (tid < TL * 3) s_sum[tid] = 0;
  __syncthreads();
  TileGeom g;
  tile_geometry(g, order, ref, off, logit, sh, b, h, Q, nH, ref_bstride, t0, warp, lane, s_sum);
  const int rowpitch = nH * THD;
#pragma unroll
  for (int i = 0; i < TQW; ++i) {      // the warp's own g_out rows (read back by this warp only)
    float2 v = make_float2(0.f, 0.f);
    if (g.q[i] >= 0) v = __ldg((const float2*)(g_out + ((int64_t)b * Q + g.q[i]) * rowpitch + h * THD) + lane);
    *reinterpret_cast<float2*>(s_go + (warp * TQW + i) * THD + 2 * lane) = v;
  }
  __syncthreads();
  if (tid < TL) tile_origin(s_sum, s_org, sh, tid);
  __syncthreads();
  const int half = lane >> 4, cl = (lane & 15) * 4;
  const float* vb = value + (int64_t)b * S * rowpitch + h * THD;
  float rgw[TQW], rgx[TQW], rgy[TQW];        // lane = (level, point): d/d weight, d/d x_pix, d/d y_pix
#pragma unroll
  for (int i = 0; i < TQW; ++i) { rgw[i] = 0.f; rgx[i] = 0.f; rgy[i] = 0.f; }
  for (int l = 0; l < TL; ++l) {
    const int W = pick(sh.w, l), H = pick(sh.h, l), start = pick(sh.start, l), wx0 = s_org[l * 2], wy0 = s_org[l * 2 + 1];
    stage_window(s_win, vb, sh, l, wx0, wy0, rowpitch, tid);
    __syncthreads();
#pragma unroll
    for (int pass = 0; pass < TQW / 2; ++pass) {
      float part[24];                         // [(ii, j)][gw, gx, gy] over this lane's four channels
#pragma unroll
      for (int ii = 0; ii < 2; ++ii) {
        const int i = pass * 2 + ii;
        const float4 go = lds4(s_go + (warp * TQW + i) * THD + cl);
#pragma unroll
        for (int j = 0; j < TP / 2; ++j) {
          const int src = l * TP + 2 * j + half;
          const float x = __shfl_sync(0xffffffffu, g.px[i], src), y = __shfl_sync(0xffffffffu, g.py[i], src);
          const float a = __shfl_sync(0xffffffffu, g.aw[i], src);
          const float xf = floorf(x), yf = floorf(y);
          const float lx = x - xf, ly = y - yf;
          const Corners c = load_corners(s_win, vb, (int)xf, (int)yf, W, H, start, wx0, wy0, rowpitch, cl);
          const float d00 = dot4(go, c.v00), d01 = dot4(go, c.v01), d10 = dot4(go, c.v10), d11 = dot4(go, c.v11);
          float* p = part + (ii * 4 + j) * 3;
          p[0] = (1.f - ly) * ((1.f - lx) * d00 + lx * d01) + ly * ((1.f - lx) * d10 + lx * d11);
          p[1] = a * ((1.f - ly) * (d01 - d00) + ly * (d11 - d10));
          p[2] = a * ((1.f - lx) * (d10 - d00) + lx * (d11 - d01));
        }
      }
      float p12[12], p6[6], p3[3];
      halve_t<24>(part, p12, 8, (lane & 8) != 0);
      halve_t<12>(p12, p6, 4, (lane & 4) != 0);
      halve_t<6>(p6, p3, 2, (lane & 2) != 0);
#pragma unroll
      for (int c = 0; c < 3; ++c) p3[c] += __shfl_xor_sync(0xffffffffu, p3[c], 1);
      // lane (half, k = (lane >> 1) & 7) holds query ii = k >> 2, point 2 * (k & 3) + half of this level;
      // hand the triple to the lane that owns that point (lane == level * 8 + point)
#pragma unroll
      for (int ii = 0; ii < 2; ++ii) {
        const int src = ((lane & 1) << 4) | ((ii * 4 + ((lane & 7) >> 1)) << 1);
        const float gw = __shfl_sync(0xffffffffu, p3[0], src), gx = __shfl_sync(0xffffffffu, p3[1], src);
        const float gy = __shfl_sync(0xffffffffu, p3[2], src);
        if ((lane >> 3) == l) { rgw[pass * 2 + ii] = gw; rgx[pass * 2 + ii] = gx; rgy[pass * 2 + ii] = gy; }
      }
    }
    __syncthreads();
  }
  const int ll = lane >> 3;
  const float Wl = (float)pick(sh.w, ll), Hl = (float)pick(sh.h, ll);
#pragma unroll
  for (int i = 0; i < TQW; ++i) {
    if (g.q[i] < 0) continue;                 // warp-uniform
    const int64_t bq = (int64_t)b * Q + g.q[i];
    const float aw = g.aw[i];
    const float dot = warp_sum(aw * rgw[i]);
    g_logit[(bq * nH + h) * (TL * TP) + lane] = aw * (rgw[i] - dot);
    // x_pix = (ref + off / W) * W - 0.5  ->  d/d off = 1, d/d ref = W_l
    *((float2*)g_off + (bq * nH + h) * (TL * TP) + lane) = make_float2(rgx[i], rgy[i]);
    if (g_ref) {
      const float grx = warp_sum(rgx[i] * Wl), gry = warp_sum(rgy[i] * Hl);
      if (lane == 0) {
        float* gr = g_ref + (int64_t)b * ref_bstride + g.q[i] * 2;
        atomicAdd(gr, grx); atomicAdd(gr + 1, gry);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward, part 2: value gradient.  Corner records -> window cells (counting sort, integer shared-memory atomics),
// one register reduction and one global red row per non-empty cell.
// ---------------------------------------------------------------------------------------------------------------
constexpr int TREC = TQ * TL * TP * 4;        // 4096 corner records per tile
constexpr int TALLC = TL * TCELLS;            // 576 cells
constexpr int TOUT = TREC / 8;                // capacity of the outside-the-window list

struct ScatterSmem {
  float go[TQ * THD];                         // g_out rows of the tile                               8 KB
  float2 entry[TREC];                         // (weight, slot) grouped by cell                       32 KB
  float2 outside[TOUT];                       // (weight, slot) of valid corners outside the window    4 KB
  int outside_pos[TOUT];                      // their value row                                      2 KB
  int count[TALLC], offset[TALLC + 1];
  int sum[TL * 3], org[TL * 2];
  int n_outside;
};

__global__ void __launch_bounds__(TTHREADS, 3) msda_tile_bwd_scatter_kernel(
    const float* __restrict__ ref, const float* __restrict__ off, const float* __restrict__ logit,
    const int* __restrict__ order, const float* __restrict__ g_out, float* __restrict__ g_value, TileShapes sh, int B,
    int S, int Q, int nH, int ref_bstride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ScatterSmem& s = *reinterpret_cast<ScatterSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t0 = blockIdx.x * TQ, h = blockIdx.y, b = blockIdx.z;
  const int rowpitch = nH * THD;
  for (int i = tid; i < TALLC; i += TTHREADS) s.count[i] = 0;
  if (tid < TL * 3) s.sum[tid] = 0;
  if (tid == 0) s.n_outside = 0;
  __syncthreads();
  TileGeom g;
  tile_geometry(g, order, ref, off, logit, sh, b, h, Q, nH, ref_bstride, t0, warp, lane, s.sum);
#pragma unroll
  for (int i = 0; i < TQW; ++i) {
    float2 v = make_float2(0.f, 0.f);
    if (g.q[i] >= 0) v = __ldg((const float2*)(g_out + ((int64_t)b * Q + g.q[i]) * rowpitch + h * THD) + lane);
    *reinterpret_cast<float2*>(s.go + (warp * TQW + i) * THD + 2 * lane) = v;
  }
  __syncthreads();
  if (tid < TL) tile_origin(s.sum, s.org, sh, tid);
  __syncthreads();
  // ---- count: lane = (level, point); rank of every corner record inside its cell -------------------------------
  const int l = lane >> 3;
  const int W = pick(sh.w, l), H = pick(sh.h, l), lstart = pick(sh.start, l), wx0 = s.org[l * 2], wy0 = s.org[l * 2 + 1];
  int code[TQW][4];                            // (cell << 12) | rank, or -1 (no record / outside list)
#pragma unroll
  for (int i = 0; i < TQW; ++i) {
    const float xf = floorf(g.px[i]), yf = floorf(g.py[i]);
    const int x0 = (int)xf, y0 = (int)yf;
    const float lx = g.px[i] - xf, ly = g.py[i] - yf;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int x = x0 + (k & 1), y = y0 + (k >> 1);
      code[i][k] = -1;
      if (g.q[i] < 0 || x < 0 || x >= W || y < 0 || y >= H) continue;
      const int cx = x - wx0, cy = y - wy0;
      if ((unsigned)cx < (unsigned)TWIN && (unsigned)cy < (unsigned)TWIN) {
        const int cell = l * TCELLS + cy * TWIN + cx;
        code[i][k] = (cell << 12) | atomicAdd(&s.count[cell], 1);
      } else {
        const int pos = atomicAdd(&s.n_outside, 1);
        const float wgt = g.aw[i] * ((k & 1) ? lx : 1.f - lx) * ((k >> 1) ? ly : 1.f - ly);
        if (pos < TOUT) {
          s.outside[pos] = make_float2(wgt, __int_as_float(warp * TQW + i));
          s.outside_pos[pos] = lstart + y * W + x;
        } else {                                 // list full (pathological offsets): this lane adds the row itself
          float* dst = g_value + ((int64_t)b * S + lstart + y * W + x) * rowpitch + h * THD;
          const float* gr = s.go + (warp * TQW + i) * THD;
          for (int c = 0; c < THD; c += 4) {
            const float4 v = lds4(gr + c);
            atomicAdd((float4*)(dst + c), make_float4(v.x * wgt, v.y * wgt, v.z * wgt, v.w * wgt));
          }
        }
      }
    }
  }
  __syncthreads();
  // ---- exclusive prefix sum over the 576 cells (one warp, 18 cells per lane) -----------------------------------
  if (warp == 0) {
    constexpr int PER = TALLC / 32;
    int local[PER], tot = 0;
#pragma unroll
    for (int i = 0; i < PER; ++i) { local[i] = tot; tot += s.count[lane * PER + i]; }
    int incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    const int base = incl - tot;
#pragma unroll
    for (int i = 0; i < PER; ++i) s.offset[lane * PER + i] = base + local[i];
    if (lane == 31) s.offset[TALLC] = incl;
  }
  __syncthreads();
  // ---- place (weight, slot) at offset[cell] + rank -------------------------------------------------------------
#pragma unroll
  for (int i = 0; i < TQW; ++i) {
    const float xf = floorf(g.px[i]), yf = floorf(g.py[i]);
    const float lx = g.px[i] - xf, ly = g.py[i] - yf;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (code[i][k] < 0) continue;
      const float wgt = g.aw[i] * ((k & 1) ? lx : 1.f - lx) * ((k >> 1) ? ly : 1.f - ly);
      s.entry[s.offset[code[i][k] >> 12] + (code[i][k] & 4095)] = make_float2(wgt, __int_as_float(warp * TQW + i));
    }
  }
  __syncthreads();
  // ---- reduce: a warp per cell, half-warps take alternate records, a lane owns 4 of the 64 channels -----------
  const int half = lane >> 4, cl = (lane & 15) * 4;
  float* gvb = g_value + (int64_t)b * S * rowpitch + h * THD + cl;
  for (int c = warp; c < TALLC; c += TWARPS) {
    const int beg = s.offset[c], end = s.offset[c + 1];
    if (beg == end) continue;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = beg + half; i < end; i += 2) {
      const float2 e = s.entry[i];
      fma4(acc, e.x, lds4(s.go + __float_as_int(e.y) * THD + cl));
    }
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, 16); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, 16);
    if (half == 0) {
      const int lv = c / TCELLS, r = c - lv * TCELLS, cy = r / TWIN, cx = r - cy * TWIN;
      const int64_t pos = pick(sh.start, lv) + (int64_t)(s.org[lv * 2 + 1] + cy) * pick(sh.w, lv) + (s.org[lv * 2] + cx);
      atomicAdd((float4*)(gvb + pos * rowpitch), acc);
    }
  }
  // ---- the few records outside their window: one 256-byte red each --------------------------------------------
  const int n_out = min(s.n_outside, TOUT);
  for (int i = warp * 2 + half; i < n_out; i += TWARPS * 2) {
    const float2 e = s.outside[i];
    const float4 v = lds4(s.go + __float_as_int(e.y) * THD + cl);
    atomicAdd((float4*)(gvb + (int64_t)s.outside_pos[i] * rowpitch), make_float4(v.x * e.x, v.y * e.x, v.z * e.x, v.w * e.x));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// query order: counting sort by (band of ref_y, bucket of ref_x)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int sort_bucket(const float* __restrict__ ref, int q, int bands, int xb) {
  const float rx = __ldg(ref + 2 * q), ry = __ldg(ref + 2 * q + 1);
  const int by = min(bands - 1, max(0, (int)floorf(fminf(fmaxf(ry, -1.f), 2.f) * (float)bands)));
  const int bx = min(xb - 1, max(0, (int)floorf(fminf(fmaxf(rx, -1.f), 2.f) * (float)xb)));
  return by * xb + bx;
}
__global__ void sort_hist_kernel(const float* __restrict__ ref, int Q, int bands, int xb, int* __restrict__ count) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < Q) atomicAdd(count + sort_bucket(ref, q, bands, xb), 1);
}
// single CTA: exclusive scan of n counters in place
__global__ void __launch_bounds__(1024) sort_scan_kernel(int* __restrict__ count, int n) {
  __shared__ int s_part[1024];
  const int t = threadIdx.x, per = (n + 1023) / 1024;
  const int beg = min(n, t * per), end = min(n, beg + per);
  int tot = 0;
  for (int i = beg; i < end; ++i) tot += count[i];
  s_part[t] = tot;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int v = t >= o ? s_part[t - o] : 0;
    __syncthreads();
    s_part[t] += v;
    __syncthreads();
  }
  int run = s_part[t] - tot;
  for (int i = beg; i < end; ++i) { const int c = count[i]; count[i] = run; run += c; }
}
__global__ void sort_place_kernel(const float* __restrict__ ref, int Q, int bands, int xb, int* __restrict__ cursor,
                                  int* __restrict__ order) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < Q) order[atomicAdd(cursor + sort_bucket(ref, q, bands, xb), 1)] = q;
}

static int fill_tile_shapes(const int* hw, int L, int S, TileShapes& sh) {
  if (L != TL) return GED_ERR_SHAPE;
  int start = 0;
  for (int l = 0; l < TL; ++l) {
    sh.h[l] = hw[2 * l]; sh.w[l] = hw[2 * l + 1]; sh.start[l] = start;
    if (sh.h[l] <= 0 || sh.w[l] <= 0 || sh.h[l] > 16384 || sh.w[l] > 16384) return GED_ERR_SHAPE;
    start += sh.h[l] * sh.w[l];
  }
  return start == S ? GED_OK : GED_ERR_SHAPE;
}

}  // namespace ged
using namespace ged;

// order (Q) int32 <- permutation of 0..Q-1 grouped by (band of ref_y, bucket of ref_x); ref (Q,2) in [0,1] (values
// outside are clamped into the edge buckets).  work: bands * xbuckets ints.  The order within a bucket is unspecified.
GED_API int ged_msda_sort_queries(const float* ref, int Q, int bands, int xbuckets, int* order, int* work,
                                  int64_t work_ints, cudaStream_t stream) {
  if (!ref || !order || !work || Q <= 0 || bands <= 0 || xbuckets <= 0) return GED_ERR_ARG;
  const int64_t n = (int64_t)bands * xbuckets;
  if (n > work_ints || n > (1 << 22)) return GED_ERR_WORKSPACE;
  if (cudaMemsetAsync(work, 0, n * sizeof(int), stream) != cudaSuccess) return GED_ERR_LAUNCH;
  sort_hist_kernel<<<cdiv(Q, 256), 256, 0, stream>>>(ref, Q, bands, xbuckets, work);
  sort_scan_kernel<<<1, 1024, 0, stream>>>(work, (int)n);
  sort_place_kernel<<<cdiv(Q, 256), 256, 0, stream>>>(ref, Q, bands, xbuckets, work, order);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

// Same tensors as ged_msda_fwd + order (Q) int32: any permutation of the queries (ged_msda_sort_queries for locality).
GED_API int ged_msda_tile_fwd(const float* value, const float* ref, int ref_batch, const float* off, const float* logit,
                              const int* order, float* out, const int* level_hw, int num_levels, int B, int S, int Q,
                              int nH, int head_dim, int num_points, cudaStream_t stream) {
  if (!value || !ref || !off || !logit || !order || !out || !level_hw) return GED_ERR_ARG;
  if (head_dim != THD || num_points != TP || (ref_batch != 1 && ref_batch != B)) return GED_ERR_SHAPE;
  TileShapes sh;
  if (int e = fill_tile_shapes(level_hw, num_levels, S, sh)) return e;
  msda_tile_fwd_kernel<<<dim3(cdiv(Q, TQ), nH, B), TTHREADS, 0, stream>>>(value, ref, off, logit, order, out, sh, B, S, Q, nH,
                                                                         ref_batch == 1 ? 0 : Q * 2);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

// g_value accumulated (zeroed by the caller); g_ref (ref_batch,Q,2) accumulated or NULL; g_off, g_logit overwritten.
GED_API int ged_msda_tile_bwd(const float* value, const float* ref, int ref_batch, const float* off, const float* logit,
                              const int* order, const float* g_out, float* g_value, float* g_ref, float* g_off,
                              float* g_logit, const int* level_hw, int num_levels, int B, int S, int Q, int nH,
                              int head_dim, int num_points, cudaStream_t stream) {
  if (!value || !ref || !off || !logit || !order || !g_out || !g_value || !g_off || !g_logit || !level_hw) return GED_ERR_ARG;
  if (head_dim != THD || num_points != TP || (ref_batch != 1 && ref_batch != B)) return GED_ERR_SHAPE;
  TileShapes sh;
  if (int e = fill_tile_shapes(level_hw, num_levels, S, sh)) return e;
  const int rbs = ref_batch == 1 ? 0 : Q * 2;
  const dim3 grid(cdiv(Q, TQ), nH, B);
  if (cudaFuncSetAttribute(msda_tile_bwd_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScatterSmem)) != cudaSuccess)
    return GED_ERR_LAUNCH;
  msda_tile_bwd_scatter_kernel<<<grid, TTHREADS, sizeof(ScatterSmem), stream>>>(ref, off, logit, order, g_out, g_value, sh, B, S,
                                                                                Q, nH, rbs);
  msda_tile_bwd_dot_kernel<<<grid, TTHREADS, 0, stream>>>(value, ref, off, logit, order, g_out, g_ref, g_off, g_logit, sh, B, S,
                                                          Q, nH, rbs);
  GED_CHECK_LAUNCH();
  return GED_OK;
}
