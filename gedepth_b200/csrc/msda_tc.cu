// Multi-scale deformable attention on the 5th-generation tensor cores, sm_100a (a11: depth/models/necks/hahi.py:280-289,
// 316-325 -> mmcv.ops.MultiScaleDeformableAttention [external, mmcv-full 1.3.13]; semantics as in msda.cu).
//
// Same tiling as msda_tile.cu (32 consecutive SORTED queries of one (batch, head) per CTA, a window of value rows per
// level in shared memory), but the bilinear gather / scatter is restated as dense algebra over the window so that it
// runs on tcgen05.mma instead of ~10^5 SIMT instructions per tile.  Per level l, with
//     A_l [32 queries x 121 window cells] = sum over the query's 8 points x 4 corners of (attention weight x
//                                           bilinear weight) at the cell the corner falls into   (sparse, built by SIMT)
//     V_l [121 cells x 64 channels]       = the window of value rows,      G [32 x 64] = the tile's g_out rows:
//   forward           out^T [64 x 32]   += V_l^T . A_l^T          (3xTF32: fp32-accurate, accumulated over levels in TMEM)
//   value gradient    dV_l  [121 x 64]   = A_l^T . G              (one pass TF32, like every other backward GEMM)
//   corner dots       D_l^T [121 x 32]   = V_l . G^T              (-> d/d weight, d/d x, d/d y per point by 4 lookups)
// Operands are written straight into 128B-swizzled K-major (and, for V_l^T, MN-major) UMMA tiles by the staging code;
// accumulators live in TMEM and are read back with tcgen05.ld.  Corners outside their window (~1-3 %) take a direct
// warp-cooperative path in fp32.
#include <cuda.h>
#include <cudaTypedefs.h>
#include "common.cuh"

namespace ged {

constexpr int CL = 4, CP = 8, CHD = 64;
constexpr int CQ = 32;                       // queries per tile
constexpr int CWARPS = 8, CTHREADS = CWARPS * 32;
constexpr int CWX = 11, CWY = 11, CCELLS = CWX * CWY;   // 121 cells, padded to the 128 TMEM lanes

struct TcShapes {
  int h[CL], w[CL], start[CL];
};
struct TcMaps {          // one 4-D tensor map (channel, x, y, batch) per level of the value tensor
  CUtensorMap m[CL];
};
constexpr uint32_t WIN_BYTES = CCELLS * 128;      // one TMA box: 121 cells x 32 channels

// ---- PTX wrappers (same forms as gemm_tcgen05.cu) ----------------------------------------------------------------
__device__ __forceinline__ uint32_t c_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void c_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(c_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void c_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "CWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra CDONE_%=;\n\t"
      "bra CWAIT_%=;\n\t"
      "CDONE_%=:\n\t}"
      :: "r"(c_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void c_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(c_smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA: one {32 channels, CWX, CWY, 1} box of a level map -> [121 rows][128 B] in the swizzle mode of the tensor map
__device__ __forceinline__ void c_tma_window(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int ch, int x, int y, int b) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :: "r"(c_smem_u32(smem_dst)), "l"((uint64_t)map), "r"(c_smem_u32(bar)), "r"(ch), "r"(x), "r"(y), "r"(b) : "memory");
}
__device__ __forceinline__ void c_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void c_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void c_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void c_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(c_smem_u32(bar)) : "memory");
}
// one lane of a CONVERGED warp: tcgen05.mma issued under it needs no per-instruction elect-and-broadcast loop (see
// gemm_tcgen05.cu::elect_one)
__device__ __forceinline__ bool c_elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void c_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void c_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void c_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// K-major 128B-swizzled tile: rows of 32 floats, 8-row atoms 1024 bytes apart (cute::UMMA::SmemDescriptor, SWIZZLE_128B)
__device__ __forceinline__ uint64_t c_desc_kmajor(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// MN-major tile (Layout_MN_SW128_32B_Atom): panels of 32 floats along MN, 4 K-rows per 512-byte atom
__device__ __forceinline__ uint64_t c_desc_mnmajor(uint32_t smem_addr, uint32_t panel_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((panel_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(512 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
__host__ __device__ constexpr uint32_t c_idesc_tf32(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// byte offset of element (row r, float column c < 32) inside a K-major 128B-swizzled tile
__device__ __forceinline__ uint32_t kmaj_off(int r, int c) { return (uint32_t)(r * 128 + ((((c >> 2) ^ (r & 7)) << 4) | ((c & 3) << 2))); }

__device__ __forceinline__ int cpick(const int (&a)[CL], int l) { return l == 0 ? a[0] : (l == 1 ? a[1] : (l == 2 ? a[2] : a[3])); }
__device__ __forceinline__ float group8_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2); v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}
__device__ __forceinline__ float group8_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1)); v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
  return v;
}

// Geometry of the tile in the (query-in-warp i = lane >> 3, point p = lane & 7) layout: every lane holds its point of
// each of the four levels.  Also accumulates the per-level sums the window origins are derived from.
struct TcGeom {
  float px[CL], py[CL], aw[CL];
  int q;                                  // original query index of this lane's slot, -1 past the end
};

__device__ __forceinline__ void tc_geometry(TcGeom& g, const int* __restrict__ order, const float* __restrict__ ref,
                                            const float* __restrict__ off, const float* __restrict__ logit, const TcShapes& sh,
                                            int b, int h, int Q, int nH, int ref_bstride, int t0, int warp, int lane, int* s_sum) {
  const int p = lane & 7, slot = warp * 4 + (lane >> 3);
  const int qs = t0 + slot;
  g.q = qs < Q ? __ldg(order + qs) : -1;
  float lg[CL];
  float rx = 0.f, ry = 0.f;
  const int64_t bq = (int64_t)b * Q + max(g.q, 0);
  if (g.q >= 0) { rx = __ldg(ref + (int64_t)b * ref_bstride + g.q * 2); ry = __ldg(ref + (int64_t)b * ref_bstride + g.q * 2 + 1); }
  float mx = -3.4e38f;
#pragma unroll
  for (int l = 0; l < CL; ++l) { lg[l] = __ldg(logit + (bq * nH + h) * (CL * CP) + l * CP + p); mx = fmaxf(mx, lg[l]); }
  mx = group8_max(mx);
  float se = 0.f;
#pragma unroll
  for (int l = 0; l < CL; ++l) { lg[l] = __expf(lg[l] - mx); se += lg[l]; }
  se = group8_sum(se);
#pragma unroll
  for (int l = 0; l < CL; ++l) {
    const int Wi = sh.w[l], Hi = sh.h[l];
    const float Wl = (float)Wi, Hl = (float)Hi;
    const float2 o = __ldg((const float2*)(off + (bq * nH + h) * (CL * CP * 2)) + l * CP + p);
    float px = fminf(fmaxf((rx + o.x / Wl) * Wl - 0.5f, -30000.f), 30000.f);
    float py = fminf(fmaxf((ry + o.y / Hl) * Hl - 0.5f, -30000.f), 30000.f);
    if (g.q < 0) { px = -30000.f; py = -30000.f; }
    g.px[l] = px; g.py[l] = py; g.aw[l] = g.q >= 0 ? lg[l] / se : 0.f;
    const int x0 = (int)floorf(px), y0 = (int)floorf(py);
    const bool in = x0 >= -1 && x0 < Wi && y0 >= -1 && y0 < Hi;
    const int sx = __reduce_add_sync(0xffffffffu, in ? x0 : 0), sy = __reduce_add_sync(0xffffffffu, in ? y0 : 0);
    const int sn = __reduce_add_sync(0xffffffffu, in ? 1 : 0);
    if (lane == 0 && sn > 0) { atomicAdd(&s_sum[l * 3], sx); atomicAdd(&s_sum[l * 3 + 1], sy); atomicAdd(&s_sum[l * 3 + 2], sn); }
  }
}

__device__ __forceinline__ void tc_origin(const int* s_sum, int* s_org, const TcShapes& sh, int l) {
  const int n = max(s_sum[l * 3 + 2], 1);
  const int mx = (int)floorf((float)s_sum[l * 3] / (float)n + 0.5f), my = (int)floorf((float)s_sum[l * 3 + 1] / (float)n + 0.5f);
  s_org[l * 2] = max(0, min(mx - (CWX / 2 - 1), cpick(sh.w, l) - CWX));
  s_org[l * 2 + 1] = max(0, min(my - (CWY / 2 - 1), cpick(sh.h, l) - CWY));
}

// Valid corners outside the window of one level, fp32, warp-cooperative (2 channels per lane): the event is handled per
// POINT (the corners of a stray point usually leave the window together).  FWD: extra[slot] += w * V[pos];
// BWD: g_value[pos] += w * g_out[q] and the corner's dot product <g_out[q], V[pos]> back to the owner lane.
template <bool FWD>
__device__ __forceinline__ void stray_corners(unsigned omask, const int (&pos)[4], const float (&wgt)[4], float (&dk)[4], int q,
                                              int slot, const float* __restrict__ vb, float* __restrict__ gvb,
                                              const float* __restrict__ gob, float* s_extra, int rowpitch, int lane) {
  unsigned m = __ballot_sync(0xffffffffu, omask != 0u);
  while (m) {
    const int src = __ffs(m) - 1;
    m &= m - 1;
    const unsigned sm = __shfl_sync(0xffffffffu, omask, src);
    const int sq = __shfl_sync(0xffffffffu, q, src), sslot = __shfl_sync(0xffffffffu, slot, src);
    float2 g2 = make_float2(0.f, 0.f);
    if (!FWD) g2 = __ldg((const float2*)(gob + (int64_t)sq * rowpitch) + lane);
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int spos = __shfl_sync(0xffffffffu, pos[k], src);
      const float sw = __shfl_sync(0xffffffffu, wgt[k], src);
      if (!((sm >> k) & 1u)) continue;                       // warp-uniform
      const float2 v2 = __ldg((const float2*)(vb + (int64_t)spos * rowpitch) + lane);
      if (FWD) { acc.x += sw * v2.x; acc.y += sw * v2.y; }
      else {
        atomicAdd((float2*)(gvb + (int64_t)spos * rowpitch) + lane, make_float2(sw * g2.x, sw * g2.y));
        const float d = warp_sum(g2.x * v2.x + g2.y * v2.y);
        if (lane == src) dk[k] = d;
      }
    }
    if (FWD) {      // rows of s_extra belong to the warp that owns the slot: plain read-modify-write
      float2* e = (float2*)(s_extra + sslot * CHD) + lane;
      float2 t = *e;
      t.x += acc.x; t.y += acc.y;
      *e = t;
    }
  }
}

// =================================================================================================================
// backward
// =================================================================================================================
// shared memory (bytes from the 1024-aligned base)
constexpr int B_V = 0;                 // V_l: 2 K-blocks [128 cells][32 ch], written by TMA (A of D^T)  32 KB; block 0 is
                                       //      reused as the corner-dot table once the MMAs have retired
constexpr int B_AT = 32768;            // A_l^T [128 cells][32 queries] swizzled (A of dV)             16 KB
constexpr int B_GT = 49152;            // G^T [64 ch][32 queries] swizzled (B of dV)                    8 KB
constexpr int B_GQ = 57344;            // G: 2 K-blocks [32 queries][32 ch] swizzled (B of D^T)         8 KB
constexpr int B_MISC = 65536;          // mbarriers, TMEM base, window sums / origins
constexpr int B_TOTAL = B_MISC + 128 + 1024;

__global__ void __launch_bounds__(CTHREADS, 3) msda_tc_bwd_kernel(
    const __grid_constant__ TcMaps maps, const float* __restrict__ value, const float* __restrict__ ref,
    const float* __restrict__ off, const float* __restrict__ logit, const int* __restrict__ order,
    const float* __restrict__ g_out, float* __restrict__ g_value, float* __restrict__ g_ref, float* __restrict__ g_off,
    float* __restrict__ g_logit, TcShapes sh, int B, int S, int Q, int nH, int ref_bstride) {
  extern __shared__ uint8_t c_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)c_smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* mma_bar = (uint64_t*)(smem + B_MISC);
  uint64_t* tma_bar = mma_bar + 1;
  uint32_t* tmem_ptr = (uint32_t*)(smem + B_MISC + 16);
  int* s_sum = (int*)(smem + B_MISC + 32);      // [12]
  int* s_org = s_sum + 12;                      // [8]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t0 = blockIdx.x * CQ, h = blockIdx.y, b = blockIdx.z;
  const int slot = warp * 4 + (lane >> 3), p = lane & 7;
  const int rowpitch = nH * CHD;

  if (tid == 0) {
    c_mbar_init(mma_bar, 1); c_mbar_init(tma_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
    for (int l = 0; l < CL; ++l) asm volatile("prefetch.tensormap [%0];" :: "l"((uint64_t)&maps.m[l]) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(c_smem_u32(tmem_ptr)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid < 12) s_sum[tid] = 0;
  for (int idx = tid; idx < 16384 / 16; idx += CTHREADS) *(float4*)(smem + B_AT + idx * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  TcGeom g;
  tc_geometry(g, order, ref, off, logit, sh, b, h, Q, nH, ref_bstride, t0, warp, lane, s_sum);
  // the tile's g_out rows, in both operand orientations (the warp stages its own four slots, 2 channels per lane)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int sl = warp * 4 + i;
    const int qi = __shfl_sync(0xffffffffu, g.q, i * 8);
    float2 v = make_float2(0.f, 0.f);
    if (qi >= 0) v = __ldg((const float2*)(g_out + ((int64_t)b * Q + qi) * rowpitch + h * CHD) + lane);
    const int ch = 2 * lane;
    *(float2*)(smem + B_GQ + (ch >> 5) * 4096 + kmaj_off(sl, ch & 31)) = v;
    *(float*)(smem + B_GT + kmaj_off(ch, sl)) = v.x;
    *(float*)(smem + B_GT + kmaj_off(ch + 1, sl)) = v.y;
  }
  c_fence_before();
  __syncthreads();
  c_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (tid < CL) tc_origin(s_sum, s_org, sh, tid);
  __syncthreads();

  const float* vb = value + (int64_t)b * S * rowpitch + h * CHD;
  float* gvb = g_value + (int64_t)b * S * rowpitch + h * CHD;
  const float* gob = g_out + (int64_t)b * Q * rowpitch + h * CHD;
  float rgw[CL], rgx[CL], rgy[CL];
  const int qd = warp & 3, hf = warp >> 2;       // TMEM lane quadrant of this warp, column half it reads
  const int ecell = qd * 32 + lane;              // the window cell (TMEM lane) this thread reads back
  const int ecy = ecell / CWX, ecx = ecell - ecy * CWX;

#pragma unroll
  for (int l = 0; l < CL; ++l) {
    const int W = sh.w[l], H = sh.h[l], start = sh.start[l], wx0 = s_org[l * 2], wy0 = s_org[l * 2 + 1];
    // ---- build A_l^T.  A warp owns the columns of its four slots (cleared by its own epilogue); two points of one
    //      query may share a cell, so the eight points of a slot take turns: plain read-modify-write, no atomics ------
    const float xf = floorf(g.px[l]), yf = floorf(g.py[l]);
    const int x0 = (int)xf, y0 = (int)yf;
    const float lx = g.px[l] - xf, ly = g.py[l] - yf, a = g.aw[l];
    int cellk[4], posk[4];
    float dk[4], wk[4];
    unsigned omask = 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int x = x0 + (k & 1), y = y0 + (k >> 1);
      const bool valid = g.q >= 0 && x >= 0 && x < W && y >= 0 && y < H;
      const int cx = x - wx0, cy = y - wy0;
      const bool inwin = valid && (unsigned)cx < (unsigned)CWX && (unsigned)cy < (unsigned)CWY;
      wk[k] = a * ((k & 1) ? lx : 1.f - lx) * ((k >> 1) ? ly : 1.f - ly);
      cellk[k] = inwin ? cy * CWX + cx : -1;
      posk[k] = start + y * W + x;
      dk[k] = 0.f;
      if (valid && !inwin) omask |= 1u << k;
    }
    {
      float* ap[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) ap[k] = (float*)(smem + B_AT + kmaj_off(max(cellk[k], 0), slot));
#pragma unroll
      for (int r = 0; r < CP; ++r) {
        if (p == r) {
          float t[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) t[k] = cellk[k] >= 0 ? *ap[k] : 0.f;
#pragma unroll
          for (int k = 0; k < 4; ++k) if (cellk[k] >= 0) *ap[k] = t[k] + wk[k];
        }
        __syncwarp();
      }
    }
    c_fence_async();
    __syncthreads();
    // ---- TMA: V_l window; tensor cores: dV_l = A_l^T . G (cols 0..63), D_l^T = V_l . G^T (cols 64..95) -----------
    if (warp == 0 && c_elect_one()) {
      c_mbar_expect_tx(tma_bar, 2 * WIN_BYTES);
      c_tma_window(smem + B_V, &maps.m[l], tma_bar, h * CHD, wx0, wy0, b);
      c_tma_window(smem + B_V + 16384, &maps.m[l], tma_bar, h * CHD + 32, wx0, wy0, b);
      c_fence_after();
      const uint32_t base = c_smem_u32(smem);
      const uint64_t a_at = c_desc_kmajor(base + B_AT), b_gt = c_desc_kmajor(base + B_GT);
      constexpr uint32_t idv = c_idesc_tf32(128, 64, false, false), idd = c_idesc_tf32(128, 32, false, false);
#pragma unroll
      for (int k = 0; k < 4; ++k) c_mma_tf32(tmem_base, a_at + 2 * k, b_gt + 2 * k, idv, k != 0);
      c_mbar_wait(tma_bar, l & 1);
      c_fence_after();
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        const uint64_t a_v = c_desc_kmajor(base + B_V + kb * 16384), b_gq = c_desc_kmajor(base + B_GQ + kb * 4096);
#pragma unroll
        for (int k = 0; k < 4; ++k) c_mma_tf32(tmem_base + 64, a_v + 2 * k, b_gq + 2 * k, idd, (kb | k) != 0);
      }
      c_commit(mma_bar);
    }
    __syncwarp();
    // corners outside the window: direct path, overlapped with the TMA / MMA latency
    stray_corners<false>(omask, posk, wk, dk, g.q, slot, vb, gvb, gob, nullptr, rowpitch, lane);
    c_mbar_wait(mma_bar, l & 1);
    c_fence_after();
    // ---- epilogue: this thread's cell row.  dV half -> red.global (skipped when the cell was not touched); corner
    //      dots -> swizzled table over V block 0 (dead now); clear the warp's own columns of A^T for the next level ---
    {
      uint32_t v[32];
      c_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(hf * 32), v);
      uint32_t any = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) any |= v[j];
      if ((any << 1) != 0u && ecell < CCELLS) {
        float* dst = gvb + (int64_t)(start + (wy0 + ecy) * W + (wx0 + ecx)) * rowpitch + hf * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          atomicAdd((float4*)dst + j, make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                   __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
      }
      uint32_t d[16];
      c_ld16(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(64 + hf * 16), d);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *(float4*)(smem + B_V + ecell * 128 + (((hf * 4 + j) ^ (ecell & 7)) << 4)) =
            make_float4(__uint_as_float(d[4 * j]), __uint_as_float(d[4 * j + 1]), __uint_as_float(d[4 * j + 2]), __uint_as_float(d[4 * j + 3]));
      if (l + 1 < CL) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {      // slots 4 warp .. 4 warp + 3 of cell row c are one 16-byte chunk
          const int c = lane + 32 * j;
          *(float4*)(smem + B_AT + c * 128 + ((warp ^ (c & 7)) << 4)) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    c_fence_before();
    __syncthreads();
    // ---- per point: d/d weight, d/d x_pix, d/d y_pix from the four corner dots -------------------------------------
#pragma unroll
    for (int k = 0; k < 4; ++k) if (cellk[k] >= 0) dk[k] = *(const float*)(smem + B_V + kmaj_off(cellk[k], slot));
    rgw[l] = (1.f - ly) * ((1.f - lx) * dk[0] + lx * dk[1]) + ly * ((1.f - lx) * dk[2] + lx * dk[3]);
    rgx[l] = a * ((1.f - ly) * (dk[1] - dk[0]) + ly * (dk[3] - dk[2]));
    rgy[l] = a * ((1.f - lx) * (dk[2] - dk[0]) + lx * (dk[3] - dk[1]));
    // no barrier here: the next level's TMA (which overwrites the table) is issued after the next __syncthreads
  }
  // ---- softmax backward, stores ------------------------------------------------------------------------------------
  float dot = 0.f, grx = 0.f, gry = 0.f;
#pragma unroll
  for (int l = 0; l < CL; ++l) { dot += g.aw[l] * rgw[l]; grx += rgx[l] * (float)sh.w[l]; gry += rgy[l] * (float)sh.h[l]; }
  dot = group8_sum(dot);
  if (g_ref) { grx = group8_sum(grx); gry = group8_sum(gry); }
  if (g.q >= 0) {
    const int64_t bq = (int64_t)b * Q + g.q;
#pragma unroll
    for (int l = 0; l < CL; ++l) {
      g_logit[(bq * nH + h) * (CL * CP) + l * CP + p] = g.aw[l] * (rgw[l] - dot);
      // x_pix = (ref + off / W) * W - 0.5  ->  d/d off = 1, d/d ref = W_l
      *((float2*)g_off + (bq * nH + h) * (CL * CP) + l * CP + p) = make_float2(rgx[l], rgy[l]);
    }
    if (g_ref && p == 0) {
      float* gr = g_ref + (int64_t)b * ref_bstride + g.q * 2;
      atomicAdd(gr, grx); atomicAdd(gr + 1, gry);
    }
  }
  c_fence_before();
  __syncthreads();
  if (warp == 0) {
    c_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(128) : "memory");
  }
}

// =================================================================================================================
// forward: out^T [64 ch x 32 queries] += V_l^T . A_l^T, 3xTF32 (hi*hi + lo*hi + hi*lo), accumulated over the four
// levels in one TMEM tile.  V_l arrives by TMA as an MN-major operand (channels contiguous), A_l is K-major.
// =================================================================================================================
constexpr int F_VHI = 0;               // V_l: 2 panels [128 cells][32 ch] (SWIZZLE_128B_ATOM_32B), raw fp32 = hi   32 KB
constexpr int F_VLO = 32768;           // x - tf32(x), pre-computed once per value tensor and fetched by TMA too     32 KB
constexpr int F_AHI = 65536;           // A_l [32 queries][128 cells]: 4 K-blocks of [32][32] swizzled, fp32 = hi  16 KB
constexpr int F_ALO = 81920;           //                                                                       16 KB
constexpr int F_EXTRA = 98304;         // [32 slots][64 ch] fp32: corners outside the windows                    8 KB
constexpr int F_MISC = 106496;
constexpr int F_TOTAL = F_MISC + 256 + 1024;

__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// lo = x - tf32(x) of a whole tensor (the tensor core truncates the fp32 operand to its top 19 bits, so hi is x itself)
__global__ void __launch_bounds__(256) split_lo_kernel(const float4* __restrict__ x, float4* __restrict__ lo, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = ldg_stream(x + i);
    lo[i] = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
  }
}

__global__ void __launch_bounds__(CTHREADS, 2) msda_tc_fwd_kernel(
    const __grid_constant__ TcMaps maps, const __grid_constant__ TcMaps maps_lo, const float* __restrict__ value,
    const float* __restrict__ ref, const float* __restrict__ off, const float* __restrict__ logit,
    const int* __restrict__ order, float* __restrict__ out, TcShapes sh, int B, int S, int Q, int nH, int ref_bstride) {
  extern __shared__ uint8_t c_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)c_smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* mma_bar = (uint64_t*)(smem + F_MISC);
  uint64_t* tma_bar = mma_bar + 1;
  uint32_t* tmem_ptr = (uint32_t*)(smem + F_MISC + 16);
  int* s_sum = (int*)(smem + F_MISC + 32);      // [12]
  int* s_org = s_sum + 12;                      // [8]
  int* s_q = s_org + 8;                         // [32] original query index of every slot
  float* s_extra = (float*)(smem + F_EXTRA);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t0 = blockIdx.x * CQ, h = blockIdx.y, b = blockIdx.z;
  const int slot = warp * 4 + (lane >> 3), p = lane & 7;
  const int rowpitch = nH * CHD;

  if (tid == 0) {
    c_mbar_init(mma_bar, 1); c_mbar_init(tma_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
    for (int l = 0; l < CL; ++l) {
      asm volatile("prefetch.tensormap [%0];" :: "l"((uint64_t)&maps.m[l]) : "memory");
      asm volatile("prefetch.tensormap [%0];" :: "l"((uint64_t)&maps_lo.m[l]) : "memory");
    }
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(c_smem_u32(tmem_ptr)), "r"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid < 12) s_sum[tid] = 0;
  // A (hi) and the stray-corner rows start at zero; so do rows 121..127 of the V tiles, which TMA never writes but the
  // contraction over 128 cells reads
  for (int idx = tid; idx < (16384 + 8192) / 16; idx += CTHREADS) {
    const int o = idx * 16;
    *(float4*)(smem + (o < 16384 ? F_AHI + o : F_EXTRA + o - 16384)) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int idx = tid; idx < 4 * 7 * 8; idx += CTHREADS) {        // 4 tiles (hi/lo x 2 panels) x 7 rows x 8 chunks
    const int t = idx / 56, r = idx - t * 56;
    *(float4*)(smem + (t >> 1) * 32768 + (t & 1) * 16384 + CCELLS * 128 + r * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  c_fence_async();
  __syncthreads();
  TcGeom g;
  tc_geometry(g, order, ref, off, logit, sh, b, h, Q, nH, ref_bstride, t0, warp, lane, s_sum);
  if (p == 0) s_q[slot] = g.q;
  c_fence_before();
  __syncthreads();
  c_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (tid < CL) tc_origin(s_sum, s_org, sh, tid);
  __syncthreads();
  const float* vb = value + (int64_t)b * S * rowpitch + h * CHD;

#pragma unroll
  for (int l = 0; l < CL; ++l) {
    const int W = sh.w[l], H = sh.h[l], start = sh.start[l], wx0 = s_org[l * 2], wy0 = s_org[l * 2 + 1];
    // ---- TMA for V_l hi / lo (the previous level's MMAs have retired), then build A_l while it is in flight -------
    if (tid == 0) {
      c_mbar_expect_tx(tma_bar, 4 * WIN_BYTES);
      c_tma_window(smem + F_VHI, &maps.m[l], tma_bar, h * CHD, wx0, wy0, b);
      c_tma_window(smem + F_VHI + 16384, &maps.m[l], tma_bar, h * CHD + 32, wx0, wy0, b);
      c_tma_window(smem + F_VLO, &maps_lo.m[l], tma_bar, h * CHD, wx0, wy0, b);
      c_tma_window(smem + F_VLO + 16384, &maps_lo.m[l], tma_bar, h * CHD + 32, wx0, wy0, b);
    }
    const float xf = floorf(g.px[l]), yf = floorf(g.py[l]);
    const int x0 = (int)xf, y0 = (int)yf;
    const float lx = g.px[l] - xf, ly = g.py[l] - yf, a = g.aw[l];
    int cellk[4], posk[4];
    float dk[4], wk[4];
    unsigned omask = 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int x = x0 + (k & 1), y = y0 + (k >> 1);
      const bool valid = g.q >= 0 && x >= 0 && x < W && y >= 0 && y < H;
      const int cx = x - wx0, cy = y - wy0;
      const bool inwin = valid && (unsigned)cx < (unsigned)CWX && (unsigned)cy < (unsigned)CWY;
      wk[k] = a * ((k & 1) ? lx : 1.f - lx) * ((k >> 1) ? ly : 1.f - ly);
      cellk[k] = inwin ? cy * CWX + cx : -1;
      posk[k] = start + y * W + x;
      dk[k] = 0.f;
      if (valid && !inwin) omask |= 1u << k;
    }
    // A warp owns the rows of its four slots: the eight points of a slot take turns (two of them may share a cell),
    // plain read-modify-write; then the warp splits its own rows (lo = x - tf32(x)).  No CTA barrier until the MMA.
    {
      float* ap[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = max(cellk[k], 0);
        ap[k] = (float*)(smem + F_AHI + (c >> 5) * 4096 + kmaj_off(slot, c & 31));
      }
#pragma unroll
      for (int r = 0; r < CP; ++r) {
        if (p == r) {
          float t[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) t[k] = cellk[k] >= 0 ? *ap[k] : 0.f;
#pragma unroll
          for (int k = 0; k < 4; ++k) if (cellk[k] >= 0) *ap[k] = t[k] + wk[k];
        }
        __syncwarp();
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {          // rows 4 warp .. 4 warp + 3: 4 K-blocks x 4 rows x 8 chunks = 128 float4
        const int i = lane + 32 * j, kb = i >> 5, rr = (i >> 3) & 3, ck = i & 7;
        const int o = kb * 4096 + (warp * 4 + rr) * 128 + ck * 16;
        const float4 x = *(const float4*)(smem + F_AHI + o);
        *(float4*)(smem + F_ALO + o) = make_float4(tf32_lo(x.x), tf32_lo(x.y), tf32_lo(x.z), tf32_lo(x.w));
      }
    }
    c_fence_async();
    __syncthreads();
    if (warp == 0 && c_elect_one()) {
      c_mbar_wait(tma_bar, l & 1);
      c_fence_after();
      const uint32_t base = c_smem_u32(smem);
      const uint64_t vhi = c_desc_mnmajor(base + F_VHI, 16384), vlo = c_desc_mnmajor(base + F_VLO, 16384);
      constexpr uint32_t idf = c_idesc_tf32(128, 32, true, false);
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        const uint64_t ahi = c_desc_kmajor(base + F_AHI + kb * 4096), alo = c_desc_kmajor(base + F_ALO + kb * 4096);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t ks = 64u * (uint32_t)(kb * 4 + k);       // 8 cells = two 512-byte atoms per k-step
          c_mma_tf32(tmem_base, vhi + ks, ahi + 2 * k, idf, (l | kb | k) != 0);
          c_mma_tf32(tmem_base, vlo + ks, ahi + 2 * k, idf, 1u);
          c_mma_tf32(tmem_base, vhi + ks, alo + 2 * k, idf, 1u);
        }
      }
      c_commit(mma_bar);
    }
    __syncwarp();
    // corners outside the window: fp32 into the warp's own rows of s_extra, overlapped with the MMAs
    stray_corners<true>(omask, posk, wk, dk, g.q, slot, vb, nullptr, nullptr, s_extra, rowpitch, lane);
    c_mbar_wait(mma_bar, l & 1);
    c_fence_after();
    if (l + 1 < CL) {      // operands are dead: the warp clears its own rows of A for the next level
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = lane + 32 * j, kb = i >> 5, rr = (i >> 3) & 3, ck = i & 7;
        *(float4*)(smem + F_AHI + kb * 4096 + (warp * 4 + rr) * 128 + ck * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncwarp();
    }
  }
  __syncthreads();         // every warp's stray rows are final
  // ---- epilogue: lanes 0..63 of the accumulator are the channels; warps (qd, hf) read 16 query columns each --------
  const int qd = warp & 3, hf = warp >> 2;
  if (qd < 2) {
    uint32_t v[16];
    c_ld16(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(hf * 16), v);
    const int ch = qd * 32 + lane;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int sl = hf * 16 + j, q = s_q[sl];
      if (q >= 0) out[((int64_t)b * Q + q) * rowpitch + h * CHD + ch] = __uint_as_float(v[j]) + s_extra[sl * CHD + ch];
    }
  }
  c_fence_before();
  __syncthreads();
  if (warp == 0) {
    c_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(32) : "memory");
  }
}

// ---- host side --------------------------------------------------------------------------------------------------
static int fill_tc_shapes(const int* hw, int L, int S, TcShapes& sh) {
  if (L != CL) return GED_ERR_SHAPE;
  int start = 0;
  for (int l = 0; l < CL; ++l) {
    sh.h[l] = hw[2 * l]; sh.w[l] = hw[2 * l + 1]; sh.start[l] = start;
    if (sh.h[l] <= 0 || sh.w[l] <= 0 || sh.h[l] > 16384 || sh.w[l] > 16384) return GED_ERR_SHAPE;
    start += sh.h[l] * sh.w[l];
  }
  return start == S ? GED_OK : GED_ERR_SHAPE;
}

// 4-D map {channel, x, y, batch} of one level of value (B, S, nH*64); box {32, CWX, CWY, 1}; out-of-range cells read 0
static int make_level_maps(TcMaps& maps, const float* value, const TcShapes& sh, int B, int S, int nH, CUtensorMapSwizzle swz) {
  static PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
  if (!enc) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
      return GED_ERR_LAUNCH;
    enc = (PFN_cuTensorMapEncodeTiled_v12000)ptr;
  }
  const cuuint64_t pitch = (cuuint64_t)nH * CHD * 4;
  for (int l = 0; l < CL; ++l) {
    cuuint64_t dims[4] = {(cuuint64_t)nH * CHD, (cuuint64_t)sh.w[l], (cuuint64_t)sh.h[l], (cuuint64_t)B};
    cuuint64_t strides[3] = {pitch, pitch * (cuuint64_t)sh.w[l], pitch * (cuuint64_t)S};
    cuuint32_t box[4] = {32, CWX, CWY, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (enc(&maps.m[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)(value + (int64_t)sh.start[l] * nH * CHD), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return GED_ERR_ARG;
  }
  return GED_OK;
}

}  // namespace ged
using namespace ged;

// Tensor-core forward: same contract as ged_msda_tile_fwd; 3xTF32 products with fp32 accumulation (fp32-accurate).
// value_lo: workspace of the size of value (B*S*nH*64 floats); it receives value - tf32(value).
GED_API int ged_msda_tc_fwd(const float* value, float* value_lo, const float* ref, int ref_batch, const float* off,
                            const float* logit, const int* order, float* out, const int* level_hw, int num_levels, int B,
                            int S, int Q, int nH, int head_dim, int num_points, cudaStream_t stream) {
  if (!value || !value_lo || !ref || !off || !logit || !order || !out || !level_hw) return GED_ERR_ARG;
  if (head_dim != CHD || num_points != CP || (ref_batch != 1 && ref_batch != B)) return GED_ERR_SHAPE;
  if (!aligned16(value) || !aligned16(value_lo)) return GED_ERR_ALIGN;
  TcShapes sh;
  if (int e = fill_tc_shapes(level_hw, num_levels, S, sh)) return e;
  TcMaps maps, maps_lo;
  if (int e = make_level_maps(maps, value, sh, B, S, nH, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return e;
  if (int e = make_level_maps(maps_lo, value_lo, sh, B, S, nH, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return e;
  const int64_t n4 = (int64_t)B * S * nH * CHD / 4;
  const int64_t split_blocks = (n4 + 255) / 256 < 148 * 16 ? (n4 + 255) / 256 : 148 * 16;
  split_lo_kernel<<<(int)split_blocks, 256, 0, stream>>>((const float4*)value, (float4*)value_lo, n4);
  if (cudaFuncSetAttribute(msda_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F_TOTAL) != cudaSuccess) return GED_ERR_LAUNCH;
  msda_tc_fwd_kernel<<<dim3(cdiv(Q, CQ), nH, B), CTHREADS, F_TOTAL, stream>>>(maps, maps_lo, value, ref, off, logit, order, out, sh, B, S,
                                                                           Q, nH, ref_batch == 1 ? 0 : Q * 2);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

// Tensor-core backward: same contract as ged_msda_tile_bwd.  The products feeding g_value / g_off / g_logit / g_ref run
// in one-pass TF32 (tcgen05.mma kind::tf32, fp32 accumulate) like the other backward GEMMs of the path;
// ged_msda_tile_bwd is the fp32 version.
GED_API int ged_msda_tc_bwd(const float* value, const float* ref, int ref_batch, const float* off, const float* logit,
                            const int* order, const float* g_out, float* g_value, float* g_ref, float* g_off,
                            float* g_logit, const int* level_hw, int num_levels, int B, int S, int Q, int nH, int head_dim,
                            int num_points, cudaStream_t stream) {
  if (!value || !ref || !off || !logit || !order || !g_out || !g_value || !g_off || !g_logit || !level_hw) return GED_ERR_ARG;
  if (head_dim != CHD || num_points != CP || (ref_batch != 1 && ref_batch != B)) return GED_ERR_SHAPE;
  if (!aligned16(value) || !aligned16(g_value) || !aligned16(g_out)) return GED_ERR_ALIGN;
  TcShapes sh;
  if (int e = fill_tc_shapes(level_hw, num_levels, S, sh)) return e;
  TcMaps maps;
  if (int e = make_level_maps(maps, value, sh, B, S, nH, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  if (cudaFuncSetAttribute(msda_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, B_TOTAL) != cudaSuccess) return GED_ERR_LAUNCH;
  msda_tc_bwd_kernel<<<dim3(cdiv(Q, CQ), nH, B), CTHREADS, B_TOTAL, stream>>>(maps, value, ref, off, logit, order, g_out, g_value,
                                                                           g_ref, g_off, g_logit, sh, B, S, Q, nH,
                                                                           ref_batch == 1 ? 0 : Q * 2);
  GED_CHECK_LAUNCH();
  return GED_OK;
}
