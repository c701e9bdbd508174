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
#include "common.cuh"

namespace ged {

constexpr int CL = 4, CP = 8, CHD = 64;
constexpr int CQ = 32;                       // queries per tile
constexpr int CWARPS = 8, CTHREADS = CWARPS * 32;
constexpr int CWX = 11, CWY = 11, CCELLS = CWX * CWY;   // 121 cells, padded to the 128 TMEM lanes
constexpr int DPITCH = 36;                   // floats per cell row of the corner-dot table (conflict-free float4 stores)

struct TcShapes {
  int h[CL], w[CL], start[CL];
};

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
__device__ __forceinline__ void c_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void c_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void c_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void c_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(c_smem_u32(bar)) : "memory");
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

// =================================================================================================================
// backward
// =================================================================================================================
// shared memory (bytes from the 1024-aligned base)
constexpr int B_V = 0;                 // V_l: 2 K-blocks [128 cells][32 ch] swizzled (A of D^T)      32 KB; reused as the
                                       //      corner-dot table [128][DPITCH] once the MMAs have retired
constexpr int B_AT = 32768;            // A_l^T [128 cells][32 queries] swizzled (A of dV)             16 KB
constexpr int B_GT = 49152;            // G^T [64 ch][32 queries] swizzled (B of dV)                    8 KB
constexpr int B_GQ = 57344;            // G: 2 K-blocks [32 queries][32 ch] swizzled (B of D^T)         8 KB
constexpr int B_MISC = 65536;          // mbarrier, TMEM base, window sums / origins
constexpr int B_TOTAL = B_MISC + 128 + 1024;

__global__ void __launch_bounds__(CTHREADS, 3) msda_tc_bwd_kernel(
    const float* __restrict__ value, const float* __restrict__ ref, const float* __restrict__ off,
    const float* __restrict__ logit, const int* __restrict__ order, const float* __restrict__ g_out,
    float* __restrict__ g_value, float* __restrict__ g_ref, float* __restrict__ g_off, float* __restrict__ g_logit,
    TcShapes sh, int B, int S, int Q, int nH, int ref_bstride) {
  extern __shared__ uint8_t c_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)c_smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = (uint64_t*)(smem + B_MISC);
  uint32_t* tmem_ptr = (uint32_t*)(smem + B_MISC + 8);
  int* s_sum = (int*)(smem + B_MISC + 16);      // [12]
  int* s_org = s_sum + 12;                      // [8]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t0 = blockIdx.x * CQ, h = blockIdx.y, b = blockIdx.z;
  const int slot = warp * 4 + (lane >> 3), p = lane & 7;
  const int rowpitch = nH * CHD;

  if (tid == 0) { c_mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(c_smem_u32(tmem_ptr)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid < 12) s_sum[tid] = 0;
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
  float rgw[CL], rgx[CL], rgy[CL];
  const int qd = warp & 3, hf = warp >> 2;       // TMEM lane quadrant of this warp, column half it reads
  const int ecell = qd * 32 + lane;              // the window cell (TMEM lane) this thread reads back
  const int ecy = ecell / CWX, ecx = ecell - ecy * CWX;

#pragma unroll
  for (int l = 0; l < CL; ++l) {
    const int W = sh.w[l], H = sh.h[l], start = sh.start[l], wx0 = s_org[l * 2], wy0 = s_org[l * 2 + 1];
    // ---- stage V_l (zero outside the map / beyond the 121 cells) and clear A_l^T --------------------------------
    for (int idx = tid; idx < 128 * 16; idx += CTHREADS) {
      const int cell = idx >> 4, part = idx & 15;
      const int cy = cell / CWX, cx = cell - cy * CWX;
      const int x = wx0 + cx, y = wy0 + cy;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (cell < CCELLS && x < W && y < H) v = __ldg((const float4*)(vb + (int64_t)(start + y * W + x) * rowpitch + part * 4));
      *(float4*)(smem + B_V + (part >> 3) * 16384 + cell * 128 + (((part & 7) ^ (cell & 7)) << 4)) = v;
    }
    for (int idx = tid; idx < 16384 / 16; idx += CTHREADS) *(float4*)(smem + B_AT + idx * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    // ---- build A_l^T: every lane adds the four corner weights of its point --------------------------------------
    const float xf = floorf(g.px[l]), yf = floorf(g.py[l]);
    const int x0 = (int)xf, y0 = (int)yf;
    const float lx = g.px[l] - xf, ly = g.py[l] - yf, a = g.aw[l];
    int cellk[4];
    float dk[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int x = x0 + (k & 1), y = y0 + (k >> 1);
      const bool valid = g.q >= 0 && x >= 0 && x < W && y >= 0 && y < H;
      const int cx = x - wx0, cy = y - wy0;
      const bool inwin = valid && (unsigned)cx < (unsigned)CWX && (unsigned)cy < (unsigned)CWY;
      const float wgt = a * ((k & 1) ? lx : 1.f - lx) * ((k >> 1) ? ly : 1.f - ly);
      cellk[k] = inwin ? cy * CWX + cx : -1;
      dk[k] = 0.f;
      if (inwin) atomicAdd((float*)(smem + B_AT + kmaj_off(cellk[k], slot)), wgt);
      // valid corners outside the window: fp32, warp-cooperative (2 channels per lane), straight to / from L2
      unsigned m = __ballot_sync(0xffffffffu, valid && !inwin);
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const int spos = __shfl_sync(0xffffffffu, start + y * W + x, src);
        const float sw = __shfl_sync(0xffffffffu, wgt, src);
        const int sq = __shfl_sync(0xffffffffu, g.q, src);
        const float2 g2 = __ldg((const float2*)(g_out + ((int64_t)b * Q + sq) * rowpitch + h * CHD) + lane);
        const float2 v2 = __ldg((const float2*)(vb + (int64_t)spos * rowpitch) + lane);
        atomicAdd((float2*)(gvb + (int64_t)spos * rowpitch) + lane, make_float2(sw * g2.x, sw * g2.y));
        const float d = warp_sum(g2.x * v2.x + g2.y * v2.y);
        if (lane == src) dk[k] = d;
      }
    }
    c_fence_async();
    __syncthreads();
    // ---- tensor cores: dV_l = A_l^T . G (cols 0..63), D_l^T = V_l . G^T (cols 64..95) ------------------------------
    if (tid == 0) {
      c_fence_after();
      const uint32_t base = c_smem_u32(smem);
      const uint64_t a_at = c_desc_kmajor(base + B_AT), b_gt = c_desc_kmajor(base + B_GT);
      constexpr uint32_t idv = c_idesc_tf32(128, 64, false, false), idd = c_idesc_tf32(128, 32, false, false);
#pragma unroll
      for (int k = 0; k < 4; ++k) c_mma_tf32(tmem_base, a_at + 2 * k, b_gt + 2 * k, idv, k != 0);
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        const uint64_t a_v = c_desc_kmajor(base + B_V + kb * 16384), b_gq = c_desc_kmajor(base + B_GQ + kb * 4096);
#pragma unroll
        for (int k = 0; k < 4; ++k) c_mma_tf32(tmem_base + 64, a_v + 2 * k, b_gq + 2 * k, idd, (kb | k) != 0);
      }
      c_commit(bar);
    }
    c_mbar_wait(bar, l & 1);
    c_fence_after();
    // ---- epilogue: this thread's cell row.  dV half -> red.global (skipped when the cell was not touched), corner
    //      dots -> shared-memory table over the (now dead) V tiles ------------------------------------------------------
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
      float* drow = (float*)(smem + B_V) + ecell * DPITCH + hf * 16;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *(float4*)(drow + 4 * j) = make_float4(__uint_as_float(d[4 * j]), __uint_as_float(d[4 * j + 1]),
                                               __uint_as_float(d[4 * j + 2]), __uint_as_float(d[4 * j + 3]));
    }
    c_fence_before();
    __syncthreads();
    // ---- per point: d/d weight, d/d x_pix, d/d y_pix from the four corner dots -------------------------------------
    {
      const float* dtab = (const float*)(smem + B_V);
#pragma unroll
      for (int k = 0; k < 4; ++k) if (cellk[k] >= 0) dk[k] = dtab[cellk[k] * DPITCH + slot];
      rgw[l] = (1.f - ly) * ((1.f - lx) * dk[0] + lx * dk[1]) + ly * ((1.f - lx) * dk[2] + lx * dk[3]);
      rgx[l] = a * ((1.f - ly) * (dk[1] - dk[0]) + ly * (dk[3] - dk[2]));
      rgy[l] = a * ((1.f - lx) * (dk[2] - dk[0]) + lx * (dk[3] - dk[1]));
    }
    __syncthreads();        // the table is overwritten by the next level's V tiles
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

}  // namespace ged
using namespace ged;

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
  if (cudaFuncSetAttribute(msda_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, B_TOTAL) != cudaSuccess) return GED_ERR_LAUNCH;
  msda_tc_bwd_kernel<<<dim3(cdiv(Q, CQ), nH, B), CTHREADS, B_TOTAL, stream>>>(value, ref, off, logit, order, g_out, g_value, g_ref,
                                                                           g_off, g_logit, sh, B, S, Q, nH,
                                                                           ref_batch == 1 ? 0 : Q * 2);
  GED_CHECK_LAUNCH();
  return GED_OK;
}
