// Swin (shifted-)window attention core on the 5th-generation tensor cores, sm_100a - forward.
// a7 + a8: depthformer_swin.py:285-360 (pad to x7, roll(-3,-3), 9-region mask of -100, partition, reverse, un-roll,
// crop) and :184-224 (q*scale @ k^T + rel-pos-bias (+mask), softmax, @ v); same contract as ged_winattn_fwd.
//
// A work item is a DUO: two consecutive (window, head) pairs.  Their 49-token Q / K / V tiles (head dim 32 = one
// 128-byte swizzle row) are gathered by coordinate from the image-ordered qkv matrix straight into UMMA operand tiles
// (padding, cyclic shift, partition are index arithmetic), split into hi + lo (3xTF32: fp32-accurate) while staging:
//     S_p [128 x 64] = [Q_A ; Q_B] . K_p^T        p = A, B   (rows of the other pair are don't-care)    tcgen05.mma, TMEM
//     P    = softmax(S + rel-pos bias + shift mask)            one accumulator row per thread, tcgen05.ld -> registers
//     O_p [128 x 32] = [P_A ; P_B] . V_p                      P written back as a K-major operand, V MN-major
// and the context rows leave from TMEM by tcgen05.ld as 128-byte stores.  Rows / keys 49..63 are padding: zero
// probabilities, never stored.
#include "common.cuh"

namespace ged {

constexpr int TW_WS = 7, TW_N = 49, TW_HD = 32;
constexpr int TW_THREADS = 256;

struct TwGeom {
  int H, W, Hp, Wp, nWx, nWin, shift;
};

// ---- PTX wrappers (forms of gemm_tcgen05.cu) ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t w_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void w_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(w_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void w_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WDONE_%=;\n\t"
      "bra WWAIT_%=;\n\t"
      "WDONE_%=:\n\t}"
      :: "r"(w_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void w_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void w_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void w_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void w_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(w_smem_u32(bar)) : "memory");
}
// one lane of a CONVERGED warp: tcgen05.mma issued under it needs no per-instruction elect-and-broadcast loop (see
// gemm_tcgen05.cu::elect_one)
__device__ __forceinline__ bool w_elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void w_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void w_ld32(uint32_t taddr, uint32_t* v) {
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
__device__ __forceinline__ uint64_t w_desc_kmajor(uint32_t smem_addr) {      // rows of 32 floats, 128B swizzle, 8-row atoms
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t w_desc_mnmajor(uint32_t smem_addr, uint32_t panel_bytes) {   // Layout_MN_SW128_32B_Atom
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((panel_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(512 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
__host__ __device__ constexpr uint32_t w_idesc_tf32(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// element (row r, float c < 32) of a K-major 128B-swizzled tile; 16-byte chunk (r, ck) of an MN-major tile whose rows are
// the contraction index (Swizzle<2,5,2>: 32-byte chunks XOR (row & 3))
__device__ __forceinline__ uint32_t w_kmaj_chunk(int r, int ck) { return (uint32_t)(r * 128 + ((ck ^ (r & 7)) << 4)); }
__device__ __forceinline__ uint32_t w_mnmaj_chunk(int r, int ck) { return (uint32_t)(r * 128 + ((((ck >> 1) ^ (r & 3)) << 5) | ((ck & 1) << 4))); }
__device__ __forceinline__ float w_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ float4 w_lo4(const float4& x) { return make_float4(w_lo(x.x), w_lo(x.y), w_lo(x.z), w_lo(x.w)); }

// token n of window (wy,wx) -> image token index, or -1 for a padding token; also its shift-mask label
__device__ __forceinline__ int tw_token(const TwGeom& g, int wy, int wx, int n, int& label) {
  const int ty = n / TW_WS, tx = n - ty * TW_WS;
  const int hs = wy * TW_WS + ty, ws = wx * TW_WS + tx;
  label = (hs < g.Hp - TW_WS ? 0 : (hs < g.Hp - g.shift ? 1 : 2)) * 3 + (ws < g.Wp - TW_WS ? 0 : (ws < g.Wp - g.shift ? 1 : 2));
  int h = hs + g.shift, w = ws + g.shift;
  if (h >= g.Hp) h -= g.Hp;
  if (w >= g.Wp) w -= g.Wp;
  return (h < g.H && w < g.W) ? h * g.W + w : -1;
}

// shared memory (bytes from the 1024-aligned base)
constexpr int W_QHI = 0;            // [128 rows: pair A 0..63, pair B 64..127][32]  K-major            16 KB
constexpr int W_QLO = 16384;
constexpr int W_KHI = 32768;        // 2 x [64 keys][32] K-major                                       16 KB
constexpr int W_KLO = 49152;
constexpr int W_PHI = 0;            // P [128 rows][64 keys] = 2 K-blocks x 16 KB, over Q / K once S has been read   32 KB
constexpr int W_PLO = 32768;
constexpr int W_VHI = 65536;        // 2 x [64 keys][32 d] MN-major                                     16 KB
constexpr int W_VLO = 81920;
constexpr int W_MISC = 98304;       // barriers, TMEM base, token / label tables, bias tables
constexpr int W_TOTAL = W_MISC + 6144 + 1024;

struct TwMeta {                     // per work item (double-buffered: producers fill the next one while consumers read this)
  int tok[2][64];                   // image token of each row, -1 = zero-padded token, -2 = not a token (rows 49..63 / no pair)
  int lab[2][64];
  float tab[2][176];                // relative-position bias of the pair's head, by (dy + 6) * 13 + (dx + 6)
  int pair_b[2], pair_head[2], pair_valid[2];
};
struct TwMisc {
  uint64_t bar_s, bar_o;
  uint32_t tmem;
  TwMeta meta[2];
};
constexpr int TW_ITEMS = 2 * TW_N * 3 * 8;                    // 16-byte chunks of Q, K, V of both pairs
constexpr int TW_PROD = 128;                                  // producer threads (warps 4..7)
constexpr int TW_ITERS = (TW_ITEMS + TW_PROD - 1) / TW_PROD;  // 19 chunks per producer thread

// producers: tables of work item `duo`
__device__ __forceinline__ void tw_prepare(TwMeta& m, const TwGeom& g, const float* __restrict__ table, int duo, int num_pairs,
                                           int nH, int ptid) {
  {
    const int p = ptid >> 6, n = ptid & 63;
    const int pair = duo * 2 + p;
    const bool valid = pair < num_pairs;
    const int head = valid ? pair % nH : 0, wlin = valid ? pair / nH : 0;
    const int win = wlin % g.nWin, b = wlin / g.nWin;
    int lab = 0, tok = -1;
    if (valid && n < TW_N) tok = tw_token(g, win / g.nWx, win % g.nWx, n, lab);
    m.tok[p][n] = (valid && n < TW_N) ? tok : -2;
    m.lab[p][n] = lab;
    if (n == 0) { m.pair_b[p] = b; m.pair_head[p] = head; m.pair_valid[p] = valid; }
  }
  for (int i = ptid; i < 2 * 169; i += TW_PROD) {
    const int p = i / 169, e = i - p * 169, pair = duo * 2 + p;
    m.tab[p][e] = pair < num_pairs ? __ldg(table + (int64_t)e * nH + pair % nH) : 0.f;
  }
}

// producers: issue every global load of the work item (Q, K, V rows of both pairs) into registers
__device__ __forceinline__ void tw_load(float4 (&v)[TW_ITERS], const TwMeta& m, const float* __restrict__ qkv,
                                        const float* __restrict__ bias, int64_t L, int C, int ptid) {
#pragma unroll
  for (int it = 0; it < TW_ITERS; ++it) {
    const int i = ptid + it * TW_PROD;
    v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < TW_ITEMS) {
      const int ck = i & 7, which = (i >> 3) % 3, r = (i >> 3) / 3;
      const int p = r / TW_N, n = r - p * TW_N;
      if (m.pair_valid[p]) {
        const int tok = m.tok[p][n], col = which * C + m.pair_head[p] * TW_HD + ck * 4;
        if (tok >= 0) v[it] = __ldg((const float4*)(qkv + ((int64_t)m.pair_b[p] * L + tok) * 3 * C + col));
        else if (bias) v[it] = __ldg((const float4*)(bias + col));       // zero-padded token: q = k = v = bias
      }
    }
  }
}

// producers: registers -> swizzled operand tiles, hi + lo
__device__ __forceinline__ void tw_store(const float4 (&v)[TW_ITERS], uint8_t* smem, float scale, int ptid) {
#pragma unroll
  for (int it = 0; it < TW_ITERS; ++it) {
    const int i = ptid + it * TW_PROD;
    if (i >= TW_ITEMS) continue;
    const int ck = i & 7, which = (i >> 3) % 3, r = (i >> 3) / 3;
    const int p = r / TW_N, n = r - p * TW_N;
    float4 x = v[it];
    if (which == 0) {
      x.x *= scale; x.y *= scale; x.z *= scale; x.w *= scale;
      const uint32_t o = w_kmaj_chunk(p * 64 + n, ck);
      *(float4*)(smem + W_QHI + o) = x; *(float4*)(smem + W_QLO + o) = w_lo4(x);
    } else if (which == 1) {
      const uint32_t o = p * 8192 + w_kmaj_chunk(n, ck);
      *(float4*)(smem + W_KHI + o) = x; *(float4*)(smem + W_KLO + o) = w_lo4(x);
    } else {
      const uint32_t o = p * 8192 + w_mnmaj_chunk(n, ck);
      *(float4*)(smem + W_VHI + o) = x; *(float4*)(smem + W_VLO + o) = w_lo4(x);
    }
  }
}
// named barriers (non-.aligned form: producer and consumer warps reach them from different program points)
__device__ __forceinline__ void tw_producer_sync() { __syncwarp(); asm volatile("barrier.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void tw_cta_sync() { __syncwarp(); asm volatile("barrier.sync 2, 256;" ::: "memory"); }

// Warp-specialised: warps 4..7 are PRODUCERS (tables, global loads of the NEXT work item into registers while the current
// one is in the softmax, then registers -> operand tiles), warps 0..3 CONSUMERS (one accumulator row per thread: softmax
// from TMEM, P back to shared memory, context rows out of TMEM); thread 0 issues the MMAs.
__global__ void __launch_bounds__(TW_THREADS, 2) winattn_tc_fwd_kernel(
    const float* __restrict__ qkv, const float* __restrict__ bias, const float* __restrict__ table, float* __restrict__ ctx,
    TwGeom g, int B, int C, int nH, float scale, int num_pairs) {
  extern __shared__ uint8_t w_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)w_smem_raw + 1023) & ~(uintptr_t)1023);
  TwMisc& ms = *reinterpret_cast<TwMisc*>(smem + W_MISC);
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool producer = warp >= 4;
  const int ptid = tid - 128;
  const int64_t L = (int64_t)g.H * g.W;

  if (tid == 0) {
    w_mbar_init(&ms.bar_s, 1); w_mbar_init(&ms.bar_o, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(w_smem_u32(&ms.tmem)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // every operand byte starts finite: padding rows / keys are never written again and only multiply zero probabilities
  for (int i = tid; i < W_MISC / 16; i += TW_THREADS) *(float4*)(smem + i * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
  w_fence_before();
  __syncthreads();
  w_fence_after();
  const uint32_t tmem = ms.tmem;
  const uint32_t sbase = w_smem_u32(smem);

  if (producer) {
    // ================= producers: separate loop so that their 19 x float4 of in-flight loads and the consumers' 64-float
    // accumulator rows never share a register allocation; the three CTA barriers per work item pair up with the
    // consumers' ones (bar.sync 2) =====================================================================================
    float4 v[TW_ITERS];
    tw_prepare(ms.meta[0], g, table, blockIdx.x, num_pairs, nH, ptid);
    tw_producer_sync();
    tw_load(v, ms.meta[0], qkv, bias, L, C, ptid);
    int buf = 0;
    for (int duo = blockIdx.x; duo * 2 < num_pairs; duo += gridDim.x, buf ^= 1) {
      tw_store(v, smem, scale, ptid);
      w_fence_async();
      tw_cta_sync();                                   // (A) operand tiles complete
      const int nxt = duo + gridDim.x;
      if (nxt * 2 < num_pairs) {                        // next work item: tables, then every global load in flight
        tw_prepare(ms.meta[buf ^ 1], g, table, nxt, num_pairs, nH, ptid);
        tw_producer_sync();
        tw_load(v, ms.meta[buf ^ 1], qkv, bias, L, C, ptid);
      }
      tw_cta_sync();                                   // (B) P written
      tw_cta_sync();                                   // (C) O read: tiles and TMEM are free
    }
  } else {
    uint32_t phase = 0;
    int buf = 0;
    const int row = tid, rp = row >> 6, ri = row & 63;
    for (int duo = blockIdx.x; duo * 2 < num_pairs; duo += gridDim.x, buf ^= 1) {
      const TwMeta& m = ms.meta[buf];
      tw_cta_sync();                                   // (A)
      // ---- S_p = [Q_A ; Q_B] . K_p^T  (3xTF32) ----------------------------------------------------------------------
      if (warp == 0 && w_elect_one()) {
        w_fence_after();
        constexpr uint32_t ids = w_idesc_tf32(128, 64, false, false);
        const uint64_t qhi = w_desc_kmajor(sbase + W_QHI), qlo = w_desc_kmajor(sbase + W_QLO);
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const uint64_t khi = w_desc_kmajor(sbase + W_KHI + p * 8192), klo = w_desc_kmajor(sbase + W_KLO + p * 8192);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            w_mma_tf32(tmem + p * 64, qhi + 2 * k, khi + 2 * k, ids, k != 0);
            w_mma_tf32(tmem + p * 64, qlo + 2 * k, khi + 2 * k, ids, 1u);
            w_mma_tf32(tmem + p * 64, qhi + 2 * k, klo + 2 * k, ids, 1u);
          }
        }
        w_commit(&ms.bar_s);
      }
      __syncwarp();
      // ---- softmax: thread = accumulator row (warps 0..3 own TMEM lane quadrants 0..3) ------------------------------
      w_mbar_wait(&ms.bar_s, phase);
      w_fence_after();
      {
        float s[64];
        w_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(rp * 64), (uint32_t*)s);
        w_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(rp * 64 + 32), (uint32_t*)(s + 32));
        const bool live = ri < TW_N && m.pair_valid[rp];
        float inv = 0.f;
        if (live) {
          const int yi = ri / TW_WS, xi = ri - yi * TW_WS, li = m.lab[rp][ri];
          const float* tb = m.tab[rp];
          const bool masked = g.shift > 0;
          float mx = -3.0e38f;
#pragma unroll
          for (int j = 0; j < TW_N; ++j) {
            const int yj = j / TW_WS, xj = j - yj * TW_WS;
            float x = s[j] + tb[(yi - yj + 6) * 13 + (xi - xj + 6)];
            if (masked && m.lab[rp][j] != li) x += -100.0f;
            s[j] = x;
            mx = fmaxf(mx, x);
          }
          float sum = 0.f;
#pragma unroll
          for (int j = 0; j < TW_N; ++j) { s[j] = exp2f((s[j] - mx) * 1.4426950408889634f); sum += s[j]; }
          inv = 1.f / sum;
        }
        // P row -> K-major operand tiles (2 K-blocks of 32 keys), hi + lo; padding rows / keys are zero
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          float4 x;
          x.x = (live && 4 * c + 0 < TW_N) ? s[4 * c + 0] * inv : 0.f;
          x.y = (live && 4 * c + 1 < TW_N) ? s[4 * c + 1] * inv : 0.f;
          x.z = (live && 4 * c + 2 < TW_N) ? s[4 * c + 2] * inv : 0.f;
          x.w = (live && 4 * c + 3 < TW_N) ? s[4 * c + 3] * inv : 0.f;
          const uint32_t o = (c >> 3) * 16384 + w_kmaj_chunk(row, c & 7);
          *(float4*)(smem + W_PHI + o) = x; *(float4*)(smem + W_PLO + o) = w_lo4(x);
        }
      }
      w_fence_async();
      w_fence_before();
      tw_cta_sync();                                   // (B)
      // ---- O_p = [P_A ; P_B] . V_p  (3xTF32) ------------------------------------------------------------------------
      if (warp == 0 && w_elect_one()) {
        w_fence_after();
        constexpr uint32_t ido = w_idesc_tf32(128, 32, false, true);
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const uint64_t vhi = w_desc_mnmajor(sbase + W_VHI + p * 8192, 8192), vlo = w_desc_mnmajor(sbase + W_VLO + p * 8192, 8192);
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t phi = w_desc_kmajor(sbase + W_PHI + kb * 16384), plo = w_desc_kmajor(sbase + W_PLO + kb * 16384);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t ks = 64u * (uint32_t)(kb * 4 + k);       // 8 keys = two 512-byte atoms of the MN-major V tile
              w_mma_tf32(tmem + 128 + p * 32, phi + 2 * k, vhi + ks, ido, (kb | k) != 0);
              w_mma_tf32(tmem + 128 + p * 32, plo + 2 * k, vhi + ks, ido, 1u);
              w_mma_tf32(tmem + 128 + p * 32, phi + 2 * k, vlo + ks, ido, 1u);
            }
          }
        }
        w_commit(&ms.bar_o);
      }
      __syncwarp();
      w_mbar_wait(&ms.bar_o, phase);
      w_fence_after();
      {
        float o[32];
        w_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(128 + rp * 32), (uint32_t*)o);
        const int tok = m.tok[rp][ri];
        if (tok >= 0) {                                            // padded rows are cropped away (:354-355)
          float4* dst = (float4*)(ctx + ((int64_t)m.pair_b[rp] * L + tok) * C + m.pair_head[rp] * TW_HD);
#pragma unroll
          for (int c = 0; c < 8; ++c) dst[c] = make_float4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
        }
      }
      phase ^= 1;
      w_fence_before();
      tw_cta_sync();                                   // (C) TMEM and the operand tiles are free for the next work item
      w_fence_after();
    }
  }
  w_fence_before();
  __syncthreads();
  if (warp == 0) {
    w_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256) : "memory");
  }
}


// =====================================================================================================================
// Backward on tcgen05 (one pass TF32, like every other backward GEMM; GEDEPTH_BWD_GEMM_PASSES=3 keeps the fp32 SIMT kernel).
// Same duo work items and stacked M = 128 accumulators as the forward.  Per pair p (rows of the other pair: don't-care):
//     S_p  [128 x 64] = [Q_A ; Q_B] . K_p^T          dP_p [128 x 64] = [dO_A ; dO_B] . V_p^T            (K-major operands)
//     P = softmax(S + bias + mask),  dS = P o (dP - rowsum(P o dP))        one accumulator row per thread, from TMEM
//     dV_p [128 x 32] = [P_A^T ; P_B^T] . dO_p       dK_p = [dS_A^T ; dS_B^T] . Q_p     (A MN-major: rows = query index)
//     dQ_p [128 x 32] = [dS_A ; dS_B] . K_p                                              (A K-major, B MN-major)
// Q, K, dO are staged twice (K-major for the products that contract over the head dimension, MN-major for the ones that
// contract over tokens); the thread that owns row (p, i) writes P and dS back as operand tiles.  TMEM: S, dP in columns
// 0..255; dV / dQ / dK re-use columns 64..255 once every row has been read.  220 KB of shared memory: one CTA per SM.
// =====================================================================================================================
constexpr int WB_QK = 0;            // [128][32]      K-major  (q * scale)
constexpr int WB_KK = 16384;        // 2 x [64][32]   K-major
constexpr int WB_VK = 32768;        // 2 x [64][32]   K-major
constexpr int WB_GK = 49152;        // [128][32]      K-major  (dO)
constexpr int WB_QM = 65536;        // 2 x [64][32]   MN-major (q * scale)
constexpr int WB_KM = 81920;        // 2 x [64][32]   MN-major
constexpr int WB_GM = 98304;        // 2 x [64][32]   MN-major (dO)
constexpr int WB_PM = 114688;       // P^T operand:  4 panels (pair, 32 keys) x [64 query rows][128 B], MN-major   32 KB
constexpr int WB_SM = 147456;       // dS^T operand, same layout                                                   32 KB
constexpr int WB_SK = 180224;       // dS operand: 2 K-blocks (32 keys) x [128 rows][32], K-major                   32 KB
constexpr int WB_MISC = 212992;
constexpr int WB_OPERANDS = WB_MISC;
constexpr int WB_TOTAL = WB_MISC + 8192 + 1024;

struct TbMisc {
  uint64_t bar_s, bar_o;
  uint32_t tmem;
  float gtab[2][176];               // rel-pos-bias gradient of the duo's two heads, by (dy + 6) * 13 + (dx + 6)
  TwMeta meta[2];
};
static_assert(sizeof(TbMisc) <= 8192, "misc area");
constexpr int TB_ITEMS = 2 * TW_N * 4 * 8;                    // 16-byte chunks of Q, K, V, dO of both pairs
constexpr int TB_ITERS = (TB_ITEMS + TW_PROD - 1) / TW_PROD;  // 25 chunks per producer thread

__device__ __forceinline__ void tb_load(float4 (&v)[TB_ITERS], const TwMeta& m, const float* __restrict__ qkv,
                                        const float* __restrict__ bias, const float* __restrict__ g_ctx, int64_t L, int C, int ptid) {
#pragma unroll
  for (int it = 0; it < TB_ITERS; ++it) {
    const int i = ptid + it * TW_PROD;
    v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < TB_ITEMS) {
      const int ck = i & 7, which = (i >> 3) & 3, r = i >> 5;
      const int p = r / TW_N, n = r - p * TW_N;
      if (m.pair_valid[p]) {
        const int tok = m.tok[p][n];
        if (which < 3) {
          const int col = which * C + m.pair_head[p] * TW_HD + ck * 4;
          if (tok >= 0) v[it] = __ldg((const float4*)(qkv + ((int64_t)m.pair_b[p] * L + tok) * 3 * C + col));
          else if (bias) v[it] = __ldg((const float4*)(bias + col));       // zero-padded token: q = k = v = bias
        } else if (tok >= 0) {                                             // cropped (padded) rows carry no gradient
          v[it] = __ldg((const float4*)(g_ctx + ((int64_t)m.pair_b[p] * L + tok) * C + m.pair_head[p] * TW_HD + ck * 4));
        }
      }
    }
  }
}

__device__ __forceinline__ void tb_store(const float4 (&v)[TB_ITERS], uint8_t* smem, float scale, int ptid) {
#pragma unroll
  for (int it = 0; it < TB_ITERS; ++it) {
    const int i = ptid + it * TW_PROD;
    if (i >= TB_ITEMS) continue;
    const int ck = i & 7, which = (i >> 3) & 3, r = i >> 5;
    const int p = r / TW_N, n = r - p * TW_N;
    float4 x = v[it];
    const uint32_t ok = w_kmaj_chunk(n, ck), om = p * 8192 + w_mnmaj_chunk(n, ck);
    if (which == 0) {
      x.x *= scale; x.y *= scale; x.z *= scale; x.w *= scale;
      *(float4*)(smem + WB_QK + w_kmaj_chunk(p * 64 + n, ck)) = x; *(float4*)(smem + WB_QM + om) = x;
    } else if (which == 1) {
      *(float4*)(smem + WB_KK + p * 8192 + ok) = x; *(float4*)(smem + WB_KM + om) = x;
    } else if (which == 2) {
      *(float4*)(smem + WB_VK + p * 8192 + ok) = x;
    } else {
      *(float4*)(smem + WB_GK + w_kmaj_chunk(p * 64 + n, ck)) = x; *(float4*)(smem + WB_GM + om) = x;
    }
  }
}

__global__ void __launch_bounds__(TW_THREADS, 1) winattn_tc_bwd_kernel(
    const float* __restrict__ qkv, const float* __restrict__ bias, const float* __restrict__ table,
    const float* __restrict__ g_ctx, float* __restrict__ g_qkv, float* __restrict__ g_bias, float* __restrict__ g_table,
    TwGeom g, int B, int C, int nH, float scale, int num_pairs) {
  extern __shared__ uint8_t w_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)w_smem_raw + 1023) & ~(uintptr_t)1023);
  TbMisc& ms = *reinterpret_cast<TbMisc*>(smem + WB_MISC);
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool producer = warp >= 4;
  const int ptid = tid - 128;
  const int64_t L = (int64_t)g.H * g.W;

  if (tid == 0) {
    w_mbar_init(&ms.bar_s, 1); w_mbar_init(&ms.bar_o, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(w_smem_u32(&ms.tmem)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // every operand byte starts finite: padding rows / keys are never written again and only meet zero probabilities
  for (int i = tid; i < WB_OPERANDS / 16; i += TW_THREADS) *(float4*)(smem + i * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = tid; i < 2 * 176; i += TW_THREADS) (&ms.gtab[0][0])[i] = 0.f;
  w_fence_before();
  __syncthreads();
  w_fence_after();
  const uint32_t tmem = ms.tmem;
  const uint32_t sbase = w_smem_u32(smem);

  if (producer) {
    float4 v[TB_ITERS];
    tw_prepare(ms.meta[0], g, table, blockIdx.x, num_pairs, nH, ptid);
    tw_producer_sync();
    tb_load(v, ms.meta[0], qkv, bias, g_ctx, L, C, ptid);
    int buf = 0;
    for (int duo = blockIdx.x; duo * 2 < num_pairs; duo += gridDim.x, buf ^= 1) {
      tb_store(v, smem, scale, ptid);
      w_fence_async();
      tw_cta_sync();                                   // (A) operand tiles complete
      const int nxt = duo + gridDim.x;
      if (nxt * 2 < num_pairs) {                        // next work item: tables, then every global load in flight
        tw_prepare(ms.meta[buf ^ 1], g, table, nxt, num_pairs, nH, ptid);
        tw_producer_sync();
        tb_load(v, ms.meta[buf ^ 1], qkv, bias, g_ctx, L, C, ptid);
      }
      tw_cta_sync();                                   // (B) P / dS written
      tw_cta_sync();                                   // (C) gradients read: tiles and TMEM are free
    }
  } else {
    uint32_t phase = 0;
    int buf = 0;
    const int row = tid, rp = row >> 6, ri = row & 63;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    for (int duo = blockIdx.x; duo * 2 < num_pairs; duo += gridDim.x, buf ^= 1) {
      const TwMeta& m = ms.meta[buf];
      tw_cta_sync();                                   // (A)
      if (warp == 0 && w_elect_one()) {
        w_fence_after();
        constexpr uint32_t ids = w_idesc_tf32(128, 64, false, false);
        const uint64_t qk = w_desc_kmajor(sbase + WB_QK), gk = w_desc_kmajor(sbase + WB_GK);
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const uint64_t kk = w_desc_kmajor(sbase + WB_KK + p * 8192), vk = w_desc_kmajor(sbase + WB_VK + p * 8192);
#pragma unroll
          for (int k = 0; k < 4; ++k) w_mma_tf32(tmem + p * 64, qk + 2 * k, kk + 2 * k, ids, k != 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) w_mma_tf32(tmem + 128 + p * 64, gk + 2 * k, vk + 2 * k, ids, k != 0);
        }
        w_commit(&ms.bar_s);
      }
      __syncwarp();
      w_mbar_wait(&ms.bar_s, phase);
      w_fence_after();
      {
        float s[64], dp[64];
        w_ld32(lane_base + (uint32_t)(rp * 64), (uint32_t*)s);
        w_ld32(lane_base + (uint32_t)(rp * 64 + 32), (uint32_t*)(s + 32));
        w_ld32(lane_base + (uint32_t)(128 + rp * 64), (uint32_t*)dp);
        w_ld32(lane_base + (uint32_t)(128 + rp * 64 + 32), (uint32_t*)(dp + 32));
        const bool live = ri < TW_N && m.pair_valid[rp];
        if (live) {
          const int yi = ri / TW_WS, xi = ri - yi * TW_WS, li = m.lab[rp][ri];
          const float* tb = m.tab[rp];
          const bool masked = g.shift > 0;
          float mx = -3.0e38f;
#pragma unroll
          for (int j = 0; j < TW_N; ++j) {
            const int yj = j / TW_WS, xj = j - yj * TW_WS;
            float x = s[j] + tb[(yi - yj + 6) * 13 + (xi - xj + 6)];
            if (masked && m.lab[rp][j] != li) x += -100.0f;
            s[j] = x;
            mx = fmaxf(mx, x);
          }
          float sum = 0.f;
#pragma unroll
          for (int j = 0; j < TW_N; ++j) { s[j] = exp2f((s[j] - mx) * 1.4426950408889634f); sum += s[j]; }
          const float inv = 1.f / sum;
          float dot = 0.f;
#pragma unroll
          for (int j = 0; j < TW_N; ++j) { s[j] *= inv; dot = fmaf(s[j], dp[j], dot); }
          float* gt = ms.gtab[rp];
#pragma unroll
          for (int j = 0; j < TW_N; ++j) {
            const int yj = j / TW_WS, xj = j - yj * TW_WS;
            dp[j] = s[j] * (dp[j] - dot);                                   // dS
            atomicAdd(gt + (yi - yj + 6) * 13 + (xi - xj + 6), dp[j]);
          }
          // P and dS rows -> operand tiles; keys 49..51 share the last written chunk and are zero, chunks beyond stay zero
#pragma unroll
          for (int c = 0; c < 13; ++c) {
            float4 pv, dv;
            pv.x = 4 * c + 0 < TW_N ? s[4 * c + 0] : 0.f; dv.x = 4 * c + 0 < TW_N ? dp[4 * c + 0] : 0.f;
            pv.y = 4 * c + 1 < TW_N ? s[4 * c + 1] : 0.f; dv.y = 4 * c + 1 < TW_N ? dp[4 * c + 1] : 0.f;
            pv.z = 4 * c + 2 < TW_N ? s[4 * c + 2] : 0.f; dv.z = 4 * c + 2 < TW_N ? dp[4 * c + 2] : 0.f;
            pv.w = 4 * c + 3 < TW_N ? s[4 * c + 3] : 0.f; dv.w = 4 * c + 3 < TW_N ? dp[4 * c + 3] : 0.f;
            const uint32_t om = (uint32_t)(rp * 2 + (c >> 3)) * 8192u + w_mnmaj_chunk(ri, c & 7);
            *(float4*)(smem + WB_PM + om) = pv;
            *(float4*)(smem + WB_SM + om) = dv;
            *(float4*)(smem + WB_SK + (c >> 3) * 16384 + w_kmaj_chunk(row, c & 7)) = dv;
          }
        }
      }
      w_fence_async();
      w_fence_before();
      tw_cta_sync();                                   // (B)
      if (warp == 0 && w_elect_one()) {
        w_fence_after();
        constexpr uint32_t id_mm = w_idesc_tf32(128, 32, true, true), id_km = w_idesc_tf32(128, 32, false, true);
        const uint64_t pm = w_desc_mnmajor(sbase + WB_PM, 8192), sm = w_desc_mnmajor(sbase + WB_SM, 8192);
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const uint64_t gm = w_desc_mnmajor(sbase + WB_GM + p * 8192, 8192), km = w_desc_mnmajor(sbase + WB_KM + p * 8192, 8192),
                         qm = w_desc_mnmajor(sbase + WB_QM + p * 8192, 8192);
#pragma unroll
          for (int k = 0; k < 8; ++k) w_mma_tf32(tmem + 64 + p * 32, pm + 64u * k, gm + 64u * k, id_mm, k != 0);       // dV
#pragma unroll
          for (int k = 0; k < 8; ++k)                                                                             // dQ
            w_mma_tf32(tmem + 128 + p * 32, w_desc_kmajor(sbase + WB_SK + (k >> 2) * 16384) + 2 * (k & 3), km + 64u * k, id_km, k != 0);
#pragma unroll
          for (int k = 0; k < 8; ++k) w_mma_tf32(tmem + 192 + p * 32, sm + 64u * k, qm + 64u * k, id_mm, k != 0);      // dK
        }
        w_commit(&ms.bar_o);
      }
      __syncwarp();
      // rel-pos-bias gradient of the two heads: every consumer passed (B), so the table is complete; flush and clear
      for (int i = tid; i < 2 * 169; i += 128) {
        const int p = i / 169, e = i - p * 169;
        const float vsum = ms.gtab[p][e];
        if (vsum != 0.f) { atomicAdd(g_table + (int64_t)e * nH + m.pair_head[p], vsum); ms.gtab[p][e] = 0.f; }
      }
      w_mbar_wait(&ms.bar_o, phase);
      w_fence_after();
      {
        const int tok = m.tok[rp][ri];
        const bool wr = tok != -2 && m.pair_valid[rp];
        const int col = m.pair_head[rp] * TW_HD;
        float* gp = tok >= 0 ? g_qkv + ((int64_t)m.pair_b[rp] * L + tok) * 3 * C + col : nullptr;
#pragma unroll
        for (int which = 0; which < 3; ++which) {             // dQ (cols 128), dK (cols 192), dV (cols 64)
          float o[32];
          const uint32_t cbase = which == 0 ? 128u : (which == 1 ? 192u : 64u);
          w_ld32(lane_base + cbase + (uint32_t)(rp * 32), (uint32_t*)o);      // warp-collective: every lane takes part
          if (wr && tok >= 0) {
            float4* dst = (float4*)(gp + which * C);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              float4 x = make_float4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
              if (which == 0) { x.x *= scale; x.y *= scale; x.z *= scale; x.w *= scale; }
              dst[c] = x;
            }
          } else if (wr && g_bias && which > 0) {            // zero-padded token: its k = v = bias (its output row is cropped: dq = 0)
#pragma unroll
            for (int d = 0; d < 32; ++d) atomicAdd(g_bias + which * C + col + d, o[d]);
          }
        }
      }
      phase ^= 1;
      w_fence_before();
      tw_cta_sync();                                   // (C) TMEM and the operand tiles are free for the next work item
      w_fence_after();
    }
  }
  w_fence_before();
  __syncthreads();
  if (warp == 0) {
    w_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256) : "memory");
  }
}

}  // namespace ged
using namespace ged;

// Same contract as ged_winattn_fwd.  `index` must be the standard Swin relative-position index (depthformer_swin.py:
// 168-172: (dy + 6) * 13 + (dx + 6)); the host wrapper checks that once per buffer.
GED_API int ged_winattn_tc_fwd(const float* qkv, const float* qkv_bias, const float* table, float* ctx, int B, int H, int W,
                               int C, int nH, int window, int shift, float scale, cudaStream_t stream) {
  if (!qkv || !table || !ctx || B <= 0) return GED_ERR_ARG;
  if (window != TW_WS || C != nH * TW_HD || H <= 0 || W <= 0 || shift < 0 || shift >= TW_WS) return GED_ERR_SHAPE;
  if (!aligned16(qkv) || !aligned16(ctx) || (qkv_bias && !aligned16(qkv_bias))) return GED_ERR_ALIGN;
  TwGeom g;
  g.H = H; g.W = W; g.shift = shift;
  g.Hp = cdiv(H, TW_WS) * TW_WS; g.Wp = cdiv(W, TW_WS) * TW_WS; g.nWx = g.Wp / TW_WS; g.nWin = (g.Hp / TW_WS) * g.nWx;
  const int64_t pairs = (int64_t)B * g.nWin * nH;
  if (pairs > 0x7fffffff) return GED_ERR_SHAPE;
  if (cudaFuncSetAttribute(winattn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, W_TOTAL) != cudaSuccess) return GED_ERR_LAUNCH;
  const int duos = (int)((pairs + 1) / 2);
  const int grid = duos < 148 * 2 * 4 ? duos : 148 * 2 * 4;       // 2 CTAs per SM, a few duos each
  winattn_tc_fwd_kernel<<<grid, TW_THREADS, W_TOTAL, stream>>>(qkv, qkv_bias, table, ctx, g, B, C, nH, scale, (int)pairs);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

// Same contract as ged_winattn_bwd (g_qkv fully overwritten, g_bias / g_table accumulated into), one pass TF32 on tcgen05.
GED_API int ged_winattn_tc_bwd(const float* qkv, const float* qkv_bias, const float* table, const float* g_ctx, float* g_qkv,
                               float* g_bias, float* g_table, int B, int H, int W, int C, int nH, int window, int shift,
                               float scale, cudaStream_t stream) {
  if (!qkv || !table || !g_ctx || !g_qkv || !g_table || B <= 0) return GED_ERR_ARG;
  if (window != TW_WS || C != nH * TW_HD || H <= 0 || W <= 0 || shift < 0 || shift >= TW_WS) return GED_ERR_SHAPE;
  if (!aligned16(qkv) || !aligned16(g_ctx) || !aligned16(g_qkv) || (qkv_bias && !aligned16(qkv_bias))) return GED_ERR_ALIGN;
  TwGeom g;
  g.H = H; g.W = W; g.shift = shift;
  g.Hp = cdiv(H, TW_WS) * TW_WS; g.Wp = cdiv(W, TW_WS) * TW_WS; g.nWx = g.Wp / TW_WS; g.nWin = (g.Hp / TW_WS) * g.nWx;
  const int64_t pairs = (int64_t)B * g.nWin * nH;
  if (pairs > 0x7fffffff) return GED_ERR_SHAPE;
  if (cudaFuncSetAttribute(winattn_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WB_TOTAL) != cudaSuccess) return GED_ERR_LAUNCH;
  const int duos = (int)((pairs + 1) / 2);
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = duos < sms ? duos : sms;                        // one CTA per SM, persistent over the duos
  winattn_tc_bwd_kernel<<<grid, TW_THREADS, WB_TOTAL, stream>>>(qkv, qkv_bias, table, g_ctx, g_qkv, g_bias, g_table, g, B, C, nH,
                                                               scale, (int)pairs);
  GED_CHECK_LAUNCH();
  return GED_OK;
}
