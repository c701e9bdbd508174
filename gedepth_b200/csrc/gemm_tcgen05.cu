// TF32 tensor-core GEMM / implicit-GEMM 3x3 convolution for sm_100a: tcgen05.mma (kind::tf32) with
// the accumulator in TMEM, operands staged by TMA into 128B-swizzled shared memory through an
// mbarrier ring, warp-specialised (TMA producer / single-thread MMA issuer / 8 epilogue warps / 4 splitter warps
// for 3xTF32), persistent over output tiles with a double-buffered TMEM accumulator so the epilogue of tile i
// overlaps the main loop of tile i+1.  Two kernels share the epilogue: gemm_tf32_kernel (one CTA per 128 x BN tile)
// and gemm2_tf32_kernel (a CTA pair, cta_group::2, per 256 x BN tile - used for the 3xTF32 arithmetic).
//
//   D[M,N] = epi( sum_t  A[m + tap_off[t], 0:Kt] . B[n, t*Kt : (t+1)*Kt]^T )
//
// * linear layers / 1x1 convs (a6, a8, a9, a11, a16 of SURVEY.md §8): one tap, tap_off = 0;
//   A = activations (rows x K, K contiguous), B = the nn.Linear / 1x1 weight as stored ([N][K]).
// * 3x3 convs (a11 fusion convs, a12/a13 PE necks, a16 decoder, a17 conv_depth): A is the
//   zero-bordered NHWC input flattened to [(n,y,x) x C]; the nine taps are the same 2-D TMA box
//   shifted by (dy-1)*(W+2) + (dx-1) rows, accumulated into one TMEM tile.  The epilogue maps the
//   padded row index back to (n,y,x) and stores interior pixels only.
// fp32 in HBM, tf32 multiply, fp32 accumulate: the reference's cuDNN convs run TF32 by default
// (torch.backends.cudnn.allow_tf32) and its fp32 linears are the precision bar.
//
// The same kernel also runs the backward: dX with the forward weight read in place as an MN-major B operand
// (ged_gemm_tf32_bt, ged_conv3x3_dx_tf32) and dW with BOTH operands MN-major, the contraction over pixels split across
// SMs and accumulated with vector atomics (ged_gemm_dw_tf32).
//
// Epilogue (fused, in order): + bias[n] -> activation -> dropout -> * row_scale[batch(m)] -> + residual[m,n].
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include "common.cuh"

namespace ged {

constexpr int BM = 128;          // UMMA_M, rows per tile (TMEM lanes)
constexpr int BK = 32;           // fp32 elements per 128-byte swizzle row
constexpr int UK = 8;            // UMMA_K for tf32
// warp0 TMA producer, warp1 MMA issuer, warp2 TMEM allocator, warp3 idle, warps 4-11 epilogue, warps 12-15 splitter
constexpr int MAX_TAPS = 9;

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2, ACT_GELU = 3, ACT_SIGMOID = 4 };

struct GemmParams {
  int M, N, K;              // K = total contraction (ntaps * Kt)
  int ntaps, kblocks_per_tap, Kt;  // Kt = contraction length of one tap
  int tap_off[MAX_TAPS];    // row offset of each tap in A
  const float* bias;        // [N] or null
  const float* residual;    // [M_out, ldd] or null
  const float* row_scale;   // [batch] or null (DropPath)
  int rows_per_batch;       // for row_scale
  float* D;
  float* D_pre;             // optional: pre-activation (x + bias) copy, same pitch as D (saved for GELU')
  int ldd;                  // row pitch of D (and of residual when ldr == 0) in floats
  int ldr;                  // row pitch of residual in floats; 0 = ldd
  int act;
  float slope;
  // dropout on the activated output, before row scale / residual (0 = off)
  float drop_p;
  uint32_t drop_seed;
  const int* drop_step;
  // conv epilogue: padded row -> (n, y, x); 0 disables
  int conv_Hp, conv_Wp;     // padded extents (H+2, W+2)
  int num_m_tiles, num_n_tiles;
  // weight-gradient modes (dW = G^T X, contraction over the rows/pixels P): the work unit is
  // (output tile, output tap, K split); partial tiles are added to D with vector atomics.
  //   mode 0: forward/dX GEMM as documented above (one unit per output tile, plain stores)
  //   mode 1: both operands MN-major as stored ([P][features]); TMA 32x32 panels with the
  //           128B/32B-atom swizzle, UMMA LayoutType SWIZZLE_128B_BASE32B (the only MN-major tf32 layout)
  int mode;
  // mode 0 only: B operand stored [K rows][N cols] (row pitch in the tensor map), i.e. MN-major - the forward's
  // weight matrix used untransposed by the dX GEMMs: dX = dY @ W with W [N_out][K_in], and the 3x3 dX conv reading
  // tap t's slab W[:, t*b_tap_cols : (t+1)*b_tap_cols] of the forward layout [Cout][9][Cin].
  int b_mn, b_tap_cols;
  int out_taps;             // 1 (linear / 1x1) or 9 (3x3): tap t reads B rows shifted by tap_off[t]
  int tap_dstride;          // element offset of tap t's output block in D
  int total_kb, kb_per_split, splits;
};

struct Unit {
  int m0, n0, kb0, kb1, tap;
};

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// One lane of a CONVERGED warp.  The MMA issuer runs its loop with all 32 lanes and issues under this predicate: inside a
// divergent `if (lane == 0)` the compiler cannot prove the descriptor / TMEM-address operands of tcgen05.mma uniform and wraps
// every UTCHMMA in an elect-and-broadcast loop (PLOP3 / ELECT / R2UR.BROADCAST / BRA.U.ANY, ~8 instructions and a branch per
// MMA), which paces the narrow tiles.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :: "r"(smem_u32(smem_dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// A operand from tensor memory (128 lanes = rows, one 32-bit column per tf32 element of K), B from shared memory
__device__ __forceinline__ void tc_mma_tf32_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 consecutive columns of this thread's TMEM lane <- 32 registers
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
         "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
         "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
         "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* v) {
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

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row swizzle atoms 1024 bytes apart.
// bits: [0,14) addr>>4 | [16,30) LBO>>4 (unused for one atom along K) | [32,46) SBO>>4 |
//       [46,48) version=1 (sm_100) | [61,64) layout: 2 = SWIZZLE_128B   (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// MN-major tf32 operand tile (cute::UMMA::Layout_MN_SW128_32B_Atom): 32-float (128 B) rows along MN, four
// consecutive K rows form one 512-byte swizzle atom (Swizzle<2,5,2>: 32-byte chunks XOR (row & 3)), which is
// what TMA writes with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B for a {32 features, BK rows} box.  One panel of
// 32 features x BK rows is BK*128 bytes; LBO = distance between MN panels, SBO = distance between K atoms.
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128_32b(uint32_t smem_addr, uint32_t panel_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((panel_bytes >> 4) & 0x3FFF) << 16;   // LBO
  d |= (uint64_t)(512 >> 4) << 32;                      // SBO
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                               // SWIZZLE_128B_BASE32B
  return d;
}
// cute::UMMA::InstrDescriptor: c_format F32 (1) @4, a/b format TF32 (2) @7/@10, a/b major @15/@16
// (0 = K-major, 1 = MN-major), n_dim = N>>3 @17, m_dim = M>>4 @24
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N, bool a_mn = false, bool b_mn = false) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// kind::f16 with bf16 operands (a/b format 1), fp32 accumulate; both operands K-major
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// SPLIT == 2 (bf16 hi/lo split): the fp32 tile TMA wrote ([rows][32 floats], 128B-swizzled) becomes a tile of the same
// size and swizzle whose rows are [32 bf16 hi | 32 bf16 lo], hi = bf16(x), lo = bf16(x - hi): x = hi + lo to 2^-17 |x|.
// Each k-block then issues hi*hi + lo*hi + hi*lo as kind::f16 MMAs (K = 16 per instruction, half the tensor-pipe time
// of the three kind::tf32 passes); the dropped lo*lo term and the residual of the split are 2^-16 relative per product,
// 32x finer than the single-pass TF32 (2^-11) the reference's PyTorch runs on Ampere+.
template <int BYTES>
__device__ __forceinline__ void split_bf16_tile(const uint8_t* src, uint8_t* dst, int st) {
  // every shared-memory load of the thread's items is issued before the first conversion (the loop is latency-bound
  // otherwise: 4 splitter warps, one per scheduler)
  constexpr int ITEMS = (BYTES / 32 + 127) / 128;
  float4 x0[ITEMS], x1[ITEMS];
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const int i = st + it * 128;
    if (i < BYTES / 32) {
      const int r = i >> 2, cp = i & 3;
      const uint32_t row = (uint32_t)r * 128u, sw = (uint32_t)r & 7u;
      x0[it] = *(const float4*)(src + row + (((2u * cp) ^ sw) << 4));
      x1[it] = *(const float4*)(src + row + (((2u * cp + 1u) ^ sw) << 4));
    }
  }
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const int i = st + it * 128;
    if (i < BYTES / 32) {
      const int r = i >> 2, cp = i & 3;
      const uint32_t row = (uint32_t)r * 128u, sw = (uint32_t)r & 7u;
      const float x[8] = {x0[it].x, x0[it].y, x0[it].z, x0[it].w, x1[it].x, x1[it].y, x1[it].z, x1[it].w};
      uint32_t hv[4], lv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(x[2 * e], x[2 * e + 1]);
        const __nv_bfloat162 l = __floats2bfloat162_rn(x[2 * e] - __low2float(h), x[2 * e + 1] - __high2float(h));
        hv[e] = *reinterpret_cast<const uint32_t*>(&h);
        lv[e] = *reinterpret_cast<const uint32_t*>(&l);
      }
      *(uint4*)(dst + row + (((uint32_t)cp ^ sw) << 4)) = make_uint4(hv[0], hv[1], hv[2], hv[3]);
      *(uint4*)(dst + row + (((4u + cp) ^ sw) << 4)) = make_uint4(lv[0], lv[1], lv[2], lv[3]);
    }
  }
}
__device__ __forceinline__ Unit decode_unit(const GemmParams& p, int u, int BN_) {
  const int tiles = p.num_m_tiles * p.num_n_tiles, per = tiles * p.out_taps;
  const int split = u / per, r = u - split * per;
  Unit t;
  t.tap = r / tiles;
  const int tile = r - t.tap * tiles;
  t.m0 = (tile / p.num_n_tiles) * 128;
  t.n0 = (tile % p.num_n_tiles) * BN_;
  t.kb0 = split * p.kb_per_split;
  t.kb1 = min(t.kb0 + p.kb_per_split, p.total_kb);
  return t;
}

__device__ __forceinline__ float apply_act(float x, int act, float slope) {
  switch (act) {
    case ACT_RELU: return fmaxf(x, 0.f);
    case ACT_LEAKY: return x > 0.f ? x : x * slope;
    case ACT_GELU: return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));   // exact GELU (nn.GELU default)
    case ACT_SIGMOID: return 1.f / (1.f + __expf(-x));
    default: return x;
  }
}

__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

constexpr int EPI_WARPS = 8;       // two warps per TMEM lane quadrant, each owning half of the columns
constexpr int EPI_PITCH = 20;      // floats; 32 rows x 16 columns staging tile per warp, conflict-free for float4
constexpr int SPLIT_WARPS = 4;

// One output tile of the epilogue for one warp: TMEM quadrant q (32 accumulator rows), column half `half` of the
// BN-wide tile.  Waits for the accumulator (full_bar, phase), then bias -> activation -> row scale -> residual and
// 128-bit coalesced stores (or vector-atomic accumulation for the split-K weight-gradient modes).
template <int BN>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, float* tile_s, uint32_t tmem_acc, uint64_t* full_bar,
                                          uint32_t full_phase, int m0, int n0, float* Dt, int q, int half, int lane) {
  const int64_t ldr = p.ldr ? p.ldr : p.ldd;
  const bool vec = ((p.ldd & 3) == 0) && ((p.N & 3) == 0) && aligned16(p.D) && (!p.residual || (aligned16(p.residual) && (ldr & 3) == 0));
  const int lr = lane >> 2, lc = (lane & 3) * 4;          // store phase: lane -> (row lr + 8 i, 4 columns from lc)
  // the four output rows this lane stores (fixed for the whole tile)
  int64_t orow[4];
  float rscale[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + q * 32 + lr + 8 * i;
    orow[i] = -1; rscale[i] = 1.f;
    if (row < p.M) {
      orow[i] = row;
      if (p.conv_Wp) {   // padded flattened pixel -> interior test and unpadded row
        const int xp = row % p.conv_Wp, t2 = row / p.conv_Wp, yp = t2 % p.conv_Hp, n = t2 / p.conv_Hp;
        orow[i] = (xp == 0 || xp == p.conv_Wp - 1 || yp == 0 || yp == p.conv_Hp - 1)
                      ? -1 : ((int64_t)n * (p.conv_Hp - 2) + (yp - 1)) * (p.conv_Wp - 2) + (xp - 1);
      }
      if (p.row_scale && orow[i] >= 0) rscale[i] = __ldg(p.row_scale + orow[i] / p.rows_per_batch);
    }
  }
  const bool drop = p.drop_p > 0.f;
  const uint32_t dseed = drop ? drop_seed_eff(p.drop_seed, p.drop_step) : 0u, dthresh = drop_threshold(p.drop_p);
  const float dinv = drop ? drop_scale(dthresh) : 1.f;
  mbar_wait(full_bar, full_phase);
  tc_fence_after();
#pragma unroll 1
  for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 16) {
    uint32_t v[16];
    tc_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *(float4*)(tile_s + lane * EPI_PITCH + 4 * j) =
          make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
    __syncwarp();
    const int col = n0 + c0 + lc;
    if (col < p.N) {
      float bias[4] = {0.f, 0.f, 0.f, 0.f};
      if (p.bias) {
#pragma unroll
        for (int e = 0; e < 4; ++e) if (col + e < p.N) bias[e] = __ldg(p.bias + col + e);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (orow[i] < 0) continue;
        const float4 t = *(const float4*)(tile_s + (lr + 8 * i) * EPI_PITCH + lc);
        float x[4] = {t.x + bias[0], t.y + bias[1], t.z + bias[2], t.w + bias[3]};
        if (p.D_pre) {
          if (vec) *(float4*)(p.D_pre + orow[i] * p.ldd + col) = make_float4(x[0], x[1], x[2], x[3]);
          else for (int e = 0; e < 4; ++e) if (col + e < p.N) p.D_pre[orow[i] * p.ldd + col + e] = x[e];
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) x[e] = apply_act(x[e], p.act, p.slope) * rscale[i];
        if (drop) {
          const uint32_t keep = drop_keep4(dseed, (uint32_t)(orow[i] * p.N + col), dthresh);   // N % 4 == 0, col % 4 == 0
#pragma unroll
          for (int e = 0; e < 4; ++e) x[e] = ((keep >> e) & 1u) ? x[e] * dinv : 0.f;
        }
        float* dst = Dt + orow[i] * p.ldd + col;
        if (p.mode != 0) {          // split-K partial: accumulate (D is zero-initialised / holds the running gradient)
          if (vec && ((p.tap_dstride & 3) == 0)) atomicAdd((float4*)dst, make_float4(x[0], x[1], x[2], x[3]));
          else for (int e = 0; e < 4; ++e) if (col + e < p.N) atomicAdd(dst + e, x[e]);
        } else if (vec) {
          if (p.residual) {
            const float4 r = __ldg((const float4*)(p.residual + orow[i] * ldr + col));
            x[0] += r.x; x[1] += r.y; x[2] += r.z; x[3] += r.w;
          }
          *(float4*)dst = make_float4(x[0], x[1], x[2], x[3]);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (col + e < p.N) dst[e] = x[e] + (p.residual ? __ldg(p.residual + orow[i] * ldr + col + e) : 0.f);
        }
      }
    }
    __syncwarp();
  }
}

// SPLIT = error-compensated "3xTF32": every fp32 operand x is used as hi = tf32(x) (the tensor core's own
// truncation) plus lo = x - hi (exact in fp32), and each k-step issues hi*hi + lo*hi + hi*lo.  The dropped
// lo*lo term is 2^-22 relative: the product is fp32-accurate while still running on tcgen05.
template <int BN, int STAGES, int SPLIT>
struct GemmSmem {
  static constexpr int A_BYTES = BM * BK * 4;       // 16 KB
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int HI_BYTES = A_BYTES + B_BYTES;
  // SPLIT == 3: 3xTF32 with the A operand (hi and lo) in tensor memory - only B keeps a lo tile in shared memory
  static constexpr int STAGE_BYTES = SPLIT == 3 ? HI_BYTES + B_BYTES : HI_BYTES * (SPLIT ? 2 : 1);
  static constexpr int A_TMEM_COLS = SPLIT == 3 ? STAGES * 2 * BK : 0;          // hi | lo, BK columns each, per stage
  static constexpr int EPI_BYTES = EPI_WARPS * 32 * EPI_PITCH * 4;
  static constexpr int BAR_BYTES = (3 * STAGES + 4) * 8 + 16;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;  // +1024 alignment slack
  static constexpr int THREADS = 128 + EPI_WARPS * 32 + (SPLIT ? SPLIT_WARPS * 32 : 0);
};

template <int BN, int STAGES, int SPLIT>
__global__ void __launch_bounds__(GemmSmem<BN, STAGES, SPLIT>::THREADS, 1) gemm_tf32_kernel(
    const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
    const GemmParams p) {
  using S = GemmSmem<BN, STAGES, SPLIT>;
  constexpr int TMEM_NEED = 2 * BN + S::A_TMEM_COLS;
  static_assert(TMEM_NEED <= 512, "tensor memory budget");
  constexpr int TMEM_COLS = (TMEM_NEED <= 32) ? 32 : (TMEM_NEED <= 64) ? 64 : (TMEM_NEED <= 128) ? 128 : (TMEM_NEED <= 256) ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B needs 1024B alignment
  uint8_t* stage_base = smem;
  float* epi_tiles = (float*)(smem + STAGES * S::STAGE_BYTES);
  uint64_t* full_bar = (uint64_t*)((uint8_t*)epi_tiles + S::EPI_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* split_bar = empty_bar + STAGES;
  uint64_t* tmem_full = split_bar + STAGES;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint32_t* tmem_ptr = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles * p.out_taps * p.splits;   // work units

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" :: "l"((uint64_t)&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"((uint64_t)&map_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(full_bar + i, 1); mbar_init(empty_bar + i, 1); mbar_init(split_bar + i, SPLIT_WARPS); }
    for (int i = 0; i < 2; ++i) { mbar_init(tmem_full + i, 1); mbar_init(tmem_empty + i, EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_ptr)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const Unit u = decode_unit(p, tile, BN);
        for (int kb = u.kb0; kb < u.kb1; ++kb) {
          mbar_wait(empty_bar + stage, phase ^ 1);
          uint8_t* sa = stage_base + stage * S::STAGE_BYTES;
          mbar_expect_tx(full_bar + stage, S::HI_BYTES);
          if (p.mode == 0) {
            const int t = kb / p.kblocks_per_tap, kc = kb - t * p.kblocks_per_tap;
            tma_load_2d(sa, &map_a, full_bar + stage, kc * BK, u.m0 + p.tap_off[t]);
            if (!p.b_mn) {
              tma_load_2d(sa + S::A_BYTES, &map_b, full_bar + stage, t * p.Kt + kc * BK, u.n0);
            } else {
#pragma unroll
              for (int pnl = 0; pnl < BN / 32; ++pnl)
                tma_load_2d(sa + S::A_BYTES + pnl * (BK * 128), &map_b, full_bar + stage,
                            t * p.b_tap_cols + u.n0 + 32 * pnl, kc * BK);
            }
          } else {
            // operands as stored: rows = contraction index (pixels), 32-feature panels side by side
#pragma unroll
            for (int pnl = 0; pnl < BM / 32; ++pnl)
              tma_load_2d(sa + pnl * (BK * 128), &map_a, full_bar + stage, u.m0 + 32 * pnl, kb * BK);
#pragma unroll
            for (int pnl = 0; pnl < BN / 32; ++pnl)
              tma_load_2d(sa + S::A_BYTES + pnl * (BK * 128), &map_b, full_bar + stage, u.n0 + 32 * pnl,
                          kb * BK + p.tap_off[u.tap]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the loop, one elected lane issues =====
    {
      const bool a_mn = p.mode == 1, b_mn = p.mode == 1 || p.b_mn;
      const uint32_t idesc = umma_idesc_tf32(BM, BN, a_mn, b_mn);
      // K-major: 8 tf32 = 32 bytes inside the 128B swizzle row, +2 in the (addr>>4) field per k-step;
      // MN-major: 8 K rows = two 512-byte atoms = 1024 bytes, +64 per k-step
      const uint32_t kstep_a = a_mn ? 64u : 2u, kstep_b = b_mn ? 64u : 2u;
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const Unit u = decode_unit(p, tile, BN);
        mbar_wait(tmem_empty + acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = u.kb0; kb < u.kb1; ++kb) {
          mbar_wait((SPLIT ? split_bar : full_bar) + stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + stage * S::STAGE_BYTES);
          const uint64_t adesc = a_mn ? umma_desc_mnmajor_sw128_32b(sa, BK * 128) : umma_desc_kmajor_sw128(sa);
          const uint64_t bdesc = b_mn ? umma_desc_mnmajor_sw128_32b(sa + S::A_BYTES, BK * 128) : umma_desc_kmajor_sw128(sa + S::A_BYTES);
          if (elect_one()) {
          if (SPLIT == 3) {
            // A hi / lo in tensor memory (columns 2 BN + stage * 64 .. +31 / +32 .. +63), B hi = the fp32 tile, B lo behind it
            const uint32_t ta = tmem_base + (uint32_t)(2 * BN + stage * 2 * BK);
            const uint64_t blo = bdesc + (S::B_BYTES >> 4);
#pragma unroll
            for (int k = 0; k < BK / UK; ++k) {
              tc_mma_tf32_ta(tmem_d, ta + UK * k, bdesc + 2 * k, idesc, ((kb - u.kb0) | k) != 0);      // hi(A) * hi(B)
              tc_mma_tf32_ta(tmem_d, ta + BK + UK * k, bdesc + 2 * k, idesc, 1u);                       // lo(A) * hi(B)
              tc_mma_tf32_ta(tmem_d, ta + UK * k, blo + 2 * k, idesc, 1u);                              // hi(A) * lo(B)
            }
          } else if (SPLIT == 2) {
            // bf16 hi/lo rows (K-major only): k-steps 0, 1 of a row are hi, 2, 3 are lo (32 bytes = 16 bf16 each)
            const uint64_t as = adesc + (S::HI_BYTES >> 4), bs = bdesc + (S::HI_BYTES >> 4);
            constexpr uint32_t idb = umma_idesc_bf16(BM, BN);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              tc_mma_bf16(tmem_d, as + 2 * k, bs + 2 * k, idb, ((kb - u.kb0) | k) != 0);   // hi(A) * hi(B)
              tc_mma_bf16(tmem_d, as + 2 * (k + 2), bs + 2 * k, idb, 1u);                    // lo(A) * hi(B)
              tc_mma_bf16(tmem_d, as + 2 * k, bs + 2 * (k + 2), idb, 1u);                    // hi(A) * lo(B)
            }
          } else {
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            tc_mma_tf32(tmem_d, adesc + kstep_a * k, bdesc + kstep_b * k, idesc, ((kb - u.kb0) | k) != 0);
            if (SPLIT) {
              const uint64_t alo = adesc + (S::HI_BYTES >> 4), blo = bdesc + (S::HI_BYTES >> 4);
              tc_mma_tf32(tmem_d, alo + kstep_a * k, bdesc + kstep_b * k, idesc, 1u);   // lo(A) * hi(B)
              tc_mma_tf32(tmem_d, adesc + kstep_a * k, blo + kstep_b * k, idesc, 1u);   // hi(A) * lo(B)
            }
          }
          }
          tc_commit(empty_bar + stage);          // frees the smem slot when these MMAs retire
          if (kb == u.kb1 - 1) tc_commit(tmem_full + acc);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 4 + EPI_WARPS) {
    // ===== epilogue: TMEM -> registers -> smem transpose -> 128-bit coalesced global accesses =====
    const int ew = warp - 4, q = ew & 3, half = ew >> 2;   // TMEM lane quadrant = warp % 4
    float* tile_s = epi_tiles + ew * 32 * EPI_PITCH;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const Unit u = decode_unit(p, tile, BN);
      const int m0 = u.m0, n0 = u.n0;
      float* const Dt = p.D + (int64_t)u.tap * p.tap_dstride;
      epilogue_tile<BN>(p, tile_s, tmem_base + (uint32_t)(acc * BN), tmem_full + acc, acc_phase, m0, n0, Dt, q, half, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty + acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (SPLIT && warp >= 4 + EPI_WARPS) {
    // ===== splitter: lo = x - tf32(x) for both operand tiles of every stage =====
    const int st = threadIdx.x - (4 + EPI_WARPS) * 32;
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const Unit u = decode_unit(p, tile, BN);
      for (int kb = u.kb0; kb < u.kb1; ++kb) {
        mbar_wait(full_bar + stage, phase);
        const float4* hi = (const float4*)(stage_base + stage * S::STAGE_BYTES);
        float4* lo = (float4*)(stage_base + stage * S::STAGE_BYTES + S::HI_BYTES);
        if (SPLIT == 3) {
          // this thread's row of the A tile (row st = TMEM lane st; splitter warp w owns lane quadrant w % 4): 8 swizzled
          // 16-byte chunks -> hi (the values as they are: the tensor core truncates) and lo = x - tf32(x) -> tensor memory
          const uint8_t* arow = (const uint8_t*)hi + st * 128;
          uint32_t vh[32], vl[32];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 x = *(const float4*)(arow + ((c ^ (st & 7)) << 4));
            const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              vh[4 * c + e] = __float_as_uint(xs[e]);
              vl[4 * c + e] = __float_as_uint(xs[e] - __uint_as_float(__float_as_uint(xs[e]) & 0xFFFFE000u));
            }
          }
          const uint32_t ta = tmem_base + ((uint32_t)((st >> 5) * 32) << 16) + (uint32_t)(2 * BN + stage * 2 * BK);
          tc_st32(ta, vh);
          tc_st32(ta + BK, vl);
          // B: lo tile behind the fp32 tile, in shared memory as before
          const float4* bh = (const float4*)((const uint8_t*)hi + S::A_BYTES);
          float4* bl = (float4*)((uint8_t*)hi + S::HI_BYTES);
#pragma unroll 4
          for (int i = st; i < S::B_BYTES / 16; i += SPLIT_WARPS * 32) {
            const float4 x = bh[i];
            float4 l;
            l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
            l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
            l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
            l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
            bl[i] = l;
          }
          tc_wait_st();
          tc_fence_before();
        } else if (SPLIT == 2) {
          split_bf16_tile<S::HI_BYTES>((const uint8_t*)hi, (uint8_t*)lo, st);
        } else {
#pragma unroll 4
        for (int i = st; i < S::HI_BYTES / 16; i += SPLIT_WARPS * 32) {
          const float4 x = hi[i];
          float4 l;
          l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
          l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
          l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
          l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
          lo[i] = l;
        }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05.mma
        __syncwarp();
        if (lane == 0) mbar_arrive(split_bar + stage);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}


// =================================================================================================
// CTA-pair variant (cta_group::2): two CTAs of a cluster (the two SMs of a TPC) compute one 256 x BN tile.
// CTA r stages its own 128 rows of A and HALF of the B tile (BN/2 rows); one tcgen05.mma.cta_group::2 issued by
// the leader (rank 0) reads A from each CTA's shared memory and B from both, and writes 128 x BN accumulators
// into each CTA's TMEM.  Per flop every SM stages and re-reads half the operand bytes of the single-CTA kernel:
// the 3xTF32 path is shared-memory-bandwidth bound and the 1-pass path L2-bandwidth bound (profiles/), so this
// is where the time goes.
//
// Signalling: TMA lands in the issuing CTA's shared memory and completes on that CTA's full barrier; the
// splitter warps (3xTF32) or a relay thread (1 pass) of each CTA then arrive on the LEADER's ready barrier
// (remote mbarrier arrive); the leader's MMA thread waits on it, issues, and multicasts tcgen05.commit to the
// empty / tmem_full barriers of both CTAs; epilogue warps of both CTAs arrive on the leader's tmem_empty.
// =================================================================================================
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `bar` (a shared::cta object of THIS CTA) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_rank(const void* bar, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(bar)), "r"(rank));
  return r;
}
// Default (.release.cta) semantics as in CUTLASS's ClusterBarrier::arrive: a cluster-scope release costs a
// ~1300-cycle fence per arrive (measured: it capped the pair kernel at one k-block per 0.7 us).  What the leader's
// tensor core reads was written by TMA (async proxy) or made visible to it by fence.proxy.async before this arrive.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAITC_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONEC_%=;\n\t"
      "bra WAITC_%=;\n\t"
      "DONEC_%=:\n\t}"
      :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {      // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               :: "r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

constexpr int BM2 = 256;   // rows per CTA-pair tile

template <int BN, int STAGES, int SPLIT>
struct Gemm2Smem {
  static constexpr int A_BYTES = BM * BK * 4;              // this CTA's 128 rows
  static constexpr int B_BYTES = (BN / 2) * BK * 4;        // this CTA's half of the B tile
  static constexpr int HI_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGE_BYTES = HI_BYTES * (SPLIT ? 2 : 1);
  static constexpr int EPI_BYTES = EPI_WARPS * 32 * EPI_PITCH * 4;
  static constexpr int BAR_BYTES = (3 * STAGES + 4) * 8 + 16;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;
  static constexpr int THREADS = 128 + EPI_WARPS * 32 + (SPLIT ? SPLIT_WARPS * 32 : 0);
};

__device__ __forceinline__ Unit decode_unit2(const GemmParams& p, int u, int BN_) {
  // same unit order as decode_unit (split-major, then output tap, then tile) with 256-row tiles
  const int tiles = p.num_m_tiles * p.num_n_tiles, per = tiles * p.out_taps;
  const int split = u / per, r = u - split * per;
  Unit t;
  t.tap = r / tiles;
  const int tile = r - t.tap * tiles;
  t.m0 = (tile / p.num_n_tiles) * BM2;
  t.n0 = (tile % p.num_n_tiles) * BN_;
  t.kb0 = split * p.kb_per_split;
  t.kb1 = min(t.kb0 + p.kb_per_split, p.total_kb);
  return t;
}

template <int BN, int STAGES, int SPLIT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Gemm2Smem<BN, STAGES, SPLIT>::THREADS, 1) gemm2_tf32_kernel(
    const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const GemmParams p) {
  using S = Gemm2Smem<BN, STAGES, SPLIT>;
  constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  constexpr uint32_t READY_COUNT = 2 * (SPLIT ? SPLIT_WARPS : 1);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem;
  float* epi_tiles = (float*)(smem + STAGES * S::STAGE_BYTES);
  uint64_t* full_bar = (uint64_t*)((uint8_t*)epi_tiles + S::EPI_BYTES);   // local: this CTA's TMA bytes
  uint64_t* empty_bar = full_bar + STAGES;                                // local: MMAs that read this slot retired
  uint64_t* ready_bar = empty_bar + STAGES;                               // leader's copy is the one that counts
  uint64_t* tmem_full = ready_bar + STAGES;                               // [2] local
  uint64_t* tmem_empty = tmem_full + 2;                                   // [2] leader's copy counts
  uint32_t* tmem_ptr = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int num_units = p.num_m_tiles * p.num_n_tiles * p.out_taps * p.splits;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" :: "l"((uint64_t)&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"((uint64_t)&map_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(full_bar + i, 1); mbar_init(empty_bar + i, 1); mbar_init(ready_bar + i, READY_COUNT); }
    for (int i = 0; i < 2; ++i) { mbar_init(tmem_full + i, 1); mbar_init(tmem_empty + i, 2 * EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_ptr)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===== TMA producer: this CTA's 128 rows of A and its half of B =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int unit = cluster_id; unit < num_units; unit += num_clusters) {
        const Unit u = decode_unit2(p, unit, BN);
        const int m0 = u.m0 + (int)rank * BM, n0 = u.n0 + (int)rank * (BN / 2);
        for (int kb = u.kb0; kb < u.kb1; ++kb) {
          mbar_wait(empty_bar + stage, phase ^ 1);
          uint8_t* sa = stage_base + stage * S::STAGE_BYTES;
          mbar_expect_tx(full_bar + stage, S::HI_BYTES);
          if (p.mode == 0) {
            const int t = kb / p.kblocks_per_tap, kc = kb - t * p.kblocks_per_tap;
            tma_load_2d(sa, &map_a, full_bar + stage, kc * BK, m0 + p.tap_off[t]);
            if (!p.b_mn) {
              tma_load_2d(sa + S::A_BYTES, &map_b, full_bar + stage, t * p.Kt + kc * BK, n0);
            } else {
#pragma unroll
              for (int pnl = 0; pnl < BN / 64; ++pnl)
                tma_load_2d(sa + S::A_BYTES + pnl * (BK * 128), &map_b, full_bar + stage,
                            t * p.b_tap_cols + n0 + 32 * pnl, kc * BK);
            }
          } else {
            // weight gradient: operands as stored (rows = pixels), this CTA's 128 of the 256 M features and its half of the N panels
#pragma unroll
            for (int pnl = 0; pnl < BM / 32; ++pnl)
              tma_load_2d(sa + pnl * (BK * 128), &map_a, full_bar + stage, m0 + 32 * pnl, kb * BK);
#pragma unroll
            for (int pnl = 0; pnl < BN / 64; ++pnl)
              tma_load_2d(sa + S::A_BYTES + pnl * (BK * 128), &map_b, full_bar + stage, n0 + 32 * pnl, kb * BK + p.tap_off[u.tap]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: warp 1 of the leader CTA runs the loop, one elected lane issues =====
    if (rank == 0) {
      const bool a_mn = p.mode == 1, b_mn = p.mode == 1 || p.b_mn != 0;
      const uint32_t idesc = umma_idesc_tf32(BM2, BN, a_mn, b_mn);
      const uint32_t kstep_a = a_mn ? 64u : 2u, kstep_b = b_mn ? 64u : 2u;
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int unit = cluster_id; unit < num_units; unit += num_clusters) {
        const Unit u = decode_unit2(p, unit, BN);
        mbar_wait_cluster(tmem_empty + acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = u.kb0; kb < u.kb1; ++kb) {
          mbar_wait_cluster(ready_bar + stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + stage * S::STAGE_BYTES);
          const uint64_t adesc = a_mn ? umma_desc_mnmajor_sw128_32b(sa, BK * 128) : umma_desc_kmajor_sw128(sa);
          const uint64_t bdesc = b_mn ? umma_desc_mnmajor_sw128_32b(sa + S::A_BYTES, BK * 128) : umma_desc_kmajor_sw128(sa + S::A_BYTES);
          if (elect_one()) {
          if (SPLIT == 2) {
            const uint64_t as = adesc + (S::HI_BYTES >> 4), bs = bdesc + (S::HI_BYTES >> 4);
            constexpr uint32_t idb = umma_idesc_bf16(BM2, BN);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              tc_mma_bf16_pair(tmem_d, as + 2 * k, bs + 2 * k, idb, ((kb - u.kb0) | k) != 0);
              tc_mma_bf16_pair(tmem_d, as + 2 * (k + 2), bs + 2 * k, idb, 1u);
              tc_mma_bf16_pair(tmem_d, as + 2 * k, bs + 2 * (k + 2), idb, 1u);
            }
          } else {
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            tc_mma_tf32_pair(tmem_d, adesc + kstep_a * k, bdesc + kstep_b * k, idesc, ((kb - u.kb0) | k) != 0);
            if (SPLIT) {
              const uint64_t alo = adesc + (S::HI_BYTES >> 4), blo = bdesc + (S::HI_BYTES >> 4);
              tc_mma_tf32_pair(tmem_d, alo + kstep_a * k, bdesc + kstep_b * k, idesc, 1u);
              tc_mma_tf32_pair(tmem_d, adesc + kstep_a * k, blo + kstep_b * k, idesc, 1u);
            }
          }
          }
          tc_commit_pair(empty_bar + stage);
          if (kb == u.kb1 - 1) tc_commit_pair(tmem_full + acc);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp == 3) {
    // ===== relay (1-pass arithmetic): this CTA's operands have landed -> tell the leader =====
    if (!SPLIT && lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int unit = cluster_id; unit < num_units; unit += num_clusters) {
        const Unit u = decode_unit2(p, unit, BN);
        for (int kb = u.kb0; kb < u.kb1; ++kb) {
          mbar_wait(full_bar + stage, phase);
          mbar_arrive_remote(mapa_rank(ready_bar + stage, 0));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4 && warp < 4 + EPI_WARPS) {
    // ===== epilogue: this CTA's 128 rows of the pair tile =====
    const int ew = warp - 4, q = ew & 3, half = ew >> 2;
    float* tile_s = epi_tiles + ew * 32 * EPI_PITCH;
    int acc = 0; uint32_t acc_phase = 0;
    for (int unit = cluster_id; unit < num_units; unit += num_clusters) {
      const Unit u = decode_unit2(p, unit, BN);
      epilogue_tile<BN>(p, tile_s, tmem_base + (uint32_t)(acc * BN), tmem_full + acc, acc_phase, u.m0 + (int)rank * BM, u.n0,
                        p.D + (int64_t)u.tap * p.tap_dstride, q, half, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(mapa_rank(tmem_empty + acc, 0));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (SPLIT && warp >= 4 + EPI_WARPS) {
    // ===== splitter: lo = x - tf32(x) for this CTA's operand tiles, then tell the leader =====
    const int st = threadIdx.x - (4 + EPI_WARPS) * 32;
    int stage = 0; uint32_t phase = 0;
    for (int unit = cluster_id; unit < num_units; unit += num_clusters) {
      const Unit u = decode_unit2(p, unit, BN);
      for (int kb = u.kb0; kb < u.kb1; ++kb) {
        mbar_wait(full_bar + stage, phase);
        const float4* hi = (const float4*)(stage_base + stage * S::STAGE_BYTES);
        float4* lo = (float4*)(stage_base + stage * S::STAGE_BYTES + S::HI_BYTES);
        if (SPLIT == 2) {
          split_bf16_tile<S::HI_BYTES>((const uint8_t*)hi, (uint8_t*)lo, st);
        } else {
#pragma unroll 4
        for (int i = st; i < S::HI_BYTES / 16; i += SPLIT_WARPS * 32) {
          const float4 x = hi[i];
          float4 l;
          l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
          l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
          l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
          l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
          lo[i] = l;
        }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(mapa_rank(ready_bar + stage, 0));
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // the peer may still read this CTA's shared memory / signal its barriers
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_cuTensorMapEncodeTiled_v12000)ptr;
  }
  return fn;
}

// 2-D fp32 tensor map: dims {inner=cols, outer=rows}, row pitch ld floats, box {32, box_rows}, 128B swizzle
static int make_map_2d(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                       CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  auto enc = get_encode();
  if (!enc) return GED_ERR_LAUNCH;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? GED_OK : GED_ERR_ARG;
}

static int g_num_sms = 0;
static int g_wide_tiles = 1;   // allow BN = 192 / 256 (fewer re-reads of the A operand through L2)
static int g_precision = 3;   // 1 = single-pass TF32, 3 = error-compensated 3xTF32 (fp32-accurate)
// 1 (default): the 3xTF32 single-CTA kernels (tiles <= 128 wide) keep the A operand's hi / lo in tensor memory (tcgen05.st by
// the splitter, tcgen05.mma with a TMEM A operand): the MMAs then read only B from shared memory.  Narrow tiles spend 43
// tensor-clocks per MMA, so they were first paced by the ISSUING THREAD (probes: no-op splitter, a third of the MMAs skipped,
// 3 / 4 / 5 ring stages, 4 / 8 splitter warps, CTA pairing all left the 64-wide 3x3 conv at 3.8-4.2 ms against 1.94 ms for
// the one-pass kernel); with the issue loop fixed (elect_one above) shared-memory traffic is what is left, and the TMEM
// operand takes the 256 -> 64 conv at 16 x 176 x 560 from 3.5-4.0 to 2.4 ms (tools/ab_gemm_narrow.py).
static int g_a_tmem = 1;

template <int BN, int STAGES, int SPLIT>
static int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mb, GemmParams& p, cudaStream_t stream) {
  using S = GemmSmem<BN, STAGES, SPLIT>;
  static_assert(S::TOTAL <= 232448, "shared memory budget (227 KB)");
  static bool attr_set[64] = {};                 // the attribute is per device
  int dev_id = 0;
  cudaGetDevice(&dev_id);
  if (!attr_set[dev_id & 63]) {
    if (cudaFuncSetAttribute(gemm_tf32_kernel<BN, STAGES, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) != cudaSuccess) return GED_ERR_LAUNCH;
    attr_set[dev_id & 63] = true;
  }
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  p.num_m_tiles = cdiv(p.M, BM);
  p.num_n_tiles = cdiv(p.N, BN);
  if (p.mode == 0) {
    p.out_taps = 1; p.tap_dstride = 0; p.splits = 1;
    p.total_kb = p.ntaps * p.kblocks_per_tap; p.kb_per_split = p.total_kb;
  } else {
    // split the contraction so that every SM gets ~2 units, but keep >= 8 k-blocks (256 rows) per unit
    const int base_units = p.num_m_tiles * p.num_n_tiles * p.out_taps;
    int splits = cdiv(2 * g_num_sms, base_units);
    const int max_splits = p.total_kb / 8 > 0 ? p.total_kb / 8 : 1;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    p.kb_per_split = cdiv(p.total_kb, splits);
    p.splits = cdiv(p.total_kb, p.kb_per_split);
  }
  const int tiles = p.num_m_tiles * p.num_n_tiles * p.out_taps * p.splits;
  const int grid = tiles < g_num_sms ? tiles : g_num_sms;
  gemm_tf32_kernel<BN, STAGES, SPLIT><<<grid, S::THREADS, S::TOTAL, stream>>>(ma, mb, p);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

// Output-tile width.  Both arithmetic modes are bound by bytes moved per k-block, which scale with (128 + BN) rows of
// 128 bytes (TMA in, splitter, operand reads), while the work scales with BN: wider tiles re-read the A operand less.
// Against that stand wave quantisation over the SMs and, for the 3xTF32 kernel, a 2-stage ring at BN > 128 (measured
// ~1.25x per k-block: tools/ab_gemm.py).  Pick the candidate with the lowest waves x (128 + BN) estimate.
static int pick_bn(int M, int N, bool split) {
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  if (N <= 96) return 96;
  if (N <= 128) return 128;
  const int sms = g_num_sms > 0 ? g_num_sms : 148;
  const int m_tiles = cdiv(M, BM);
  const int cands[4] = {128, 96, 192, 256};
  int best = 128;
  double best_cost = 1e30;
  for (int i = 0; i < 4; ++i) {
    const int bn = cands[i];
    if (bn > 128 && !g_wide_tiles) continue;
    if (bn > 128 && (int64_t)cdiv(N, bn) * bn * 100 > (int64_t)N * 115) continue;   // > 15 % padded columns: not worth it
    const int64_t tiles = (int64_t)m_tiles * cdiv(N, bn);
    const double waves = (double)((tiles + sms - 1) / sms);
    const double cost = waves * (128 + bn) * ((split && bn > 128) ? 1.25 : 1.0);
    if (cost < best_cost * 0.999) { best_cost = cost; best = bn; }
  }
  return best;
}

static int dispatch(int bn, const CUtensorMap& ma, const CUtensorMap& mb, GemmParams& p, cudaStream_t stream) {
  if (g_precision == 2 && p.mode == 0 && !p.b_mn) {      // bf16 hi/lo split: K-major operands only
    switch (bn) {
      case 32: return launch_gemm<32, 4, 2>(ma, mb, p, stream);
      case 64: return launch_gemm<64, 4, 2>(ma, mb, p, stream);
      case 96: return launch_gemm<96, 3, 2>(ma, mb, p, stream);
      case 192: return launch_gemm<192, 2, 2>(ma, mb, p, stream);
      case 256: return launch_gemm<256, 2, 2>(ma, mb, p, stream);
      default: return launch_gemm<128, 3, 2>(ma, mb, p, stream);
    }
  }
  if (g_precision == 3 && g_a_tmem && p.mode == 0 && !p.b_mn && bn <= 128) {   // 3xTF32 with A hi / lo in tensor memory
    switch (bn) {
      case 32: return launch_gemm<32, 6, 3>(ma, mb, p, stream);
      case 64: return launch_gemm<64, 5, 3>(ma, mb, p, stream);
      case 96: return launch_gemm<96, 4, 3>(ma, mb, p, stream);
      default: return launch_gemm<128, 3, 3>(ma, mb, p, stream);
    }
  }
  if (g_precision >= 2) {
    switch (bn) {
      case 32: return launch_gemm<32, 4, 1>(ma, mb, p, stream);
      case 64: return launch_gemm<64, 4, 1>(ma, mb, p, stream);
      case 96: return launch_gemm<96, 3, 1>(ma, mb, p, stream);
      case 192: return launch_gemm<192, 2, 1>(ma, mb, p, stream);
      case 256: return launch_gemm<256, 2, 1>(ma, mb, p, stream);
      default: return launch_gemm<128, 3, 1>(ma, mb, p, stream);
    }
  }
  switch (bn) {
    case 32: return launch_gemm<32, 8, 0>(ma, mb, p, stream);
    case 64: return launch_gemm<64, 6, 0>(ma, mb, p, stream);
    case 96: return launch_gemm<96, 6, 0>(ma, mb, p, stream);
    case 192: return launch_gemm<192, 5, 0>(ma, mb, p, stream);
    case 256: return launch_gemm<256, 4, 0>(ma, mb, p, stream);
    default: return launch_gemm<128, 5, 0>(ma, mb, p, stream);
  }
}


static int g_pair = 2;        // 2 = CTA-pair kernel for large problems in both arithmetic modes (measured: one-pass dX GEMMs
                              // 45.4 -> 38.1 ms in config 3), 1 = for 3xTF32 only, 0 = never

static int g_pair_dw = 1;     // weight-gradient GEMMs on the CTA-pair kernel where the output has >= 256 rows

template <int BN, int STAGES, int SPLIT>
static int launch_gemm2(const CUtensorMap& ma, const CUtensorMap& mb, GemmParams& p, cudaStream_t stream) {
  using S = Gemm2Smem<BN, STAGES, SPLIT>;
  static_assert(S::TOTAL <= 232448, "shared memory budget (227 KB)");
  static bool attr_set[64] = {};                 // the attribute is per device
  int dev_id = 0;
  cudaGetDevice(&dev_id);
  if (!attr_set[dev_id & 63]) {
    if (cudaFuncSetAttribute(gemm2_tf32_kernel<BN, STAGES, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) != cudaSuccess) return GED_ERR_LAUNCH;
    attr_set[dev_id & 63] = true;
  }
  p.num_m_tiles = cdiv(p.M, BM2);
  p.num_n_tiles = cdiv(p.N, BN);
  const int max_clusters = g_num_sms / 2;
  if (p.mode == 0) {
    p.out_taps = 1; p.tap_dstride = 0; p.splits = 1;
    p.total_kb = p.ntaps * p.kblocks_per_tap; p.kb_per_split = p.total_kb;
  } else {
    // split the contraction so that every SM pair gets ~2 units, but keep >= 8 k-blocks (256 pixel rows) per unit
    const int base_units = p.num_m_tiles * p.num_n_tiles * p.out_taps;
    int splits = cdiv(2 * max_clusters, base_units);
    const int max_splits = p.total_kb / 8 > 0 ? p.total_kb / 8 : 1;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    p.kb_per_split = cdiv(p.total_kb, splits);
    p.splits = cdiv(p.total_kb, p.kb_per_split);
  }
  const int units = p.num_m_tiles * p.num_n_tiles * p.out_taps * p.splits;
  const int clusters = units < max_clusters ? units : max_clusters;
  gemm2_tf32_kernel<BN, STAGES, SPLIT><<<2 * clusters, S::THREADS, S::TOTAL, stream>>>(ma, mb, p);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

// Pair-tile width: the widest of 256 / 192 / 128 that wastes <= 15 % of the columns; 0 = do not use the pair kernel.
// Measured (tools/ab_pair.py): 1.3-1.5x for the 3xTF32 arithmetic at N >= 128 (shared-memory bound: the pair halves
// the operand bytes each SM stages and re-reads); no gain for single-pass TF32 (already ~600 TFLOP/s on the wide
// single-CTA tiles), for 64-column outputs (the A operand dominates) or for problems with fewer tiles than SM pairs.
static int pick_bn_pair(int M, int N) {
  if (!g_pair || (g_precision < 2 && g_pair != 2) || N < 128) return 0;     // g_pair == 2: also for one-pass TF32 (A/B switch)
  const int sms = g_num_sms > 0 ? g_num_sms : 148;
  const int cands[3] = {256, 192, 128};
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    const int n_tiles = cdiv(N, bn);
    if ((int64_t)n_tiles * bn * 100 > (int64_t)N * 115) continue;
    if ((int64_t)cdiv(M, BM2) * n_tiles < sms / 2) return 0;
    return bn;
  }
  return 0;
}

static int dispatch2(int bn, const CUtensorMap& ma, const CUtensorMap& mb, GemmParams& p, cudaStream_t stream) {
  if (g_precision == 2 && p.mode == 0 && !p.b_mn) {
    switch (bn) {
      case 64: return launch_gemm2<64, 4, 2>(ma, mb, p, stream);
      case 128: return launch_gemm2<128, 4, 2>(ma, mb, p, stream);
      case 192: return launch_gemm2<192, 3, 2>(ma, mb, p, stream);
      default: return launch_gemm2<256, 3, 2>(ma, mb, p, stream);
    }
  }
  if (g_precision >= 2) {
    switch (bn) {
      case 64: return launch_gemm2<64, 4, 1>(ma, mb, p, stream);
      case 128: return launch_gemm2<128, 4, 1>(ma, mb, p, stream);
      case 192: return launch_gemm2<192, 3, 1>(ma, mb, p, stream);
      default: return launch_gemm2<256, 3, 1>(ma, mb, p, stream);
    }
  }
  switch (bn) {
    case 64: return launch_gemm2<64, 8, 0>(ma, mb, p, stream);
    case 128: return launch_gemm2<128, 8, 0>(ma, mb, p, stream);
    case 192: return launch_gemm2<192, 7, 0>(ma, mb, p, stream);
    default: return launch_gemm2<256, 6, 0>(ma, mb, p, stream);
  }
}

static int run(const float* A, int64_t a_rows, int lda, const float* Bw, int ldb, GemmParams& p, cudaStream_t stream) {
  if (!aligned16(A) || !aligned16(Bw) || (lda % 4) || (ldb % 4)) return GED_ERR_ALIGN;
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int bn2 = pick_bn_pair(p.M, p.N);
  const int bn = bn2 ? bn2 : pick_bn(p.M, p.N, g_precision >= 2);
  CUtensorMap ma, mb;
  const int Kt = p.K / p.ntaps;
  if (int e = make_map_2d(&ma, A, a_rows, Kt, lda, BM)) return e;
  if (!p.b_mn) {
    if (int e = make_map_2d(&mb, Bw, p.N, p.K, ldb, bn2 ? bn2 / 2 : bn)) return e;
  } else {      // [Kt rows][ntaps * b_tap_cols cols] as stored, 32-column panels
    const int64_t cols = p.ntaps > 1 ? (int64_t)p.ntaps * p.b_tap_cols : p.N;
    if (int e = make_map_2d(&mb, Bw, Kt, cols, ldb, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return e;
  }
  p.kblocks_per_tap = cdiv(Kt, BK);
  p.Kt = Kt;
  if (bn2) return dispatch2(bn2, ma, mb, p, stream);
  return dispatch(bn, ma, mb, p, stream);
}

// dW[N][tap][K] += sum_p G[p][n] * X[p + tap_off[tap]][k]  (weight gradients).
// G = [P][N], X = [Px][K] as the forward stored them (MN-major UMMA operands, no copies).
static int run_dw(const float* G, int ldg, const float* X, int ldx, int64_t P, int64_t Px, GemmParams& p, cudaStream_t stream) {
  if (!aligned16(G) || !aligned16(X) || (ldg % 4) || (ldx % 4)) return GED_ERR_ALIGN;
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  // B panels of 32 features; wide tiles (192 / 256 columns) halve the re-reads of the G operand through L2
  int bn = p.N <= 32 ? 32 : p.N <= 64 ? 64 : p.N <= 96 ? 96 : (p.N % 128 != 0 && p.N % 96 == 0) ? 96 : 128;
  if (g_wide_tiles && p.N >= 192) bn = (p.N % 256 == 0) ? 256 : (p.N % 192 == 0) ? 192 : bn;
  CUtensorMap ma, mb;
  if (int e = make_map_2d(&ma, G, P, p.M, ldg, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return e;
  if (int e = make_map_2d(&mb, X, Px, p.N, ldx, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return e;
  p.ntaps = 1; p.Kt = (int)P; p.kblocks_per_tap = cdiv((int)P, BK); p.total_kb = p.kblocks_per_tap;
  // CTA-pair tiles (256 output features x bn): the X operand is re-read once per 256 instead of per 128 output features
  if (g_pair_dw && g_precision == 1 && p.M >= 256 && (p.M % 256 == 0 || p.M >= 1024) && bn >= 64 && (bn % 64) == 0)
    return dispatch2(bn, ma, mb, p, stream);
  return dispatch(bn, ma, mb, p, stream);
}

}  // namespace ged
using namespace ged;

// precision: 1 = TF32 (10-bit mantissa products), 3 = 3xTF32 split (fp32-accurate; default), 2 = bf16 hi/lo split
// (three kind::f16 products per k-step, 2^-16 relative per product; K-major operands only - MN-major ones use 3xTF32).
// Returns the previous setting.  Process-wide; not a per-call argument so call sites stay those of F.linear / conv2d.
GED_API int ged_set_gemm_precision(int passes) {
  const int prev = g_precision;
  if (passes == 1 || passes == 2 || passes == 3) g_precision = passes;
  return prev;
}

// 2 = CTA-pair (cta_group::2) kernel for large forward / dX problems in both arithmetic modes (default), 1 = for the
// 3xTF32 arithmetic only, 0 = single-CTA kernels only.
GED_API int ged_set_gemm_pair(int on) {
  const int prev = g_pair;
  g_pair = on == 2 ? 2 : (on ? 1 : 0);
  return prev;
}

// 1 (default) = the 3xTF32 single-CTA kernels feed the A operand (hi and lo) from tensor memory, 0 = from shared memory.
GED_API int ged_set_gemm_a_tmem(int on) {
  const int prev = g_a_tmem;
  g_a_tmem = on ? 1 : 0;
  return prev;
}

// 1 = weight-gradient GEMMs use the CTA-pair kernel where the output has >= 256 rows (default), 0 = single-CTA kernels.
GED_API int ged_set_gemm_pair_dw(int on) {
  const int prev = g_pair_dw;
  g_pair_dw = on ? 1 : 0;
  return prev;
}

// 1 = allow 192/256-wide output tiles (default), 0 = at most 128.  Returns the previous setting.
GED_API int ged_set_gemm_wide_tiles(int on) {
  const int prev = g_wide_tiles;
  g_wide_tiles = on ? 1 : 0;
  return prev;
}

// Weight gradient of a linear / 1x1 conv (ntaps = 1) or a 3x3 conv (ntaps = 9) on tcgen05:
//   D[n][t][k] += sum_p G[p][n] * X[p + tap_off[t]][k],   n < N, k < K, p < P  (rows of X outside [0, Px) read as 0)
// D element (n, t, k) lives at D[n * ldd + t * tap_dstride + k]; D must hold the running gradient (or zeros): the
// contraction over P is split across SMs and partial tiles are added with vector atomics.
// G is [P][N] (pitch ldg) and X is [Px][K] (pitch ldx) exactly as the forward stored them (MN-major UMMA operands,
// TMA panels of 32 features, no transposed copies).
// Replaces the autograd weight-gradient GEMMs / cuDNN wgrad of every nn.Linear / Conv2d on the path.
GED_API int ged_gemm_dw_tf32(const float* G, int ldg, const float* X, int ldx, float* D, int ldd, int N, int K,
                             int64_t P, int64_t Px, int ntaps, const int* tap_off, int tap_dstride,
                             cudaStream_t stream) {
  if (!G || !X || !D || N <= 0 || K <= 0 || P <= 0 || Px <= 0 || ntaps < 1 || ntaps > MAX_TAPS) return GED_ERR_ARG;
  if ((N % 4) || (K % 4) || P > 0x7fffffff || Px > 0x7fffffff) return GED_ERR_SHAPE;
  GemmParams p{};
  p.M = N; p.N = K; p.K = (int)P; p.mode = 1; p.out_taps = ntaps; p.tap_dstride = tap_dstride;
  for (int t = 0; t < ntaps; ++t) p.tap_off[t] = tap_off ? tap_off[t] : 0;
  p.bias = nullptr; p.residual = nullptr; p.row_scale = nullptr; p.rows_per_batch = 1;
  p.D = D; p.D_pre = nullptr; p.ldd = ldd; p.act = 0; p.slope = 0.f; p.conv_Hp = 0; p.conv_Wp = 0;
  return run_dw(G, ldg, X, ldx, P, Px, p, stream);
}

// D[M,N] = A[M,K] @ Wt[K,N]: the dX GEMM of a linear layer / 1x1 conv reading the forward weight matrix
// ([N_out = K][K_in = N], row pitch ldw) in place as an MN-major B operand - no transposed copy.
GED_API int ged_gemm_tf32_bt(const float* A, int lda, const float* Wt, int ldw, float* D, int ldd, int M, int N,
                             int K, cudaStream_t stream) {
  if (!A || !Wt || !D || M <= 0 || N <= 0 || K <= 0) return GED_ERR_ARG;
  if ((K % 4) || (N % 4)) return GED_ERR_SHAPE;
  GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.ntaps = 1; p.tap_off[0] = 0; p.b_mn = 1; p.b_tap_cols = 0;
  p.rows_per_batch = 1; p.D = D; p.ldd = ldd;
  return run(A, M, lda, Wt, ldw, p, stream);
}

// The same plus a residual: D = A @ Wt + R (R with row pitch ldr, 0 = D's; R == D accumulates in place).  Lets a gradient that
// fans in from several consumers be summed inside the GEMM epilogues instead of by separate elementwise passes.
GED_API int ged_gemm_tf32_bt_acc(const float* A, int lda, const float* Wt, int ldw, float* D, int ldd, int M, int N,
                                 int K, const float* residual, int ldr, cudaStream_t stream) {
  if (!A || !Wt || !D || M <= 0 || N <= 0 || K <= 0 || ldr < 0 || (ldr > 0 && ldr < N)) return GED_ERR_ARG;
  if ((K % 4) || (N % 4)) return GED_ERR_SHAPE;
  GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.ntaps = 1; p.tap_off[0] = 0; p.b_mn = 1; p.b_tap_cols = 0;
  p.rows_per_batch = 1; p.D = D; p.ldd = ldd; p.residual = residual; p.ldr = ldr;
  return run(A, M, lda, Wt, ldw, p, stream);
}

// dX of the 3x3/s1/p1 conv: DX[B,H,W,Cin] = sum_t Gpad[p - off_t, :] @ Wk[:, t, :] with Gpad the zero-bordered dY
// [B,H+2,W+2,Cout] and Wk the FORWARD weights [Cout][3][3][Cin] read in place (MN-major B operand, one 32-column
// panel set per tap) - no flipped / transposed weight copy.
GED_API int ged_conv3x3_dx_tf32(const float* Gpad, const float* Wk, float* DX, int ldx, int B, int H, int W, int Cin,
                                int Cout, cudaStream_t stream) {
  if (!Gpad || !Wk || !DX || B <= 0 || H <= 0 || W <= 0) return GED_ERR_ARG;
  if ((Cin % 4) || (Cout % 4)) return GED_ERR_SHAPE;
  const int Hp = H + 2, Wp = W + 2;
  GemmParams p{};
  p.M = B * Hp * Wp; p.N = Cin; p.K = 9 * Cout; p.ntaps = 9; p.b_mn = 1; p.b_tap_cols = Cin;
  for (int ky = 0; ky < 3; ++ky)
    for (int kx = 0; kx < 3; ++kx) p.tap_off[ky * 3 + kx] = -((ky - 1) * Wp + (kx - 1));
  p.rows_per_batch = 1; p.D = DX; p.ldd = ldx; p.conv_Hp = Hp; p.conv_Wp = Wp;
  return run(Gpad, (int64_t)p.M, Cout, Wk, 9 * Cin, p, stream);
}

// D[M,N] = epi(A[M,K] @ W[N,K]^T).  A row pitch lda, W row pitch ldw, D/residual row pitch ldd (floats).
// act: 0 none, 1 relu, 2 leaky(slope), 3 gelu(erf), 4 sigmoid.  row_scale[b] multiplies rows of batch b
// (rows_per_batch rows each) before the residual is added.  Replaces F.linear / 1x1 Conv2d call sites
// depthformer_swin.py:96,119,174-176,193,222 ; mmcv FFN ; hahi.py:122-165 (1x1) ; MSDA linears.
GED_API int ged_gemm_tf32(const float* A, int lda, const float* W, int ldw, float* D, int ldd, int M,
                          int N, int K, const float* bias, int act, float slope, const float* residual,
                          const float* row_scale, int rows_per_batch, float* D_pre, float drop_p,
                          unsigned drop_seed, const int* drop_step, cudaStream_t stream) {
  if (!A || !W || !D || M <= 0 || N <= 0 || K <= 0) return GED_ERR_ARG;
  if (K % 4 || drop_p < 0.f || drop_p >= 1.f || (drop_p > 0.f && ((int64_t)M * N > 0xFFFFFFFFll || (N % 4)))) return GED_ERR_SHAPE;
  GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.ntaps = 1; p.tap_off[0] = 0;
  p.bias = bias; p.residual = residual; p.row_scale = row_scale;
  p.rows_per_batch = rows_per_batch > 0 ? rows_per_batch : 1;
  p.D = D; p.D_pre = D_pre; p.ldd = ldd; p.act = act; p.slope = slope; p.conv_Hp = 0; p.conv_Wp = 0;
  p.drop_p = drop_p; p.drop_seed = drop_seed; p.drop_step = drop_step;
  return run(A, M, lda, W, ldw, p, stream);
}

// 3x3 / stride 1 / pad 1 convolution as nine shifted GEMMs.  Xpad: zero-bordered NHWC input
// [B, H+2, W+2, Cin] (ged_pad_nhwc writes it); Wk: weights [Cout][ky][kx][Cin]; Y: NHWC [B,H,W,Cout]
// with channel pitch ldy.  Replaces the nn.Conv2d(k=3,p=1) sites hahi.py:138-165, pemask_neck.py:36-42,
// dynamicpe_neck.py:497-502, densedepth_head.py:21-22, decode_head.py:391.
GED_API int ged_conv3x3_tf32(const float* Xpad, const float* Wk, float* Y, int ldy, int B, int H, int W,
                             int Cin, int Cout, const float* bias, int act, float slope,
                             cudaStream_t stream) {
  if (!Xpad || !Wk || !Y || B <= 0 || H <= 0 || W <= 0) return GED_ERR_ARG;
  if (Cin % 4) return GED_ERR_SHAPE;
  const int Hp = H + 2, Wp = W + 2;
  GemmParams p{};
  p.M = B * Hp * Wp; p.N = Cout; p.K = 9 * Cin; p.ntaps = 9;
  for (int ky = 0; ky < 3; ++ky)
    for (int kx = 0; kx < 3; ++kx) p.tap_off[ky * 3 + kx] = (ky - 1) * Wp + (kx - 1);
  p.bias = bias; p.residual = nullptr; p.row_scale = nullptr; p.rows_per_batch = 1;
  p.D = Y; p.D_pre = nullptr; p.ldd = ldy; p.act = act; p.slope = slope; p.conv_Hp = Hp; p.conv_Wp = Wp;
  // when Cin is not a multiple of 32 the K blocks of one tap would straddle into the next tap's
  // weights; the A side reads zeros there (TMA out-of-bounds fill past column Cin), so it is exact.
  return run(Xpad, (int64_t)p.M, Cin, Wk, 9 * Cin, p, stream);
}
