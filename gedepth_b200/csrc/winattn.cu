// Swin (shifted-)window attention core, sm_100a.
// a7 + a8: depthformer_swin.py:285-360 (pad to x7, roll(-3,-3), 9-region mask of -100, partition,
// reverse, un-roll, crop) and :184-224 (q*scale @ k^T + rel-pos-bias (+mask), softmax, @ v).
//
// The QKV and proj linears commute with the window shuffle and run as GEMMs on the image-ordered
// token matrix; this kernel is only the 49x49 core.  One CTA per (window, head) gathers its 49
// tokens from the image-ordered qkv matrix BY COORDINATE - padding, cyclic shift, partition,
// reverse, un-shift and crop are index arithmetic here, not the 4-6 full-tensor copies per block of
// the reference.  Zero-padded tokens are real keys whose q = k = v = the qkv bias;
// the shift mask is evaluated from region labels.  head_dim = 32, window = 7 (every GE config).
// < 1 % of the model's FLOPs: SIMT fp32, exact softmax.
#include "common.cuh"

namespace ged {

constexpr int WS = 7, WN = 49, HD = 32, KP = 36, SP = 49;   // KP: 16-byte aligned row pitch (float4 reads, conflict-free
                                                            // for 8 consecutive rows); SP: odd pitch of the 49x49 tiles
constexpr int WA_THREADS = 256;

struct WinGeom {
  int H, W, Hp, Wp, nWx, shift;
};

// token n of window (wy,wx) -> image token index, or -1 for a padding token; also its mask label
__device__ __forceinline__ int token_index(const WinGeom& g, int wy, int wx, int n, int& label) {
  const int ty = n / WS, tx = n - ty * WS;
  const int hs = wy * WS + ty, ws = wx * WS + tx;            // coordinates in the rolled, padded map
  label = (hs < g.Hp - WS ? 0 : (hs < g.Hp - g.shift ? 1 : 2)) * 3 +
          (ws < g.Wp - WS ? 0 : (ws < g.Wp - g.shift ? 1 : 2));
  int h = hs + g.shift, w = ws + g.shift;                    // undo roll(-shift)
  if (h >= g.Hp) h -= g.Hp;
  if (w >= g.Wp) w -= g.Wp;
  return (h < g.H && w < g.W) ? h * g.W + w : -1;
}

__device__ __forceinline__ void load_qkv(const float* __restrict__ qkv, const float* __restrict__ bias,
                                         int64_t batch_off, int C, int head, const int* s_tok,
                                         float (*s_q)[KP], float (*s_k)[KP], float (*s_v)[KP], float scale) {
  // 49 tokens x 3 x 32 floats; consecutive threads read consecutive channels (128 B segments)
  for (int i = threadIdx.x; i < WN * 3 * HD; i += WA_THREADS) {
    const int n = i / (3 * HD), r = i - n * 3 * HD, which = r / HD, d = r - which * HD;
    const int col = which * C + head * HD + d;
    const int t = s_tok[n];
    const float v = t >= 0 ? __ldg(qkv + batch_off + (int64_t)t * 3 * C + col) : (bias ? __ldg(bias + col) : 0.f);
    if (which == 0) s_q[n][d] = v * scale; else if (which == 1) s_k[n][d] = v; else s_v[n][d] = v;
  }
}

// 49x49 product a b^T (rows of 32 floats, pitch KP) with a 2 x 4 register tile per thread: rows {ti, ti+25},
// columns {tj, tj+13, tj+26, tj+39}.  Neighbouring lanes take neighbouring columns, so the float4 reads of b are
// bank-conflict free and a is a 2-3 address broadcast; per 8 outputs a thread issues 6 LDS.128 per 4 k instead of 16.
template <class Store>
__device__ __forceinline__ void tile_abt(const float (*a)[KP], const float (*b)[KP], Store store) {
  for (int t = threadIdx.x; t < 25 * 13; t += WA_THREADS) {
    const int ti = t / 13, tj = t - ti * 13;
    const int i1 = min(ti + 25, WN - 1);
    int j[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) j[c] = min(tj + 13 * c, WN - 1);
    float acc[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
#pragma unroll
    for (int d = 0; d < HD / 4; ++d) {
      const float4 a0 = ((const float4*)a[ti])[d], a1 = ((const float4*)a[i1])[d];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 y = ((const float4*)b[j[c]])[d];
        acc[0][c] += a0.x * y.x + a0.y * y.y + a0.z * y.z + a0.w * y.w;
        acc[1][c] += a1.x * y.x + a1.y * y.y + a1.z * y.z + a1.w * y.w;
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (tj + 13 * c >= WN) break;
      store(ti, tj + 13 * c, acc[0][c]);
      if (ti + 25 < WN) store(ti + 25, tj + 13 * c, acc[1][c]);
    }
  }
}

// S = q k^T + bias + mask, then row softmax, in place in s_s
__device__ __forceinline__ void scores_softmax(const float (*s_q)[KP], const float (*s_k)[KP],
                                               float (*s_s)[SP], const float* __restrict__ table,
                                               const long long* __restrict__ index, int nH, int head,
                                               const int* s_lab, bool masked) {
  tile_abt(s_q, s_k, [&](int i, int j, float acc) {
    acc += __ldg(table + (int64_t)__ldg(index + i * WN + j) * nH + head);
    if (masked && s_lab[i] != s_lab[j]) acc += -100.0f;
    s_s[i][j] = acc;
  });
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < WN; i += WA_THREADS / 32) {
    const float a = s_s[i][lane], b = lane + 32 < WN ? s_s[i][lane + 32] : -INFINITY;
    const float mx = warp_max(fmaxf(a, b));
    const float ea = expf(a - mx), eb = lane + 32 < WN ? expf(b - mx) : 0.f;
    const float inv = 1.f / warp_sum(ea + eb);
    s_s[i][lane] = ea * inv;
    if (lane + 32 < WN) s_s[i][lane + 32] = eb * inv;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(WA_THREADS, 4) winattn_fwd_kernel(
    const float* __restrict__ qkv, const float* __restrict__ bias, const float* __restrict__ table,
    const long long* __restrict__ index, float* __restrict__ ctx, WinGeom g, int C, int nH, float scale) {
  __shared__ __align__(16) float s_q[WN][KP], s_k[WN][KP], s_v[WN][KP];
  __shared__ float s_s[WN][SP];
  __shared__ int s_tok[WN], s_lab[WN];
  const int win = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int wy = win / g.nWx, wx = win - wy * g.nWx;
  if (threadIdx.x < WN) { int lab; s_tok[threadIdx.x] = token_index(g, wy, wx, threadIdx.x, lab); s_lab[threadIdx.x] = lab; }
  __syncthreads();
  const int64_t L = (int64_t)g.H * g.W;
  load_qkv(qkv, bias, (int64_t)b * L * 3 * C, C, head, s_tok, s_q, s_k, s_v, scale);
  __syncthreads();
  scores_softmax(s_q, s_k, s_s, table, index, nH, head, s_lab, g.shift > 0);
  for (int e = threadIdx.x; e < WN * (HD / 4); e += WA_THREADS) {
    const int i = e >> 3, d4 = e & 7;
    const int t = s_tok[i];
    if (t < 0) continue;                       // padded rows are cropped away (:354-355)
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 7
    for (int j = 0; j < WN; ++j) {
      const float pij = s_s[i][j];
      const float4 v = ((const float4*)s_v[j])[d4];
      acc.x += pij * v.x; acc.y += pij * v.y; acc.z += pij * v.z; acc.w += pij * v.w;
    }
    *(float4*)(ctx + ((int64_t)b * L + t) * C + head * HD + d4 * 4) = acc;
  }
}

// Backward: recompute P; dV = P^T dO; dP = dO V^T; dS = P o (dP - rowsum(dP o P));
// dQ = dS K * scale; dK = dS^T (Q*scale); d table[index] += dS.  Padded tokens send their dK, dV
// to the qkv-bias gradient (their k, v ARE the bias).
__global__ void __launch_bounds__(WA_THREADS, 4) winattn_bwd_kernel(
    const float* __restrict__ qkv, const float* __restrict__ bias, const float* __restrict__ table,
    const long long* __restrict__ index, const float* __restrict__ g_ctx, float* __restrict__ g_qkv,
    float* __restrict__ g_bias, float* __restrict__ g_table, WinGeom g, int C, int nH, float scale) {
  __shared__ __align__(16) float s_q[WN][KP], s_k[WN][KP], s_v[WN][KP], s_o[WN][KP];
  __shared__ float s_s[WN][SP], s_d[WN][SP];
  __shared__ float s_tab[(2 * WS - 1) * (2 * WS - 1)];
  __shared__ int s_tok[WN], s_lab[WN];
  const int win = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int wy = win / g.nWx, wx = win - wy * g.nWx;
  if (threadIdx.x < WN) { int lab; s_tok[threadIdx.x] = token_index(g, wy, wx, threadIdx.x, lab); s_lab[threadIdx.x] = lab; }
  for (int i = threadIdx.x; i < (2 * WS - 1) * (2 * WS - 1); i += WA_THREADS) s_tab[i] = 0.f;
  __syncthreads();
  const int64_t L = (int64_t)g.H * g.W;
  load_qkv(qkv, bias, (int64_t)b * L * 3 * C, C, head, s_tok, s_q, s_k, s_v, scale);
  for (int e = threadIdx.x; e < WN * HD; e += WA_THREADS) {
    const int i = e >> 5, d = e & 31, t = s_tok[i];
    s_o[i][d] = t >= 0 ? __ldg(g_ctx + ((int64_t)b * L + t) * C + head * HD + d) : 0.f;
  }
  __syncthreads();
  scores_softmax(s_q, s_k, s_s, table, index, nH, head, s_lab, g.shift > 0);
  // dP -> s_d
  tile_abt(s_o, s_v, [&](int i, int j, float acc) { s_d[i][j] = acc; });
  __syncthreads();
  // dS = P o (dP - sum_j dP P), in place in s_d; bias-table gradient
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = warp; i < WN; i += WA_THREADS / 32) {
      const float pa = s_s[i][lane], da = s_d[i][lane];
      const float pb = lane + 32 < WN ? s_s[i][lane + 32] : 0.f, dbv = lane + 32 < WN ? s_d[i][lane + 32] : 0.f;
      const float dot = warp_sum(pa * da + pb * dbv);
      const float sa = pa * (da - dot), sb = pb * (dbv - dot);
      s_d[i][lane] = sa;
      atomicAdd(&s_tab[(int)__ldg(index + i * WN + lane)], sa);
      if (lane + 32 < WN) { s_d[i][lane + 32] = sb; atomicAdd(&s_tab[(int)__ldg(index + i * WN + lane + 32)], sb); }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (2 * WS - 1) * (2 * WS - 1); i += WA_THREADS)
    if (s_tab[i] != 0.f) atomicAdd(g_table + (int64_t)i * nH + head, s_tab[i]);
  // dQ, dK, dV
  const int64_t boff = (int64_t)b * L * 3 * C;
  for (int e = threadIdx.x; e < WN * (HD / 4); e += WA_THREADS) {
    const int n = e >> 3, d4 = e & 7, t = s_tok[n];
    float4 dq = make_float4(0.f, 0.f, 0.f, 0.f), dk = dq, dv = dq;
#pragma unroll 7
    for (int j = 0; j < WN; ++j) {
      const float a = s_d[n][j], bq = s_d[j][n], c = s_s[j][n];
      const float4 k = ((const float4*)s_k[j])[d4], q = ((const float4*)s_q[j])[d4], o = ((const float4*)s_o[j])[d4];
      dq.x += a * k.x; dq.y += a * k.y; dq.z += a * k.z; dq.w += a * k.w;
      dk.x += bq * q.x; dk.y += bq * q.y; dk.z += bq * q.z; dk.w += bq * q.w;
      dv.x += c * o.x; dv.y += c * o.y; dv.z += c * o.z; dv.w += c * o.w;
    }
    const int col = head * HD + d4 * 4;
    if (t >= 0) {
      float* gp = g_qkv + boff + (int64_t)t * 3 * C;
      *(float4*)(gp + col) = make_float4(dq.x * scale, dq.y * scale, dq.z * scale, dq.w * scale);
      *(float4*)(gp + C + col) = dk;
      *(float4*)(gp + 2 * C + col) = dv;
    } else if (g_bias) {
      atomicAdd(g_bias + C + col, dk.x); atomicAdd(g_bias + C + col + 1, dk.y);
      atomicAdd(g_bias + C + col + 2, dk.z); atomicAdd(g_bias + C + col + 3, dk.w);
      atomicAdd(g_bias + 2 * C + col, dv.x); atomicAdd(g_bias + 2 * C + col + 1, dv.y);
      atomicAdd(g_bias + 2 * C + col + 2, dv.z); atomicAdd(g_bias + 2 * C + col + 3, dv.w);
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Backward with the five 49 x 49 x 32 products on the tensor cores through warp-level mma.sync (m16n8k8, TF32 operands by
// truncation, fp32 accumulate) - the arithmetic of every other backward GEMM (one pass TF32); GEDEPTH_BWD_GEMM_PASSES=3 keeps
// the fp32 SIMT kernel above.  Same CTA-per-(window, head) structure (3 CTAs per SM hide the gather latency): the tiles
// are far below the 128-row tcgen05 shape - the tcgen05 version (winattn_tc.cu) needs 220 KB of operand tiles per pair of
// windows and leaves the SM 91 % idle - while the SIMT kernel is bound by shared-memory loads (6 wavefronts per 384 FMAs).
// Here a warp owns a 16-row block of each product and feeds fragments straight from the padded (64-row) tiles.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int MR = 64;            // padded token count (rows 49..63 are zero)
constexpr int MP = 36;            // pitch of the 32-float q / k / v / dO rows: fragment loads [row g][col t] hit 32 distinct banks
constexpr int MS = 68;            // pitch of the 64 x 64 P / dS tiles
constexpr int WM_SMEM = (4 * MR * MP + 2 * MR * MS) * 4;

__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// A fragment (16 x 8) of a row-major tile: element (m, k) at t[m * pitch + k]
__device__ __forceinline__ void frag_a_rm(uint32_t (&a)[4], const float* t, int pitch, int m0, int k0, int g, int tq) {
  a[0] = __float_as_uint(t[(m0 + g) * pitch + k0 + tq]);
  a[1] = __float_as_uint(t[(m0 + g + 8) * pitch + k0 + tq]);
  a[2] = __float_as_uint(t[(m0 + g) * pitch + k0 + tq + 4]);
  a[3] = __float_as_uint(t[(m0 + g + 8) * pitch + k0 + tq + 4]);
}
// A fragment of the TRANSPOSE of a row-major tile: element (m, k) at t[k * pitch + m]
__device__ __forceinline__ void frag_a_tr(uint32_t (&a)[4], const float* t, int pitch, int m0, int k0, int g, int tq) {
  a[0] = __float_as_uint(t[(k0 + tq) * pitch + m0 + g]);
  a[1] = __float_as_uint(t[(k0 + tq) * pitch + m0 + g + 8]);
  a[2] = __float_as_uint(t[(k0 + tq + 4) * pitch + m0 + g]);
  a[3] = __float_as_uint(t[(k0 + tq + 4) * pitch + m0 + g + 8]);
}
// B fragment (8 x 8): element (k, n) at t[n * pitch + k] ("col": the operand stored as [n][k]) or at t[k * pitch + n]
__device__ __forceinline__ void frag_b_nk(uint32_t (&b)[2], const float* t, int pitch, int k0, int n0, int g, int tq) {
  b[0] = __float_as_uint(t[(n0 + g) * pitch + k0 + tq]);
  b[1] = __float_as_uint(t[(n0 + g) * pitch + k0 + tq + 4]);
}
__device__ __forceinline__ void frag_b_kn(uint32_t (&b)[2], const float* t, int pitch, int k0, int n0, int g, int tq) {
  b[0] = __float_as_uint(t[(k0 + tq) * pitch + n0 + g]);
  b[1] = __float_as_uint(t[(k0 + tq + 4) * pitch + n0 + g]);
}

// relative-position index of (query i, key j): the reference's buffer (depthformer_swin.py:168-172), or its closed form
// (dy + 6) * 13 + (dx + 6) when the host has verified that the buffer is the standard one (STD: no global lookups)
template <bool STD>
__device__ __forceinline__ int rel_index(const long long* __restrict__ index, int i, int j) {
  if (STD) {
    const int yi = i / WS, xi = i - yi * WS, yj = j / WS, xj = j - yj * WS;
    return (yi - yj + WS - 1) * (2 * WS - 1) + (xi - xj + WS - 1);
  }
  return (int)__ldg(index + i * WN + j);
}

template <bool STD>
__global__ void __launch_bounds__(WA_THREADS, 3) winattn_bwd_mma_kernel(
    const float* __restrict__ qkv, const float* __restrict__ bias, const float* __restrict__ table,
    const long long* __restrict__ index, const float* __restrict__ g_ctx, float* __restrict__ g_qkv,
    float* __restrict__ g_bias, float* __restrict__ g_table, WinGeom g, int C, int nH, float scale) {
  extern __shared__ __align__(16) float wm_smem[];
  float* s_q = wm_smem;                     // [MR][MP]  q * scale
  float* s_k = s_q + MR * MP;
  float* s_v = s_k + MR * MP;
  float* s_o = s_v + MR * MP;               // dO
  float* s_s = s_o + MR * MP;               // [MR][MS]  S -> P
  float* s_d = s_s + MR * MS;               // [MR][MS]  dP -> dS
  __shared__ float s_tab[(2 * WS - 1) * (2 * WS - 1)], s_bias[(2 * WS - 1) * (2 * WS - 1)];
  __shared__ int s_tok[WN], s_lab[WN];
  const int win = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int wy = win / g.nWx, wx = win - wy * g.nWx;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  if (tid < WN) { int lab; s_tok[tid] = token_index(g, wy, wx, tid, lab); s_lab[tid] = lab; }
  for (int i = tid; i < (2 * WS - 1) * (2 * WS - 1); i += WA_THREADS) { s_tab[i] = 0.f; s_bias[i] = __ldg(table + (int64_t)i * nH + head); }
  // padding: rows 49..63 of q / k / v / dO (the gather writes every element of rows 0..48) and the whole P / dS tiles
  for (int i = tid; i < 4 * (MR - WN) * MP / 4; i += WA_THREADS) {
    const int tIdx = i / ((MR - WN) * MP / 4), r = i - tIdx * ((MR - WN) * MP / 4);
    ((float4*)(wm_smem + tIdx * MR * MP + WN * MP))[r] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int i = tid; i < 2 * MR * MS / 4; i += WA_THREADS) ((float4*)s_s)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  const int64_t L = (int64_t)g.H * g.W, boff = (int64_t)b * L * 3 * C;
  // gather q (scaled), k, v, dO: 49 tokens x 4 x 8 float4 chunks
  for (int i = tid; i < WN * 32; i += WA_THREADS) {
    const int n = i >> 5, which = (i >> 3) & 3, c4 = i & 7, t = s_tok[n];
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (which < 3) {
      const int col = which * C + head * HD + c4 * 4;
      if (t >= 0) x = __ldg((const float4*)(qkv + boff + (int64_t)t * 3 * C + col));
      else if (bias) x = __ldg((const float4*)(bias + col));
      if (which == 0) { x.x *= scale; x.y *= scale; x.z *= scale; x.w *= scale; }
    } else if (t >= 0) {
      x = __ldg((const float4*)(g_ctx + ((int64_t)b * L + t) * C + head * HD + c4 * 4));
    }
    float* dst = (which == 0 ? s_q : which == 1 ? s_k : which == 2 ? s_v : s_o) + n * MP + c4 * 4;
    *(float4*)dst = x;
  }
  __syncthreads();
  // ---- S = q k^T (+ bias, mask) and dP = dO v^T: warp = 16 rows x 32 columns of both ---------------------------------
  {
    const int m0 = (warp >> 1) * 16, nb = (warp & 1) * 32;
    float accS[4][4], accP[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) { accS[n][e] = 0.f; accP[n][e] = 0.f; }
#pragma unroll
    for (int k0 = 0; k0 < HD; k0 += 8) {
      uint32_t aq[4], ao[4];
      frag_a_rm(aq, s_q, MP, m0, k0, gq, tq);
      frag_a_rm(ao, s_o, MP, m0, k0, gq, tq);
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        uint32_t bk[2], bv[2];
        frag_b_nk(bk, s_k, MP, k0, nb + 8 * n, gq, tq);
        frag_b_nk(bv, s_v, MP, k0, nb + 8 * n, gq, tq);
        mma_tf32_16x8x8(accS[n], aq, bk);
        mma_tf32_16x8x8(accP[n], ao, bv);
      }
    }
    const bool masked = g.shift > 0;
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = m0 + gq + (e >> 1) * 8, j = nb + 8 * n + 2 * tq + (e & 1);
        if (i < WN && j < WN) {
          float x = accS[n][e] + s_bias[rel_index<STD>(index, i, j)];
          if (masked && s_lab[i] != s_lab[j]) x += -100.0f;
          s_s[i * MS + j] = x;
          s_d[i * MS + j] = accP[n][e];
        }
      }
  }
  __syncthreads();
  // ---- softmax rows, dS = P o (dP - sum_j dP P), bias-table gradient: a warp takes TWO rows at a time so that the three
  // shuffle reductions of one row overlap those of the other -----------------------------------------------------------
  for (int i0 = 2 * warp; i0 < WN; i0 += 2 * (WA_THREADS / 32)) {
    float a[2], b2[2], da[2], dbv[2];
    const bool hi = lane + 32 < WN;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int i = min(i0 + r, WN - 1);
      a[r] = s_s[i * MS + lane]; b2[r] = hi ? s_s[i * MS + lane + 32] : -INFINITY;
      da[r] = s_d[i * MS + lane]; dbv[r] = hi ? s_d[i * MS + lane + 32] : 0.f;
    }
    float mx[2] = {fmaxf(a[0], b2[0]), fmaxf(a[1], b2[1])};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mx[0] = fmaxf(mx[0], __shfl_xor_sync(0xffffffffu, mx[0], o)); mx[1] = fmaxf(mx[1], __shfl_xor_sync(0xffffffffu, mx[1], o)); }
    float ea[2], eb[2], sum[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) { ea[r] = expf(a[r] - mx[r]); eb[r] = hi ? expf(b2[r] - mx[r]) : 0.f; sum[r] = ea[r] + eb[r]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sum[0] += __shfl_xor_sync(0xffffffffu, sum[0], o); sum[1] += __shfl_xor_sync(0xffffffffu, sum[1], o); }
    float pa[2], pb[2], dot[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) { const float inv = 1.f / sum[r]; pa[r] = ea[r] * inv; pb[r] = eb[r] * inv; dot[r] = pa[r] * da[r] + pb[r] * dbv[r]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { dot[0] += __shfl_xor_sync(0xffffffffu, dot[0], o); dot[1] += __shfl_xor_sync(0xffffffffu, dot[1], o); }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int i = i0 + r;
      if (i >= WN) break;
      const float sa = pa[r] * (da[r] - dot[r]), sb = pb[r] * (dbv[r] - dot[r]);
      s_s[i * MS + lane] = pa[r];
      s_d[i * MS + lane] = sa;
      atomicAdd(&s_tab[rel_index<STD>(index, i, lane)], sa);
      if (hi) {
        s_s[i * MS + lane + 32] = pb[r];
        s_d[i * MS + lane + 32] = sb;
        atomicAdd(&s_tab[rel_index<STD>(index, i, lane + 32)], sb);
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < (2 * WS - 1) * (2 * WS - 1); i += WA_THREADS)
    if (s_tab[i] != 0.f) atomicAdd(g_table + (int64_t)i * nH + head, s_tab[i]);
  // ---- dV = P^T dO, dQ = dS k, dK = dS^T q: warp = 16 output rows x 16 channels of all three ---------------------------
  {
    const int m0 = (warp >> 1) * 16, nb = (warp & 1) * 16;
    float accV[2][4], accQ[2][4], accK[2][4];
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) { accV[n][e] = 0.f; accQ[n][e] = 0.f; accK[n][e] = 0.f; }
#pragma unroll 2
    for (int k0 = 0; k0 < 56; k0 += 8) {          // contraction over the 49 (-> 56) tokens; rows beyond are zero
      uint32_t apt[4], asr[4], ast[4];
      frag_a_tr(apt, s_s, MS, m0, k0, gq, tq);    // P^T
      frag_a_rm(asr, s_d, MS, m0, k0, gq, tq);    // dS
      frag_a_tr(ast, s_d, MS, m0, k0, gq, tq);    // dS^T
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        uint32_t bo[2], bk[2], bq[2];
        frag_b_kn(bo, s_o, MP, k0, nb + 8 * n, gq, tq);
        frag_b_kn(bk, s_k, MP, k0, nb + 8 * n, gq, tq);
        frag_b_kn(bq, s_q, MP, k0, nb + 8 * n, gq, tq);
        mma_tf32_16x8x8(accV[n], apt, bo);
        mma_tf32_16x8x8(accQ[n], asr, bk);
        mma_tf32_16x8x8(accK[n], ast, bq);
      }
    }
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      const int r = m0 + gq + 8 * h2;
      if (r >= WN) continue;
      const int t = s_tok[r];
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        const int col = head * HD + nb + 8 * n + 2 * tq;
        const float2 dq = make_float2(accQ[n][2 * h2] * scale, accQ[n][2 * h2 + 1] * scale);
        const float2 dk = make_float2(accK[n][2 * h2], accK[n][2 * h2 + 1]);
        const float2 dv = make_float2(accV[n][2 * h2], accV[n][2 * h2 + 1]);
        if (t >= 0) {
          float* gp = g_qkv + boff + (int64_t)t * 3 * C + col;
          *(float2*)gp = dq;
          *(float2*)(gp + C) = dk;
          *(float2*)(gp + 2 * C) = dv;
        } else if (g_bias) {
          atomicAdd(g_bias + C + col, dk.x); atomicAdd(g_bias + C + col + 1, dk.y);
          atomicAdd(g_bias + 2 * C + col, dv.x); atomicAdd(g_bias + 2 * C + col + 1, dv.y);
        }
      }
    }
  }
}

}  // namespace ged
using namespace ged;

static int make_geom(int H, int W, int shift, WinGeom& g) {
  if (H <= 0 || W <= 0 || shift < 0 || shift >= WS) return GED_ERR_SHAPE;
  g.H = H; g.W = W; g.shift = shift;
  g.Hp = cdiv(H, WS) * WS; g.Wp = cdiv(W, WS) * WS; g.nWx = g.Wp / WS;
  return GED_OK;
}

// qkv (B, H*W, 3C) image order; bias (3C) or NULL; table (169, nH); index (49,49) int64; ctx (B, H*W, C)
GED_API int ged_winattn_fwd(const float* qkv, const float* qkv_bias, const float* table,
                            const long long* index, float* ctx, int B, int H, int W, int C, int nH,
                            int window, int shift, float scale, cudaStream_t stream) {
  if (!qkv || !table || !index || !ctx || B <= 0) return GED_ERR_ARG;
  if (window != WS || C != nH * HD) return GED_ERR_SHAPE;
  WinGeom g;
  if (int e = make_geom(H, W, shift, g)) return e;
  dim3 grid((g.Hp / WS) * g.nWx, nH, B);
  winattn_fwd_kernel<<<grid, WA_THREADS, 0, stream>>>(qkv, qkv_bias, table, index, ctx, g, C, nH, scale);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

// g_qkv is fully overwritten; g_bias (3C) and g_table (169,nH) are ACCUMULATED into.
GED_API int ged_winattn_bwd(const float* qkv, const float* qkv_bias, const float* table,
                            const long long* index, const float* g_ctx, float* g_qkv, float* g_bias,
                            float* g_table, int B, int H, int W, int C, int nH, int window, int shift,
                            float scale, cudaStream_t stream) {
  if (!qkv || !table || !index || !g_ctx || !g_qkv || !g_table || B <= 0) return GED_ERR_ARG;
  if (window != WS || C != nH * HD) return GED_ERR_SHAPE;
  WinGeom g;
  if (int e = make_geom(H, W, shift, g)) return e;
  dim3 grid((g.Hp / WS) * g.nWx, nH, B);
  // per device, so set on every call (cheap); a failure only costs occupancy, the launch below still reports errors
  if (cudaFuncSetAttribute(winattn_bwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100) != cudaSuccess) (void)cudaGetLastError();
  winattn_bwd_kernel<<<grid, WA_THREADS, 0, stream>>>(qkv, qkv_bias, table, index, g_ctx, g_qkv, g_bias, g_table, g, C, nH, scale);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

// ged_winattn_bwd with the five products on the tensor cores through warp-level mma.sync (one pass TF32, the arithmetic of
// the other backward GEMMs); same contract.
// std_index != 0: `index` has been verified (host) to be Swin's standard relative-position index; it is then evaluated in
// closed form instead of being read.
GED_API int ged_winattn_bwd_mma(const float* qkv, const float* qkv_bias, const float* table, const long long* index,
                                const float* g_ctx, float* g_qkv, float* g_bias, float* g_table, int B, int H, int W, int C,
                                int nH, int window, int shift, float scale, int std_index, cudaStream_t stream) {
  if (!qkv || !table || !index || !g_ctx || !g_qkv || !g_table || B <= 0) return GED_ERR_ARG;
  if (window != WS || C != nH * HD) return GED_ERR_SHAPE;
  if (!aligned16(qkv) || !aligned16(g_ctx) || !aligned16(g_qkv) || (qkv_bias && !aligned16(qkv_bias))) return GED_ERR_ALIGN;
  WinGeom g;
  if (int e = make_geom(H, W, shift, g)) return e;
  dim3 grid((g.Hp / WS) * g.nWx, nH, B);
  if (cudaFuncSetAttribute(winattn_bwd_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WM_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(winattn_bwd_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WM_SMEM) != cudaSuccess)
    return GED_ERR_LAUNCH;
  if (std_index)
    winattn_bwd_mma_kernel<true><<<grid, WA_THREADS, WM_SMEM, stream>>>(qkv, qkv_bias, table, index, g_ctx, g_qkv, g_bias, g_table, g, C, nH, scale);
  else
    winattn_bwd_mma_kernel<false><<<grid, WA_THREADS, WM_SMEM, stream>>>(qkv, qkv_bias, table, index, g_ctx, g_qkv, g_bias, g_table, g, C, nH, scale);
  GED_CHECK_LAUNCH();
  return GED_OK;
}
