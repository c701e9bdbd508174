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
