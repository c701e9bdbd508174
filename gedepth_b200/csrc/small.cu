// The narrow ends of the path, sm_100a SIMT (HBM-bound; a tensor-core tile would be > 90 % padding):
//  * backward of the 3x3 convs with 1 or 11 output channels - conv_depth 64->1 (decode_head.py:391,489), convfinal of
//    LightPEMASKNeck 64->1 (pemask_neck.py:36,64) and DynamicPENeckSOFT 64->11 (dynamicpe_neck.py:497,539): activation
//    derivative + bias gradient, dX and dW;
//  * Linear 512->2 (+sigmoid) of HAHIHeteroNeck.reference_points (hahi.py:176,299-300), forward and backward;
//  * query + positional encoding (+ level embedding) of the two deformable-attention modules (hahi.py:252-270,280-325,
//    mmcv MultiScaleDeformableAttention: `query = query + query_pos`) and its adjoint (level-embedding column sums fused
//    with the fan-in sum of the query gradient).
#include "common.cuh"

namespace ged {

__device__ __forceinline__ float small_act_grad(float y, int act, float slope) {
  switch (act) {
    case 1: return y > 0.f ? 1.f : 0.f;            // relu (y = output)
    case 2: return y > 0.f ? 1.f : slope;          // leaky relu
    case 4: return y * (1.f - y);                  // sigmoid
    default: return 1.f;
  }
}

// gz[r][c] = g[r][c] * act'(y[r][c]);  db[c] += sum_r gz[r][c].   CO <= 16 channels, one thread per row.
template <int CO>
__global__ void __launch_bounds__(256) small_act_bwd_kernel(const float* __restrict__ g, const float* __restrict__ y,
                                                            float* __restrict__ gz, float* __restrict__ db, int64_t rows,
                                                            int act, float slope) {
  __shared__ float s_db[CO];
  if (threadIdx.x < CO) s_db[threadIdx.x] = 0.f;
  __syncthreads();
  float part[CO];
#pragma unroll
  for (int c = 0; c < CO; ++c) part[c] = 0.f;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < CO; ++c) {
      const float v = g[r * CO + c] * (act ? small_act_grad(y[r * CO + c], act, slope) : 1.f);
      gz[r * CO + c] = v;
      part[c] += v;
    }
  }
  if (db) {
#pragma unroll
    for (int c = 0; c < CO; ++c) {
      const float s = warp_sum(part[c]);
      if ((threadIdx.x & 31) == 0) atomicAdd(&s_db[c], s);
    }
    __syncthreads();
    if (threadIdx.x < CO) atomicAdd(db + threadIdx.x, s_db[threadIdx.x]);
  }
}

// dX[b,y,x,ci] = sum_{co,ky,kx} gz[b, y+1-ky, x+1-kx, co] * w[co][ky][kx][ci]   (zero outside the map)
template <int CO>
__global__ void __launch_bounds__(256) conv3x3_small_dx_kernel(const float* __restrict__ gz, const float* __restrict__ w,
                                                               float* __restrict__ dx, int B, int H, int W, int CI) {
  extern __shared__ __align__(16) float s_w[];          // [CO][9][CI]
  for (int i = threadIdx.x; i < CO * 9 * CI; i += blockDim.x) s_w[i] = w[i];
  __syncthreads();
  const int quads = CI >> 2;
  const int64_t total = (int64_t)B * H * W * quads;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int qd = (int)(idx % quads);
    const int64_t p = idx / quads;
    const int x = (int)(p % W), yy = (int)((p / W) % H);
    const int64_t b = p / ((int64_t)W * H);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int sy = yy + 1 - ky;
      if (sy < 0 || sy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int sx = x + 1 - kx;
        if (sx < 0 || sx >= W) continue;
        const float* gp = gz + ((b * H + sy) * W + sx) * CO;
#pragma unroll
        for (int co = 0; co < CO; ++co) {
          const float gv = __ldg(gp + co);
          const float4 wv = *reinterpret_cast<const float4*>(s_w + (co * 9 + ky * 3 + kx) * CI + qd * 4);
          acc.x += gv * wv.x; acc.y += gv * wv.y; acc.z += gv * wv.z; acc.w += gv * wv.w;
        }
      }
    }
    *reinterpret_cast<float4*>(dx + p * CI + qd * 4) = acc;
  }
}

// dW[co][ky][kx][ci] += sum_{b,y,x} gz[b,y,x,co] * xp[b, y+ky, x+kx, ci]   (xp zero-bordered: (B, H+2, W+2, CI))
// block = (CI, 256 / CI): thread (ci, lane) walks every blockDim.y-th pixel of the block's pixel range with CO x 9
// accumulators, UNR pixels per iteration: all of their loads are issued before the first product.  (One pixel per
// iteration with 2 CTAs per SM left 10 dependent-latency loads per warp in flight: 1.3 ms per launch at 16 x 176 x 560 x 64
// for 0.4 GB of operands.)
template <int CO>
struct SmallDw {
  static constexpr int UNR = 4;                       // CO = 11: 99 accumulators + 4 x 20 operands, one CTA of 8 warps per SM
  static constexpr int MINB = CO == 1 ? 3 : (CO == 2 ? 2 : 1);
};
template <int CO>
__global__ void __launch_bounds__(256, SmallDw<CO>::MINB) conv3x3_small_dw_kernel(
    const float* __restrict__ gz, const float* __restrict__ xp, float* __restrict__ dw, int B, int H, int W, int CI,
    int rows_per_block, int sw_log2) {
  constexpr int UNR = SmallDw<CO>::UNR;
  const int DW_SW = 1 << sw_log2;       // columns per strip
  float acc[CO][9];
#pragma unroll
  for (int co = 0; co < CO; ++co)
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[co][t] = 0.f;
  // A CTA owns a strip of DW_SW (16 or 64) columns x `rows_per_block` rows of one sample and walks it row by row: the three input
  // rows a pixel row needs are re-used by the next two pixel rows while still in L1 / L2.  (Walking a flat pixel range, all
  // CTAs together kept 3 rows x 144 KB x 444 CTAs = 190 MB live - more than the L2 - and every tap row came from DRAM again:
  // 1.18 GB of DRAM traffic for a 0.41 GB input.)
  const int ci = threadIdx.x;
  const int Wp = W + 2, Hp = H + 2;
  const int strips = (W + DW_SW - 1) / DW_SW, ychunks = (H + rows_per_block - 1) / rows_per_block;
  int bid = blockIdx.x;
  const int sx_i = bid % strips; bid /= strips;
  const int yc = bid % ychunks;
  const int64_t b = bid / ychunks;
  const int xs = sx_i * DW_SW, y0 = yc * rows_per_block, y1 = min(H, y0 + rows_per_block);
  const int n_local = (y1 - y0) * DW_SW;
  const int step = blockDim.y;
  for (int pb = threadIdx.y; pb < n_local; pb += step * UNR) {
    float gv[UNR][CO], xv[UNR][9];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int pl = pb + u * step;
      const int yl = pl >> sw_log2, xl = pl - (yl << sw_log2);
      const bool ok = pl < n_local && xs + xl < W;
      const int x = ok ? xs + xl : xs, yy = ok ? y0 + yl : y0;      // a valid address; its products are discarded through gv = 0
      const int64_t p = (b * H + yy) * W + x;
#pragma unroll
      for (int co = 0; co < CO; ++co) gv[u][co] = ok ? __ldg(gz + p * CO + co) : 0.f;
      const float* xb = xp + ((b * Hp + yy) * Wp + x) * CI + ci;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) xv[u][ky * 3 + kx] = __ldg(xb + ((int64_t)ky * Wp + kx) * CI);
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u)
#pragma unroll
      for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int co = 0; co < CO; ++co) acc[co][t] += gv[u][co] * xv[u][t];
  }
#pragma unroll
  for (int co = 0; co < CO; ++co)
#pragma unroll
    for (int t = 0; t < 9; ++t) atomicAdd(dw + (int64_t)(co * 9 + t) * CI + ci, acc[co][t]);
}

// y[m][n] = act(x[m][:] . w[n][:] + b[n]),  N <= 4: one warp per row
template <int N>
__global__ void __launch_bounds__(256) linear_small_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                               const float* __restrict__ b, float* __restrict__ y, int64_t M,
                                                               int K, int act) {
  const int lane = threadIdx.x & 31;
  const int64_t m = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (m >= M) return;
  float acc[N];
#pragma unroll
  for (int n = 0; n < N; ++n) acc[n] = 0.f;
  for (int k4 = lane; k4 < K / 4; k4 += 32) {
    const float4 xv = ldg_stream(reinterpret_cast<const float4*>(x + m * K) + k4);
#pragma unroll
    for (int n = 0; n < N; ++n) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(w + (int64_t)n * K) + k4);
      acc[n] += xv.x * wv.x + xv.y * wv.y + xv.z * wv.z + xv.w * wv.w;
    }
  }
#pragma unroll
  for (int n = 0; n < N; ++n) {
    float v = warp_sum(acc[n]) + (b ? __ldg(b + n) : 0.f);
    if (act == 4) v = 1.f / (1.f + __expf(-v));
    if (lane == 0) y[m * N + n] = v;
  }
}

// gz = g * act'(y);  dw[n][k] += sum_m gz[m][n] x[m][k];  db[n] += sum_m gz[m][n];  dx[m][k] = sum_n gz[m][n] w[n][k]
// block: one thread per 4 consecutive k, a chunk of rows per block
template <int N>
__global__ void __launch_bounds__(256) linear_small_bwd_kernel(const float* __restrict__ g, const float* __restrict__ y,
                                                               const float* __restrict__ x, const float* __restrict__ w,
                                                               float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db,
                                                               int64_t M, int K, int act, int64_t rows_per_block) {
  const int k4 = threadIdx.x;
  if (k4 >= K / 4) return;
  float4 acc[N], wv[N];
  float dbp[N];
#pragma unroll
  for (int n = 0; n < N; ++n) {
    acc[n] = make_float4(0.f, 0.f, 0.f, 0.f);
    wv[n] = __ldg(reinterpret_cast<const float4*>(w + (int64_t)n * K) + k4);
    dbp[n] = 0.f;
  }
  const int64_t m0 = (int64_t)blockIdx.x * rows_per_block, m1 = min(M, m0 + rows_per_block);
  for (int64_t m = m0; m < m1; ++m) {
    const float4 xv = ldg_stream(reinterpret_cast<const float4*>(x + m * K) + k4);
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int n = 0; n < N; ++n) {
      const float gzv = __ldg(g + m * N + n) * (act ? small_act_grad(__ldg(y + m * N + n), act, 0.f) : 1.f);
      acc[n].x += gzv * xv.x; acc[n].y += gzv * xv.y; acc[n].z += gzv * xv.z; acc[n].w += gzv * xv.w;
      d.x += gzv * wv[n].x; d.y += gzv * wv[n].y; d.z += gzv * wv[n].z; d.w += gzv * wv[n].w;
      dbp[n] += gzv;
    }
    if (dx) *(reinterpret_cast<float4*>(dx + m * K) + k4) = d;
  }
#pragma unroll
  for (int n = 0; n < N; ++n) {
    if (dw) atomicAdd(reinterpret_cast<float4*>(dw + (int64_t)n * K) + k4, acc[n]);
    if (db && k4 == 0) atomicAdd(db + n, dbp[n]);
  }
}

struct LevelStarts {
  int start[5];          // token index where each of the four levels begins, start[4] = S
};
__device__ __forceinline__ int level_of(const LevelStarts& ls, int s) { return (s >= ls.start[1]) + (s >= ls.start[2]) + (s >= ls.start[3]); }

// q[b][s][:] = query[b][s][:] + pos[s][:] (+ level_embed[level(s)][:])
// IT = unsigned when the element count fits 32 bits (always on this path): a 64-bit division costs ~5x a 32-bit one and
// there are two per float4
template <typename IT>
__global__ void __launch_bounds__(256) add_pos_kernel(const float4* __restrict__ query, const float4* __restrict__ pos,
                                                      const float4* __restrict__ level_embed, float4* __restrict__ q, int B, int S,
                                                      int C4, LevelStarts ls) {
  const IT total = (IT)B * (IT)S * (IT)C4;
  for (IT i = (IT)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (IT)gridDim.x * blockDim.x) {
    const IT row = i / (IT)C4;
    const int c = (int)(i - row * (IT)C4);
    const int s = (int)(row % (IT)S);
    float4 v = ldg_stream(query + i);
    const float4 pv = __ldg(pos + (int64_t)s * C4 + c);
    v.x += pv.x; v.y += pv.y; v.z += pv.z; v.w += pv.w;
    if (level_embed) {
      const float4 e = __ldg(level_embed + level_of(ls, s) * C4 + c);
      v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
    }
    stg_stream(q + i, v);
  }
}

// Adjoint of add_pos fused with the fan-in of the query gradient: g_le[level(s)][:] += sum_b dq[b][s][:] and, when
// `extra` is given, dq += extra in place.  block = (C4, rows): a block walks a contiguous chunk of token rows.
__global__ void __launch_bounds__(256) add_pos_bwd_kernel(float4* __restrict__ dq, const float4* __restrict__ extra,
                                                          float* __restrict__ g_le, int B, int S, int C4, LevelStarts ls,
                                                          int rows_per_block) {
  const int c = threadIdx.x;          // blockDim.x == C4
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min((int64_t)B * S, r0 + rows_per_block);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int cur = -1;
  for (int64_t r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
    const int lv = level_of(ls, (int)(r % S));
    if (lv != cur) {
      if (cur >= 0 && g_le) atomicAdd(reinterpret_cast<float4*>(g_le) + cur * C4 + c, acc);
      acc = make_float4(0.f, 0.f, 0.f, 0.f);
      cur = lv;
    }
    float4 v = dq[r * C4 + c];
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    if (extra) {
      const float4 e = ldg_stream(extra + r * C4 + c);
      v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
      dq[r * C4 + c] = v;
    }
  }
  if (cur >= 0 && g_le) atomicAdd(reinterpret_cast<float4*>(g_le) + cur * C4 + c, acc);
}

static inline int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }

}  // namespace ged
using namespace ged;

#define SMALL_CO_SWITCH(CO, ...)          \
  switch (CO) {                            \
    case 1: { constexpr int C_ = 1; __VA_ARGS__; break; }   \
    case 2: { constexpr int C_ = 2; __VA_ARGS__; break; }   \
    case 11: { constexpr int C_ = 11; __VA_ARGS__; break; } \
    default: return GED_ERR_SHAPE;         \
  }

// Backward of a 3x3 / stride 1 / pad 1 conv with Cout in {1, 2, 11}.  g, y (output after `act`, may be NULL when act == 0),
// gz (workspace = g's size): (B,H,W,Cout) contiguous; xp (B,H+2,W+2,Cin) zero-bordered input; w [Cout][3][3][Cin].
// dx (B,H,W,Cin) overwritten or NULL; dw [Cout][3][3][Cin] and db [Cout] ACCUMULATED or NULL.
GED_API int ged_conv3x3_small_bwd(const float* g, const float* y, float* gz, const float* xp, const float* w, float* dx,
                                  float* dw, float* db, int B, int H, int W, int Cin, int Cout, int act, float slope,
                                  cudaStream_t stream) {
  if (!g || !gz || !w || (act && !y) || (dw && !xp) || B <= 0 || H <= 0 || W <= 0) return GED_ERR_ARG;
  if (Cin % 4 || Cin > 256 || Cin < 4) return GED_ERR_SHAPE;
  if ((dx && !aligned16(dx)) || !aligned16(w)) return GED_ERR_ALIGN;
  const int64_t rows = (int64_t)B * H * W;
  const int blocks = (int)imin64((rows + 255) / 256, 148 * 8);
  SMALL_CO_SWITCH(Cout, (small_act_bwd_kernel<C_><<<blocks, 256, 0, stream>>>(g, y, gz, db, rows, act, slope)));
  if (dx) {
    const size_t smem = (size_t)Cout * 9 * Cin * sizeof(float);
    const int dblocks = (int)imin64((rows * (Cin / 4) + 255) / 256, 148 * 8);
    SMALL_CO_SWITCH(Cout, {
      if (smem > 48 * 1024 && cudaFuncSetAttribute(conv3x3_small_dx_kernel<C_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return GED_ERR_LAUNCH;
      conv3x3_small_dx_kernel<C_><<<dblocks, 256, smem, stream>>>(gz, w, dx, B, H, W, Cin);
    });
  }
  if (dw) {
    const int ny = max(1, 256 / Cin);
    // CTAs = B x strips x row chunks.  Cout = 11 (248 registers, one CTA per SM, 6336 atomics per CTA at the end): 64-column
    // strips and about ONE wave of CTAs; Cout <= 2 (three / two CTAs per SM): 16-column strips and about four waves, so that
    // the last wave's idle SMs cost little.
    const int sw_log2 = Cout > 2 ? 6 : 4;
    const int target = Cout > 2 ? 148 : 148 * (Cout == 1 ? 3 : 2) * 4;
    const int strips = cdiv(W, 1 << sw_log2);
    int ychunks = max(1, min(H, cdiv(target, B * strips)));
    const int rpb = cdiv(H, ychunks);
    ychunks = cdiv(H, rpb);
    const int wblocks = B * strips * ychunks;
    SMALL_CO_SWITCH(Cout, (conv3x3_small_dw_kernel<C_><<<wblocks, dim3(Cin, ny), 0, stream>>>(gz, xp, dw, B, H, W, Cin, rpb, sw_log2)));
  }
  GED_CHECK_LAUNCH();
  return GED_OK;
}

#define SMALL_N_SWITCH(N, ...)            \
  switch (N) {                             \
    case 1: { constexpr int N_ = 1; __VA_ARGS__; break; }   \
    case 2: { constexpr int N_ = 2; __VA_ARGS__; break; }   \
    case 3: { constexpr int N_ = 3; __VA_ARGS__; break; }   \
    case 4: { constexpr int N_ = 4; __VA_ARGS__; break; }   \
    default: return GED_ERR_SHAPE;         \
  }

// y[M][N] = act(x[M][K] @ w[N][K]^T + b), N <= 4, K % 4 == 0; act: 0 none, 4 sigmoid
GED_API int ged_linear_small_fwd(const float* x, const float* w, const float* b, float* y, int64_t M, int N, int K, int act,
                                 cudaStream_t stream) {
  if (!x || !w || !y || M <= 0 || K <= 0) return GED_ERR_ARG;
  if (K % 4 || (act != 0 && act != 4)) return GED_ERR_SHAPE;
  if (!aligned16(x) || !aligned16(w)) return GED_ERR_ALIGN;
  const int64_t blocks = (M * 32 + 255) / 256;
  SMALL_N_SWITCH(N, (linear_small_fwd_kernel<N_><<<(unsigned)blocks, 256, 0, stream>>>(x, w, b, y, M, K, act)));
  GED_CHECK_LAUNCH();
  return GED_OK;
}

// dw [N][K] and db [N] ACCUMULATED (or NULL); dx [M][K] overwritten (or NULL); y = forward output (needed when act != 0)
GED_API int ged_linear_small_bwd(const float* g, const float* y, const float* x, const float* w, float* dx, float* dw,
                                 float* db, int64_t M, int N, int K, int act, cudaStream_t stream) {
  if (!g || !x || !w || (act && !y) || M <= 0 || K <= 0) return GED_ERR_ARG;
  if (K % 4 || K > 1024 || (act != 0 && act != 4)) return GED_ERR_SHAPE;
  if (!aligned16(x) || !aligned16(w) || (dx && !aligned16(dx)) || (dw && !aligned16(dw))) return GED_ERR_ALIGN;
  const int blocks = (int)imin64(M, 148 * 4);
  const int64_t per = (M + blocks - 1) / blocks;
  SMALL_N_SWITCH(N, (linear_small_bwd_kernel<N_><<<blocks, 256, 0, stream>>>(g, y, x, w, dx, dw, db, M, K, act, per)));
  GED_CHECK_LAUNCH();
  return GED_OK;
}

// q (B,S,C) = query (B,S,C) + pos (S,C) [+ level_embed (4,C) of the level token s belongs to]; level_start[5]
GED_API int ged_add_pos_fwd(const float* query, const float* pos, const float* level_embed, const int* level_start, float* q,
                            int B, int S, int C, cudaStream_t stream) {
  if (!query || !pos || !q || (level_embed && !level_start) || B <= 0 || S <= 0) return GED_ERR_ARG;
  if (C % 4) return GED_ERR_SHAPE;
  if (!aligned16(query) || !aligned16(pos) || !aligned16(q) || (level_embed && !aligned16(level_embed))) return GED_ERR_ALIGN;
  LevelStarts ls{};
  for (int i = 0; i < 5; ++i) ls.start[i] = level_start ? level_start[i] : (i == 0 ? 0 : S);
  const int64_t total = (int64_t)B * S * (C / 4);
  const int blocks = (int)imin64((total + 255) / 256, 148 * 16);
  if (total + (int64_t)blocks * 256 < (1ll << 32))
    add_pos_kernel<unsigned><<<blocks, 256, 0, stream>>>((const float4*)query, (const float4*)pos, (const float4*)level_embed, (float4*)q,
                                                       B, S, C / 4, ls);
  else
    add_pos_kernel<int64_t><<<blocks, 256, 0, stream>>>((const float4*)query, (const float4*)pos, (const float4*)level_embed, (float4*)q,
                                                      B, S, C / 4, ls);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

// g_level_embed (4,C) += per-level column sums of dq (B,S,C) (NULL: skip); dq += extra in place (NULL: skip)
GED_API int ged_add_pos_bwd(float* dq, const float* extra, float* g_level_embed, const int* level_start, int B, int S, int C,
                            cudaStream_t stream) {
  if (!dq || (!extra && !g_level_embed) || (g_level_embed && !level_start) || B <= 0 || S <= 0) return GED_ERR_ARG;
  if (C % 4 || C / 4 > 256) return GED_ERR_SHAPE;
  if (!aligned16(dq) || (extra && !aligned16(extra)) || (g_level_embed && !aligned16(g_level_embed))) return GED_ERR_ALIGN;
  LevelStarts ls{};
  for (int i = 0; i < 5; ++i) ls.start[i] = level_start ? level_start[i] : (i == 0 ? 0 : S);
  const int C4 = C / 4, ny = max(1, 256 / C4);
  const int64_t rows = (int64_t)B * S;
  const int blocks = (int)imin64((rows + 63) / 64, 148 * 8);
  const int per = (int)((rows + blocks - 1) / blocks);
  add_pos_bwd_kernel<<<blocks, dim3(C4, ny), 0, stream>>>((float4*)dq, (const float4*)extra, g_level_embed, B, S, C4, ls, per);
  GED_CHECK_LAUNCH();
  return GED_OK;
}
