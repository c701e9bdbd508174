// Train-mode BatchNorm2d (+ReLU) over channels-last feature maps viewed as (rows = B*H*W, C), sm_100a.
// a4 / a11: the stem's bn1 (depthformer_swin.py:1040-1041,1153) and the 15 ConvModule BNs of HAHIHeteroNeck
// (hahi.py:122-165); per-GPU batch statistics exactly as the reference (SyncBN is never activated),
// eps 1e-5, momentum 0.1, running_var updated with the unbiased estimate.  HBM-bound column reductions:
// statistics in fp64 partials, a thread owns 4 consecutive channels (128-bit accesses).
#include "common.cuh"

namespace ged {

constexpr int BN_ROWS_PER_BLOCK = 256;

// sums[0:C] += sum_r x, sums[C:2C] += sum_r x^2         grid (ceil(C/128), row chunks), block (32, 8)
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ x, double* __restrict__ sums,
                                                        int64_t rows, int C) {
  __shared__ float4 s1[8][32], s2[8][32];
  const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
  const int64_t r0 = (int64_t)blockIdx.y * BN_ROWS_PER_BLOCK, r1 = min(rows, r0 + BN_ROWS_PER_BLOCK);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), q = a;
  if (c < C)
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
      const float4 v = __ldg((const float4*)(x + r * C + c));
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      q.x += v.x * v.x; q.y += v.y * v.y; q.z += v.z * v.z; q.w += v.w * v.w;
    }
  s1[threadIdx.y][threadIdx.x] = a; s2[threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    double A[4] = {0, 0, 0, 0}, Q[4] = {0, 0, 0, 0};
    for (int i = 0; i < 8; ++i) {
      const float4 t = s1[i][threadIdx.x], u = s2[i][threadIdx.x];
      A[0] += t.x; A[1] += t.y; A[2] += t.z; A[3] += t.w;
      Q[0] += u.x; Q[1] += u.y; Q[2] += u.z; Q[3] += u.w;
    }
    for (int e = 0; e < 4; ++e) { atomicAdd(sums + c + e, A[e]); atomicAdd(sums + C + c + e, Q[e]); }
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, float* __restrict__ mean, float* __restrict__ rstd,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, int64_t rows, int C,
                                   float eps, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double n = (double)rows, m = sums[c] / n;
  double var = sums[C + c] / n - m * m;
  if (var < 0) var = 0;
  mean[c] = (float)m;
  rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(var * (n / (n > 1 ? n - 1 : 1)));
  }
}

// IT = unsigned when the element count fits 32 bits (a 64-bit modulo per float4 otherwise)
template <typename IT>
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ b, const float* __restrict__ mean,
                                                        const float* __restrict__ rstd, float* __restrict__ y,
                                                        int64_t total4_, int C, int relu) {
  const IT C4 = (IT)(C >> 2), total4 = (IT)total4_;
  for (IT i = (IT)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (IT)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const float4 v = __ldg((const float4*)x + i), mu = __ldg((const float4*)(mean + c)), rs = __ldg((const float4*)(rstd + c));
    const float4 ww = __ldg((const float4*)(w + c)), bb = __ldg((const float4*)(b + c));
    float4 o;
    o.x = (v.x - mu.x) * rs.x * ww.x + bb.x; o.y = (v.y - mu.y) * rs.y * ww.y + bb.y;
    o.z = (v.z - mu.z) * rs.z * ww.z + bb.z; o.w = (v.w - mu.w) * rs.w * ww.w + bb.w;
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    *((float4*)y + i) = o;
  }
}

// sums[0:C] += sum gz (= db), sums[C:2C] += sum gz*xhat (= dw);  gz = g * [y > 0] when relu
__global__ void __launch_bounds__(256) bn_bwd_sums_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                           const float* __restrict__ y, const float* __restrict__ mean,
                                                           const float* __restrict__ rstd, double* __restrict__ sums,
                                                           int64_t rows, int C, int relu) {
  __shared__ float4 s1[8][32], s2[8][32];
  const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
  const int64_t r0 = (int64_t)blockIdx.y * BN_ROWS_PER_BLOCK, r1 = min(rows, r0 + BN_ROWS_PER_BLOCK);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), q = a;
  if (c < C) {
    const float4 mu = __ldg((const float4*)(mean + c)), rs = __ldg((const float4*)(rstd + c));
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
      float4 gg = __ldg((const float4*)(g + r * C + c));
      const float4 v = __ldg((const float4*)(x + r * C + c));
      if (relu) {
        const float4 yy = __ldg((const float4*)(y + r * C + c));
        gg.x = yy.x > 0.f ? gg.x : 0.f; gg.y = yy.y > 0.f ? gg.y : 0.f; gg.z = yy.z > 0.f ? gg.z : 0.f; gg.w = yy.w > 0.f ? gg.w : 0.f;
      }
      a.x += gg.x; a.y += gg.y; a.z += gg.z; a.w += gg.w;
      q.x += gg.x * (v.x - mu.x) * rs.x; q.y += gg.y * (v.y - mu.y) * rs.y;
      q.z += gg.z * (v.z - mu.z) * rs.z; q.w += gg.w * (v.w - mu.w) * rs.w;
    }
  }
  s1[threadIdx.y][threadIdx.x] = a; s2[threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    double A[4] = {0, 0, 0, 0}, Q[4] = {0, 0, 0, 0};
    for (int i = 0; i < 8; ++i) {
      const float4 t = s1[i][threadIdx.x], u = s2[i][threadIdx.x];
      A[0] += t.x; A[1] += t.y; A[2] += t.z; A[3] += t.w;
      Q[0] += u.x; Q[1] += u.y; Q[2] += u.z; Q[3] += u.w;
    }
    for (int e = 0; e < 4; ++e) { atomicAdd(sums + c + e, A[e]); atomicAdd(sums + C + c + e, Q[e]); }
  }
}

// dx = w * rstd * (gz - db/n - xhat * dw/n);  also writes db, dw (float) from the fp64 sums (thread 0..C-1 of block 0)
template <typename IT>
__global__ void __launch_bounds__(256) bn_bwd_dx_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                         const float* __restrict__ y, const float* __restrict__ w,
                                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                                         const double* __restrict__ sums, float* __restrict__ dx,
                                                         float* __restrict__ dw, float* __restrict__ db, int64_t rows,
                                                         int C, int relu) {
  const IT C4 = (IT)(C >> 2), total4 = (IT)rows * C4;
  const float invn = 1.f / (float)rows;
  if (blockIdx.x == 0)
    for (int c = threadIdx.x; c < C; c += blockDim.x) { db[c] = (float)sums[c]; dw[c] = (float)sums[C + c]; }
  for (IT i = (IT)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (IT)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    float4 gg = __ldg((const float4*)g + i);
    const float4 v = __ldg((const float4*)x + i);
    if (relu) {
      const float4 yy = __ldg((const float4*)y + i);
      gg.x = yy.x > 0.f ? gg.x : 0.f; gg.y = yy.y > 0.f ? gg.y : 0.f; gg.z = yy.z > 0.f ? gg.z : 0.f; gg.w = yy.w > 0.f ? gg.w : 0.f;
    }
    const float4 mu = __ldg((const float4*)(mean + c)), rs = __ldg((const float4*)(rstd + c)), ww = __ldg((const float4*)(w + c));
    const float sb[4] = {(float)sums[c] * invn, (float)sums[c + 1] * invn, (float)sums[c + 2] * invn, (float)sums[c + 3] * invn};
    const float sw[4] = {(float)sums[C + c] * invn, (float)sums[C + c + 1] * invn, (float)sums[C + c + 2] * invn, (float)sums[C + c + 3] * invn};
    float4 o;
    o.x = ww.x * rs.x * (gg.x - sb[0] - (v.x - mu.x) * rs.x * sw[0]);
    o.y = ww.y * rs.y * (gg.y - sb[1] - (v.y - mu.y) * rs.y * sw[1]);
    o.z = ww.z * rs.z * (gg.z - sb[2] - (v.z - mu.z) * rs.z * sw[2]);
    o.w = ww.w * rs.w * (gg.w - sb[3] - (v.w - mu.w) * rs.w * sw[3]);
    *((float4*)dx + i) = o;
  }
}

}  // namespace ged
using namespace ged;

static inline unsigned bn_grid(int64_t total) {
  const int64_t b = (total + 255) / 256;
  return (unsigned)(b < 148 * 16 ? b : 148 * 16);
}

// x, y: (rows, C) channels-last; sums: double[2C] scratch; save_mean / save_rstd: float[C] kept for the backward;
// running_mean / running_var updated in place (may be NULL).
GED_API int ged_bn_train_fwd(const float* x, const float* w, const float* b, float* running_mean, float* running_var,
                             float* y, float* save_mean, float* save_rstd, double* sums, int64_t rows, int C, float eps,
                             float momentum, int relu, cudaStream_t stream) {
  if (!x || !w || !b || !y || !save_mean || !save_rstd || !sums || rows <= 0 || C <= 0) return GED_ERR_ARG;
  if (C % 4) return GED_ERR_SHAPE;
  if (!aligned16(x) || !aligned16(y) || !aligned16(w) || !aligned16(b) || !aligned16(save_mean) || !aligned16(save_rstd)) return GED_ERR_ALIGN;
  if (cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, stream) != cudaSuccess) return GED_ERR_LAUNCH;
  dim3 grid(cdiv(C, 128), (unsigned)((rows + BN_ROWS_PER_BLOCK - 1) / BN_ROWS_PER_BLOCK));
  bn_stats_kernel<<<grid, dim3(32, 8), 0, stream>>>(x, sums, rows, C);
  bn_finalize_kernel<<<cdiv(C, 128), 128, 0, stream>>>(sums, save_mean, save_rstd, running_mean, running_var, rows, C, eps, momentum);
  const int64_t total4 = rows * (C / 4);
  if (total4 + (int64_t)bn_grid(total4) * 256 < (1ll << 32))
    bn_apply_kernel<unsigned><<<bn_grid(total4), 256, 0, stream>>>(x, w, b, save_mean, save_rstd, y, total4, C, relu);
  else
    bn_apply_kernel<int64_t><<<bn_grid(total4), 256, 0, stream>>>(x, w, b, save_mean, save_rstd, y, total4, C, relu);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

// g: gradient w.r.t. the (ReLU'd) output y.  dx overwritten; dw, db overwritten (float[C]).
GED_API int ged_bn_train_bwd(const float* g, const float* x, const float* y, const float* w, const float* save_mean,
                             const float* save_rstd, float* dx, float* dw, float* db, double* sums, int64_t rows, int C,
                             int relu, cudaStream_t stream) {
  if (!g || !x || !w || !save_mean || !save_rstd || !dx || !dw || !db || !sums || rows <= 0 || (relu && !y)) return GED_ERR_ARG;
  if (C % 4) return GED_ERR_SHAPE;
  if (!aligned16(g) || !aligned16(x) || !aligned16(dx) || (y && !aligned16(y))) return GED_ERR_ALIGN;
  if (cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, stream) != cudaSuccess) return GED_ERR_LAUNCH;
  dim3 grid(cdiv(C, 128), (unsigned)((rows + BN_ROWS_PER_BLOCK - 1) / BN_ROWS_PER_BLOCK));
  bn_bwd_sums_kernel<<<grid, dim3(32, 8), 0, stream>>>(g, x, y, save_mean, save_rstd, sums, rows, C, relu);
  const int64_t total4 = rows * (C / 4);
  if (total4 + (int64_t)bn_grid(total4) * 256 < (1ll << 32))
    bn_bwd_dx_kernel<unsigned><<<bn_grid(total4), 256, 0, stream>>>(g, x, y, w, save_mean, save_rstd, sums, dx, dw, db, rows, C, relu);
  else
    bn_bwd_dx_kernel<int64_t><<<bn_grid(total4), 256, 0, stream>>>(g, x, y, w, save_mean, save_rstd, sums, dx, dw, db, rows, C, relu);
  GED_CHECK_LAUNCH();
  return GED_OK;
}
