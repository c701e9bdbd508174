// Optimizer-side kernels over the flat fp32 parameter / gradient arenas:
//   ged_sumsq        grad-norm^2 of the whole arena in one pass (fp64 accumulate) - the L2 clip of
//                    optimizer_config.grad_clip (configs/depthformer/depthformer_v.py:148)
//   ged_adamw_step   AdamW (lr 1e-4, betas (0.9,0.999), wd 0.01 with per-tensor decay_mult folded into a
//                    per-element wd mask segment table; depthformer_v.py:128-139) with the clip
//                    coefficient and the 1/world_size gradient average applied on the fly.
// HBM-bound: 16 B read + 12 B written per parameter.
#include "common.cuh"

namespace ged {

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, int64_t n,
                                                     double* __restrict__ out) {
  __shared__ double s_red[8];
  double acc = 0.0;
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg((const float4*)g + i);
    acc += (double)(v.x * v.x + v.y * v.y) + (double)(v.z * v.z + v.w * v.w);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float v = g[(n4 << 2) + threadIdx.x]; acc += (double)v * v; }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int w = 0; w < 8; ++w) a += s_red[w];
    atomicAdd(out, a);
  }
}

// wd_mask[i] in {0,1}: stored as one float per parameter TENSOR boundary table would need a search;
// a per-element uint8 mask costs 1 B/param and keeps the kernel branch-free.
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                     float* __restrict__ m, float* __restrict__ v,
                                                     const uint8_t* __restrict__ wd_mask, int64_t n,
                                                     const double* __restrict__ sumsq, float max_norm,
                                                     float grad_scale, float lr, float beta1, float beta2,
                                                     float eps, float wd, float bc1, float bc2,
                                                     const int* __restrict__ step_dev, const float* __restrict__ lr_dev) {
  if (lr_dev) lr = *lr_dev;   // learning rate of this step lives on the device (schedules under CUDA-graph replay)
  if (step_dev) {   // step count lives on the device (CUDA-graph replays must not bake it in)
    const float t = (float)(*step_dev);
    bc1 = 1.f - powf(beta1, t);
    bc2 = 1.f - powf(beta2, t);
  }
  // clip coefficient: max_norm / (norm + 1e-6), clamped to 1 (torch.nn.utils.clip_grad_norm_)
  float coef = grad_scale;
  if (sumsq && max_norm > 0.f) {
    const float norm = (float)sqrt(*sumsq) * grad_scale;
    coef *= fminf(max_norm / (norm + 1e-6f), 1.f);
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    float pi = p[i];
    if (wd_mask == nullptr || wd_mask[i]) pi *= 1.f - lr * wd;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

}  // namespace ged
using namespace ged;

GED_API int ged_sumsq(const float* g, int64_t n, double* out, cudaStream_t stream) {
  if (!g || !out || n <= 0) return GED_ERR_ARG;
  if (!aligned16(g)) return GED_ERR_ALIGN;
  if (cudaMemsetAsync(out, 0, sizeof(double), stream) != cudaSuccess) return GED_ERR_LAUNCH;
  const int64_t n4 = n >> 2;
  const unsigned grid = (unsigned)(n4 / 256 + 1 < 1184 ? n4 / 256 + 1 : 1184);   // 148 SMs x 8 CTAs
  sumsq_kernel<<<grid, 256, 0, stream>>>(g, n, out);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

// step >= 1, or step_dev != NULL: the 1-based step count is read from device memory at run time (for
// CUDA-graph replay); likewise lr_dev != NULL overrides lr with a device scalar (warm-up / cosine schedules keep
// working when the step is a replayed graph).  sumsq may be NULL (no clipping).  grad_scale folds the 1/world_size average.
GED_API int ged_adamw_step(float* p, const float* g, float* m, float* v, const uint8_t* wd_mask, int64_t n,
                           const double* sumsq, float max_norm, float grad_scale, float lr, float beta1,
                           float beta2, float eps, float weight_decay, int step, const int* step_dev,
                           const float* lr_dev, cudaStream_t stream) {
  if (!p || !g || !m || !v || n <= 0 || (step < 1 && !step_dev)) return GED_ERR_ARG;
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  const unsigned grid = (unsigned)(n / 256 + 1 < 2368 ? n / 256 + 1 : 2368);
  adamw_kernel<<<grid, 256, 0, stream>>>(p, g, m, v, wd_mask, n, sumsq, max_norm, grad_scale, lr, beta1, beta2, eps, weight_decay, bc1, bc2, step_dev, lr_dev);
  GED_CHECK_LAUNCH();
  return GED_OK;
}
