// Loss kernels, sm_100a.
//   ged_silog_*   a18+a19  decode_head.py:586-599 (bilinear to the gt size, align_corners=True) fused
//                          with sigloss.py:36-53 (g = log(pred+eps) - log(gt+eps) over gt>0;
//                          sqrt(var_unbiased(g) + lam*mean(g)^2)).  No boolean-index gather, no host sync.
//   ged_ce_*      a18      celoss.py:355-413: mean over non-ignored pixels of (logsumexp - logit[label]).
// Reductions accumulate in fp64 (three numbers per launch), finalised on the device.
#include "common.cuh"
#include "tile.cuh"

namespace ged {

// stats layout (double[8]): 0 n, 1 sum g, 2 sum g^2, 3 loss, 4 mean, 5 dLoss/d(sum-term scale) ...
__global__ void __launch_bounds__(TX * TILE_H) silog_fwd_kernel(
    const float* __restrict__ pred, const float* __restrict__ gt, double* __restrict__ stats, int H,
    int W, int hp, int wp, float sy, float sx, float eps, float max_depth, int upsample) {
  __shared__ float s_p[ST_H][ST_W];
  __shared__ double s_red[3][TX * TILE_H / 32];
  const int b = blockIdx.z, oy0 = blockIdx.y * TILE_H, ox0 = blockIdx.x * TILE_W;
  SrcWindow sw{0, 0, 0, 0};
  if (upsample) {
    sw = src_window(oy0, ox0, H, W, hp, wp, sy, sx, true);
    for (int i = threadIdx.y * TX + threadIdx.x; i < sw.h * sw.w; i += TX * TILE_H) {
      int r = i / sw.w, c = i - r * sw.w;
      s_p[r][c] = __ldg(pred + ((int64_t)b * hp + sw.y0 + r) * wp + sw.x0 + c);
    }
    __syncthreads();
  }
  const int oy = oy0 + threadIdx.y;
  double n = 0.0, s1 = 0.0, s2 = 0.0;
  if (oy < H) {
    Tap ty{0, 0, 0.f, 0.f};
    if (upsample) ty = tap(oy, sy, true, hp);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ox = ox0 + i * TX + threadIdx.x;
      if (ox >= W) continue;
      const int64_t o = ((int64_t)b * H + oy) * W + ox;
      const float g_t = __ldg(gt + o);
      if (!(g_t > 0.f) || (max_depth > 0.f && !(g_t <= max_depth))) continue;
      float p;
      if (upsample) {
        const Tap tx = tap(ox, sx, true, wp);
        const int r0 = ty.i0 - sw.y0, r1 = ty.i1 - sw.y0, c0 = tx.i0 - sw.x0, c1 = tx.i1 - sw.x0;
        p = ty.l0 * (tx.l0 * s_p[r0][c0] + tx.l1 * s_p[r0][c1]) +
            ty.l1 * (tx.l0 * s_p[r1][c0] + tx.l1 * s_p[r1][c1]);
      } else {
        p = __ldg(pred + o);
      }
      const float g = logf(p + eps) - logf(g_t + eps);
      n += 1.0; s1 += (double)g; s2 += (double)g * (double)g;
    }
  }
  n = warp_sum(n); s1 = warp_sum(s1); s2 = warp_sum(s2);
  const int tid = threadIdx.y * TX + threadIdx.x, wid = tid >> 5, lane = tid & 31;
  if (lane == 0) { s_red[0][wid] = n; s_red[1][wid] = s1; s_red[2][wid] = s2; }
  __syncthreads();
  if (tid < 3) {
    double a = 0.0;
    for (int w = 0; w < TX * TILE_H / 32; ++w) a += s_red[tid][w];
    if (a != 0.0) atomicAdd(stats + tid, a);
  }
}

__global__ void silog_finalize_kernel(double* __restrict__ stats, float* __restrict__ loss, float lam) {
  const double n = stats[0], s1 = stats[1], s2 = stats[2];
  const double mean = s1 / n;
  const double var = (s2 - s1 * s1 / n) / (n - 1.0);       // torch.var: unbiased
  const double l = sqrt(var + (double)lam * mean * mean);
  stats[3] = l; stats[4] = mean;
  *loss = (float)l;
}

// d loss / d g_i = (1/(2 loss)) * ( 2 (g_i - mean)/(n-1) + 2 lam mean / n );  d g_i / d pred_i = 1/(pred_i+eps)
__global__ void __launch_bounds__(TX * TILE_H) silog_bwd_kernel(
    const float* __restrict__ pred, const float* __restrict__ gt, const double* __restrict__ stats,
    const float* __restrict__ g_loss, float* __restrict__ g_pred, int H, int W, int hp, int wp, float sy,
    float sx, float eps, float lam, float max_depth, int upsample) {
  __shared__ float s_p[ST_H][ST_W];
  __shared__ float s_g[1][TILE_H][TILE_W + 1];
  const int b = blockIdx.z, oy0 = blockIdx.y * TILE_H, ox0 = blockIdx.x * TILE_W;
  SrcWindow sw{0, 0, 0, 0};
  if (upsample) {
    sw = src_window(oy0, ox0, H, W, hp, wp, sy, sx, true);
    for (int i = threadIdx.y * TX + threadIdx.x; i < sw.h * sw.w; i += TX * TILE_H) {
      int r = i / sw.w, c = i - r * sw.w;
      s_p[r][c] = __ldg(pred + ((int64_t)b * hp + sw.y0 + r) * wp + sw.x0 + c);
    }
    __syncthreads();
  }
  const double n = stats[0], loss = stats[3], mean = stats[4];
  const float c_var = (float)((double)__ldg(g_loss) / (loss * (n - 1.0)));
  const float c_mean = (float)((double)__ldg(g_loss) * (double)lam * mean / (loss * n));
  const float fmean = (float)mean;
  const int oy = oy0 + threadIdx.y;
  if (oy < H) {
    Tap ty{0, 0, 0.f, 0.f};
    if (upsample) ty = tap(oy, sy, true, hp);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int lx = i * TX + threadIdx.x, ox = ox0 + lx;
      float gp = 0.f;
      if (ox < W) {
        const int64_t o = ((int64_t)b * H + oy) * W + ox;
        const float g_t = __ldg(gt + o);
        if (g_t > 0.f && !(max_depth > 0.f && !(g_t <= max_depth))) {
          float p;
          if (upsample) {
            const Tap tx = tap(ox, sx, true, wp);
            const int r0 = ty.i0 - sw.y0, r1 = ty.i1 - sw.y0, c0 = tx.i0 - sw.x0, c1 = tx.i1 - sw.x0;
            p = ty.l0 * (tx.l0 * s_p[r0][c0] + tx.l1 * s_p[r0][c1]) +
                ty.l1 * (tx.l0 * s_p[r1][c0] + tx.l1 * s_p[r1][c1]);
          } else {
            p = __ldg(pred + o);
          }
          const float g = logf(p + eps) - logf(g_t + eps);
          gp = (c_var * (g - fmean) + c_mean) / (p + eps);
        }
        if (!upsample) g_pred[o] = gp;
      }
      s_g[0][threadIdx.y][lx] = gp;
    }
  }
  if (!upsample) return;
  __syncthreads();
  tile_adjoint_upsample<1>(s_g, g_pred + (int64_t)b * hp * wp, 0, oy0, ox0, H, W, hp, wp, sy, sx, true, sw);
}

// ---- cross entropy over C channels (planar NCHW logits), float labels, ignore_index ------------
template <int C>
__global__ void __launch_bounds__(256) ce_fwd_kernel(const float* __restrict__ logits,
                                                      const float* __restrict__ target,
                                                      double* __restrict__ stats, int64_t HW,
                                                      float ignore_index) {
  __shared__ double s_red[2][8];
  const int b = blockIdx.y;
  double n = 0.0, s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (int64_t)gridDim.x * blockDim.x) {
    const float t = __ldg(target + b * HW + i);
    if (t == ignore_index) continue;
    const int lab = (int)t;     // .long() truncation (decode_head.py:525)
    if (lab < 0 || lab >= C) { s += (double)NAN; continue; }   // torch's CrossEntropyLoss asserts on such a label: fail loudly (NaN loss)
    float L[C], mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < C; ++c) { L[c] = __ldg(logits + ((int64_t)b * C + c) * HW + i); mx = fmaxf(mx, L[c]); }
    float se = 0.f, lt = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { se += expf(L[c] - mx); if (c == lab) lt = L[c]; }
    n += 1.0; s += (double)((mx + logf(se)) - lt);
  }
  n = warp_sum(n); s = warp_sum(s);
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_red[0][wid] = n; s_red[1][wid] = s; }
  __syncthreads();
  if (threadIdx.x < 2) {
    double a = 0.0;
    for (int w = 0; w < 8; ++w) a += s_red[threadIdx.x][w];
    if (a != 0.0) atomicAdd(stats + threadIdx.x, a);
  }
}

__global__ void ce_finalize_kernel(const double* __restrict__ stats, float* __restrict__ loss) {
  *loss = (float)(stats[1] / stats[0]);
}

template <int C>
__global__ void __launch_bounds__(256) ce_bwd_kernel(const float* __restrict__ logits,
                                                      const float* __restrict__ target,
                                                      const double* __restrict__ stats,
                                                      const float* __restrict__ g_loss,
                                                      float* __restrict__ g_logits, int64_t HW,
                                                      float ignore_index) {
  const int b = blockIdx.y;
  const float scale = (float)((double)__ldg(g_loss) / stats[0]);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (int64_t)gridDim.x * blockDim.x) {
    const float t = __ldg(target + b * HW + i);
    const bool ign = (t == ignore_index);
    const int lab = (int)t;
    float L[C], mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < C; ++c) { L[c] = ign ? 0.f : __ldg(logits + ((int64_t)b * C + c) * HW + i); mx = fmaxf(mx, L[c]); }
    float se = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) { L[c] = expf(L[c] - mx); se += L[c]; }
    const float inv = scale / se;
#pragma unroll
    for (int c = 0; c < C; ++c)
      g_logits[((int64_t)b * C + c) * HW + i] = ign ? 0.f : (L[c] * inv - (c == lab ? scale : 0.f));
  }
}

}  // namespace ged
using namespace ged;

static inline bool window_ok(int H, int W, int hp, int wp) {
  const float sy = resize_scale(hp, H, true), sx = resize_scale(wp, W, true);
  const int th = H < TILE_H ? H : TILE_H, tw = W < TILE_W ? W : TILE_W;
  return (float)th * sy + 3.f <= (float)ST_H && (float)tw * sx + 3.f <= (float)ST_W;
}

// stats: device double[8] scratch owned by the caller (kept for the backward).
GED_API int ged_silog_fwd(const float* pred, const float* gt, double* stats, float* loss, int B, int H,
                          int W, int hp, int wp, float eps, float lam, float max_depth, int upsample,
                          cudaStream_t stream) {
  if (!pred || !gt || !stats || !loss || B <= 0) return GED_ERR_ARG;
  if (upsample && !window_ok(H, W, hp, wp)) return GED_ERR_SHAPE;
  if (!upsample && (hp != H || wp != W)) return GED_ERR_SHAPE;
  if (cudaMemsetAsync(stats, 0, 8 * sizeof(double), stream) != cudaSuccess) return GED_ERR_LAUNCH;
  const float sy = resize_scale(hp, H, true), sx = resize_scale(wp, W, true);
  dim3 block(TX, TILE_H), grid(cdiv(W, TILE_W), cdiv(H, TILE_H), B);
  silog_fwd_kernel<<<grid, block, 0, stream>>>(pred, gt, stats, H, W, hp, wp, sy, sx, eps, max_depth, upsample);
  silog_finalize_kernel<<<1, 1, 0, stream>>>(stats, loss, lam);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

GED_API int ged_silog_bwd(const float* pred, const float* gt, const double* stats, const float* g_loss,
                          float* g_pred, int B, int H, int W, int hp, int wp, float eps, float lam,
                          float max_depth, int upsample, cudaStream_t stream) {
  if (!pred || !gt || !stats || !g_loss || !g_pred || B <= 0) return GED_ERR_ARG;
  if (upsample && !window_ok(H, W, hp, wp)) return GED_ERR_SHAPE;
  const float sy = resize_scale(hp, H, true), sx = resize_scale(wp, W, true);
  if (upsample && cudaMemsetAsync(g_pred, 0, sizeof(float) * (size_t)B * hp * wp, stream) != cudaSuccess) return GED_ERR_LAUNCH;
  dim3 block(TX, TILE_H), grid(cdiv(W, TILE_W), cdiv(H, TILE_H), B);
  silog_bwd_kernel<<<grid, block, 0, stream>>>(pred, gt, stats, g_loss, g_pred, H, W, hp, wp, sy, sx, eps, lam, max_depth, upsample);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

GED_API int ged_ce_fwd(const float* logits, const float* target, double* stats, float* loss, int B, int C,
                       int H, int W, float ignore_index, cudaStream_t stream) {
  if (!logits || !target || !stats || !loss || B <= 0) return GED_ERR_ARG;
  if (C != NSLOPE) return GED_ERR_SHAPE;
  if (cudaMemsetAsync(stats, 0, 8 * sizeof(double), stream) != cudaSuccess) return GED_ERR_LAUNCH;
  const int64_t HW = (int64_t)H * W;
  const unsigned gx = (unsigned)min((int64_t)2048, (HW + 255) / 256);
  ce_fwd_kernel<NSLOPE><<<dim3(gx, B), 256, 0, stream>>>(logits, target, stats, HW, ignore_index);
  ce_finalize_kernel<<<1, 1, 0, stream>>>(stats, loss);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

GED_API int ged_ce_bwd(const float* logits, const float* target, const double* stats, const float* g_loss,
                       float* g_logits, int B, int C, int H, int W, float ignore_index,
                       cudaStream_t stream) {
  if (!logits || !target || !stats || !g_loss || !g_logits || B <= 0) return GED_ERR_ARG;
  if (C != NSLOPE) return GED_ERR_SHAPE;
  const int64_t HW = (int64_t)H * W;
  const unsigned gx = (unsigned)min((int64_t)2048, (HW + 255) / 256);
  ce_bwd_kernel<NSLOPE><<<dim3(gx, B), 256, 0, stream>>>(logits, target, stats, g_loss, g_logits, HW, ignore_index);
  GED_CHECK_LAUNCH();
  return GED_OK;
}
