// On-device evaluation: the nine depth metrics of depth/core/evaluation/metrics.py:8-45
// over the evaluation mask of depth/datasets/kitti.py:355-385 (min/max depth AND the Garg / Eigen crop rectangle),
// and the flip-TTA average of depth/models/depther/encoder_decoder.py:249-274 - without the per-image
// .cpu().numpy() round trip of depth/apis/test.py:209-218.  HBM-bound: 8 B / pixel read, 80 B / image written.
#include "common.cuh"

namespace ged {

constexpr int MET_N = 10;   // count, #a1, #a2, #a3, sum|g-p|/g, sum (g-p)^2/g, sum (g-p)^2, sum (ln g - ln p)^2, sum (ln p - ln g), sum |log10 g - log10 p|

__global__ void __launch_bounds__(256) depth_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                             double* __restrict__ sums, int H, int W, int y0, int y1,
                                                             int x0, int x1, float min_depth, float max_depth) {
  __shared__ double s_red[8][MET_N];
  const int b = blockIdx.y;
  const int64_t HW = (int64_t)H * W;
  double acc[MET_N];
#pragma unroll
  for (int k = 0; k < MET_N; ++k) acc[k] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (int64_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / W), x = (int)(i - (int64_t)y * W);
    if (y < y0 || y >= y1 || x < x0 || x >= x1) continue;
    const float g = __ldg(gt + b * HW + i);
    if (!(g > min_depth && g < max_depth)) continue;
    const float p = __ldg(pred + b * HW + i);
    // float32 arithmetic per element as numpy does on float32 arrays; the means are accumulated in fp64
    const float thresh = fmaxf(g / p, p / g);
    const float d = g - p;
    const float lg = logf(g), lp = logf(p);
    acc[0] += 1.0;
    acc[1] += thresh < 1.25f ? 1.0 : 0.0;
    acc[2] += thresh < 1.5625f ? 1.0 : 0.0;
    acc[3] += thresh < 1.953125f ? 1.0 : 0.0;
    acc[4] += (double)(fabsf(d) / g);
    acc[5] += (double)((d * d) / g);
    acc[6] += (double)(d * d);
    acc[7] += (double)((lg - lp) * (lg - lp));
    acc[8] += (double)(lp - lg);
    acc[9] += (double)fabsf(log10f(g) - log10f(p));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < MET_N; ++k) {
    const double v = warp_sum(acc[k]);
    if (lane == 0) s_red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < MET_N) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += s_red[w][threadIdx.x];
    atomicAdd(sums + (int64_t)b * MET_N + threadIdx.x, v);
  }
}

// out = 0.5 * (a + hflip(b))   (B,H,W): prediction of the plain image + un-flipped prediction of the mirrored one
__global__ void __launch_bounds__(256) tta_merge_kernel(const float* __restrict__ a, const float* __restrict__ bflip,
                                                         float* __restrict__ out, int W, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / W;
    const int x = (int)(i - row * W);
    out[i] = 0.5f * (__ldg(a + i) + __ldg(bflip + row * W + (W - 1 - x)));
  }
}

// Test-time input (configs/depthformer/depthformer_v.py:33-53): KBCrop window (transforms.py:176-197) -> optional
// horizontal flip (RandomFlip) -> Normalize = mmcv.imnormalize [external]: BGR->RGB, then cv2.subtract / cv2.multiply
// on a float32 image with float64 scalars, i.e. each op in double, rounded to float32 (verified against cv2).
// src: uint8 (H0, W0, 3) BGR as cv2 / mmcv.imfrombytes deliver it; dst: planes 0..2 of a (5, H, W) float32 image.
__global__ void __launch_bounds__(256) rgb_crop_normalize_kernel(const uint8_t* __restrict__ src, int W0, int top, int left,
                                                                  int flip, int to_rgb, double m0, double m1, double m2,
                                                                  double i0, double i1, double i2, float* __restrict__ dst,
                                                                  int H, int W) {
  const int64_t total = (int64_t)H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / W), x = (int)(i - (int64_t)y * W);
    const int sx = left + (flip ? W - 1 - x : x);
    const uint8_t* px = src + ((int64_t)(top + y) * W0 + sx) * 3;
    const double mean[3] = {m0, m1, m2}, inv[3] = {i0, i1, i2};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = (float)px[to_rgb ? 2 - c : c];
      const float r1 = (float)((double)v - mean[c]);
      dst[(int64_t)c * total + i] = (float)((double)r1 * inv[c]);
    }
  }
}

}  // namespace ged
using namespace ged;

// dst (5,H,W): planes 0..2 <- normalised RGB of the crop window [top, top+H) x [left, left+W) of the uint8 BGR image
// (H0, W0, 3); flip = horizontal mirror of the cropped image.  mean / std: 3 values each (as in img_norm_cfg).
// Planes 3 / 4 (the ground-plane channels) come from ged_ground_plane with u0 = left (or left+W-1, su = -1 when flipped).
GED_API int ged_rgb_crop_normalize(const unsigned char* bgr, int H0, int W0, int top, int left, int flip, int to_rgb,
                                   const float* mean3, const float* std3, float* dst, int H, int W, cudaStream_t stream) {
  if (!bgr || !mean3 || !std3 || !dst || H <= 0 || W <= 0) return GED_ERR_ARG;
  if (top < 0 || left < 0 || top + H > H0 || left + W > W0) return GED_ERR_SHAPE;
  const int64_t total = (int64_t)H * W;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  rgb_crop_normalize_kernel<<<blocks, 256, 0, stream>>>(bgr, W0, top, left, flip, to_rgb, (double)mean3[0], (double)mean3[1],
                                                       (double)mean3[2], 1.0 / (double)std3[0], 1.0 / (double)std3[1],
                                                       1.0 / (double)std3[2], dst, H, W);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

// sums (B,10) fp64 must be zeroed by the caller (accumulated, so several crops / chunks can be added up).
// Mask: y0 <= y < y1, x0 <= x < x1 and min_depth < gt < max_depth (kitti.py:366-385, metrics.py:35-41).
GED_API int ged_depth_metrics(const float* pred, const float* gt, double* sums, int B, int H, int W, int y0, int y1,
                              int x0, int x1, float min_depth, float max_depth, cudaStream_t stream) {
  if (!pred || !gt || !sums || B <= 0 || H <= 0 || W <= 0) return GED_ERR_ARG;
  const int64_t HW = (int64_t)H * W;
  int blocks = (int)((HW + 256 * 8 - 1) / (256 * 8));
  if (blocks > 296) blocks = 296;
  if (blocks < 1) blocks = 1;
  depth_metrics_kernel<<<dim3(blocks, B), 256, 0, stream>>>(pred, gt, sums, H, W, y0, y1, x0, x1, min_depth, max_depth);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

GED_API int ged_tta_merge(const float* a, const float* b_flipped, float* out, int B, int H, int W, cudaStream_t stream) {
  if (!a || !b_flipped || !out || B <= 0 || H <= 0 || W <= 0) return GED_ERR_ARG;
  const int64_t total = (int64_t)B * H * W;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  tta_merge_kernel<<<blocks, 256, 0, stream>>>(a, b_flipped, out, W, total);
  GED_CHECK_LAUNCH();
  return GED_OK;
}
