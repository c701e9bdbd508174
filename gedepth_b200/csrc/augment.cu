// Train-time augmentation of the 5-channel input on the device, sm_100a.
//
// Reference (CPU data-loader workers): configs/depthformer/depthformer_v.py:13-33 -
//   KBCrop -> Resize(ratio 0.5..2) -> Padding -> RandomRotate(p .5, 2.5 deg) -> RandomFlip -> RandomCrop(352, 704) -> ColorAug
//   -> Normalize  (depth/datasets/pipelines/transforms.py:149-205, 484-732, 64-109, 208-296, 299-353, 356-417, 420-481, 12-61),
// which run through mmcv 1.3.13 into OpenCV (cv2.resize, cv2.warpAffine on float32 H x W x 5 images, nearest for the depth
// map and the slope labels).  Two kernels restate that chain with OpenCV's own arithmetic, so the result is bit-identical
// to the reference's on the same inputs and the same drawn parameters (oracle/augment.py documents the algorithms and is
// pinned against cv2 and against the reference's transform classes):
//
//   ged_aug_resize_pad        KB crop window -> bilinear / nearest resize -> placement on the (zero / 255) canvas.
//                             cv2.resize(INTER_LINEAR, float32): horizontal pass then vertical pass, every product and sum
//                             rounded to fp32 separately (no FMA contraction: __fmul_rn / __fadd_rn); border columns take
//                             weight 0, border rows keep their weights; exact 2x down-scaling is the 2 x 2 average.
//   ged_aug_warp_crop_norm    rotate (cv2.warpAffine: inverse matrix in double, 1/32-pixel fixed-point coordinates with 10
//                             fractional bits, table weights, ((v0 w0 + v1 w1) + v2 w2) + v3 w3) -> flip -> crop ->
//                             colour augmentation (pow through double, the float32 / float64 casts of the numpy code) ->
//                             uint8 truncation, BGR -> RGB, (x - mean) / std as mmcv.imnormalize -> CHW planes.
// Only the 352 x 704 crop of the rotated image is ever evaluated.  HBM-bound streaming work, one thread per output pixel.
#include "common.cuh"

namespace ged {

__device__ __forceinline__ float lerp_cv(float s0, float s1, float a) {     // S0 * (1 - a) + S1 * a, three roundings
  return __fadd_rn(__fmul_rn(s0, __fsub_rn(1.f, a)), __fmul_rn(s1, a));
}

// cv2.resize tap of destination index d: source index and weight (resize.cpp: fx = (float)((dx + 0.5) * scale - 0.5))
__device__ __forceinline__ void cv_linear_tap(int d, double scale, int ssize, bool clamp, int& s, float& a) {
  const float f = (float)__dsub_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), 0.5);
  s = (int)floorf(f);
  a = __fsub_rn(f, (float)s);
  if (clamp) {
    if (s < 0) { s = 0; a = 0.f; }
    if (s >= ssize - 1) { s = ssize - 1; a = 0.f; }
  }
}

// src5: 5 planes [H0][W0]; depth / label [H0][W0]; the KB window is rows [top, top + sh), columns [left, left + sw).
struct AugResize {
  const float* src5; const float* depth; const float* label;
  float* canvas5; float* canvas_d; float* canvas_l;
  double scale_x, scale_y;
  int H0, W0, top, left, sh, sw, nw, nh, pad_x, pad_y, cw, ch, area2, pad_;
};

__device__ __forceinline__ void aug_resize_pad_body(const AugResize& a, int64_t first, int64_t stride) {
  const float* __restrict__ src5 = a.src5;
  const int W0 = a.W0, sw = a.sw, sh = a.sh, cw = a.cw;
  const int64_t total = (int64_t)a.cw * a.ch, plane0 = (int64_t)a.H0 * W0;
  for (int64_t i = first; i < total; i += stride) {
    const int y = (int)(i / cw), x = (int)(i - (int64_t)y * cw);
    const int dx = x - a.pad_x, dy = y - a.pad_y;
    float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, d = 0.f, l = 255.f;       // Padding :80-103
    if (dx >= 0 && dx < a.nw && dy >= 0 && dy < a.nh) {
      const float* base = src5 + (int64_t)a.top * W0 + a.left;
      if (a.area2) {                                                  // INTER_LINEAR -> INTER_AREA when exactly 2x down
        const int64_t o = (int64_t)(2 * dy) * W0 + 2 * dx;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          const float* p = base + c * plane0 + o;
          v[c] = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(__ldg(p), __ldg(p + 1)), __ldg(p + W0)), __ldg(p + W0 + 1)), 0.25f);
        }
      } else {
        int sx, sy;
        float ax, ay;
        cv_linear_tap(dx, a.scale_x, sw, true, sx, ax);
        cv_linear_tap(dy, a.scale_y, sh, false, sy, ay);
        const int x1 = min(sx + 1, sw - 1), y0 = min(max(sy, 0), sh - 1), y1 = min(max(sy + 1, 0), sh - 1);
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          const float* p = base + c * plane0;
          const float r0 = lerp_cv(__ldg(p + (int64_t)y0 * W0 + sx), __ldg(p + (int64_t)y0 * W0 + x1), ax);
          const float r1 = lerp_cv(__ldg(p + (int64_t)y1 * W0 + sx), __ldg(p + (int64_t)y1 * W0 + x1), ax);
          v[c] = lerp_cv(r0, r1, ay);
        }
      }
      // INTER_NEAREST: sx = min(floor(dx * scale), w - 1)
      const int qx = min((int)floor(__dmul_rn((double)dx, a.scale_x)), sw - 1), qy = min((int)floor(__dmul_rn((double)dy, a.scale_y)), sh - 1);
      const int64_t q = (int64_t)(a.top + qy) * W0 + a.left + qx;
      d = __ldg(a.depth + q);
      l = __ldg(a.label + q);
    }
#pragma unroll
    for (int c = 0; c < 5; ++c) a.canvas5[c * total + i] = v[c];
    a.canvas_d[i] = d;
    a.canvas_l[i] = l;
  }
}

__global__ void __launch_bounds__(256) aug_resize_pad_kernel(const AugResize a) {
  aug_resize_pad_body(a, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x);
}

struct AugWarp {
  double m[6];          // inverted affine matrix (cv2.warpAffine inverts the forward one in double)
  double colors[3];     // ColorAug per-channel factors (BGR order), float64 as numpy draws them
  double mean[3], inv_std[3];     // RGB order, as mmcv.imnormalize holds them
  const float* canvas5; const float* canvas_d; const float* canvas_l;
  float* img; float* depth; float* label;
  float gamma, brightness, depth_scale;
  int rotate, flip, color, crop_x, crop_y, cw, ch, out_w, out_h;
};
struct AugFrame { AugResize r; AugWarp w; };    // one frame of a batch (host-packed, copied to device memory)

__device__ __forceinline__ void aug_warp_crop_norm_body(const AugWarp& a, int64_t first, int64_t stride) {
  const float* __restrict__ canvas5 = a.canvas5;
  const float* __restrict__ canvas_d = a.canvas_d;
  const float* __restrict__ canvas_l = a.canvas_l;
  float* __restrict__ img = a.img;
  float* __restrict__ depth = a.depth;
  float* __restrict__ label = a.label;
  const int64_t total = (int64_t)a.out_w * a.out_h, cplane = (int64_t)a.cw * a.ch;
  for (int64_t i = first; i < total; i += stride) {
    const int oy = (int)(i / a.out_w), ox = (int)(i - (int64_t)oy * a.out_w);
    const int ry = a.crop_y + oy;
    int rx = a.crop_x + ox;
    if (a.flip) rx = a.cw - 1 - rx;                                   // crop of the mirrored image
    float v[5], d, l;
    if (a.rotate) {
      // imgwarp.cpp WarpAffineInvoker: AB_BITS = 10, INTER_BITS = 5
      const long long adx = llrint(__dmul_rn(__dmul_rn(a.m[0], (double)rx), 1024.0));
      const long long bdx = llrint(__dmul_rn(__dmul_rn(a.m[3], (double)rx), 1024.0));
      const long long X0 = llrint(__dmul_rn(__dadd_rn(__dmul_rn(a.m[1], (double)ry), a.m[2]), 1024.0));
      const long long Y0 = llrint(__dmul_rn(__dadd_rn(__dmul_rn(a.m[4], (double)ry), a.m[5]), 1024.0));
      {
        const long long X = (X0 + 16 + adx) >> 5, Y = (Y0 + 16 + bdx) >> 5;
        const int sx = (int)(X >> 5), sy = (int)(Y >> 5);
        const float fx = (float)(X & 31) / 32.f, fy = (float)(Y & 31) / 32.f;
        const float w0 = __fmul_rn(1.f - fy, 1.f - fx), w1 = __fmul_rn(1.f - fy, fx), w2 = __fmul_rn(fy, 1.f - fx), w3 = __fmul_rn(fy, fx);
        const bool x0ok = sx >= 0 && sx < a.cw, x1ok = sx + 1 >= 0 && sx + 1 < a.cw;
        const bool y0ok = sy >= 0 && sy < a.ch, y1ok = sy + 1 >= 0 && sy + 1 < a.ch;
        const int64_t o = (int64_t)sy * a.cw + sx;
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          const float* p = canvas5 + c * cplane + o;
          const float v0 = (x0ok && y0ok) ? __ldg(p) : 0.f, v1 = (x1ok && y0ok) ? __ldg(p + 1) : 0.f;
          const float v2 = (x0ok && y1ok) ? __ldg(p + a.cw) : 0.f, v3 = (x1ok && y1ok) ? __ldg(p + a.cw + 1) : 0.f;
          v[c] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v0, w0), __fmul_rn(v1, w1)), __fmul_rn(v2, w2)), __fmul_rn(v3, w3));
        }
      }
      {
        const long long X = (X0 + 512 + adx) >> 10, Y = (Y0 + 512 + bdx) >> 10;
        const bool ok = X >= 0 && X < a.cw && Y >= 0 && Y < a.ch;
        d = ok ? __ldg(canvas_d + Y * a.cw + X) : 0.f;                 // depth_pad_val = 0 (transforms.py:232)
        l = ok ? __ldg(canvas_l + Y * a.cw + X) : 255.f;               // "pe" in key -> 255 (:285)
      }
    } else {
      const int64_t o = (int64_t)ry * a.cw + rx;
#pragma unroll
      for (int c = 0; c < 5; ++c) v[c] = __ldg(canvas5 + c * cplane + o);
      d = __ldg(canvas_d + o);
      l = __ldg(canvas_l + o);
    }
    if (a.color) {                                                    // ColorAug :449-474 on the BGR planes
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float t = (float)pow((double)v[c], (double)a.gamma);          // float32 ** float32 (powf, correctly rounded via double)
        t = __fmul_rn(t, a.brightness);
        t = (float)__dmul_rn((double)t, a.colors[c]);                 // float32 array *= float64 array
        v[c] = fminf(fmaxf(t, 0.f), 255.f);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {                                     // Normalize :41-45: astype(uint8), BGR -> RGB, (x - mean) / std
      const float u = (float)(unsigned char)(int)v[2 - c];
      const float r1 = (float)__dsub_rn((double)u, a.mean[c]);
      img[(int64_t)c * total + i] = (float)__dmul_rn((double)r1, a.inv_std[c]);
    }
    img[3 * total + i] = v[3] > 0.f ? __fdiv_rn(v[3], a.depth_scale) : v[3];
    img[4 * total + i] = v[4];
    depth[i] = d;
    label[i] = l;
  }
}

__global__ void __launch_bounds__(256) aug_warp_crop_norm_kernel(const AugWarp a) {
  aug_warp_crop_norm_body(a, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x);
}

// batch: blockIdx.y = frame
__global__ void __launch_bounds__(256) aug_resize_pad_batch_kernel(const AugFrame* __restrict__ frames) {
  aug_resize_pad_body(frames[blockIdx.y].r, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x);
}
__global__ void __launch_bounds__(256) aug_warp_crop_norm_batch_kernel(const AugFrame* __restrict__ frames) {
  aug_warp_crop_norm_body(frames[blockIdx.y].w, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x);
}

// uint8 (H, W, 3) interleaved -> three float planes
__global__ void __launch_bounds__(256) aug_u8_to_planes_kernel(const unsigned char* __restrict__ src, float* __restrict__ dst, int64_t hw) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) dst[c * hw + i] = (float)src[i * 3 + c];
  }
}

static inline int aug_blocks(int64_t total) {
  int64_t b = (total + 255) / 256;
  return (int)(b < 148 * 16 ? (b < 1 ? 1 : b) : 148 * 16);
}

}  // namespace ged
using namespace ged;

// planes 0..2 of a (5, H, W) float32 frame <- the uint8 BGR image as cv2.imread delivers it (the loader's
// np.concatenate of the uint8 image with the float32 plane maps, loading.py:362,524-527)
GED_API int ged_aug_u8_to_planes(const unsigned char* bgr, float* planes, int H, int W, cudaStream_t stream) {
  if (!bgr || !planes || H <= 0 || W <= 0) return GED_ERR_ARG;
  const int64_t hw = (int64_t)H * W;
  aug_u8_to_planes_kernel<<<aug_blocks(hw), 256, 0, stream>>>(bgr, planes, hw);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

static int aug_fill_resize(AugResize& a, const float* src5, const float* depth, const float* label, int H0, int W0, int top, int left,
                           int sh, int sw, int nw, int nh, int pad_x, int pad_y, int cw, int ch, float* canvas5, float* canvas_d,
                           float* canvas_l) {
  if (!src5 || !depth || !label || !canvas5 || !canvas_d || !canvas_l) return GED_ERR_ARG;
  if (sh <= 0 || sw <= 0 || nw <= 0 || nh <= 0 || top < 0 || left < 0 || top + sh > H0 || left + sw > W0 || pad_x < 0 ||
      pad_y < 0 || pad_x + nw > cw || pad_y + nh > ch)
    return GED_ERR_SHAPE;
  a.src5 = src5; a.depth = depth; a.label = label; a.canvas5 = canvas5; a.canvas_d = canvas_d; a.canvas_l = canvas_l;
  a.scale_x = 1.0 / ((double)nw / (double)sw); a.scale_y = 1.0 / ((double)nh / (double)sh);
  a.H0 = H0; a.W0 = W0; a.top = top; a.left = left; a.sh = sh; a.sw = sw; a.nw = nw; a.nh = nh; a.pad_x = pad_x; a.pad_y = pad_y;
  a.cw = cw; a.ch = ch; a.area2 = (sw == 2 * nw && sh == 2 * nh) ? 1 : 0; a.pad_ = 0;
  return GED_OK;
}

static int aug_fill_warp(AugWarp& a, const float* canvas5, const float* canvas_d, const float* canvas_l, int cw, int ch,
                         const double* minv6, int rotate, int flip, int crop_x, int crop_y, int out_w, int out_h, int color,
                         float gamma, float brightness, const double* colors3, const float* mean3, const float* std3,
                         float depth_scale, float* img, float* depth, float* label) {
  if (!canvas5 || !canvas_d || !canvas_l || !img || !depth || !label || !mean3 || !std3 || (rotate && !minv6) || (color && !colors3))
    return GED_ERR_ARG;
  if (crop_x < 0 || crop_y < 0 || crop_x + out_w > cw || crop_y + out_h > ch || out_w <= 0 || out_h <= 0) return GED_ERR_SHAPE;
  for (int i = 0; i < 6; ++i) a.m[i] = rotate ? minv6[i] : 0.0;
  for (int i = 0; i < 3; ++i) {
    a.colors[i] = color ? colors3[i] : 1.0;
    a.mean[i] = (double)mean3[i];
    a.inv_std[i] = 1.0 / (double)std3[i];
  }
  a.canvas5 = canvas5; a.canvas_d = canvas_d; a.canvas_l = canvas_l; a.img = img; a.depth = depth; a.label = label;
  a.gamma = gamma; a.brightness = brightness; a.depth_scale = depth_scale;
  a.rotate = rotate; a.flip = flip; a.color = color; a.crop_x = crop_x; a.crop_y = crop_y; a.cw = cw; a.ch = ch;
  a.out_w = out_w; a.out_h = out_h;
  return GED_OK;
}

// KBCrop window (top, left, sh x sw) of src5 / depth / label -> resized to nw x nh -> placed at (pad_x, pad_y) on the
// cw x ch canvas (5 float planes + depth + label planes; background 0 / 0 / 255).
GED_API int ged_aug_resize_pad(const float* src5, const float* depth, const float* label, int H0, int W0, int top, int left,
                               int sh, int sw, int nw, int nh, int pad_x, int pad_y, int cw, int ch, float* canvas5,
                               float* canvas_d, float* canvas_l, cudaStream_t stream) {
  AugResize a;
  if (int e = aug_fill_resize(a, src5, depth, label, H0, W0, top, left, sh, sw, nw, nh, pad_x, pad_y, cw, ch, canvas5, canvas_d, canvas_l)) return e;
  aug_resize_pad_kernel<<<aug_blocks((int64_t)cw * ch), 256, 0, stream>>>(a);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

// canvas -> (rotate) -> (flip) -> crop (crop_x, crop_y, out_w x out_h) -> (colour augmentation) -> Normalize -> img (5, out_h,
// out_w), depth (out_h, out_w), label (out_h, out_w).  minv6: the INVERTED rotation matrix in double (host); colors3:
// float64 per-channel factors in BGR order; mean3 / std3 in RGB order.
GED_API int ged_aug_warp_crop_norm(const float* canvas5, const float* canvas_d, const float* canvas_l, int cw, int ch,
                                   const double* minv6, int rotate, int flip, int crop_x, int crop_y, int out_w, int out_h,
                                   int color, float gamma, float brightness, const double* colors3, const float* mean3,
                                   const float* std3, float depth_scale, float* img, float* depth, float* label,
                                   cudaStream_t stream) {
  AugWarp a;
  if (int e = aug_fill_warp(a, canvas5, canvas_d, canvas_l, cw, ch, minv6, rotate, flip, crop_x, crop_y, out_w, out_h, color, gamma,
                            brightness, colors3, mean3, std3, depth_scale, img, depth, label)) return e;
  aug_warp_crop_norm_kernel<<<aug_blocks((int64_t)out_w * out_h), 256, 0, stream>>>(a);
  GED_CHECK_LAUNCH();
  return GED_OK;
}

// ---- whole batch in two launches -----------------------------------------------------------------------------------------
// The caller owns a host array of ged_aug_frame_bytes() * B bytes, fills frame i with ged_aug_pack_frame (the union of
// the arguments of the two single-frame calls; `canvas` = 7 * cw * ch floats of workspace private to the frame) and hands it
// to ged_aug_train_batch together with a device buffer of the same size: one host -> device copy of the descriptors, then
// the resize / pad kernel and the rotate / flip / crop / colour / normalise kernel over grid (blocks, B).
GED_API int ged_aug_frame_bytes(void) { return (int)sizeof(AugFrame); }

GED_API int ged_aug_pack_frame(void* frames_host, int index, const float* src5, const float* depth, const float* label,
                               float* canvas, float* img, float* depth_out, float* label_out, int H0, int W0, int top, int left,
                               int sh, int sw, int nw, int nh, int pad_x, int pad_y, int cw, int ch, const double* minv6,
                               int rotate, int flip, int crop_x, int crop_y, int out_w, int out_h, int color, float gamma,
                               float brightness, const double* colors3, const float* mean3, const float* std3,
                               float depth_scale) {
  if (!frames_host || index < 0 || !canvas) return GED_ERR_ARG;
  AugFrame& f = reinterpret_cast<AugFrame*>(frames_host)[index];
  const int64_t n = (int64_t)cw * ch;
  if (int e = aug_fill_resize(f.r, src5, depth, label, H0, W0, top, left, sh, sw, nw, nh, pad_x, pad_y, cw, ch, canvas, canvas + 5 * n,
                              canvas + 6 * n)) return e;
  return aug_fill_warp(f.w, canvas, canvas + 5 * n, canvas + 6 * n, cw, ch, minv6, rotate, flip, crop_x, crop_y, out_w, out_h, color,
                       gamma, brightness, colors3, mean3, std3, depth_scale, img, depth_out, label_out);
}

GED_API int ged_aug_train_batch(const void* frames_host, int B, void* frames_dev, cudaStream_t stream) {
  if (!frames_host || !frames_dev || B <= 0 || B > 65535) return GED_ERR_ARG;
  const AugFrame* fh = reinterpret_cast<const AugFrame*>(frames_host);
  int64_t max_canvas = 0, max_out = 0;
  for (int i = 0; i < B; ++i) {
    const int64_t c = (int64_t)fh[i].r.cw * fh[i].r.ch, o = (int64_t)fh[i].w.out_w * fh[i].w.out_h;
    max_canvas = c > max_canvas ? c : max_canvas;
    max_out = o > max_out ? o : max_out;
  }
  if (cudaMemcpyAsync(frames_dev, frames_host, sizeof(AugFrame) * (size_t)B, cudaMemcpyHostToDevice, stream) != cudaSuccess) return GED_ERR_LAUNCH;
  const AugFrame* fd = reinterpret_cast<const AugFrame*>(frames_dev);
  int bx = aug_blocks(max_canvas);
  if ((int64_t)bx * B > 148 * 32) bx = (148 * 32 + B - 1) / B;
  aug_resize_pad_batch_kernel<<<dim3(bx, B), 256, 0, stream>>>(fd);
  int bo = aug_blocks(max_out);
  if ((int64_t)bo * B > 148 * 32) bo = (148 * 32 + B - 1) / B;
  aug_warp_crop_norm_batch_kernel<<<dim3(bo, B), 256, 0, stream>>>(fd);
  GED_CHECK_LAUNCH();
  return GED_OK;
}
