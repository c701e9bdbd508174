// Shared device helpers for the GEDepth sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define GED_OK 0
#define GED_ERR_ARG (-1)
#define GED_ERR_SHAPE (-2)
#define GED_ERR_ALIGN (-3)
#define GED_ERR_LAUNCH (-4)
#define GED_ERR_WORKSPACE (-5)

#define GED_API extern "C" __attribute__((visibility("default")))

#define GED_CHECK_LAUNCH()                                   \
  do {                                                       \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) return GED_ERR_LAUNCH;           \
  } while (0)

namespace ged {

// Bilinear source taps, the arithmetic of ATen's area_pixel_compute_source_index /
// upsample_bilinear2d (what F.interpolate(mode='bilinear') runs in the reference:
// encoder_decoder.py:83,114 align_corners=False; decode_head.py:491-502 align_corners=True).
struct Tap {
  int i0, i1;
  float l0, l1;
};

__host__ __device__ __forceinline__ float resize_scale(int in, int out, bool align) {
  if (align) return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
  return (float)in / (float)out;
}

__device__ __forceinline__ Tap tap(int d, float scale, bool align, int in) {
  float s = align ? scale * (float)d : fmaxf(scale * ((float)d + 0.5f) - 0.5f, 0.f);
  Tap t;
  t.i0 = min((int)s, in - 1);
  t.i1 = t.i0 + (t.i0 < in - 1 ? 1 : 0);
  t.l1 = s - (float)t.i0;
  t.l0 = 1.f - t.l1;
  return t;
}

// Destination index range [lo, hi] whose taps can touch source index j (superset; callers
// re-evaluate tap() on every candidate, so only "no miss" matters).
__device__ __forceinline__ void adjoint_range(int j, float scale, bool align, int in, int out,
                                              int& lo, int& hi) {
  if (scale <= 0.f) { lo = 0; hi = out - 1; return; }
  float inv = 1.f / scale;
  float a, b;
  if (align) { a = ((float)j - 1.f) * inv; b = ((float)j + 1.f) * inv; }
  else { a = ((float)j - 0.5f) * inv - 0.5f; b = ((float)j + 1.5f) * inv - 0.5f; }
  lo = max(0, (int)floorf(a) - 1);
  hi = min(out - 1, (int)ceilf(b) + 1);
  if (j == 0) lo = 0;
  if (j >= in - 1) hi = out - 1;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// streaming 128-bit accesses: inputs/outputs of the per-pixel kernels are touched once
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Counter-based dropout mask shared by the GEMM epilogue (forward) and ged_dropout_bwd.  One 32-bit hash serves two
// neighbouring elements (16 bits each): element idx is kept iff its 16 bits >= round(p * 65536); the survivors are
// scaled by 1 / (1 - round(p*65536)/65536), so the estimator stays unbiased for the quantised p (0.1 -> 0.100006).
// seed = host seed + step counter (device) * odd constant, so a captured CUDA graph draws a fresh mask at every
// replay.  Replaces nn.Dropout(0.1) of mmcv's MultiScaleDeformableAttention (hahi.py:179-188 [external default]);
// same distribution, not the same stream as torch.
__host__ __device__ __forceinline__ uint32_t drop_mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) {      // 16-bit threshold
  const int t = (int)(p * 65536.f + 0.5f);
  return (uint32_t)(t < 0 ? 0 : (t > 65535 ? 65535 : t));
}
__host__ __device__ __forceinline__ float drop_scale(uint32_t thresh16) { return 65536.f / (float)(65536u - thresh16); }
// idx must be a multiple of 4: keep flags of elements idx .. idx+3 as bits 0..3
__host__ __device__ __forceinline__ uint32_t drop_keep4(uint32_t seed, uint32_t idx, uint32_t thresh16) {
  const uint32_t h0 = drop_mix((idx >> 1) * 0x9E3779B9u + seed), h1 = drop_mix(((idx >> 1) + 1u) * 0x9E3779B9u + seed);
  return ((h0 & 0xFFFFu) >= thresh16 ? 1u : 0u) | ((h0 >> 16) >= thresh16 ? 2u : 0u) |
         ((h1 & 0xFFFFu) >= thresh16 ? 4u : 0u) | ((h1 >> 16) >= thresh16 ? 8u : 0u);
}
__device__ __forceinline__ uint32_t drop_seed_eff(uint32_t seed, const int* step_dev) {
  return seed + (step_dev ? (uint32_t)__ldg(step_dev) * 0x9E3779B1u : 0u);
}

__host__ __device__ __forceinline__ int cdiv(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace ged
