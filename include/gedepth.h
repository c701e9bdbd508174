/* gedepth.h - C ABI of libgedepth_sm100.so: the sm_100a kernels of the GEDepth hot path.
 *
 * The reference (qcraftai/gedepth) is pure Python on PyTorch + mmcv; it has no FFI of its own for
 * this path (its only native code, depth/models/_cdht, is dead and used pybind:
 * depth/models/_cdht/deep_hough_cuda.cpp:99-102).  Each entry point below therefore cites the
 * reference call site whose arithmetic it replaces.  A maintainer binds them with ctypes inside
 * the corresponding module's forward (INTEGRATION.md shows the stubs).
 *
 * Conventions: plain device pointers and sizes, no framework types; every function enqueues work
 * on `stream` and returns immediately with 0 (GED_OK) or a negative error code; nothing here
 * allocates, synchronises or throws.  All tensors are fp32 and contiguous in the stated layout
 * unless a stride argument says otherwise.  "accumulated" outputs must be zeroed by the caller.
 */
#ifndef GEDEPTH_H_
#define GEDEPTH_H_

#include <stdint.h>
#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GED_OK 0
#define GED_ERR_ARG (-1)       /* null pointer / non-positive size */
#define GED_ERR_SHAPE (-2)     /* shape outside what the kernel is built for */
#define GED_ERR_ALIGN (-3)     /* pointer or pitch not 16-byte aligned */
#define GED_ERR_LAUNCH (-4)    /* CUDA launch / runtime error */
#define GED_ERR_WORKSPACE (-5)

int ged_version(void);                 /* 10000*major + 100*minor + patch */
const char* ged_arch(void);            /* "sm_100a" */

/* ---- ground embedding ------------------------------------------------------------------------ */
/* tools/preprocess_data_kitti.py:47-53 (+ loading.py:388-403, transforms.py:40-48).
 * pe[v,u] = coef[0] / (coef[1]*(u0+su*x) + coef[2]*(v0+sv*y) + coef[3]) in fp64 on the integer grid;
 * ch4 <- (float)pe; ch3 <- pe with >clamp_max -> 0, <0 -> 0, then /depth_scale where >0.
 * ch3/ch4: first pixel of the channel plane of batch 0; batch_stride in floats (5*H*W inside img). */
int ged_ground_plane(float* ch3, float* ch4, int64_t batch_stride3, int64_t batch_stride4, int B, int H,
                     int W, const double* coef4, double u0, double v0, double su, double sv,
                     float depth_scale, float clamp_max, cudaStream_t stream);
/* the int64 meshgrid of preprocess_data_kitti.py:52 (u[y,x] = u0+x, v[y,x] = v0+y) */
int ged_pixel_grid(long long* u, long long* v, int H, int W, int u0, int v0, cudaStream_t stream);

/* depth/models/depther/encoder_decoder.py:112-123 (Vanilla).  pe_norm = img[:,3] (batch stride in
 * floats), y_half (B,1,h2,w2) -> y (B,1,H,W) = bilinear(align_corners=False), pe_mask = pe_norm*y*200. */
int ged_ge_vanilla_fwd(const float* pe_norm, int64_t pe_batch_stride, const float* y_half, float* y,
                       float* pe_mask, int B, int H, int W, int h2, int w2, cudaStream_t stream);
/* g_y / g_pe_mask may be NULL; g_y_half is overwritten. */
int ged_ge_vanilla_bwd(const float* pe_norm, int64_t pe_batch_stride, const float* g_y,
                       const float* g_pe_mask, float* g_y_half, int B, int H, int W, int h2, int w2,
                       cudaStream_t stream);

/* depth/models/depther/encoder_decoder.py:79-102 (Adaptive).  pe_raw = img[:,4]; logits_half
 * (B,11,h2,w2) planar; height (B) per-sample camera height or NULL -> height_scalar;
 * logits_full (B,11,H,W) may be NULL (inference). */
int ged_ge_adaptive_fwd(const float* pe_raw, int64_t pe_batch_stride, const float* y_half,
                        const float* logits_half, const float* height, float height_scalar,
                        float depth_scale, float* y, float* pe_mask, float* logits_full, int B, int H,
                        int W, int h2, int w2, cudaStream_t stream);
int ged_ge_adaptive_bwd(const float* pe_raw, int64_t pe_batch_stride, const float* y_half,
                        const float* logits_half, const float* height, float height_scalar,
                        float depth_scale, const float* g_y, const float* g_pe_mask,
                        const float* g_logits_full, float* g_y_half, float* g_logits_half, int B, int H,
                        int W, int h2, int w2, cudaStream_t stream);
/* 1 (default): when H == 2*h2 and W == 2*w2 (every GE config) the Vanilla backward and both Adaptive kernels run their
 * closed-form x2 versions (csrc/ge_adaptive_x2.cu), the forward staged by TMA when W % 8 == 0; 2: the same kernels staged by
 * per-thread asynchronous copies only; 0: generic bilinear kernels only.  Returns the previous setting. */
int ged_set_ge_x2(int on);

/* depth/models/decode_heads/decode_head.py:489-508.  d = relu(conv_depth(feat)) (B,1,h2,w2);
 * out = d*(1-y_h) + pe_h + min_depth with y_h, pe_h = bilinear(align_corners=True) of y, pe_mask. */
int ged_fuse_head_fwd(const float* d, const float* pe_mask, const float* y, float* out, float* y_h,
                      float min_depth, int B, int H, int W, int h2, int w2, cudaStream_t stream);
/* g_yh_extra: gradient arriving on the returned y_h (may be NULL). */
int ged_fuse_head_bwd(const float* g_out, const float* g_yh_extra, const float* d, const float* y_h,
                      float* g_d, float* g_pe_mask, float* g_y, int B, int H, int W, int h2, int w2,
                      cudaStream_t stream);

/* tools/preprocess_data_kitti.py:59-63,86-89 (truncate=0: round half even) and
 * tools/preprocess_data_ddad.py:47-51,77-82 (truncate=1).  k_out in {-5..5} or 255 where gt==0. */
int ged_find_k(const float* gt, const float* pe, int64_t pe_batch_stride, float* k_out, int B, int H,
               int W, double cam_height, int truncate, cudaStream_t stream);

/* ---- losses ---------------------------------------------------------------------------------- */
/* decode_head.py:586-599 + losses/sigloss.py:36-53.  upsample=1: pred is (B,1,hp,wp) and is
 * bilinearly resized (align_corners=True) to the gt size on the fly.  stats: double[8] scratch kept
 * for the backward.  max_depth <= 0 disables the upper bound. */
int ged_silog_fwd(const float* pred, const float* gt, double* stats, float* loss, int B, int H, int W,
                  int hp, int wp, float eps, float lam, float max_depth, int upsample, cudaStream_t stream);
int ged_silog_bwd(const float* pred, const float* gt, const double* stats, const float* g_loss,
                  float* g_pred, int B, int H, int W, int hp, int wp, float eps, float lam,
                  float max_depth, int upsample, cudaStream_t stream);
/* losses/celoss.py:355-413 with decode_head.py:523-525: logits (B,C=11,H,W) planar, float labels. */
int ged_ce_fwd(const float* logits, const float* target, double* stats, float* loss, int B, int C, int H,
               int W, float ignore_index, cudaStream_t stream);
int ged_ce_bwd(const float* logits, const float* target, const double* stats, const float* g_loss,
               float* g_logits, int B, int C, int H, int W, float ignore_index, cudaStream_t stream);

/* ---- Swin ------------------------------------------------------------------------------------ */
/* nn.LayerNorm sites: embed.py:299-300, depthformer_swin.py:118,463,469,1178. */
int ged_layernorm_fwd(const float* x, const float* w, const float* b, float* y, float* mean, float* rstd,
                      int64_t rows, int C, float eps, cudaStream_t stream);
/* dw / db accumulated (may both be NULL). */
/* g_add (optional): the residual-branch gradient of x, added to dx in the same pass (SwinBlock: x is both the
 * LayerNorm input and the identity, depthformer_swin.py:461-472). */
int ged_layernorm_bwd(const float* g, const float* x, const float* w, const float* mean, const float* rstd,
                      const float* g_add, float* dx, float* dw, float* db, int64_t rows, int C, cudaStream_t stream);
/* 1 (default): rows of C <= 768 channels are held in registers (one global read per operand; the backward's weight / bias
 * gradients come out of the same kernel); 0: the three-pass forward and the two-kernel backward.  Returns the previous setting. */
int ged_set_layernorm_reg(int on);
/* depthformer_swin.py:285-360 + :184-224 minus the two linears.  qkv (B,H*W,3C) image order. */
int ged_winattn_fwd(const float* qkv, const float* qkv_bias, const float* table, const long long* index,
                    float* ctx, int B, int H, int W, int C, int nH, int window, int shift, float scale,
                    cudaStream_t stream);
/* g_qkv overwritten; g_bias (3C, may be NULL) and g_table (169,nH) accumulated. */
/* ged_winattn_fwd on the tensor cores (csrc/winattn_tc.cu): two (window, head) pairs per work item, S = Q K^T and O = P V
 * as tcgen05.mma (3xTF32, accumulators in TMEM), softmax one accumulator row per thread.  The relative-position index is
 * Swin's closed form (dy + 6) * 13 + (dx + 6) (depthformer_swin.py:168-172). */
int ged_winattn_tc_fwd(const float* qkv, const float* qkv_bias, const float* table, float* ctx, int B, int H, int W, int C,
                       int nH, int window, int shift, float scale, cudaStream_t stream);
/* ged_winattn_bwd with the five 49 x 49 x 32 products on the tensor cores through warp-level mma.sync (m16n8k8 TF32, one
 * pass like the other backward GEMMs): the default backward of the path (csrc/winattn.cu). */
int ged_winattn_bwd_mma(const float* qkv, const float* qkv_bias, const float* table, const long long* index,
                        const float* g_ctx, float* g_qkv, float* g_bias, float* g_table, int B, int H, int W, int C, int nH,
                        int window, int shift, float scale, int std_index /* index verified standard: closed form */,
                        cudaStream_t stream);
/* ged_winattn_bwd on the tensor cores (csrc/winattn_tc.cu): S, dP, dV, dQ, dK as tcgen05.mma (one pass TF32, like the
 * other backward GEMMs), softmax / dS one TMEM row per thread.  Standard Swin relative-position index only. */
int ged_winattn_tc_bwd(const float* qkv, const float* qkv_bias, const float* table, const float* g_ctx, float* g_qkv,
                       float* g_bias, float* g_table, int B, int H, int W, int C, int nH, int window, int shift, float scale,
                       cudaStream_t stream);
int ged_winattn_bwd(const float* qkv, const float* qkv_bias, const float* table, const long long* index,
                    const float* g_ctx, float* g_qkv, float* g_bias, float* g_table, int B, int H, int W,
                    int C, int nH, int window, int shift, float scale, cudaStream_t stream);

/* Train-mode BatchNorm2d (+ReLU) over channels-last maps as (rows=B*H*W, C): stem bn1 (depthformer_swin.py:1040-1041)
 * and the ConvModule BNs of hahi.py:122-165.  Per-GPU batch statistics, biased variance for normalisation, unbiased
 * for running_var, momentum 0.1.  sums: double[2C] scratch.  save_mean / save_rstd (float[C]) feed the backward. */
int ged_bn_train_fwd(const float* x, const float* w, const float* b, float* running_mean, float* running_var, float* y,
                     float* save_mean, float* save_rstd, double* sums, int64_t rows, int C, float eps, float momentum,
                     int relu, cudaStream_t stream);
int ged_bn_train_bwd(const float* g, const float* x, const float* y, const float* w, const float* save_mean,
                     const float* save_rstd, float* dx, float* dw, float* db, double* sums, int64_t rows, int C, int relu,
                     cudaStream_t stream);

/* ---- tensor-core GEMM / conv (tcgen05, TF32) --------------------------------------------------- */
/* D[M,N] = epi(A[M,K] @ W[N,K]^T): F.linear / 1x1 Conv2d sites depthformer_swin.py:96,119,174-176,
 * 193,222; mmcv FFN; hahi.py:122-165; MSDA linears.  act: 0 none 1 relu 2 leaky 3 gelu 4 sigmoid.
 * D_pre (optional, pitch ldd): copy of the pre-activation x+bias, kept for the GELU derivative. */
int ged_gemm_tf32(const float* A, int lda, const float* W, int ldw, float* D, int ldd, int M, int N,
                  int K, const float* bias, int act, float slope, const float* residual,
                  const float* row_scale, int rows_per_batch, float* D_pre, float drop_p, unsigned drop_seed,
                  const int* drop_step, cudaStream_t stream);
/* drop_p > 0: dropout on the activated output before row scale / residual (nn.Dropout(0.1) after output_proj in
 * mmcv's MultiScaleDeformableAttention [external]); counter-based mask from (drop_seed, *drop_step, element index).
 * ged_dropout_bwd re-draws the same mask: gz = g * keep / (1-p), db += column sums (db may be NULL). */
int ged_dropout_bwd(const float* g, int64_t ldg, float* gz, float* db, int64_t rows, int N, float drop_p,
                    unsigned drop_seed, const int* drop_step, cudaStream_t stream);
/* dX of a linear / 1x1 conv: D[M,N] = A[M,K] @ Wt[K,N], Wt = the forward weight [N_out=K][K_in=N] read in place
 * as an MN-major UMMA operand (no transposed copy; torch.autograd's grad_output @ weight). */
int ged_gemm_tf32_bt(const float* A, int lda, const float* Wt, int ldw, float* D, int ldd, int M, int N, int K,
                     cudaStream_t stream);
/* D = A @ Wt + R: the same GEMM with a residual of row pitch ldr (0 = D's; R may be D itself: accumulate in place; R may be
 * a channel slice of a wider gradient), so that a gradient fanning in from several consumers is summed in the GEMM epilogues. */
int ged_gemm_tf32_bt_acc(const float* A, int lda, const float* Wt, int ldw, float* D, int ldd, int M, int N, int K,
                         const float* residual, int ldr, cudaStream_t stream);
/* dX of the 3x3/s1/p1 conv from the zero-bordered dY [B,H+2,W+2,Cout] and the FORWARD weights [Cout][3][3][Cin]
 * read in place (cuDNN dgrad in the reference); DX [B,H,W,*] with channel pitch ldx. */
int ged_conv3x3_dx_tf32(const float* Gpad, const float* Wk, float* DX, int ldx, int B, int H, int W, int Cin,
                        int Cout, cudaStream_t stream);
/* Process-wide GEMM/conv arithmetic: 3 (default) = error-compensated 3xTF32 (each fp32 operand split into
 * tf32 hi + lo, three tcgen05 MMAs per k-step: fp32-accurate, the parity mode); 1 = single-pass TF32 (what
 * PyTorch 1.8 / cuDNN run by default on Ampere+ for the reference).  Returns the previous value. */
int ged_set_gemm_precision(int passes);
/* 1 (default) = large forward / dX problems run on CTA pairs (tcgen05 cta_group::2: a 256 x BN tile per two SMs,
 * each staging half of the operands), 0 = single-CTA kernels only.  Returns the previous value. */
int ged_set_gemm_pair(int on);
/* 1 (default): weight-gradient GEMMs with >= 256 output rows run on the CTA-pair kernel; 0: single-CTA kernels only */
int ged_set_gemm_pair_dw(int on);
/* 1 (default): the 3xTF32 single-CTA kernels (tiles <= 128 wide) keep the A operand (hi and lo) in tensor memory; 0: shared memory */
int ged_set_gemm_a_tmem(int on);
/* 1 (default) = allow 192/256-column output tiles, 0 = at most 128.  Returns the previous value. */
int ged_set_gemm_wide_tiles(int on);
/* Weight gradients (autograd of every nn.Linear / Conv2d on the path; the reference gets them from cuBLAS /
 * cuDNN wgrad through torch.autograd):  D[n*ldd + t*tap_dstride + k] += sum_p G[p][n] * X[p + tap_off[t]][k]
 * for n < N, k < K, t < ntaps (1: linear / 1x1, 9: 3x3 with tap_off = (ky-1)*(W+2) + (kx-1) over zero-bordered
 * NHWC operands); rows of X outside [0,Px) read as zero.  D holds the running gradient (or zeros): the
 * contraction over P is split across SMs and partial tiles are added with vector atomics.
 * G [P][N] pitch ldg, X [Px][K] pitch ldx exactly as the forward stored them: MN-major UMMA operands
 * (SWIZZLE_128B_BASE32B, TMA panels of 32 features), no transposed copies. */
int ged_gemm_dw_tf32(const float* G, int ldg, const float* X, int ldx, float* D, int ldd, int N, int K,
                     int64_t P, int64_t Px, int ntaps, const int* tap_off, int tap_dstride, cudaStream_t stream);
/* 3x3/s1/p1 conv, NHWC: hahi.py:138-165, pemask_neck.py:36-42, dynamicpe_neck.py:497-502,
 * densedepth_head.py:21-22, decode_head.py:391.  Xpad [B,H+2,W+2,Cin] zero-bordered;
 * Wk [Cout][3][3][Cin]; Y [B,H,W,*] with channel pitch ldy. */
int ged_conv3x3_tf32(const float* Xpad, const float* Wk, float* Y, int ldy, int B, int H, int W, int Cin,
                     int Cout, const float* bias, int act, float slope, cudaStream_t stream);

/* ---- data movement around the convs (NHWC fp32) ------------------------------------------------ */
/* 1 (default): ged_prep_conv_input / ged_upsample_nhwc_bwd run one CTA per output row with the bilinear taps tabulated once
 * in shared memory; 0: the flat grid-stride kernels (also taken for rows wider than 1024, tap tables beyond 40 KB or ratios above x10).  Returns the
 * previous setting. */
int ged_set_layout_rows(int on);
/* dst [B,H+2,W+2,C0+C1] = zero border | [bilinear(src0 (B,h0,w0,C0) -> HxW, align_corners=True), src1 (B,H,W,C1)]:
 * F.interpolate + torch.cat + padding of densedepth_head.py:24-27 / hahi.py:329-353 in one pass.  src*_bstride: batch stride
 * in floats, 0 = dense (a source may be one level's slice of the (B, S, C) token tensor: hahi.py:338-353). */
int ged_prep_conv_input(const float* src0, int C0, int h0, int w0, const float* src1, int C1, float* dst, int B,
                        int H, int W, int64_t src0_bstride, int64_t src1_bstride, cudaStream_t stream);
/* out (B,h0,w0,C0) = resize^T of channels [0,C0) of g (B,H,W,ldg). */
int ged_upsample_nhwc_bwd(const float* g, int ldg, float* out, int C0, int B, int H, int W, int h0, int w0,
                          cudaStream_t stream);
/* out (B,H,W,C) = base + bilinear(t (B,h0,w0,C) -> HxW, align_corners=True); out may alias base: pemask_neck.py:52-63. */
int ged_resize_add_nhwc(const float* t, const float* base, float* out, int C, int B, int H, int W, int h0, int w0,
                        cudaStream_t stream);
/* gz = g * act'(ref) * row_scale[row / rows_per_batch]; db[c] += column sums (db may be NULL).  g has row pitch ldg
 * (a channel slice of a wider gradient is read in place); gz / ref are dense [rows][N].
 * ref = layer output for relu(1)/leaky(2)/sigmoid(4), pre-activation for gelu(3); act 0: copy/scale only. */
int ged_act_bwd(const float* g, int64_t ldg, const float* ref, float* gz, float* db, const float* row_scale,
                int rows_per_batch, int64_t rows, int N, int act, float slope, cudaStream_t stream);

/* tokens (B, ceil(H/P)*ceil(W/P), Cin*P*P) in Conv2d weight order from channels [0,Cin) of an NCHW batch, zero
 * padded bottom/right (embed.py:282-297): the patch-embedding conv becomes ged_gemm_tf32. */
int ged_patchify(const float* img, int64_t batch_stride, float* tok, int B, int Cin, int H, int W, int P,
                 cudaStream_t stream);
/* im2col of a strided/padded conv over channels [0,Cin) of an NCHW batch: tok (B*Ho*Wo, Kp), column
 * k = (c*kh+ky)*kw+kx (Conv2d weight order), zeros outside the image and in columns >= Cin*kh*kw (Kp % 4 == 0).
 * The 7x7/s2/p3 stem conv (depthformer_swin.py:1032-1039, 1152) becomes ged_gemm_tf32 with K = 148. */
int ged_im2col(const float* img, int64_t batch_stride, float* tok, int B, int Cin, int H, int W, int kh, int kw,
               int stride, int pad, int Kp, cudaStream_t stream);
/* nn.Unfold(2,2) gather of PatchMerging (depthformer_swin.py:98-117): (B,H,W,C) -> (B,ceil(H/2)*ceil(W/2),4C),
 * feature = c*4 + ky*2 + kx; backward=1 applies the adjoint. */
int ged_merge_patches(const float* src, float* dst, int B, int H, int W, int C, int backward, cudaStream_t stream);
/* encoder_decoder.py:132-138: clamp to [lo,hi] then bilinear (align_corners=True) to HxW; x (B,1,h0,w0). */
int ged_clamp_resize(const float* x, float* out, int B, int h0, int w0, int H, int W, float lo, float hi,
                     cudaStream_t stream);

/* ---- deformable attention sampling ----------------------------------------------------------- */
/* mmcv.ops.MultiScaleDeformableAttention core (hahi.py:280-289,316-325).  value (B,S,nH,64);
 * ref (ref_batch,Q,2); off (B,Q,nH,4,8,2); logit (B,Q,nH,32); out (B,Q,nH*64); level_hw = {h0,w0,...}. */
int ged_msda_fwd(const float* value, const float* ref, int ref_batch, const float* off,
                 const float* logit, float* out, const int* level_hw, int num_levels, int B, int S, int Q,
                 int nH, int head_dim, int num_points, cudaStream_t stream);
/* g_value accumulated; g_ref accumulated or NULL; g_off, g_logit overwritten. */
int ged_msda_bwd(const float* value, const float* ref, int ref_batch, const float* off,
                 const float* logit, const float* g_out, float* g_value, float* g_ref, float* g_off,
                 float* g_logit, const int* level_hw, int num_levels, int B, int S, int Q, int nH,
                 int head_dim, int num_points, cudaStream_t stream);

/* Locality-aware versions (csrc/msda_tile.cu): `order` (Q) int32 is any permutation of the queries; a CTA takes 32
 * consecutive entries of one (batch, head), stages a 12 x 12 window of value rows per level in shared memory and
 * reduces the value gradient per window cell before it leaves the SM.  ged_msda_sort_queries groups the queries by
 * reference point (band of y, bucket of x; ref (Q,2), work: bands * xbuckets ints) - only locality depends on the
 * order.  g_ref is (ref_batch,Q,2): with ref_batch == 1 the gradient is summed over the batch. */
int ged_msda_sort_queries(const float* ref, int Q, int bands, int xbuckets, int* order, int* work, int64_t work_ints,
                          cudaStream_t stream);
int ged_msda_tile_fwd(const float* value, const float* ref, int ref_batch, const float* off, const float* logit,
                      const int* order, float* out, const int* level_hw, int num_levels, int B, int S, int Q, int nH,
                      int head_dim, int num_points, cudaStream_t stream);
int ged_msda_tile_bwd(const float* value, const float* ref, int ref_batch, const float* off, const float* logit,
                      const int* order, const float* g_out, float* g_value, float* g_ref, float* g_off, float* g_logit,
                      const int* level_hw, int num_levels, int B, int S, int Q, int nH, int head_dim, int num_points,
                      cudaStream_t stream);

/* The same on the tensor cores (csrc/msda_tc.cu): per tile and level the sampling is a tcgen05.mma GEMM of a sparse
 * (query x window cell) weight matrix against the 11 x 11 window of value rows (TMA-staged); forward in 3xTF32
 * (fp32-accurate), backward (value gradient and corner dot products) in one-pass TF32, fp32 accumulate in TMEM. */
int ged_msda_tc_fwd(const float* value, float* value_lo /* workspace, same size as value */, const float* ref,
                    int ref_batch, const float* off, const float* logit, const int* order, float* out,
                    const int* level_hw, int num_levels, int B, int S, int Q, int nH, int head_dim, int num_points,
                    cudaStream_t stream);
int ged_msda_tc_bwd(const float* value, const float* ref, int ref_batch, const float* off, const float* logit,
                    const int* order, const float* g_out, float* g_value, float* g_ref, float* g_off, float* g_logit,
                    const int* level_hw, int num_levels, int B, int S, int Q, int nH, int head_dim, int num_points,
                    cudaStream_t stream);

/* ---- the narrow ends of the path (csrc/small.cu, SIMT) ---- */
/* Backward of a 3x3/s1/p1 conv with Cout in {1, 2, 11} (conv_depth decode_head.py:391, convfinal pemask_neck.py:36 /
 * dynamicpe_neck.py:497).  g, y (output after act; NULL when act == 0), gz (workspace of g's size): (B,H,W,Cout);
 * xp (B,H+2,W+2,Cin) zero-bordered input; w [Cout][3][3][Cin].  dx overwritten or NULL; dw, db ACCUMULATED or NULL. */
int ged_conv3x3_small_bwd(const float* g, const float* y, float* gz, const float* xp, const float* w, float* dx,
                          float* dw, float* db, int B, int H, int W, int Cin, int Cout, int act, float slope,
                          cudaStream_t stream);
/* Linear with N <= 4 outputs (HAHIHeteroNeck.reference_points 512 -> 2, hahi.py:176,299-300); act 0 none | 4 sigmoid. */
int ged_linear_small_fwd(const float* x, const float* w, const float* b, float* y, int64_t M, int N, int K, int act,
                         cudaStream_t stream);
int ged_linear_small_bwd(const float* g, const float* y, const float* x, const float* w, float* dx, float* dw, float* db,
                         int64_t M, int N, int K, int act, cudaStream_t stream);
/* q (B,S,C) = query + pos (S,C) [+ level_embed (4,C) of the level each token belongs to; level_start[5]] - the
 * `query + query_pos` of mmcv MultiScaleDeformableAttention with hahi.py:252-270's level embedding folded in. */
int ged_add_pos_fwd(const float* query, const float* pos, const float* level_embed, const int* level_start, float* q,
                    int B, int S, int C, cudaStream_t stream);
/* adjoint: g_level_embed (4,C) += per-level column sums of dq (or NULL); dq += extra in place (or NULL). */
int ged_add_pos_bwd(float* dq, const float* extra, float* g_level_embed, const int* level_start, int B, int S, int C,
                    cudaStream_t stream);

/* Roofline probe: the scatter pattern of ged_msda_bwd alone (one 256-byte red.global.add.v4.f32 row per half-warp at
 * pseudo-random rows of a (rows, heads*64) buffer), `iters` per half-warp.  Returns the number of warps launched
 * (payload = warps * iters * 512 bytes) or a negative error; the caller times it. */
int ged_msda_atomic_probe(float* buf, int rows, int heads, int iters, cudaStream_t stream);
/* Work mapping of the two kernels above.  bit0: a CTA takes one query x 8 heads (default: 8 queries x one head);
 * bit1: backward as two kernels (g_value scatter, then offset/weight gradients); -1 (default) = pick bit0 per call
 * (one query x 8 heads when Q >= 2 S).  Returns the previous value. */
int ged_set_msda_variant(int v);

/* ---- evaluation on the device (depth/core/evaluation/metrics.py:8-45, depth/datasets/kitti.py:355-385) ---- */
/* Per-image sums over the mask {y0<=y<y1, x0<=x<x1, min_depth < gt < max_depth}; sums (B,10) fp64, ACCUMULATED:
 * [count, #thresh<1.25, #<1.25^2, #<1.25^3, sum|g-p|/g, sum(g-p)^2/g, sum(g-p)^2, sum(ln g-ln p)^2, sum(ln p-ln g),
 *  sum|log10 g-log10 p|] from which a1,a2,a3,abs_rel,rmse,log_10,rmse_log,silog,sq_rel follow. */
int ged_depth_metrics(const float* pred, const float* gt, double* sums, int B, int H, int W, int y0, int y1, int x0,
                      int x1, float min_depth, float max_depth, cudaStream_t stream);
/* Flip test-time augmentation (encoder_decoder.py:226-233,262-270): out = (a + hflip(b_flipped)) / 2. */
int ged_tta_merge(const float* a, const float* b_flipped, float* out, int B, int H, int W, cudaStream_t stream);

/* Test-time input on the device (configs/depthformer/depthformer_v.py:33-53: KBCrop -> flip -> Normalize): planes
 * 0..2 of dst (5,H,W) <- mmcv.imnormalize of the [top,left] crop of a uint8 BGR image (H0,W0,3), optionally mirrored;
 * bit-exact with cv2's float32 arithmetic.  mean3 / std3 are HOST pointers.  Planes 3/4: ged_ground_plane. */
int ged_rgb_crop_normalize(const unsigned char* bgr, int H0, int W0, int top, int left, int flip, int to_rgb,
                           const float* mean3, const float* std3, float* dst, int H, int W, cudaStream_t stream);

/* ---- optimizer (configs/depthformer/depthformer_v.py:128-148) ---------------------------------- */
int ged_sumsq(const float* g, int64_t n, double* out, cudaStream_t stream);
int ged_adamw_step(float* p, const float* g, float* m, float* v, const uint8_t* wd_mask, int64_t n,
                   const double* sumsq, float max_norm, float grad_scale, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int step, const int* step_dev,
                   const float* lr_dev /* NULL, or the step's learning rate in device memory */, cudaStream_t stream);

/* ---- train-time augmentation of the 5-channel input on the device (csrc/augment.cu) -------------------
 * Replaces the CPU data-loader transforms of configs/depthformer/depthformer_v.py:13-33 - KBCrop, Resize, Padding,
 * RandomRotate, RandomFlip, RandomCrop, ColorAug, Normalize (depth/datasets/pipelines/transforms.py:149-205, 484-732, 64-109,
 * 208-296, 299-353, 356-417, 420-481, 12-61) - which run through mmcv into cv2.resize / cv2.warpAffine on float32 H x W x 5
 * images.  Bit-identical to the reference for the same drawn parameters (OpenCV's own fixed-point / fp32 arithmetic). */
/* planes 0..2 of a (5, H, W) float32 frame <- uint8 BGR (H, W, 3) */
int ged_aug_u8_to_planes(const unsigned char* bgr, float* planes, int H, int W, cudaStream_t stream);
/* KB window (top, left, sh x sw) of src5 (5 planes [H0][W0]) / depth / label -> resized to nw x nh (bilinear / nearest) ->
 * placed at (pad_x, pad_y) on the cw x ch canvas (background 0 / 0 / 255) */
int ged_aug_resize_pad(const float* src5, const float* depth, const float* label, int H0, int W0, int top, int left, int sh,
                       int sw, int nw, int nh, int pad_x, int pad_y, int cw, int ch, float* canvas5, float* canvas_d,
                       float* canvas_l, cudaStream_t stream);
/* canvas -> rotate (minv6: inverted 2x3 matrix, double, host) -> flip -> crop -> ColorAug (colors3: float64 BGR factors, host)
 * -> Normalize -> img (5, out_h, out_w), depth (out_h, out_w), label (out_h, out_w); mean3 / std3 in RGB order (host) */
int ged_aug_warp_crop_norm(const float* canvas5, const float* canvas_d, const float* canvas_l, int cw, int ch,
                           const double* minv6, int rotate, int flip, int crop_x, int crop_y, int out_w, int out_h, int color,
                           float gamma, float brightness, const double* colors3, const float* mean3, const float* std3,
                           float depth_scale, float* img, float* depth, float* label, cudaStream_t stream);
/* Whole batch in two launches: fill frame `index` of a host array of B * ged_aug_frame_bytes() bytes with the union of the
 * arguments above (`canvas`: 7 * cw * ch floats of workspace private to the frame), then ged_aug_train_batch copies the
 * descriptors to `frames_dev` (same size) and runs both kernels over grid (blocks, B). */
int ged_aug_frame_bytes(void);
int ged_aug_pack_frame(void* frames_host, int index, const float* src5, const float* depth, const float* label, float* canvas,
                       float* img, float* depth_out, float* label_out, int H0, int W0, int top, int left, int sh, int sw, int nw,
                       int nh, int pad_x, int pad_y, int cw, int ch, const double* minv6, int rotate, int flip, int crop_x,
                       int crop_y, int out_w, int out_h, int color, float gamma, float brightness, const double* colors3,
                       const float* mean3, const float* std3, float depth_scale);
int ged_aug_train_batch(const void* frames_host, int B, void* frames_dev, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GEDEPTH_H_ */
