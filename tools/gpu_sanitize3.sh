#!/bin/bash
# compute-sanitizer over the kernels of round 2c: streaming Vanilla ground embedding (forward / backward), LayerNorm rows in
# registers (+ fused dw / db), row-structured prep_conv_input / upsample adjoint / im2col / merge_patches, act_bwd (both kernels,
# 128- and 512-row chunks are too large for the sanitizer: the short one only), small-Cout conv dW strips, strided-residual dX GEMM,
# adaptive backward with the deferred staging wait.  Summaries -> gpurun_out/sanitizer3_summary.txt.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
SEL='test_ge_vanilla_x2_streaming_equals_tiled and (64-160 or 6-12 or 66-516 or 2-4) or test_ge_vanilla_fwd_bwd and (64-160 or 4-8) or test_ge_adaptive_fwd_bwd and (64-160 or 36-84) or test_layernorm_fwd_bwd and (640-96 or 333-384 or 1237-192 or 77-512 or 65-640 or 3-4) or test_prep_conv_input_and_adjoint and (11-35-22-70 or 5-7-16-22 or 3-4-11-15 or 4-9-32-72 or 2-3-40-50) or test_prep_conv_input_batch_strided_sources or test_stem_conv_as_im2col_gemm and 35-83 and 3xtf32 or test_merge_patches_fwd_bwd and (9-21-192 or 5-11-384) or test_act_bwd and not long or test_conv3x3 and (20-48-64-1 or 17-33-64-11) and 3xtf32 or test_gemm_pair_kernel and 3xtf32 and 20000-64-64'
: > gpurun_out/sanitizer3_summary.txt
for tool in ${TOOLS:-memcheck synccheck racecheck}; do
  log=gpurun_out/sanitizer3_${tool}.log
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 \
      python -m pytest tests/test_ops_gpu.py -q -m gpu -x -p no:cacheprovider -k "$SEL" > $log 2>&1
  rc=$?
  echo "$tool: exit $rc | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' $log | tr '\n' ' ')" >> gpurun_out/sanitizer3_summary.txt
done
cat gpurun_out/sanitizer3_summary.txt
