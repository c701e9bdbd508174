#!/bin/bash
# compute-sanitizer over the kernels added late in round 2: closed-form x2 ground embedding (TMA / cp.async staging), tcgen05
# window-attention backward, bf16-split GEMM, device augmentation.  Summaries -> gpurun_out/sanitizer2_summary.txt.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
SEL='test_ge_adaptive_fwd_bwd and (64-160 or 6-12 or 36-84 or 30-244) or test_ge_adaptive_x2_equals_generic and 64-160 or test_ge_vanilla_fwd_bwd and 64-160 or test_window_attention_fwd_bwd and tcgen05 and (16-40-3-3 or 11-35-24) or test_gemm_plain and bf16x3 and (256-96-96 or 1000-288-96) or test_gemm_pair_kernel and bf16x3 and 33000'
: > gpurun_out/sanitizer2_summary.txt
for tool in ${TOOLS:-memcheck synccheck racecheck}; do
  log=gpurun_out/sanitizer2_${tool}.log
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 \
      python -m pytest tests/test_ops_gpu.py tests/test_augment_gpu.py -q -m gpu -x -p no:cacheprovider -k "$SEL or test_device_augmentation_is_bit_identical_to_the_reference and (21 or 5)" > $log 2>&1
  rc=$?
  echo "$tool: exit $rc | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' $log | tr '\n' ' ')" >> gpurun_out/sanitizer2_summary.txt
done
cat gpurun_out/sanitizer2_summary.txt
