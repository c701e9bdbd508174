#!/bin/bash
# compute-sanitizer over one small parity test per kernel family (memcheck, racecheck, synccheck).  Logs -> gpurun_out/,
# one-line summaries -> gpurun_out/sanitizer_summary.txt (copied to profiles/ by hand).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
SEL='test_msda_fwd_bwd and (path or tile) and 850 or test_msda_module_fwd_bwd or test_window_attention_fwd_bwd and 16-40-3-3 or test_linear_small or test_conv3x3 and 11-35-64-64 and 3xtf32 or test_conv3x3 and 20-48-64-1 and 3xtf32 or test_gemm_dw or test_linear_autograd or test_ge_vanilla_fwd_bwd and 64-160'
: > gpurun_out/sanitizer_summary.txt
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  log=gpurun_out/sanitizer_${tool}.log
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 \
      python -m pytest tests/test_ops_gpu.py -q -m gpu -x -p no:cacheprovider -k "$SEL" > $log 2>&1
  rc=$?
  echo "$tool: exit $rc | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' $log | tr '\n' ' ')" >> gpurun_out/sanitizer_summary.txt
done
cat gpurun_out/sanitizer_summary.txt
