#!/bin/bash
# dW / wide-tile bring-up: targeted tests under a hang guard, then the A/B timings, then the whole suite + bench.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T="timeout 300 python -m pytest -q -m gpu -p no:cacheprovider"
$T tests/test_ops_gpu.py -k "gemm_dw or conv3x3_dw" > gpurun_out/t_dw.log 2>&1; echo "rc=$?" >> gpurun_out/t_dw.log
$T tests/test_ops_gpu.py -k "gemm_plain or gemm_epilogue or linear_autograd or conv3x3 or conv2d_cat or conv1x1 or patch_embed" > gpurun_out/t_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/t_gemm.log
timeout 300 python tools/ab_gemm.py > gpurun_out/ab_gemm.log 2>&1
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/t_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
tail -n 25 gpurun_out/t_dw.log; tail -n 8 gpurun_out/t_gemm.log; cat gpurun_out/ab_gemm.log; tail -n 8 gpurun_out/t_gpu.log; tail -c 1500 gpurun_out/bench.log
