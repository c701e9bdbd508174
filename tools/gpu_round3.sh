#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T="timeout 300 python -m pytest -q -m gpu -p no:cacheprovider"
$T tests/test_ops_gpu.py -k "msda" > gpurun_out/t_msda.log 2>&1; echo "rc=$?" >> gpurun_out/t_msda.log
timeout 300 python tools/ab_msda.py > gpurun_out/ab_msda.log 2>&1
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/t_gpu.log
timeout 300 python tools/ab_gemm.py > gpurun_out/ab_gemm.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
tail -n 4 gpurun_out/t_msda.log; cat gpurun_out/ab_msda.log; tail -n 12 gpurun_out/t_gpu.log; grep -v "^fwd\|^dW\|^conv" gpurun_out/ab_gemm.log | head -80; tail -c 1500 gpurun_out/bench.log
