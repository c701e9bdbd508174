#!/bin/bash
# the routine check: whole GPU suite + default bench (no CPU baseline) + library-op audit
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/t_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
timeout 600 python tools/profile_ops.py > gpurun_out/profile_ops.log 2>&1
tail -n 8 gpurun_out/t_gpu.log; tail -c 1300 gpurun_out/bench.log; echo; grep -v Warning gpurun_out/profile_ops.log | tail -n 18
