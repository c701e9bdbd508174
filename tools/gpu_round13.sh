#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/t_gpu.log
timeout 300 python tools/ge_sweep.py > gpurun_out/ge_sweep.csv 2> gpurun_out/ge_sweep.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
timeout 600 python tools/profile_ops.py > gpurun_out/profile_ops.log 2>&1
tail -n 8 gpurun_out/t_gpu.log; tail -n 3 gpurun_out/ge_sweep.err; grep -E "352,1120|1024,2048" gpurun_out/ge_sweep.csv; tail -c 1400 gpurun_out/bench.log; echo; grep -v Warning gpurun_out/profile_ops.log | tail -n 22
