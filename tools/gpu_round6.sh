#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 90 python tools/pair_probe.py > gpurun_out/pair_probe.log 2>&1
rc=$?
echo "probe rc=$rc" >> gpurun_out/pair_probe.log
if [ $rc -ne 0 ] || ! grep -q "probe ok" gpurun_out/pair_probe.log; then export GEDEPTH_GEMM_PAIR=0; echo "PAIR KERNEL DISABLED" >> gpurun_out/pair_probe.log; nvidia-smi > gpurun_out/smi_after_probe.txt 2>&1; fi
T="timeout 300 python -m pytest -q -m gpu -p no:cacheprovider"
if [ -z "$GEDEPTH_GEMM_PAIR" ]; then $T tests/test_ops_gpu.py -k "pair_kernel" > gpurun_out/t_pair.log 2>&1; echo "rc=$?" >> gpurun_out/t_pair.log; fi
$T tests/test_ops_gpu.py -k "stem_conv or linear_autograd or conv3x3 or conv1x1 or conv2d_cat" > gpurun_out/t_new.log 2>&1; echo "rc=$?" >> gpurun_out/t_new.log
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/t_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
GEDEPTH_GEMM_PAIR=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nopair.log 2>&1
cat gpurun_out/pair_probe.log; tail -n 6 gpurun_out/t_pair.log; tail -n 6 gpurun_out/t_new.log; tail -n 12 gpurun_out/t_gpu.log; tail -c 1300 gpurun_out/bench.log; echo; tail -c 1300 gpurun_out/bench_nopair.log
