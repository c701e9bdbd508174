#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/t_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv -s 8500 -c 3800 \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm2_tf32 -c 4 -f -o gpurun_out/r01b_gemm_pair python tools/prof_kernels.py gemm > gpurun_out/ncu_a.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_kernel -c 3 -f -o gpurun_out/r01b_gemm_dw python tools/prof_kernels.py dw > gpurun_out/ncu_b.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:msda -c 4 -f -o gpurun_out/r01b_msda python tools/prof_kernels.py msda > gpurun_out/ncu_c.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:winattn -c 2 -f -o gpurun_out/r01b_winattn python tools/prof_kernels.py attn > gpurun_out/ncu_d.log 2>&1
tail -n 6 gpurun_out/t_gpu.log; tail -c 2500 gpurun_out/bench.log; echo; wc -l gpurun_out/launches.csv; ls -la gpurun_out/*.ncu-rep
