#!/bin/bash
# One gpurun call: whole GPU parity suite, smoke, the default bench line, and ncu --set full captures of the
# GEMM variants.  Logs land in gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -m gedepth_b200.build > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x --durations=15 > gpurun_out/t_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1
if [ -n "$NCU_GEMM" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -c 6 -f -o gpurun_out/gemm_3x \
      python tools/prof_kernels.py gemm > gpurun_out/ncu_gemm3.log 2>&1
  GEDEPTH_GEMM_PASSES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -c 6 -f -o gpurun_out/gemm_1x \
      python tools/prof_kernels.py gemm > gpurun_out/ncu_gemm1.log 2>&1
fi
tail -n 6 gpurun_out/t_gpu.log gpurun_out/smoke.log
tail -c 3000 gpurun_out/bench.log
