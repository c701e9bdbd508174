#!/bin/bash
# One gpurun call: unit parity (safe kernels first, tcgen05 in its own process), model parity,
# smoke, short bench.  Logs land in gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -m gedepth_b200.build > gpurun_out/build.log 2>&1
T="timeout 600 python -m pytest -q -m gpu -p no:cacheprovider"
$T tests/test_ops_gpu.py -k "not gemm and not conv and not linear" > gpurun_out/t_ops_simt.log 2>&1
$T tests/test_ops_gpu.py -k "gemm or linear" > gpurun_out/t_ops_gemm.log 2>&1
$T tests/test_ops_gpu.py -k "conv" > gpurun_out/t_ops_conv.log 2>&1

$T tests/test_model_gpu.py -s > gpurun_out/t_model.log 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 ${BENCH_ARGS:---no-cpu-baseline} > gpurun_out/bench.log 2>&1
tail -n 5 gpurun_out/t_ops_simt.log gpurun_out/t_ops_gemm.log gpurun_out/t_ops_conv.log gpurun_out/t_model.log gpurun_out/smoke.log gpurun_out/bench.log
