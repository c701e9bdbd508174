#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider -k "window_attention" > gpurun_out/t_wa.log 2>&1; echo "rc=$?" >> gpurun_out/t_wa.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
GEDEPTH_DW_STREAM=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dwstream.log 2>&1
GEDEPTH_DW_STREAM=1 timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu -p no:cacheprovider -k "graph or arena" > gpurun_out/t_dwstream.log 2>&1; echo "rc=$?" >> gpurun_out/t_dwstream.log
tail -n 3 gpurun_out/t_wa.log; tail -n 4 gpurun_out/t_dwstream.log
python - <<'PY'
import json
for f in ['bench','bench_dwstream']:
    l=[x for x in open(f'gpurun_out/{f}.log') if x.startswith('{')]
    if not l: print(f,'NO JSON', open(f'gpurun_out/{f}.log').read()[-800:]); continue
    d=json.loads(l[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], {k:v for k,v in list(d['kernel_ms'].items())[:10]})
PY
