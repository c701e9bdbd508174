mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/t_gpu.log; tail -n 4 gpurun_out/t_gpu.log
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_h9.log 2>&1
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_h9.log').read().splitlines() if x.startswith('{')]
d=json.loads(l[-1]); print(d['value'], d['ms_per_step'], d['clocks']); print({k:v for k,v in d['kernel_ms'].items() if 'small' in k or 'act' in k or 'upsample' in k})
PY
