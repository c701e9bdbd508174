#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/t_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "rc=$?" >> gpurun_out/bench.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --ncu-step --warmup 3 > gpurun_out/ncu_launches.log 2>&1
tail -n 3 gpurun_out/t_gpu.log; tail -n 2 gpurun_out/smoke.log; wc -l gpurun_out/launches.csv
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench.log') if x.startswith('{')]
d=json.loads(l[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'], d['clocks']); print(json.dumps(d['roofline'])[:1800]); print(d['cpu_baseline'])
PY
