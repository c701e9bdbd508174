#!/bin/bash
# end-of-round check, as the driver runs it: GPU suite, smoke, default bench, reference arm, 2-GPU bench, configs 3/4
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/t_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "rc=$?" >> gpurun_out/bench.log
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2>&1
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2>&1
  echo "rc=$?" >> gpurun_out/bench_n2.log
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_n2.log 2>&1
fi
timeout 900 python bench.py --workload config2 --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_config2.log 2>&1
timeout 900 python bench.py --passes 2 --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_p2.log 2>&1
tail -n 4 gpurun_out/t_gpu.log; tail -n 2 gpurun_out/smoke.log
python - <<'PY'
import json
for f in ['bench','bench_ref','bench_n2','bench_ref_n2','bench_config2','bench_p2']:
    try:
        txt=open(f'gpurun_out/{f}.log').read()
    except Exception as e:
        print(f, 'missing'); continue
    l=[x for x in txt.splitlines() if x.startswith('{')]
    if not l: print(f,'NO JSON', txt[-600:]); continue
    d=json.loads(l[-1]); print(f, 'value', d.get('value'), 'ms', d.get('ms_per_step'), 'e2e', (d.get('e2e') or {}).get('value'), 'n', d.get('n_gpus'), 'launches', d.get('gpu_launches'), 'clocks', d.get('clocks'), 'cpu', d.get('cpu_baseline'))
PY
