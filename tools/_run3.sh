mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/t_gpu.log; tail -n 4 gpurun_out/t_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "rc=$?" >> gpurun_out/bench.log
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench.log').read().splitlines() if x.startswith('{')]
d=json.loads(l[-1])
for k in ('value','ms_per_step','e2e','clocks','gpu_launches','roofline','cpu_baseline','gpu_library_baseline','vs_library_gpu','ground_embed','swin_window_attention','other_configs','roofline_hbm_kernels','train_augment'):
    print(k, json.dumps(d.get(k))[:1500])
print(d['kernel_ms'])
PY
