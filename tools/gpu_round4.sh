#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2>&1
echo "n2 rc=$?" >> gpurun_out/bench_n2.log
timeout 900 python bench.py --backbone swin_l --variant a --batch 16 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_swinl_a.log 2>&1
echo "rc=$?" >> gpurun_out/bench_swinl_a.log
timeout 900 python bench.py --backbone swin_l --variant a --dataset ddad --batch 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_swinl_a_ddad.log 2>&1
echo "rc=$?" >> gpurun_out/bench_swinl_a_ddad.log
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2>&1
for f in bench_n2 bench_swinl_a bench_swinl_a_ddad bench_ref; do echo "== $f"; tail -c 1200 gpurun_out/$f.log; echo; done
