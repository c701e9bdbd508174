#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python tools/ge_sweep.py > gpurun_out/ge_sweep.csv 2> gpurun_out/ge_sweep.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2>&1
echo "n2 rc=$?" >> gpurun_out/bench_n2.log
tail -n 5 gpurun_out/ge_sweep.err; cat gpurun_out/ge_sweep.csv | head -120; tail -c 600 gpurun_out/bench_n2.log
