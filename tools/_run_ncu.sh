mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ge_vanilla -c 4 -o gpurun_out/r02c_ge_vanilla_stream python tools/ge_once.py 8 1024 2048 > gpurun_out/r02c_ncu_ge.log 2>&1
tail -n 3 gpurun_out/r02c_ncu_ge.log; ls -la gpurun_out/r02c_ge_vanilla_stream.ncu-rep
nvidia-smi --query-gpu=name,clocks.sm --format=csv
