mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02c_launches_config3.csv python bench.py --ncu-step > gpurun_out/r02c_ncu_step.log 2>&1
tail -n 3 gpurun_out/r02c_ncu_step.log
gzip -f gpurun_out/r02c_launches_config3.csv
ls -la gpurun_out/r02c_launches_config3.csv.gz
timeout 300 python tools/ge_sweep.py 2>&1 | grep -v Warn > gpurun_out/r02c_ge_sweep.csv
tail -n 14 gpurun_out/r02c_ge_sweep.csv
