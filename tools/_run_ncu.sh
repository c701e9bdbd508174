mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02d_launches_config3.csv python bench.py --ncu-step > gpurun_out/r02d_ncu_step.log 2>&1
tail -n 2 gpurun_out/r02d_ncu_step.log
gzip -f gpurun_out/r02d_launches_config3.csv
ls -la gpurun_out/r02d_launches_config3.csv.gz
