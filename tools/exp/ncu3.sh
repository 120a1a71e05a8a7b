python tools/ncu_calls.py 128 > gpurun_out/ncu3_plain.log 2>&1 &&
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/launches_B128.csv python tools/ncu_calls.py 128 > gpurun_out/ncu3.log 2>&1
