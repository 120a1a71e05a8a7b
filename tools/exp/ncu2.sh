set -x
tools/exp/umma_rate cal > gpurun_out/cal_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:rate_kernel -c 4 -f -o /tmp/cal tools/exp/umma_rate cal > gpurun_out/cal_ncu.log 2>&1
ncu -i /tmp/cal.ncu-rep --page raw --csv > gpurun_out/cal.raw.csv
