export VB_B=64 VB_REPS=3
CMD="python tools/conv_micro.py"
mkdir -p /tmp/nc
VB_ONLY=0 VB_EPI=r3nss VB_TUNE=66 $CMD > gpurun_out/ncu4_plain.log 2>&1 &&
VB_ONLY=0 VB_EPI=r3nss VB_TUNE=66 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 3 -c 1 -f -o /tmp/nc/r3nss $CMD > gpurun_out/ncu4.log 2>&1
VB_ONLY=0 VB_EPI=mod VB_TUNE=66 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 3 -c 1 -f -o /tmp/nc/mod $CMD >> gpurun_out/ncu4.log 2>&1
for n in r3nss mod; do
  ncu -i /tmp/nc/$n.ncu-rep --page raw --csv > gpurun_out/ncu_$n.raw.csv 2>/dev/null
  ncu -i /tmp/nc/$n.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu_$n.sass.csv 2>/dev/null
done
