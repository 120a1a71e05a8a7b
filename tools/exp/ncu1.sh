set -x
export VB_B=64 VB_REPS=3
CMD="python tools/conv_micro.py"
mkdir -p /tmp/nc
VB_ONLY=0 VB_EPI=r1s VB_TUNE=66 $CMD > gpurun_out/ncu1_plain.log 2>&1 &&
VB_ONLY=0 VB_EPI=r1s VB_TUNE=66 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 3 -c 1 -f -o /tmp/nc/r1s_sr $CMD > gpurun_out/ncu1.log 2>&1
VB_ONLY=14 VB_EPI=r1s VB_TUNE=0 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 3 -c 1 -f -o /tmp/nc/r1s_proj $CMD >> gpurun_out/ncu1.log 2>&1
VB_ONLY=7 VB_EPI=qkv VB_TUNE=0 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 3 -c 1 -f -o /tmp/nc/qkv $CMD >> gpurun_out/ncu1.log 2>&1
for n in r1s_sr r1s_proj qkv; do
  ncu -i /tmp/nc/$n.ncu-rep --page raw --csv > gpurun_out/ncu_$n.raw.csv 2>/dev/null
  ncu -i /tmp/nc/$n.ncu-rep --page source --csv --print-source sass > gpurun_out/ncu_$n.sass.csv 2>/dev/null
  ncu -i /tmp/nc/$n.ncu-rep --page source --csv --print-source cuda > gpurun_out/ncu_$n.cuda.csv 2>/dev/null
done
ls -la /tmp/nc gpurun_out
