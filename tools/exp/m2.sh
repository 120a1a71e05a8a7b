python -m pytest tests -m gpu -x -q > gpurun_out/m2_tests.log 2>&1; echo rc=$? >> gpurun_out/m2_tests.log
export VB_B=64 VB_REPS=20
VB_DBG=32 VB_ONLY=14,20,0 VB_EPI=simple,r1s python tools/conv_micro.py > gpurun_out/m2_micro.log 2>&1
VB_WGT_NOWAIT=0 VB_DBG=32 VB_ONLY=14,20,0 VB_EPI=simple,r1s python tools/conv_micro.py >> gpurun_out/m2_micro.log 2>&1
python tools/insitu_ops.py 128 gpurun_out/m2_insitu128.csv > gpurun_out/m2_insitu128.log 2>&1
VB_WGT_NOWAIT=0 python tools/insitu_ops.py 128 gpurun_out/m2_insitu128_wait.csv 2>&1 | grep "^==" > gpurun_out/m2_insitu128_wait.log
