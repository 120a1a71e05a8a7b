python -m pytest tests -m gpu -x -q > gpurun_out/m4_tests.log 2>&1; echo rc=$? >> gpurun_out/m4_tests.log
export VB_B=64 VB_REPS=20
for st in 0 1; do
VB_QKV_STAGED=$st VB_DBG=32 VB_ONLY=7,16,17,18 VB_EPI=qkv python tools/conv_micro.py >> gpurun_out/m4_micro.log 2>&1
VB_B=128 VB_QKV_STAGED=$st VB_DBG=32 VB_ONLY=7,16,17,18 VB_EPI=qkv python tools/conv_micro.py >> gpurun_out/m4_micro.log 2>&1
done
python tools/insitu_ops.py 128 gpurun_out/m4_insitu128.csv > gpurun_out/m4_insitu128.log 2>&1
