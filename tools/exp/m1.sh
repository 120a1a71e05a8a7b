set -x
export VB_B=64 VB_REPS=20
for t in 0 64 128 192; do
VB_TUNE=$t VB_DBG=32 VB_ONLY=14,20 VB_EPI=simple,r1s python tools/conv_micro.py
done
VB_DBG=32 VB_ONLY=7,16,17 VB_EPI=qkv python tools/conv_micro.py
VB_TUNE=128 VB_DBG=32 VB_ONLY=7,16,17 VB_EPI=qkv python tools/conv_micro.py
VB_DBG=32 VB_ONLY=0 VB_EPI=simple,mod,r1s,r2nss python tools/conv_micro.py
VB_TUNE=66 VB_DBG=32 VB_ONLY=0 VB_EPI=simple,mod,r1s,r2nss python tools/conv_micro.py
echo ==== PROF
export VB_LIB_PATH=$PWD/vivid_b200/libvb_prof.so VB_REPS=1
for t in 0 64; do
VB_TUNE=$t VB_ONLY=14,20 VB_EPI=simple,r1s python tools/conv_micro.py
done
VB_TUNE=66 VB_ONLY=0 VB_EPI=simple,r1s,r2nss python tools/conv_micro.py
VB_TUNE=2 VB_ONLY=0 VB_EPI=simple,r1s,r2nss python tools/conv_micro.py
