python -m pytest tests -m gpu -x -q > gpurun_out/m9_tests.log 2>&1; echo rc=$? >> gpurun_out/m9_tests.log
P=$PWD/vivid_b200/libvb_prev.so
for i in 1 2; do
python tools/sustained.py vivid-sr 128 5 2>&1 | tail -1 | sed "s/$/ lib=new/" >> gpurun_out/m9_sustained.log
VB_LIB_PATH=$P python tools/sustained.py vivid-sr 128 5 2>&1 | tail -1 | sed "s/$/ lib=prev/" >> gpurun_out/m9_sustained.log
done
python tools/sustained.py vivid-base 128 4 2>&1 | tail -1 | sed "s/$/ lib=new/" >> gpurun_out/m9_sustained.log
VB_LIB_PATH=$P python tools/sustained.py vivid-base 128 4 2>&1 | tail -1 | sed "s/$/ lib=prev/" >> gpurun_out/m9_sustained.log
