for i in 1 2; do
for v in cur h0 h20000; do
if [ $v = cur ]; then L=""; else L=$PWD/vivid_b200/libvb_$v.so; fi
VB_LIB_PATH=$L python tools/sustained.py vivid-sr 128 4 2>&1 | tail -1 | sed "s/$/ lib=$v/" >> gpurun_out/m11_sustained.log
done
done
for v in cur h0 h20000; do
if [ $v = cur ]; then L=""; else L=$PWD/vivid_b200/libvb_$v.so; fi
VB_LIB_PATH=$L python tools/sustained.py vivid-base 128 3 2>&1 | tail -1 | sed "s/$/ lib=$v/" >> gpurun_out/m11_sustained.log
VB_LIB_PATH=$L python tools/sustained.py vivid-base 32 3 2>&1 | tail -1 | sed "s/$/ lib=$v/" >> gpurun_out/m11_sustained.log
done
