python -m pytest tests -m gpu -x -q > gpurun_out/m7_tests.log 2>&1; echo rc=$? >> gpurun_out/m7_tests.log
for f in 1 0 1 0; do
VB_FOLD_RES=$f python tools/sustained.py vivid-sr 128 5 2>&1 | tail -1 | sed "s/$/ FOLD_RES=$f/" >> gpurun_out/m7_sustained.log
done
for f in 1 0; do
VB_FOLD_RES=$f python tools/sustained.py vivid-base 128 4 2>&1 | tail -1 | sed "s/$/ FOLD_RES=$f/" >> gpurun_out/m7_sustained.log
done
