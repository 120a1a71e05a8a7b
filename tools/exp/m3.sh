python -m pytest tests -m gpu -x -q > gpurun_out/m3_tests.log 2>&1; echo rc=$? >> gpurun_out/m3_tests.log
python tools/insitu_ops.py 128 gpurun_out/m3_insitu128.csv > gpurun_out/m3_insitu128.log 2>&1
python bench.py --no-cpu > gpurun_out/m3_bench.json 2> gpurun_out/m3_bench.err
