for cfg in "1 128" "0 128" "1 96" "1 160" "1 128" "0 128"; do
set -- $cfg
VB_DUAL_STREAM=$1 python bench.py --no-cpu --no-insitu --no-e2e --steps 2 --warmup 2 --batch $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('dual_stream=$1 batch=$2', d['value'], 'images/s', d['ms_per_step'], 'ms/step', d['clocks']['sm_mhz'], 'MHz')
" >> gpurun_out/m8.log
done
