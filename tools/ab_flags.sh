# A/B runs of bench.py under environment switches (debug helper): usage  bash tools/ab_flags.sh
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary --no-op-profile > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err; python -c "
import json,sys
try:
    d=json.loads([l for l in open('gpurun_out/ab_$name.json') if l.startswith('{')][-1]); print('$name', round(d['ms_per_step'],3), round(d['value']), d['gpu_launches'])
except Exception as e: print('$name FAILED', e)
"; }
run pair1 RCGAN_TC_PAIR=1
run pair2 RCGAN_TC_PAIR=2
run pair3 RCGAN_TC_PAIR=3
run pair1b RCGAN_TC_PAIR=1
run pair3b RCGAN_TC_PAIR=3
