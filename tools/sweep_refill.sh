#!/usr/bin/env bash
# Sweeps the traversal refill thresholds on one config (run on the GPU box): tools/sweep_refill.sh cfg4 "--spp 4"
cfg=${1:-cfg4}; extra=${2:---spp 4}
for any in 32 24 16 8; do for clo in 32 16; do
  B200RT_REFILL_ANY=$any B200RT_REFILL_CLOSEST=$clo python bench.py --config $cfg $extra --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null \
   | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stage_ms_per_step']; print('any=$any closest=$clo', round(d['ms_per_step'],2), 'extend', round(s['ms_extend'],2), 'shadow', round(s['ms_shadow'],2))"
done; done
