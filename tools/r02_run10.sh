#!/bin/bash
# round 2, tenth device run: PPM deposit variants (entries per scan step, occupancy, queues in shared memory)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in default l1b8 l2b6 s_l1b8 s_l2b8 s_l4b5 s_l1b10 s_l2b6; do
  if [ $v = default ]; then lib=""; else lib="B200RT_LIB=build/variants/libb200rt_$v.so"; fi
  env $lib X=1 python bench.py --config cfg5 --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/r02n_$v.err | tail -1 > gpurun_out/r02n_$v.json
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r02n_$v.json')); print('$v', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['stage_ms_per_step'].items()})
except Exception as e: print('$v FAILED', e, open('gpurun_out/r02n_$v.err').read()[-300:])"
done
