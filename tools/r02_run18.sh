#!/bin/bash
# round 2: register budget of the HW7 shade kernels (they wait on memory at 16 warps per SM, profiles/r02zz): 2 / 3 / 4
# resident blocks of 256 threads = 128 / 80 / 64 registers.  The -m gpu suite then runs on the fastest build.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
best=new; best_ms=1000000
for v in new mb3 mb4; do
  lib=""; [ "$v" != "new" ] && lib="build/variants/libb200rt_$v.so"
  B200RT_LIB=$lib timeout 100 python bench.py --config cfg4 --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/r02y_$v.err | grep '^{' | tail -1 > gpurun_out/r02y_${v}_cfg4.json
  ms=$(python -c "
import json
d=json.load(open('gpurun_out/r02y_${v}_cfg4.json')); print(round(d['ms_per_step'],3)); import sys; print('$v', round(d['ms_per_step'],3), {a: round(b,3) for a,b in d['stage_ms_per_step'].items()}, file=sys.stderr)")
  if python -c "import sys; sys.exit(0 if float('$ms') < 0.985 * float('$best_ms') else 1)"; then best=$v; best_ms=$ms; fi
done
echo "fastest by more than 1.5 %: $best ($best_ms ms)"
if [ "$best" != "new" ]; then
  B200RT_LIB=build/variants/libb200rt_$best.so timeout 120 python -m pytest tests -m gpu -q -x --ignore=tests/test_z_baseline_size.py --ignore=tests/test_z_multi_gpu.py 2>&1 | tail -2
  B200RT_LIB=build/variants/libb200rt_$best.so timeout 60 python -m pytest tests/test_z_baseline_size.py -m gpu -q -x -k cfg4 2>&1 | tail -2
fi
