#!/bin/bash
# round 2: two-level filtered gate (flat boxes of axis-aligned walls) and class-ordered tiles in the HW7 shade kernels,
# each against the build without it (tools/build_variant.sh: old = neither, gate = gate only, r2 / r8 = tile rounds).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q --ignore=tests/test_z_baseline_size.py --ignore=tests/test_z_multi_gpu.py 2>&1 | tail -4
one() {  # variant config steps warmup
  local lib=""
  [ "$1" != "new" ] && lib="build/variants/libb200rt_$1.so"
  B200RT_LIB=$lib timeout 200 python bench.py --config $2 --steps $3 --warmup $4 --no-cpu-baseline 2> gpurun_out/r02w_$1_$2.err | grep '^{' | tail -1 > gpurun_out/r02w_$1_$2.json
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r02w_$1_$2.json')); print('$1 $2', round(d['ms_per_step'],3), {a: round(b,3) for a,b in d['stage_ms_per_step'].items()})
except Exception as e: print('$1 $2 FAILED', e, open('gpurun_out/r02w_$1_$2.err').read()[-600:])"
}
for v in old new; do one $v cfg4 3 3; done
for v in gate r2 r8; do one $v cfg4 3 3; done
for v in old new; do one $v cfg2 20 5; one $v cfg3 3 3; one $v cfg5 10 3; done
