#!/bin/bash
# round 2: class-ordered shade tiles on cfg4 — tile size and class key (tools/build_variant.sh: r16 = 16 rounds,
# obj64 / obj128 = shade_b ordered by object like shade_a with 64 / 128 classes, c128 = 128 classes, material key).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
one() {  # variant config steps warmup
  local lib=""
  [ "$1" != "new" ] && lib="build/variants/libb200rt_$1.so"
  B200RT_LIB=$lib timeout 200 python bench.py --config $2 --steps $3 --warmup $4 --no-cpu-baseline 2> gpurun_out/r02x_$1_$2.err | grep '^{' | tail -1 > gpurun_out/r02x_$1_$2.json
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r02x_$1_$2.json')); print('$1 $2', round(d['ms_per_step'],3), {a: round(b,3) for a,b in d['stage_ms_per_step'].items()})
except Exception as e: print('$1 $2 FAILED', e, open('gpurun_out/r02x_$1_$2.err').read()[-600:])"
}
for v in new r16 obj64 obj128 c128; do one $v cfg4 3 3; done
