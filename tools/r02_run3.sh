#!/bin/bash
# round 2, third device run: top-mask pre-pass + eager refill + pop in phase 1 + contexts — correctness, then sweeps
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_edge_cases.py tests/test_ppm.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02c_pytest.log; tail -4 gpurun_out/r02c_pytest.log
B200RT_CONTEXTS=3 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4
run() { # name, env...
  local name=$1; shift
  env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/r02c_$name.err | tail -1 > gpurun_out/r02c_$name.json
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/r02c_{n}.json"))
    s = d["stage_ms_per_step"]
    print(f"{n:28s} ms/frame {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  extend {s['ms_extend']:.3f} shadow {s['ms_shadow']:.3f} shade {s['ms_shade']:.3f}")
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/r02c_{n}.err").read()[-300:])
PY
}
for b in 8 6; do for r in 26 12 6 3; do
  run b${b}_r${r} B200RT_WIDE_BLOCKS=$b B200RT_REFILL_CLOSEST=$r B200RT_REFILL_ANY=$r
done; done
for v in 2 8 12; do run b8_r6_v$v B200RT_WIDE_BLOCKS=8 B200RT_REFILL_CLOSEST=6 B200RT_REFILL_ANY=6 B200RT_NODE_VOTE=$v; done
for k in 2 3 4 6; do run b8_r6_ctx$k B200RT_WIDE_BLOCKS=8 B200RT_REFILL_CLOSEST=6 B200RT_REFILL_ANY=6 B200RT_CONTEXTS=$k; done
run binary_ctx1 B200RT_NO_WIDE=1
run binary_ctx3 B200RT_NO_WIDE=1 B200RT_CONTEXTS=3
for v in "B200RT_NO_WIDE=1" "B200RT_WIDE_BLOCKS=8 B200RT_REFILL_CLOSEST=26 B200RT_REFILL_ANY=26" "B200RT_WIDE_BLOCKS=8 B200RT_REFILL_CLOSEST=6 B200RT_REFILL_ANY=6" "B200RT_WIDE_BLOCKS=6 B200RT_REFILL_CLOSEST=6 B200RT_REFILL_ANY=6"; do
  env $v python bench.py --config cfg4 --spp 4 --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/r02c_cfg4.err | tail -1 > gpurun_out/r02c_cfg4.json
  python -c "
import json; d=json.load(open('gpurun_out/r02c_cfg4.json')); print('cfg4 [$v]', round(d['ms_per_step'],2), round(d['value']), {k: round(v,2) for k,v in d['stage_ms_per_step'].items()})"
done
