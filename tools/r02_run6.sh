#!/bin/bash
# round 2, sixth device run: node test without conversions (byte-permute magic floats), mask building once per node
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_edge_cases.py tests/test_refit.py tests/test_z_dropin_binary.py tests/test_z_baseline_size.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r02h_pytest.log; tail -4 gpurun_out/r02h_pytest.log
run() { # name, env...
  local name=$1; shift
  env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/r02h_$name.err | tail -1 > gpurun_out/r02h_$name.json
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/r02h_{n}.json"))
    s = d["stage_ms_per_step"]
    print(f"{n:28s} ms/frame {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  extend {s['ms_extend']:.3f} shadow {s['ms_shadow']:.3f} shade {s['ms_shade']:.3f}")
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/r02h_{n}.err").read()[-300:])
PY
}
run default X=1
for b in 4 8; do run b$b B200RT_WIDE_BLOCKS=$b; done
for v in 8 12 16; do run v$v B200RT_NODE_VOTE=$v; done
for r in 16 20 30; do run r$r B200RT_REFILL_CLOSEST=$r B200RT_REFILL_ANY=$r; done
run v12_r20 B200RT_NODE_VOTE=12 B200RT_REFILL_CLOSEST=20 B200RT_REFILL_ANY=20
run ctx2 B200RT_CONTEXTS=2
run binary B200RT_NO_WIDE=1
for c in cfg1 cfg3 cfg5; do
  python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/r02h_$c.err | tail -1 > gpurun_out/r02h_$c.json
  python -c "
import json; d=json.load(open('gpurun_out/r02h_$c.json')); print('$c', round(d['ms_per_step'],3), round(d['value']), {k: round(v,2) for k,v in d['stage_ms_per_step'].items()})"
done
python bench.py --config cfg4 --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/r02h_cfg4.err | tail -1 > gpurun_out/r02h_cfg4.json
python -c "
import json; d=json.load(open('gpurun_out/r02h_cfg4.json')); print('cfg4', round(d['ms_per_step'],3), round(d['value']), {k: round(v,2) for k,v in d['stage_ms_per_step'].items()})"
ncu --set full --clock-control none --import-source on -k regex:k_trace_wide -c 2 -f -o gpurun_out/r02h_wide \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_ncu.log 2>&1; tail -1 gpurun_out/r02h_ncu.log
