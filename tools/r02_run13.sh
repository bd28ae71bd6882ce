#!/bin/bash
# round 2, thirteenth device run: k_trace_coop (8 lanes per ray) — parity with every wave forced through it, then the
# size limit swept
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B200RT_COOP_MAX=2000000000 python -m pytest tests/test_gpu_parity.py tests/test_edge_cases.py tests/test_z_open_fuzz.py tests/test_ppm.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r02t_pytest.log; tail -3 gpurun_out/r02t_pytest.log
run() { # name cfg steps env...
  local name=$1 cfg=$2 steps=$3; shift 3
  env "$@" python bench.py --config $cfg --steps $steps --warmup 3 --no-cpu-baseline 2> gpurun_out/r02t_$name.err | tail -1 > gpurun_out/r02t_$name.json
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/r02t_{n}.json"))
    print(f"{n:20s} ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  value {d['value']:.0f}", {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()})
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/r02t_{n}.err").read()[-300:])
PY
}
for m in 0 30000 100000 300000 600000 1500000; do run cfg2_c$m cfg2 20 B200RT_COOP_MAX=$m; done
for m in 0 100000 600000; do run cfg5_c$m cfg5 10 B200RT_COOP_MAX=$m; done
for m in 0 100000; do run cfg1_c$m cfg1 20 B200RT_COOP_MAX=$m; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02t_launches_cfg2.csv env B200RT_COOP_MAX=600000 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
