#!/bin/bash
# round 2, twelfth device run: child boxes as bfloat16 pairs (conversions on the FMA pipe) against the 8-bit encoding
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_edge_cases.py tests/test_z_open_fuzz.py tests/test_z_baseline_size.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r02q_pytest.log; tail -3 gpurun_out/r02q_pytest.log
run() { # name cfg steps env...
  local name=$1 cfg=$2 steps=$3; shift 3
  env "$@" python bench.py --config $cfg --steps $steps --warmup 3 --no-cpu-baseline 2> gpurun_out/r02q_$name.err | tail -1 > gpurun_out/r02q_$name.json
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/r02q_{n}.json"))
    print(f"{n:20s} ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  value {d['value']:.0f}", {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()}, {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d["roofline"]["per_ray"].items() if k in ("nodes", "tris", "shadow_nodes", "node_bytes")})
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/r02q_{n}.err").read()[-300:])
PY
}
run cfg2_bf16_b8 cfg2 20 X=1
run cfg2_bf16_b6 cfg2 20 B200RT_WIDE_BLOCKS=6
run cfg2_q8_b8 cfg2 20 B200RT_LIB=build/variants/libb200rt_q8.so
run cfg2_q8_b6 cfg2 20 B200RT_LIB=build/variants/libb200rt_q8.so B200RT_WIDE_BLOCKS=6
run cfg4_bf16 cfg4 3 X=1
run cfg4_bf16_b6 cfg4 3 B200RT_WIDE_BLOCKS=6
run cfg4_q8 cfg4 3 B200RT_LIB=build/variants/libb200rt_q8.so
run cfg3_bf16 cfg3 3 X=1
run cfg3_q8 cfg3 3 B200RT_LIB=build/variants/libb200rt_q8.so
ncu --set full --clock-control none --import-source on -k regex:k_trace_wide -s 18 -c 2 -f -o gpurun_out/r02q_wide \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02q_ncu.log 2>&1; tail -1 gpurun_out/r02q_ncu.log
