#!/bin/bash
# round 2, seventh device run: conversion split between the ALU and conversion pipes (variants), PPM deposit cell by cell
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_ppm.py tests/test_z_multi_gpu.py "tests/test_z_baseline_size.py::test_cfg5_ppm_1080p_eye_pass_and_a_million_photons_match_the_oracle" -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r02i_pytest.log; tail -4 gpurun_out/r02i_pytest.log
run() { # name cfg env...
  local name=$1 cfg=$2; shift 2
  env "$@" python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/r02i_$name.err | tail -1 > gpurun_out/r02i_$name.json
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/r02i_{n}.json"))
    print(f"{n:24s} ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  value {d['value']:.0f}", {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()})
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/r02i_{n}.err").read()[-300:])
PY
}
for v in i0 i1 i2; do run cfg2_$v cfg2 B200RT_LIB=build/variants/libb200rt_$v.so; done
for v in i0 i1 i2; do run cfg2_${v}_b8v8 cfg2 B200RT_LIB=build/variants/libb200rt_$v.so B200RT_WIDE_BLOCKS=8 B200RT_NODE_VOTE=8; done
run cfg5_sort cfg5 X=1
run cfg5_nosort cfg5 B200RT_PPM_CELL_SORT=0
run cfg5_sort_pixel cfg5 B200RT_PPM_ORDER=pixel
run cfg5_sort_nodir cfg5 B200RT_PPM_NO_SORT=1
ncu --set full --clock-control none --import-source on -k regex:k_ppm_photon_shade -c 2 -f -o gpurun_out/r02i_ppm \
  python bench.py --config cfg5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02i_ncu.log 2>&1; tail -1 gpurun_out/r02i_ncu.log
