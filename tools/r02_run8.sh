#!/bin/bash
# round 2, eighth device run: PPM deposit in gather form (hit-point-major, no atomics) against the scatter form
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_ppm.py tests/test_z_multi_gpu.py "tests/test_z_baseline_size.py::test_cfg5_ppm_1080p_eye_pass_and_a_million_photons_match_the_oracle" tests/test_z_dropin_binary.py -m gpu -q -x -s 2>&1 | grep -v "^    " | tail -25 > gpurun_out/r02j_pytest.log; tail -12 gpurun_out/r02j_pytest.log
run() { # name cfg env...
  local name=$1 cfg=$2; shift 2
  env "$@" python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/r02j_$name.err | tail -1 > gpurun_out/r02j_$name.json
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/r02j_{n}.json"))
    print(f"{n:24s} ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  value {d['value']:.0f}", {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()}, d["roofline"]["per_ray"])
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/r02j_{n}.err").read()[-300:])
PY
}
run cfg5_gather cfg5 X=1
run cfg5_scatter cfg5 B200RT_PPM_SCATTER=1
run cfg5_gather_pixel cfg5 B200RT_PPM_ORDER=pixel
ncu --set full --clock-control none --import-source on -k regex:k_ppm_gather -c 1 -f -o gpurun_out/r02j_gather \
  python bench.py --config cfg5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_ncu.log 2>&1; tail -1 gpurun_out/r02j_ncu.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02j_launches_cfg5.csv python bench.py --config cfg5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_ncu2.log 2>&1
