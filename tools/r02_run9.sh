#!/bin/bash
# round 2, ninth device run: PPM scan grid rebuilt as the radii shrink
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_ppm.py tests/test_z_multi_gpu.py "tests/test_z_baseline_size.py::test_cfg5_ppm_1080p_eye_pass_and_a_million_photons_match_the_oracle" tests/test_z_dropin_binary.py tests/test_refit.py -m gpu -q -x -s 2>&1 | grep -v "^    " | tail -25 > gpurun_out/r02l_pytest.log; tail -8 gpurun_out/r02l_pytest.log
run() { # name cfg steps env...
  local name=$1 cfg=$2 steps=$3; shift 3
  env "$@" python bench.py --config $cfg --steps $steps --warmup 3 --no-cpu-baseline 2> gpurun_out/r02l_$name.err | tail -1 > gpurun_out/r02l_$name.json
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/r02l_{n}.json"))
    print(f"{n:24s} ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  value {d['value']:.0f}", {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()}, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d["roofline"]["per_ray"].items()})
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/r02l_{n}.err").read()[-300:])
PY
}
run cfg5_adaptive cfg5 10 X=1
run cfg5_static cfg5 10 B200RT_PPM_STATIC_GRID=1
run cfg5_adaptive_100 cfg5 97 X=1
run cfg2 cfg2 20 X=1
