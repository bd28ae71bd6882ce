#!/bin/bash
# round 2, eleventh device run: deposit with the R(n) head in shared memory and the candidate's loads issued together
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_ppm.py "tests/test_z_baseline_size.py::test_cfg5_ppm_1080p_eye_pass_and_a_million_photons_match_the_oracle" -m gpu -q -x 2>&1 | tail -3
for v in default; do
  python bench.py --config cfg5 --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/r02o_$v.err | tail -1 > gpurun_out/r02o_$v.json
  python -c "
import json
d=json.load(open('gpurun_out/r02o_$v.json')); print('$v', round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['stage_ms_per_step'].items()})"
done
ncu --set full --clock-control none --import-source on -k "regex:k_ppm_photon_shade" -s 16 -c 1 -f -o gpurun_out/r02o_ppm python bench.py --config cfg5 --steps 1 --warmup 8 --no-cpu-baseline > gpurun_out/r02o_ncu.log 2>&1; tail -1 gpurun_out/r02o_ncu.log
