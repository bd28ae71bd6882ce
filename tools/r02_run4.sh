#!/bin/bash
# round 2, fourth device run: full -m gpu suite with the multi-GPU entry points, cfg2 / cfg5 benches (PPM hit-point
# order sweep), and one ncu capture of the traversal kernels with source
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -s 2>&1 | grep -v "^    " | tail -40 > gpurun_out/r02f_pytest.log; tail -15 gpurun_out/r02f_pytest.log
python bench.py --steps 20 --warmup 5 2> gpurun_out/r02f_cfg2.err | tail -1 > gpurun_out/r02f_cfg2.json
for o in none pixel cell; do
  B200RT_PPM_ORDER=$o python bench.py --config cfg5 --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/r02f_cfg5_$o.err | tail -1 > gpurun_out/r02f_cfg5_$o.json
done
python - <<'PY'
import json
for n in ("cfg2", "cfg5_none", "cfg5_pixel", "cfg5_cell"):
    try:
        d = json.load(open(f"gpurun_out/r02f_{n}.json"))
        print(n, "ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e ms", round(d["e2e"]["ms_per_step"], 3), d["stage_ms_per_step"])
    except Exception as e:
        print(n, "FAILED", e, open(f"gpurun_out/r02f_{n}.err").read()[-400:])
PY
ncu --set full --clock-control none --import-source on -k regex:k_trace_wide -c 2 -f -o gpurun_out/r02f_wide \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_ncu.log 2>&1; tail -2 gpurun_out/r02f_ncu.log
