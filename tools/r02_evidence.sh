#!/bin/bash
# round 2, evidence run (1 GPU): the -m gpu suite, bench.py of every BASELINE config, launch lists and ncu --set full
# captures of the dominant kernels.  Summaries go to profiles/ (tools/make_traffic.py, profiles/summarize.py).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02p}
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_pytest.log
python bench.py --steps 20 --warmup 5 2> gpurun_out/${T}_bench_cfg2.err | tail -1 > gpurun_out/${T}_bench_cfg2.json
python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/${T}_bench_cfg2_reference.err | tail -1 > gpurun_out/${T}_bench_cfg2_reference.json
python bench.py --config cfg1 --steps 20 --warmup 5 2> gpurun_out/${T}_bench_cfg1.err | tail -1 > gpurun_out/${T}_bench_cfg1.json
python bench.py --config cfg3 --steps 3 --warmup 3 2> gpurun_out/${T}_bench_cfg3.err | tail -1 > gpurun_out/${T}_bench_cfg3.json
python bench.py --config cfg4 --steps 3 --warmup 3 2> gpurun_out/${T}_bench_cfg4.err | tail -1 > gpurun_out/${T}_bench_cfg4.json
python bench.py --config cfg5 --steps 10 --warmup 3 2> gpurun_out/${T}_bench_cfg5.err | tail -1 > gpurun_out/${T}_bench_cfg5.json
python - "$T" <<'PY'
import json, sys
T = sys.argv[1]
for c in ("cfg2", "cfg2_reference", "cfg1", "cfg3", "cfg4", "cfg5"):
    try:
        d = json.load(open(f"gpurun_out/{T}_bench_{c}.json"))
        r = d.get("roofline", {})
        print(f"{c:16s} ms/step {d['ms_per_step']:.3f} ms/frame {d.get('ms_per_frame', 0):.2f} value {d['value']:.1f} e2e {d['e2e']['value']:.1f} "
              f"roofline {r.get('bound')} frac {r.get('frac')} cpu {d.get('cpu_baseline', {}).get('value')}")
    except Exception as e:
        print(c, "FAILED", e)
PY
# launch lists (the same commands as the bench lines above, fewer steps)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${T}_launches_cfg5.csv python bench.py --config cfg5 --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${T}_launches_cfg4.csv python bench.py --config cfg4 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
# ncu --set full of the dominant kernels, one frame / pass after the warm-up
K='regex:k_trace_wide|k_trace_hw1_fast|k_ppm_photon_shade'
ncu --set full --clock-control none --import-source on -k "$K" -s 18 -c 6 -f -o gpurun_out/${T}_cfg2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_cfg2.log 2>&1
ncu --set full --clock-control none -k "$K" -s 42 -c 14 -f -o gpurun_out/${T}_cfg1 python bench.py --config cfg1 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_cfg1.log 2>&1
ncu --set full --clock-control none -k "$K" -s 84 -c 28 -f -o gpurun_out/${T}_cfg5 python bench.py --config cfg5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_cfg5.log 2>&1
ncu --set full --clock-control none -k "$K" -s 24 -c 8 -f -o gpurun_out/${T}_cfg3 python bench.py --config cfg3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_cfg3.log 2>&1
ncu --set full --clock-control none -k "$K" -s 30 -c 10 -f -o gpurun_out/${T}_cfg4 python bench.py --config cfg4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_cfg4.log 2>&1
ls -la gpurun_out/${T}_*.ncu-rep
