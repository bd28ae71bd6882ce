#!/bin/bash
# round 2, evidence run (1 GPU): bench.py of every BASELINE config, launch lists and ncu --set full captures of the
# dominant kernels.  The .ncu-rep files are summarised ON the box (profiles/summarize.py, tools/make_traffic.py) and
# only the text comes back (gpurun_out/ is capped at 64 MiB).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02p}
python bench.py --steps 20 --warmup 5 2> gpurun_out/${T}_bench_cfg2.err | tail -1 > gpurun_out/${T}_bench_cfg2.json
python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/${T}_bench_cfg2_reference.err | tail -1 > gpurun_out/${T}_bench_cfg2_reference.json
python bench.py --config cfg1 --steps 20 --warmup 5 2> gpurun_out/${T}_bench_cfg1.err | tail -1 > gpurun_out/${T}_bench_cfg1.json
python bench.py --config cfg3 --steps 3 --warmup 3 2> gpurun_out/${T}_bench_cfg3.err | tail -1 > gpurun_out/${T}_bench_cfg3.json
python bench.py --config cfg4 --steps 3 --warmup 3 2> gpurun_out/${T}_bench_cfg4.err | tail -1 > gpurun_out/${T}_bench_cfg4.json
python bench.py --config cfg5 --steps 10 --warmup 3 2> gpurun_out/${T}_bench_cfg5.err | tail -1 > gpurun_out/${T}_bench_cfg5.json
python bench.py --config cfg5 --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/${T}_bench_cfg5b.err | tail -1 > gpurun_out/${T}_bench_cfg5b.json
python - "$T" <<'PY'
import json, sys
T = sys.argv[1]
for c in ("cfg2", "cfg2_reference", "cfg1", "cfg3", "cfg4", "cfg5", "cfg5b"):
    try:
        d = json.load(open(f"gpurun_out/{T}_bench_{c}.json"))
        r = d.get("roofline", {})
        print(f"{c:16s} ms/step {d['ms_per_step']:.3f} ms/frame {d.get('ms_per_frame', 0):.2f} value {d['value']:.1f} e2e {d['e2e']['value']:.1f} "
              f"roofline {r.get('bound')} frac {r.get('frac')} cpu {d.get('cpu_baseline', {}).get('value')} {d.get('stage_ms_per_step')}")
    except Exception as e:
        print(c, "FAILED", e)
PY
# launch lists (the same commands as the bench lines above, fewer steps)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${T}_launches_cfg5.csv python bench.py --config cfg5 --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${T}_launches_cfg4.csv python bench.py --config cfg4 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
# ncu --set full of the dominant kernels, one frame / pass after the warm-up; kept in /tmp, summarised here
K='regex:k_trace_wide|k_trace_hw1_fast|k_ppm_photon_shade'
R=/tmp/${T}_reps; mkdir -p $R
ncu --set full --clock-control none -k "$K" -s 18 -c 6 -f -o $R/cfg2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_cfg2.log 2>&1
ncu --set full --clock-control none -k "$K" -s 42 -c 14 -f -o $R/cfg1 python bench.py --config cfg1 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_cfg1.log 2>&1
ncu --set full --clock-control none -k "$K" -s 84 -c 28 -f -o $R/cfg5 python bench.py --config cfg5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_cfg5.log 2>&1
ncu --set full --clock-control none -k "$K" -s 24 -c 4 -f -o $R/cfg3 python bench.py --config cfg3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_cfg3.log 2>&1
ncu --set full --clock-control none -k "$K" -s 30 -c 4 -f -o $R/cfg4 python bench.py --config cfg4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_cfg4.log 2>&1
for c in cfg1 cfg2 cfg3 cfg4 cfg5; do python profiles/summarize.py raw $R/$c.ncu-rep > gpurun_out/${T}_${c}_ncu_full.txt 2>/dev/null; done
python tools/make_traffic.py cfg1=$R/cfg1.ncu-rep cfg2=$R/cfg2.ncu-rep cfg3=$R/cfg3.ncu-rep cfg4=$R/cfg4.ncu-rep cfg5=$R/cfg5.ncu-rep > /dev/null 2>&1
cp profiles/traffic.json gpurun_out/${T}_traffic.json
rm -f gpurun_out/*.ncu-rep
du -sh gpurun_out
