#!/bin/bash
# round 2, final confirmation run (1 GPU): cfg3's capture again (it walks the binary tree now), the -m gpu suite, and
# the bench lines of the round's last state with roofline.traffic filled from profiles/traffic.json
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02z
K='regex:k_trace_wide|k_trace_fast|k_trace_hw1_fast|k_ppm_photon_shade'
ncu --set full --clock-control none -k "$K" -s 24 -c 4 -f -o /tmp/${T}_cfg3 python bench.py --config cfg3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_cfg3.log 2>&1
python profiles/summarize.py raw /tmp/${T}_cfg3.ncu-rep > gpurun_out/${T}_cfg3_ncu_full.txt 2>/dev/null
B200RT_TRAFFIC_MERGE=1 python tools/make_traffic.py cfg3=/tmp/${T}_cfg3.ncu-rep > /dev/null 2>&1
cp profiles/traffic.json gpurun_out/${T}_traffic.json
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/${T}_pytest.log; tail -2 gpurun_out/${T}_pytest.log
python bench.py --steps 20 --warmup 5 2> gpurun_out/${T}_bench_cfg2.err | tail -1 > gpurun_out/${T}_bench_cfg2.json
python bench.py --impl reference --steps 2 --warmup 1 2> /dev/null | tail -1 > gpurun_out/${T}_bench_cfg2_reference.json
python bench.py --config cfg1 --steps 20 --warmup 5 2> /dev/null | tail -1 > gpurun_out/${T}_bench_cfg1.json
python bench.py --config cfg3 --steps 3 --warmup 3 2> /dev/null | tail -1 > gpurun_out/${T}_bench_cfg3.json
python bench.py --config cfg4 --steps 3 --warmup 3 2> /dev/null | tail -1 > gpurun_out/${T}_bench_cfg4.json
python bench.py --config cfg5 --steps 10 --warmup 3 2> /dev/null | tail -1 > gpurun_out/${T}_bench_cfg5.json
python - <<'PY'
import json
for c in ("cfg2", "cfg2_reference", "cfg1", "cfg3", "cfg4", "cfg5"):
    try:
        d = json.load(open(f"gpurun_out/r02z_bench_{c}.json"))
        r = d.get("roofline", {})
        print(f"{c:16s} ms/step {d['ms_per_step']:.3f} value {d['value']:.1f} e2e {d['e2e']['value']:.1f} roofline {r.get('bound')} frac {r.get('frac')} traffic {r.get('traffic')} dram_frac {r.get('dram_frac')} cpu {d.get('cpu_baseline', {}).get('value')}")
    except Exception as e:
        print(c, "FAILED", e)
PY
