#!/bin/bash
# round 2, last confirmation run (1 GPU) after the class-ordered shade tiles and the two-level gate: the -m gpu suite,
# the bench lines, cfg4's launch list and one full capture of the two HW7 shade kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02zz
timeout 420 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/${T}_pytest.log; tail -2 gpurun_out/${T}_pytest.log
python bench.py --steps 20 --warmup 5 2> gpurun_out/${T}_bench_cfg2.err | tail -1 > gpurun_out/${T}_bench_cfg2.json
python bench.py --config cfg4 --steps 3 --warmup 3 2> /dev/null | tail -1 > gpurun_out/${T}_bench_cfg4.json
python bench.py --config cfg1 --steps 20 --warmup 5 --no-cpu-baseline 2> /dev/null | tail -1 > gpurun_out/${T}_bench_cfg1.json
python bench.py --config cfg3 --steps 3 --warmup 3 --no-cpu-baseline 2> /dev/null | tail -1 > gpurun_out/${T}_bench_cfg3.json
python bench.py --config cfg5 --steps 10 --warmup 3 --no-cpu-baseline 2> /dev/null | tail -1 > gpurun_out/${T}_bench_cfg5.json
python - <<'PY'
import json
for c in ("cfg2", "cfg4", "cfg1", "cfg3", "cfg5"):
    try:
        d = json.load(open(f"gpurun_out/r02zz_bench_{c}.json"))
        r = d.get("roofline", {})
        print(f"{c:6s} ms/step {d['ms_per_step']:.3f} value {d['value']:.1f} e2e {d['e2e']['value']:.1f} ({d['e2e'].get('ms_per_step')}) frac {r.get('frac')} cpu {d.get('cpu_baseline', {}).get('value')} stages {d.get('stage_ms_per_step')}")
    except Exception as e:
        print(c, "FAILED", e)
PY
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${T}_launches_cfg4.csv python bench.py --config cfg4 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 150 ncu --set full --clock-control none -k "regex:k_shade7" -s 8 -c 2 -f -o /tmp/${T}_shade7 python bench.py --config cfg4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu_shade7.log 2>&1
python profiles/summarize.py raw /tmp/${T}_shade7.ncu-rep > gpurun_out/${T}_k_shade7_ncu_full.txt 2>/dev/null
tail -3 gpurun_out/${T}_k_shade7_ncu_full.txt
