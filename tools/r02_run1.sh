#!/bin/bash
# round 2, first device run: GPU tests on the 8-wide traversal, A/B against the binary walk, occupancy / vote sweep, ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02a_pytest.log; tail -5 gpurun_out/r02a_pytest.log
run() { # name, env...
  local name=$1; shift
  env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/r02a_$name.err | tail -1 > gpurun_out/r02a_$name.json
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/r02a_{n}.json"))
    r = d["roofline"]
    print(n, "ms/frame %.3f" % d["ms_per_step"], "Mrays/s %.0f" % d["value"], d["stage_ms_per_step"], "frac %.3f" % r["frac"], r["per_ray"])
except Exception as e:
    print(n, "FAILED", e)
PY
}
run binary B200RT_NO_WIDE=1
run wide6 B200RT_WIDE_BLOCKS=6
run wide8 B200RT_WIDE_BLOCKS=8
run wide4 B200RT_WIDE_BLOCKS=4
run wide6_vote1 B200RT_WIDE_BLOCKS=6 B200RT_NODE_VOTE=1
run wide6_vote8 B200RT_WIDE_BLOCKS=6 B200RT_NODE_VOTE=8
run wide6_vote16 B200RT_WIDE_BLOCKS=6 B200RT_NODE_VOTE=16
run wide6_refill16 B200RT_WIDE_BLOCKS=6 B200RT_REFILL_CLOSEST=16 B200RT_REFILL_ANY=16
run wide6_refill32 B200RT_WIDE_BLOCKS=6 B200RT_REFILL_CLOSEST=32 B200RT_REFILL_ANY=32
for c in binary wide; do
  e=B200RT_WIDE_BLOCKS=6; [ $c = binary ] && e=B200RT_NO_WIDE=1
  env $e python bench.py --config cfg4 --spp 4 --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/r02a_cfg4_$c.err | tail -1 > gpurun_out/r02a_cfg4_$c.json
  python -c "
import json; d=json.load(open('gpurun_out/r02a_cfg4_$c.json')); print('cfg4 $c', d['ms_per_step'], d['value'], d['stage_ms_per_step'], d['roofline']['per_ray'])"
done
# ncu: launch list of the default path, then full captures of the wide kernel and (for the record) the binary walk's local-memory traffic
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_launches_cfg2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace_wide -s 8 -c 4 -o gpurun_out/r02a_wide \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_ncu2.log 2>&1
B200RT_NO_WIDE=1 ncu --set full --clock-control none -k regex:k_trace_fast -s 8 -c 2 -o gpurun_out/r02a_binary \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_ncu3.log 2>&1
ls -la gpurun_out | tail -20
