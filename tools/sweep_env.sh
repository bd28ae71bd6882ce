#!/usr/bin/env bash
# Sweeps one tuning environment variable of libb200rt over a config (run on the GPU box):
#   tools/sweep_env.sh B200RT_NODE_VOTE "1 4 8 16" cfg4 "--spp 4"
var=$1; vals=$2; cfg=${3:-cfg2}; extra=${4:-}
for v in $vals; do
  env $var=$v python bench.py --config $cfg $extra --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null \
   | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stage_ms_per_step']; print('$cfg $var=$v', round(d['ms_per_step'],3), {k: round(x,3) for k,x in s.items()})"
done
