#!/bin/bash
# round 2: wavefront contexts under the multi-GPU frame split (does overlapping the small waves of one band with the
# large waves of another pay once every GPU has only a slice of the frame?)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-4}
for k in 1 2 3 4; do
  B200RT_CONTEXTS=$k python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/r02v_ctx$k.err | grep '^{' | tail -1 > gpurun_out/r02v_n${N}_ctx$k.json
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r02v_n${N}_ctx$k.json')); print('N=$N contexts $k', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), {a: round(b,3) for a,b in d['stage_ms_per_step'].items()})
except Exception as e: print('ctx$k FAILED', e, open('gpurun_out/r02v_ctx$k.err').read()[-400:])"
done
