#!/bin/bash
# round 2, multi-GPU run (gpurun --gpus N): peer-memory probe, 2-device tests, bench under torchrun at 2..N ranks
cd "$(dirname "$0")/.."
mkdir -p gpurun_out /tmp/ipc
N=${1:-2}
nvidia-smi topo -m 2>&1 | head -12 > gpurun_out/r02g_topo.txt
(./build/ipc_probe owner /tmp/ipc > gpurun_out/r02g_ipc_owner.log 2>&1 &); ./build/ipc_probe peer /tmp/ipc > gpurun_out/r02g_ipc_peer.log 2>&1; sleep 1; cat gpurun_out/r02g_ipc_*.log
python -m pytest tests/test_z_multi_gpu.py -m gpu -q -x -s 2>&1 | grep -v "^    " | tail -25 > gpurun_out/r02g_pytest.log; tail -12 gpurun_out/r02g_pytest.log
run() { # name n args...
  local name=$1 n=$2; shift 2
  if [ "$n" = 1 ]; then python bench.py --gpus 1 "$@" --no-cpu-baseline 2> gpurun_out/r02g_$name.err | tail -1 > gpurun_out/r02g_$name.json
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n "$@" 2> gpurun_out/r02g_$name.err | grep '^{' | tail -1 > gpurun_out/r02g_$name.json; fi
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/r02g_{n}.json"))
    print(f"{n:16s} gpus {d['n_gpus']} ms/step {d['ms_per_step']:.3f} ms/frame {d['ms_per_frame']:.3f} value {d['value']:.0f} e2e ms {d['e2e']['ms_per_step']:.3f} e2e {d['e2e']['value']:.0f} {d['scaling']}")
except Exception as e:
    print(n, "FAILED", e, open(f"gpurun_out/r02g_{n}.err").read()[-600:])
PY
}
run cfg2_n1 1 --steps 20 --warmup 5
k=2; while [ $k -le $N ]; do run cfg2_n$k $k --steps 20 --warmup 5; k=$((k*2)); done
run cfg2_cam_n$N $N --steps 20 --warmup 5 --split cameras
run cfg5_n1 1 --config cfg5 --steps 10 --warmup 3
k=2; while [ $k -le $N ]; do run cfg5_n$k $k --config cfg5 --steps 10 --warmup 3; k=$((k*2)); done
run cfg4_n1 1 --config cfg4 --steps 3 --warmup 3
run cfg4_n$N $N --config cfg4 --steps 3 --warmup 3
run cfg3_n$N $N --config cfg3 --steps 3 --warmup 3
