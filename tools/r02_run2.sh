#!/bin/bash
# round 2, second device run: profiles of the PRIMARY-wave launches (closest hit d=0, any hit d=0) of both traversal kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export B200RT_WIDE_BLOCKS=8
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_trace_wide -s 12 -c 2 -o gpurun_out/r02b_wide8_d0 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_ncu1.log 2>&1
B200RT_NO_WIDE=1 ncu --set full --clock-control none --import-source on -k regex:k_trace_fast -s 12 -c 2 -o gpurun_out/r02b_binary_d0 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_ncu2.log 2>&1
B200RT_NO_WIDE=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches_cfg2_binary.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_ncu3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches_cfg2_wide8.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_ncu4.log 2>&1
ls -la gpurun_out | grep r02b
