#!/bin/bash
# usage: gpu_ncu_kernel.sh <kernel regex> <skip> <count> <out name>
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:$1" -s "$2" -c "$3" -f -o "gpurun_out/$4" \
    python bench.py --steps 1 --warmup 3 --no-stream --no-cpu-baseline --no-parity > gpurun_out/ncu_kernel.log 2>&1
tail -2 gpurun_out/ncu_kernel.log | cut -c1-200
