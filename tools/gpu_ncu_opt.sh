#!/bin/bash
# usage: gpu_ncu_opt.sh <kernel regex> <skip> <count> <out name> <bench opt flags...>
mkdir -p gpurun_out
k=$1; s=$2; c=$3; o=$4; shift 4
ncu --set full --clock-control none --import-source on -k "regex:$k" -s "$s" -c "$c" -f -o "gpurun_out/$o" \
    python bench.py --steps 1 --warmup 3 --no-stream --no-cpu-baseline --no-parity "$@" > gpurun_out/ncu_kernel.log 2>&1
tail -1 gpurun_out/ncu_kernel.log | cut -c1-200
