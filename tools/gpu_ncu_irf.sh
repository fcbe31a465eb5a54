#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:irf_s2 -s 3 -c 1 -f -o gpurun_out/r2_irf_v4 \
    python bench.py --steps 1 --warmup 3 --no-stream --no-cpu-baseline --no-parity > gpurun_out/ncu_irf.log 2>&1
tail -2 gpurun_out/ncu_irf.log | cut -c1-300
