#!/bin/bash
# ncu full capture of the fused IRF kernel at the bench workload + a step launch table
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:irf_s2 -s 3 -c 1 -f -o gpurun_out/r2_irf_v1 \
    python bench.py --steps 1 --warmup 3 --no-stream --no-cpu-baseline --no-parity > gpurun_out/ncu_irf.log 2>&1
tail -3 gpurun_out/ncu_irf.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none --launch-skip 236 --launch-count 80 --csv --log-file gpurun_out/r2_step_metrics_v1.csv \
    python bench.py --steps 1 --warmup 3 --no-stream --no-cpu-baseline --no-parity > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log
python -m pytest tests/test_gpu_parity.py -q -x -k "batch256 or update or smooth or graph or compat or workspace" 2>&1 | tail -5
