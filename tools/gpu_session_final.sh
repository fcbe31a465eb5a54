#!/bin/bash
# Final-build evidence: GPU suite, default bench line, ncu step table (time + DRAM + L2 per launch), irf/corr details.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r2_gputest_final.log; tail -3 gpurun_out/r2_gputest_final.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none --launch-skip 140 --launch-count 120 --csv --log-file gpurun_out/r2_step_metrics_final.csv \
    python bench.py --steps 1 --warmup 3 --no-stream --no-cpu-baseline --no-parity > gpurun_out/ncu_step.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:corr_ts -s 3 -c 1 -f -o gpurun_out/r2_corr_final \
    python bench.py --steps 1 --warmup 3 --no-stream --no-cpu-baseline --no-parity > gpurun_out/ncu_corr.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:irf_s2 -s 3 -c 1 -f -o gpurun_out/r2_irf_final \
    python bench.py --steps 1 --warmup 3 --no-stream --no-cpu-baseline --no-parity > gpurun_out/ncu_irf.log 2>&1
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_final.json"))
print("fps", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches_per_step"], "roof", round(d["roofline"]["frac"],3), "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], "stream", d["stream"]["modes"]["gpu_crop"]["value"])
PY
head -c 400 gpurun_out/r2_bench_reference.json
