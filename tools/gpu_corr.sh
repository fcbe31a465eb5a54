#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tcgen05.py -q -x -k corr 2>&1 | tail -3
timeout 300 python tests/tc_check.py corrperf tcgen05 ffma 2>&1 | tail -1 | cut -c1-400
timeout 600 python bench.py --steps 20 --warmup 5 --no-stream --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_q.json"))
print("fps", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "parity", d["parity_check"]["argmax_exact"], d["parity_check"]["max_rel"])
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "us_per_launch")})
PY
