#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2_gputest3.log; tail -6 gpurun_out/r2_gputest3.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench3.json"))
print("fps", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches_per_step"], "parity", d["parity_check"]["argmax_exact"], d["parity_check"]["max_rel"], "roof", round(d["roofline"]["frac"],3), "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
print({k: round(v["ms_per_step"], 3) for k, v in d["stages"].items()})
print("stream", json.dumps(d["stream"])[:1500])
PY
