#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tests/tc_check.py irf 19 > gpurun_out/irf19.log 2>&1; echo "irf rc=$?"; tail -c 1200 gpurun_out/irf19.log
python bench.py --steps 20 --warmup 5 --no-stream --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_q.json"))
print("fps", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches_per_step"], "parity", d["parity_check"]["argmax_exact"], d["parity_check"]["max_rel"])
print({k: round(v["ms_per_step"], 3) for k, v in d["stages"].items()})
PY
