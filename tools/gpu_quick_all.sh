#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_tcgen05.py -q -x 2>&1 | tail -4
python -m pytest tests/test_gpu_parity.py -q -x -k "block_by_block or seed0 or synthetic or chunking or fused or batch256" 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 --no-stream --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_q.json"))
print("fps", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches_per_step"], "parity", d["parity_check"]["argmax_exact"], d["parity_check"]["max_rel"])
print({k: round(v["ms_per_step"], 3) for k, v in d["stages"].items()})
PY
