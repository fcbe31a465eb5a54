#!/bin/bash
# Round-2 GPU session 1: validate the fused xif2_0 kernel in isolation, then the GPU suite and the bench.
mkdir -p gpurun_out
timeout 400 python tests/tc_check.py irf 19 > gpurun_out/r2_irf19.log 2>&1
echo "irf rc=$?"; tail -c 2500 gpurun_out/r2_irf19.log
if grep -q '"bit_identical": true, "features_bit_identical": true' gpurun_out/r2_irf19.log && ! grep -q '"bit_identical": false' gpurun_out/r2_irf19.log; then
  echo "IRF_OK"
else
  echo "IRF_BROKEN -> rebuilding with fuse_irf off by default for the rest of this session"
  sed -i 's/int fuse_irf = 1;/int fuse_irf = 0;/' feartracker_b200/csrc/fear_context.cu
  python -m feartracker_b200.build --force > gpurun_out/rebuild.log 2>&1
fi
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_tcgen05.py::test_fused_irf_block 2>&1 | tail -40 > gpurun_out/r2_gputest1.log
tail -6 gpurun_out/r2_gputest1.log
python bench.py --steps 20 --warmup 5 --opt fuse_irf=0 --no-stream --no-cpu-baseline > gpurun_out/r2_bench_unfused.json 2> gpurun_out/r2_bench_unfused.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
python - <<'PY'
import json
for f in ("r2_bench_unfused", "r2_bench1"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "fps", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches_per_step"],
              "parity", d.get("parity_check"), "roof", d["roofline"]["frac"] if d.get("roofline") else None)
        print({k: round(v["ms_per_step"], 3) for k, v in d["stages"].items()})
        if d.get("stream"): print("stream", json.dumps(d["stream"])[:900])
    except Exception as e:
        print(f, "unreadable", e)
PY
