#!/bin/bash
# A/B of library options on the bench workload: bash tools/gpu_ab.sh "optA=1" "optB=2 optC=3" ...   ("-" = defaults)
mkdir -p gpurun_out
python - "$@" <<'PY'
import sys, json, subprocess, torch
sys.path.insert(0, ".")
import feartracker_b200 as fb
from bench import load_state
from oracle import fear_oracle as fo
net = fb.FEARNet(**fb.FEAR_XS_MODEL_KWARGS); net.load_state_dict(load_state(), strict=True); net = net.cuda().eval(); net.reserve(8)
zt, xt, _, _ = fo.synthetic_crops(5)
ref = None
for arg in sys.argv[1:]:
    opts = [] if arg == "-" else arg.split()
    for kv in opts:
        k, v = kv.split("="); net.set_option(k, v)
    zf = net.get_features(zt.cuda()); out = net.track(xt.cuda(), zf); torch.cuda.synchronize()
    cur = (zf.clone(), out[fo.TARGET_REGRESSION_LABEL_KEY].clone(), out[fo.TARGET_CLASSIFICATION_KEY].clone())
    if ref is None: ref = cur
    print("BITS", arg, all(torch.equal(a, b) for a, b in zip(ref, cur)))
PY
for arg in "$@"; do
  flags=""; if [ "$arg" != "-" ]; then for kv in $arg; do flags="$flags --opt $kv"; done; fi
  python bench.py --steps 20 --warmup 5 --no-stream --no-cpu-baseline --no-parity $flags > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - "$arg" <<'PY'
import json, sys
d = json.load(open("gpurun_out/ab.json"))
print("AB", sys.argv[1], "| fps", round(d["value"]), "ms", round(d["ms_per_step"], 3), {k: round(v["ms_per_step"], 3) for k, v in d["stages"].items() if v["ms_per_step"] > 0.03})
PY
done
