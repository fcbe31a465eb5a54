#!/bin/bash
python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
import feartracker_b200 as fb
from bench import load_state
from oracle import fear_oracle as fo
net = fb.FEARNet(**fb.FEAR_XS_MODEL_KWARGS); net.load_state_dict(load_state(), strict=True); net = net.cuda().eval(); net.reserve(8)
_, xt, _, _ = fo.synthetic_crops(3)
x = xt.cuda()
for n in (5, 6, 7, 8, 9):
    net.set_option("fuse_dwpw", "3"); a = net.backbone_prefix(x, n)
    net.set_option("fuse_dwpw", "7"); b = net.backbone_prefix(x, n)
    d = (a - b).abs()
    print("blocks", n, "shape", tuple(a.shape), "equal", bool(torch.equal(a, b)), "max abs diff", float(d.max()), "ref max", float(a.abs().max()),
          "bad frac", float((d > 0).float().mean()))
    if n == 6 and not torch.equal(a, b):
        bad = (d[0] > 0).any(0)   # (H, W)
        print("bad rows", bad.any(1).nonzero().flatten().tolist()[:40]); print("bad cols", bad.any(0).nonzero().flatten().tolist()[:40])
        print("bad channels", (d[0] > 0).flatten(1).any(1).nonzero().flatten().tolist())
PY
