"""GPU: the tcgen05 (tensor-core, 3xTF32) kernels against an fp64 reference, each case in its own
process so a device trap cannot poison the rest of the suite."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "gpurun_out")


def _run(*args, timeout=180):
    proc = subprocess.run([sys.executable, os.path.join(HERE, "tc_check.py"), *map(str, args)], capture_output=True,
                          text=True, timeout=timeout)
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "tc_check_" + "_".join(map(str, args)) + ".log"), "w") as f:
        f.write(proc.stdout + "\n--- stderr ---\n" + proc.stderr)
    lines = [l for l in proc.stdout.splitlines() if l.startswith("TC_CHECK ")]
    assert proc.returncode == 0 and lines, f"tc_check {args} failed: {proc.stderr[-2000:]}"
    return json.loads(lines[-1][len("TC_CHECK "):])


@pytest.mark.parametrize("impl", ["tcgen05"])
@pytest.mark.parametrize("B,Bz", [(1, 1), (3, 3), (5, 1), (200, 200), (256, 256), (300, 1)])
def test_corr_tcgen05(B, Bz, impl):
    res = _run("corr", B, Bz, impl)
    assert res["ffma"]["max_err_rel"] < 1e-5, res["ffma"]
    tc = res["tcgen05"]
    assert tc["x_intact"], "correlation kernel must not touch the x channels"
    assert tc["max_err_rel"] < 1e-5, tc  # 3xTF32 keeps fp32-level accuracy


@pytest.mark.parametrize("pw,corr", [("tcgen05", "ffma"), ("tcgen05", "tcgen05"), ("ffma", "ffma")])
def test_network_with_tensor_core_kernels(pw, corr):
    """Every 1x1 conv (all layer shapes of FEAR-XS) on the tcgen05 GEMM: block-by-block and final maps."""
    res = _run("net", pw, corr, timeout=400)
    bad = {k: v for k, v in res["blocks"].items() if v[1] > 1e-5}
    assert not bad, bad
    assert res["reg"][0] <= 1e-3 and res["reg"][1] <= 1e-3, res
    assert res["cls"][0] <= 1e-3 and res["cls"][1] <= 1e-3, res
    assert res["argmax_same"]


@pytest.mark.parametrize("B", [2, 19])
def test_fused_irf_block(B):
    """xif2_0 as one tcgen05 kernel (expanded tensor on-chip): bit-identical to the three-kernel path, <= 2e-5 of the
    fp64 oracle; B = 19 gives every persistent CTA several tiles (608 search tiles on 148 CTAs)."""
    res = _run("irf", B, timeout=400)
    for name in ("search", "template"):
        r = res[name]
        assert r["vs_oracle"][1] < 2e-5, (name, r)
        assert r["bit_identical"] and r["features_bit_identical"], (name, r)
