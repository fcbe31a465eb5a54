#!/usr/bin/env python
"""Profiling aid (needs a library built with FEAR_NVCC_FLAGS=-DFEAR_IRF_TIMING): per-warp cycle breakdown of the
worker warps of irf_s2_fused_kernel on the bench workload.  python tools/irf_timing.py [B]"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import feartracker_b200 as fb  # noqa: E402
from feartracker_b200 import _lib  # noqa: E402
from bench import load_state  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
net = fb.FEARNet(**fb.FEAR_XS_MODEL_KWARGS)
net.load_state_dict(load_state(), strict=True)
net = net.cuda().eval()
net.reserve(B)
x = torch.randn(B, 3, 256, 256, device="cuda")
net.get_features(x)
lib = _lib.load()
fn = lib.fear_debug_irf_timing
fn.restype, fn.argtypes = ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
out = np.zeros(148 * 128, dtype=np.uint64)
assert fn(net._handle, out.ctypes.data, out.size) == 1
for _ in range(3):
    net.get_features(x)
torch.cuda.synchronize()
assert fn(net._handle, out.ctypes.data, out.size) == 0
t = out.reshape(148, 8, 16).astype(np.int64)
names = ["wait acc_full", "epilogue", "bar slab done", "wait a2_empty", "depthwise", "bar slab free", "total", "in tcgen05.wait::ld", "epi:arrive", "epi:prefetch issue", "epi:process", "-"]
print("cycles per worker warp over the kernel (mean over 148 CTAs); tiles per CTA ~", B * 32 / 148)
for g in range(2):
    m = t[:, g * 4:(g + 1) * 4, :].mean(axis=(0, 1))
    print(f"group {g}: " + ", ".join(f"{n} {m[i]:.0f} ({100 * m[i] / m[6]:.0f}%)" for i, n in enumerate(names)))
print("per-quadrant epilogue cycles (group 0):", t[:, :4, 1].mean(0).round())
print("per-quadrant depthwise cycles (group 0):", t[:, :4, 4].mean(0).round())
