"""Profiling build only (FEAR_NVCC_FLAGS=-DFEAR_PW_TIMING): per-launch, per-role cycle breakdown of pw_tc_kernel over one
track() step of the bench workload (cycles summed over CTAs; printed per CTA)."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feartracker_b200 as fb  # noqa: E402
from bench import load_state  # noqa: E402
from feartracker_b200 import _lib  # noqa: E402
from oracle import fear_oracle as fo  # noqa: E402

B = 256
net = fb.FEARNet(**fb.FEAR_XS_MODEL_KWARGS)
net.load_state_dict(load_state(), strict=True)
net = net.cuda().eval()
net.reserve(B)
lib = _lib.load()
fn = lib.fear_debug_pw_timing
fn.restype, fn.argtypes = ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]
zt, xt, _, _ = fo.synthetic_crops(B)
x = xt.cuda()
zf = net.get_features(zt.cuda())
for _ in range(2):
    net.track_boxes(x, zf)
cnt = np.zeros((64, 32), dtype=np.uint64)
info = np.zeros((64, 8), dtype=np.int32)
fn(cnt.ctypes.data, info.ctypes.data)
net.track_boxes(x, zf)
n = fn(cnt.ctypes.data, info.ctypes.data)
c = cnt.astype(np.float64)
print(f"{n} pw_tc launches in one step; cycles per CTA (k = 1000 clk)")
print(" id      M    N    K dw mw tiles ch | prod:wE  tot | mma:wAcc wSplit(wA)  issue   tot | grp0:wFull compute wBoxE wEmpty write   tot | epi:wFull  work")
for i in range(min(n, 64)):
    M, N, K, dwk, mw, grid, tiles, chunks = info[i]
    g = max(grid, 1)
    v = c[i] / g / 1000.0
    print(f"{i:3d} {M:7d} {N:4d} {K:4d} {dwk:2d} {mw:2d} {tiles:5d} {chunks:2d} | {v[0]:7.1f} {v[1]:5.1f} | {v[2]:8.1f} {v[3]:6.1f}({v[21]:5.1f}) {v[4]:6.1f} {v[5]:5.1f} |"
          f" {v[6]:10.1f} {v[7]:7.1f} {v[8]:5.1f} {v[9]:6.1f} {v[10]:5.1f} {v[11]:5.1f} | {v[18]:9.1f} {v[19]:5.1f}")
