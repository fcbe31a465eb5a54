#!/usr/bin/env python
"""bench.py -- FEAR-XS per-frame inference throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

A "step" = one pass of the hot path (FEARNet.track + box decode) over one batch of synthetic
crops: 256 search crops (3x256x256 fp32, ImageNet-normalised uniform uint8, seed 20260924) with
their 256 template feature maps, per GPU (BASELINE config 2; at N GPUs each rank owns a contiguous
256-frame shard of the 256*N batch -- config 4 at N=8 -- and the step ends with ONE all-gather of
the 48-byte box records).  One JSON line on stdout (rank 0).  See DESIGN.md section "Measurement".
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "FEAR-XS frames/sec (track + decode, batch 256 per GPU)"
CORR_BYTES_PER_FRAME = 2 * (65536 + 262144 + 65536)  # 2 branches x (z + x + s) fp32, SURVEY.md 8(d)
PATH_BYTES_PER_FRAME = 14_947_328  # block-fused budget of the whole track(), SURVEY.md 8(d)
SEED = 20260924


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_state():
    with np.load(os.path.join(ROOT, "tests", "golden", "fear_xs_hotpath_state.npz")) as f:
        hot = {k: torch.from_numpy(f[k]) for k in f.files}
    with open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")) as f:
        keys = json.load(f)
    return {k: hot[k] if k in hot else torch.zeros(s, dtype=getattr(torch, d)) for k, (s, d) in keys.items()}


def synthetic_batch(batch, rank, with_u8=False):
    """uint8-derived, ImageNet-normalised crops; every rank regenerates its own shard from the seed."""
    g = torch.Generator().manual_seed(SEED + rank)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1) * 255.0
    inv = 1.0 / (torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1) * 255.0)
    zu = torch.randint(0, 256, (batch, 3, 128, 128), generator=g, dtype=torch.uint8)
    xu = torch.randint(0, 256, (batch, 3, 256, 256), generator=g, dtype=torch.uint8)
    if with_u8:
        return (zu.float() - mean) * inv, (xu.float() - mean) * inv, xu
    return (zu.float() - mean) * inv, (xu.float() - mean) * inv


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, str(gpu_index)
        self.mark_a = self.mark_b = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", self.gpu], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def wait_first_sample(self, timeout=5.0):
        t0 = time.time()
        while not self.rows and time.time() - t0 < timeout and self.proc is not None:
            time.sleep(0.05)

    def begin(self):
        self.mark_a = len(self.rows)

    def end(self):
        time.sleep(0.25)  # let the sample(s) covering the end of the timed region arrive
        self.mark_b = len(self.rows)

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        a = max(0, (self.mark_a or 0) - 1)
        rows = self.rows[a:self.mark_b] if self.mark_b is not None else self.rows[a:]
        rows = rows or self.rows[-3:]
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) >= 9:
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def pick_cpu_threads():
    """Thread count for the CPU baseline: the host may expose far more logical CPUs than this container
    can actually run on (oversubscription makes torch's intra-op pool collapse), so probe a few counts
    on a 4-frame slice and keep the fastest."""
    from oracle import fear_oracle as fo

    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    sd = {k: v for k, v in load_state().items() if v.is_floating_point()}
    zt, xt = synthetic_batch(4, 0)
    zf = fo.get_features(sd, zt)
    best, best_t = 1, float("inf")
    for n in sorted({min(avail, c) for c in (8, 16, 32, 64, avail)}):
        torch.set_num_threads(n)
        fo.track(sd, xt[:1], zf[:1])
        t0 = time.perf_counter()
        fo.track(sd, xt, zf)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = n, dt
        if dt > 20:
            break
    return best, 4 / best_t


def time_cpu_oracle(batch, steps, warmup, threads):
    """The oracle port (reference source restated, torch CPU fp32) on the host cores: frames/s."""
    from oracle import fear_oracle as fo

    torch.set_num_threads(threads)
    sd = {k: v for k, v in load_state().items() if v.is_floating_point()}
    zt, xt = synthetic_batch(batch, 0)
    zf = fo.get_features(sd, zt)
    for _ in range(warmup):
        fo.decode(*[fo.track(sd, xt, zf)[k] for k in (fo.TARGET_REGRESSION_LABEL_KEY, fo.TARGET_CLASSIFICATION_KEY)])
    t0 = time.perf_counter()
    for _ in range(steps):
        out = fo.track(sd, xt, zf)
        fo.decode(out[fo.TARGET_REGRESSION_LABEL_KEY], out[fo.TARGET_CLASSIFICATION_KEY])
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path (oracle port) on all host
    threads, same metric/config; rank 0 only."""
    if rank != 0:
        return
    threads, probe_fps = pick_cpu_threads()
    # bounded sample: size the per-step slice so warm-up + K steps stay around a minute
    sample = int(max(1, min(32, probe_fps * 60.0 / (args.steps + 2))))
    fps, ms = time_cpu_oracle(sample, args.steps, max(1, min(args.warmup, 2)), threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "FEAR-XS batched inference, batch=256 synthetic crops (BASELINE config 2)",
                   "sample": f"{sample}-frame slice of the workload per step"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"oracle port (reference source restated, torch CPU fp32), track()+decode on "
                                   f"{sample} frames x {args.steps} steps"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=256, help="frames per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--corr", default=None, help="correlation kernel implementation: ffma | tcgen05")
    ap.add_argument("--pw", default=None, help="1x1-conv implementation: ffma | tcgen05")
    ap.add_argument("--dw", default=None, help="depthwise implementation: pixel | strip | roll | auto")
    ap.add_argument("--dw-wide", type=int, default=None)
    ap.add_argument("--fuse", type=int, default=None, help="1 = fused pw-expand+dw kernels for the stride-2 blocks")
    ap.add_argument("--fuse-stem", type=int, default=None, help="0 = separate stem / xif1_0 kernels")
    ap.add_argument("--early-sub", type=int, default=None, help="sub-batch (frames) of the high-resolution blocks")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE",
                    help="extra fear_set_option pairs (experiments), e.g. --opt small_const=0")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import feartracker_b200 as fb
    from feartracker_b200 import _lib, sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the FEAR hot path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout must carry exactly ONE JSON line.  NCCL prints its version banner on stdout when the communicator is
        # created (NCCL_DEBUG_FILE is read too early to help from here), so fd 1 points at stderr while the process
        # group and its communicator come up (the barrier forces the lazy creation), then it is restored.
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier(device_ids=[local_rank])
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    B, total = args.batch, args.batch * world

    net = fb.FEARNet(**fb.FEAR_XS_MODEL_KWARGS)
    net.load_state_dict(load_state(), strict=True)
    net = net.to(dev).eval()
    net.reserve(B)
    if args.corr:
        net.set_option("corr", args.corr)
        _lib.check(_lib.load().fear_set_option(None, b"corr", args.corr.encode()), "fear_set_option")
    if args.pw:
        net.set_option("pw", args.pw)
    if args.dw:
        net.set_option("dw", args.dw)
    if args.dw_wide is not None:
        net.set_option("dw_wide", str(args.dw_wide))
    if args.fuse is not None:
        net.set_option("fuse", str(args.fuse))
    if args.fuse_stem is not None:
        net.set_option("fuse_stem", str(args.fuse_stem))
    for kv in args.opt:
        k, v = kv.split("=", 1)
        net.set_option(k, v)
    if args.early_sub is not None:
        net.set_option("early_sub", str(args.early_sub))

    zt, xt, xu = synthetic_batch(B, rank, with_u8=True)
    x_host, z_dev = xt.pin_memory(), net.get_features(zt.to(dev))
    xu_host = xu.permute(0, 2, 3, 1).contiguous().pin_memory()  # raw uint8 HWC crops, as the tracker holds them
    zf_host = z_dev.cpu().pin_memory()
    x_dev = x_host.to(dev)
    box_host = torch.empty((total, 48), dtype=torch.uint8).pin_memory()
    stream = torch.cuda.current_stream(dev)

    def step_device():
        boxes = net.track_boxes(x_dev, z_dev)
        return sharding.all_gather_boxes(boxes, total)

    def step_e2e(src=None):
        # public API on pinned HOST buffers: chunked H2D overlapped with compute, boxes copied back
        boxes = net.track_boxes_from_host(xu_host if src is None else src, zf_host)
        boxes = sharding.all_gather_boxes(boxes, total)
        box_host.copy_(boxes, non_blocking=True)
        return boxes

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        a.record(stream)
        for _ in range(steps):
            fn()
        b.record(stream)
        sync_all()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    clocks = ClockSampler(local_rank).__enter__()
    for _ in range(args.warmup):
        step_device()
    clocks.wait_first_sample()
    l0 = net.launch_count()
    clocks.begin()
    ms_total = timed(step_device, args.steps)
    if ms_total < 400.0:  # keep the GPU under the same load until nvidia-smi (100 ms period) has sampled it
        extra = int(400.0 / (ms_total / args.steps)) + 1
        for _ in range(extra):
            step_device()
        torch.cuda.synchronize(dev)
    clocks.end()
    clocks.__exit__()
    launches = (net.launch_count() - l0) // (args.steps + (extra if ms_total < 400.0 else 0))
    fps = total * args.steps / (ms_total * 1e-3)

    for _ in range(args.warmup):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    fps_e2e = total * args.steps / (ms_e2e * 1e-3)
    for _ in range(args.warmup):
        step_e2e(x_host)
    ms_e2e32 = timed(lambda: step_e2e(x_host), args.steps)

    # ---- per-stage device time over another K steps (CUDA events around every launch, same stream) ----
    net.profile(True)
    sync_all()
    for _ in range(args.steps):
        step_device()
    sync_all()
    stages = net.stage_times()
    net.profile(False)
    corr_ms, corr_n = stages["corr"]
    peak, peak_src = load_peaks()
    roofline = None
    if corr_n:
        per_launch_s = corr_ms * 1e-3 / corr_n
        bytes_per_launch = CORR_BYTES_PER_FRAME * B / (corr_n / args.steps)  # one launch per branch (per chunk)
        achieved = bytes_per_launch / per_launch_s / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "corr_traffic.json")
        if os.path.isfile(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        roofline = {
            "kernel": "pixel-wise correlation (fear_corr_nhwc_f32, one launch per branch)", "bound": "hbm",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "peak_source": peak_src, "us_per_launch": per_launch_s * 1e6,
            "algorithmic_bytes_per_launch": int(bytes_per_launch),
        }
    step_ms_sum = sum(v[0] for v in stages.values()) / args.steps
    path_gbs = PATH_BYTES_PER_FRAME * B / (ms_total / args.steps * 1e-3) / 1e9
    stage_report = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps,
                        "share": (v[0] / args.steps) / step_ms_sum if step_ms_sum else None}
                    for k, v in stages.items()}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads, probe_fps = pick_cpu_threads()
        sample = int(max(1, min(32, probe_fps * 3.0)))
        cfps, cms = time_cpu_oracle(sample, 6, 1, threads)
        cpu = {"value": cfps, "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": f"oracle port (reference source restated, torch CPU fp32): track()+decode on a {sample}-frame "
                         f"slice of the workload, 1 warm-up + 6 timed steps ({cms:.0f} ms/step), best of probed "
                         "thread counts"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": "FEAR-XS batched inference, batch=256 synthetic crops per GPU (BASELINE config 2; "
                            "config 4 at 8 GPUs), FEAR-XS checkpoint weights",
                "global_batch": total, "per_gpu_batch": B, "parallelism": f"frames sharded over {world} rank(s)",
                "l2": "inputs larger than L2 (201 MB search batch per step; >2 GB of workspace traffic per step)",
                "impl": {"corr": args.corr or "default", "pw": args.pw or "default", "dw": args.dw or "default",
                         "early_sub": args.early_sub},
            },
            "e2e": {"value": fps_e2e, "unit": "frames/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(xu_host.numel() + zf_host.numel() * 4),
                    "d2h_bytes_per_step": int(box_host.numel()),
                    "api": "FEARNet.track_boxes_from_host on pinned host buffers: raw uint8 HWC crops (what "
                           "FEARTracker holds; ImageNet normalisation fused into the stem kernel, bit-identical to "
                           "host normalisation) + fp32 template features; double-buffered staging, so the copy of step "
                           "i+1 overlaps the kernels of step i; box records copied back every step"},
            "e2e_fp32_inputs": {"value": total * args.steps / (ms_e2e32 * 1e-3), "unit": "frames/s",
                                "ms_per_step": ms_e2e32 / args.steps,
                                "h2d_bytes_per_step": int(x_host.numel() * 4 + zf_host.numel() * 4),
                                "d2h_bytes_per_step": int(box_host.numel()),
                                "api": "same call with host-normalised fp32 (B,3,256,256) crops (PCIe-bound)"},
            "gpu_launches": int(launches * args.steps),
            "gpu_launches_per_step": int(launches),
            "clocks": clocks.summary(),
            "roofline": roofline,
            "roofline_path": {"bound": "hbm", "achieved": path_gbs, "peak": peak, "unit": "GB/s",
                              "frac": path_gbs / peak, "what": "whole track() against the block-fused byte budget "
                                                               "(14,947,328 B/frame)"},
            "stages": stage_report,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
