#!/usr/bin/env python
"""Micro-benchmark of the implicit-GEMM kernel on the cfg4 convolution shapes (GPU only).

  python tools/gemm_probe.py [--batch 4] [--reps 20] [--variants v1,v2,...]

For every (shape, variant) prints time, algorithmic TFLOP/s and cycles per k-step per SM.
Variants are `key=value` overrides joined with '+': stages=N, stats=0, repk=R (repeat the k-table
R times: isolates the steady-state k-step cost from per-tile overhead), bn=N.  Used to decide
kernel design questions; numbers quoted in profiles/README.md come from here."""
import argparse
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_image_generation_b200 import _lib, plan as P  # noqa: E402

SHAPES = [
    # name, spatial, [cin...], cout, k
    ("L0 128->128", (40, 48, 40), [128], 128, 3),
    ("L0 256->128", (40, 48, 40), [128, 128], 128, 3),
    ("L1 256->256", (20, 24, 20), [256], 256, 3),
    ("L2 512->512", (10, 12, 10), [512], 512, 3),
    # epilogue-bound shapes: thin K (a 1x1x1 convolution, n_kb = 2, like in_conv / out_conv.taps /
    # the attention logits) and the 2D models' top levels (9 taps x 1-2 slabs)
    ("T0 1x1 128->128", (40, 48, 40), [128], 128, 1),
    ("2D 64->64 240", (240, 240), [64], 64, 3),
    ("2D 128->128 120", (120, 120), [128], 128, 3),
]


def build(B, sp, cins, cout, k, var):
    dev = "cuda"
    acts = [torch.randn(B, *sp, c, device=dev).to(torch.bfloat16) for c in cins]
    cin = sum(cins)
    w = torch.randn(cout, cin, *([k] * len(sp)), device=dev) / (cin * k ** len(sp)) ** 0.5
    wm = P.pack_conv_weight(w, splits=cins)
    y = torch.zeros(B, *sp, cout, dtype=torch.bfloat16, device=dev)
    bias = torch.randn(cout, device=dev)
    stats = torch.zeros(B, 8, 2, device=dev, dtype=torch.float64) if var.get("stats", 1) else None
    pl = P.conv_plan([P.ConvSource(a) for a in acts], wm, y, k, bias=bias, stats=stats,
                     stats_cpg=cout // 8 if stats is not None else 0,
                     block_n=var.get("bn") or None, stages=var.get("stages", 0),
                     xreuse=int(var["xr"]) if "xr" in var else None)
    if "sched" in var:
        pl.sched = var["sched"]
    if "swap" in var:
        pl.swap_ab = bool(var["swap"])
    r = var.get("repk", 1)
    if r > 1:
        pl.ktable = np.tile(pl.ktable, (1, r, 1))
    for key in ("mt", "cluster"):
        if key in var and hasattr(pl, key):
            setattr(pl, key, var[key])
    if var.get("trace"):
        pl.trace = torch.zeros(148 * 3, 16, dtype=torch.int64, device=dev)
    pl.materialize(dev)
    return pl, (acts, wm, y, bias, stats)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--ms", type=float, default=300.0, help="sustained run length per variant")
    ap.add_argument("--variants", default="base")
    ap.add_argument("--shapes", default="")
    args = ap.parse_args()
    _lib.require_device()
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    for name, sp, cins, cout, k in SHAPES:
        if args.shapes and not any(s in name for s in args.shapes.split(",")):
            continue
        for vs in args.variants.split(","):
            var = {}
            if vs != "base":
                for kv in vs.split("+"):
                    a, b = kv.split("=")
                    var[a] = int(b)
            try:
                pl, keep = build(args.batch, sp, cins, cout, k, var)
            except Exception as e:  # noqa: BLE001
                print(f"{name:14s} {vs:24s} build failed: {e}")
                continue
            for _ in range(3):
                pl.launch()
            torch.cuda.synchronize()
            # sustained: back-to-back launches for ~args.ms milliseconds, SM clock sampled by NVML
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pl.launch()
            e1.record()
            torch.cuda.synchronize()
            n = max(3, int(args.ms / max(e0.elapsed_time(e1), 1e-3)))
            clocks = []
            stop = threading.Event()

            def sample():
                while not stop.is_set():
                    clocks.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                    time.sleep(0.005)

            th = threading.Thread(target=sample)
            e0.record()
            th.start()
            for _ in range(n):
                pl.launch()
            e1.record()
            torch.cuda.synchronize()
            stop.set()
            th.join()
            ms = e0.elapsed_time(e1) / n
            clocks = sorted(clocks[len(clocks) // 3:]) or [0]
            mhz = clocks[len(clocks) // 2]
            flops = pl.flops * var.get("repk", 1)
            tiles = pl.grid()
            ksteps = pl.flops * var.get("repk", 1) / (2.0 * 128 * 128 * 64)
            cyc = ms * 1e-3 * mhz * 1e6 * 148 / ksteps
            print(f"{name:14s} {vs:24s} {ms:8.4f} ms  {flops / ms / 1e9:8.1f} TF/s  grid {tiles:5d} "
                  f"n_kb {pl.n_kb:4d} bn {pl.block_n:3d} sched {pl._args.sched} {mhz:5d} MHz "
                  f"{cyc:6.0f} cyc/(128x128x64 step)/SM  pipe {25600.0 / cyc:5.1f}%", flush=True)
            if pl.trace is not None:
                pl.trace.zero_()
                pl.launch()
                torch.cuda.synchronize()
                full = pl.trace.cpu().numpy().astype(np.int64)
                G = min(148, pl.grid())
                tr = full[:148]
                if full[G:3 * G].any():  # fine epilogue trace (library built with -DMRI_GEMM_FINE_TRACE)
                    names = ["decode_tile", "wait tmem_full", "box head + bar 1", "residual / add_c",
                             "chunk loop", "stats tail", "bar 2", "store issue", "tile tail", "proxy fence"]
                    for who, rows in (("thread 0", full[G:2 * G]), ("thread 96", full[2 * G:3 * G])):
                        tot = rows[:, :10].sum(1)
                        print(f"    fine trace, epilogue {who}: total {int(np.median(tot))} cycles (median over CTAs)")
                        for k, nm in enumerate(names):
                            print(f"        {nm:20s} {int(np.median(rows[:, k])):9d}")
                tr = tr[tr[:, 0] > 0]
                t0 = tr[:, 0].min()
                print(f"    trace over {len(tr)} CTAs (ns from first CTA entry / cycles from CTA setup):")
                print(f"    CTA entry spread {int((tr[:, 0] - t0).max())} ns; exit (last) {int((tr[:, 6] - t0).max())} ns,"
                      f" exit (median) {int(np.median(tr[:, 6] - t0))} ns")
                for nm, col in (("first stage landed", 2), ("last MMA issued", 3), ("last accumulator done", 4),
                                ("epilogue done", 5)):
                    d = tr[:, col] - tr[:, 1]
                    print(f"    {nm:22s} min {int(d.min()):8d}  median {int(np.median(d)):8d}  max {int(d.max()):8d} cycles")
                for nm, col in (("MMA waits for operands", 7), ("MMA waits for epilogue", 8),
                                ("producer waits (ring full)", 9), ("producer waits (x ring)", 10), ("segments", 11),
                                ("epilogue waits for MMA", 12), ("epilogue waits for staging", 13)):
                    d = tr[:, col]
                    print(f"    {nm:26s} min {int(d.min()):8d}  median {int(np.median(d)):8d}  max {int(d.max()):8d}")
            del pl, keep


if __name__ == "__main__":
    main()
