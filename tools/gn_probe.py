#!/usr/bin/env python
"""Achieved HBM GB/s of the GroupNorm kernels (forward apply, backward reduce / apply) at the
cfg5 level shapes (GPU only).  Algorithmic bytes: apply 2+2 B/elt, bwd reduce 2+2, bwd apply 2+2+2.
Buffers are rotated so that consecutive launches never find their input in the 126 MB L2.
  B=8 python tools/gn_probe.py"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_image_generation_b200 import ops

B = int(os.environ.get("B", "8"))
PEAK = 6555.2
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
dev = "cuda"
shapes = [(76800, 128), (76800, 256), (9600, 256), (9600, 512), (1200, 512), (1200, 1024)]
if os.environ.get("CFG") == "2d":
    shapes = [(240 * 240, 64), (240 * 240, 128), (120 * 120, 128), (60 * 60, 256), (30 * 30, 512)]


def timed(fn, nrot, iters=20):
    for i in range(3):
        fn(i % nrot)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for i in range(iters):
        fn(i % nrot)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3


rows = []
for spatial, C in shapes:
    n = B * spatial * C
    nrot = max(2, min(8, int(400e6 // (n * 2)) + 1))  # rotate > L2 worth of input
    xs = [torch.randn(B, spatial, C, device=dev).bfloat16() for _ in range(nrot)]
    dys = [torch.randn(B, spatial, C, device=dev).bfloat16() for _ in range(nrot)]
    outs = [torch.empty_like(xs[0]) for _ in range(nrot)]
    groups, cpg = 8, C // 8
    stats = torch.zeros(B, groups, 2, device=dev, dtype=torch.float64)
    ops.gn_stats(xs[0], stats, B, spatial, C, cpg)
    gamma = torch.rand(C, device=dev) + 0.5
    beta = torch.randn(C, device=dev) * 0.1
    sums = torch.zeros(3, B, C, device=dev, dtype=torch.float64)
    col = torch.zeros(B, C, device=dev, dtype=torch.float64)
    t_fwd = timed(lambda i: ops.gn_apply(xs[i], outs[i], stats, gamma, beta, B, spatial, C, groups, cpg, 1e-5, True), nrot)
    t_st = timed(lambda i: ops.gn_stats(xs[i], stats, B, spatial, C, cpg), nrot)
    t_red = timed(lambda i: ops.gn_bwd_reduce(xs[i], dys[i], stats, gamma, beta, sums, B, spatial, C, groups, cpg, 1e-5, True), nrot)
    t_app = timed(lambda i: ops.gn_bwd_apply(xs[i], dys[i], None, outs[i], stats, gamma, beta, sums, B, spatial, C, groups, cpg, 1e-5, True), nrot)
    t_appc = timed(lambda i: ops.gn_bwd_apply(xs[i], dys[i], dys[(i + 1) % nrot], outs[i], stats, gamma, beta, sums, B, spatial, C, groups, cpg, 1e-5, True, colsum=col), nrot)
    t_col = timed(lambda i: ops.colsum(dys[i], sums, B, spatial, C), nrot)
    gb = n * 1e-9
    rows.append((spatial, C, t_fwd, t_st, t_red, t_app, t_appc, t_col))
    print("B=%d S=%6d C=%4d | apply %7.1f us %5.0f GB/s (%4.1f%%) | stats %7.1f us %5.0f GB/s | bwd_reduce %7.1f us %5.0f GB/s (%4.1f%%) | "
          "bwd_apply %7.1f us %5.0f GB/s (%4.1f%%) | bwd_apply+add+colsum %7.1f us %5.0f GB/s | colsum %7.1f us %5.0f GB/s" % (
              B, spatial, C, t_fwd * 1e6, 4 * gb / t_fwd, 400 * gb / t_fwd / PEAK, t_st * 1e6, 2 * gb / t_st,
              t_red * 1e6, 4 * gb / t_red, 400 * gb / t_red / PEAK, t_app * 1e6, 6 * gb / t_app, 600 * gb / t_app / PEAK,
              t_appc * 1e6, 8 * gb / t_appc, t_col * 1e6, 2 * gb / t_col), flush=True)
    del xs, dys, outs
