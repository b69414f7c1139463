#!/usr/bin/env python
"""Per-op CUDA-event timing of the cfg2 2D UNet forward (B=64, 1x240x240), GPU only."""
import collections, contextlib, io, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_image_generation_b200.model_scripts.slice_cond_2d_ddpm.unet import UNet
B = int(os.environ.get("B2", "64"))
with contextlib.redirect_stdout(io.StringIO()):
    m = UNet(img_channels=1, base_channels=64, channel_mults=(1, 2, 4, 8), time_emb_dim=256).cuda().eval()
prog = m.program(B, (240, 240), 1, 0)
def run():
    prog._arena[:max(prog._arena_used, 4)].zero_()
    evs = []
    for n, fn in zip(prog.op_names, prog.ops):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); evs.append((n, a, b))
    torch.cuda.synchronize()
    return [(n, a.elapsed_time(b)) for n, a, b in evs]
run(); r = run()
tot = sum(t for _, t in r)
agg = collections.defaultdict(float)
for n, t in r:
    agg["gemm" if n.startswith("gemm") else n.split(".")[-1]] += t
print("total %.2f ms" % tot, {k: round(v, 2) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])})
plans = {("gemm:" + p.name): p for p in prog.plans}
for n, t in r:
    if n.startswith("gemm"):
        p = plans[n]
        print("%-34s %7.3f ms %7.1f TF/s tiles %6d n_kb %3d swap %d sched %d n_total %d box %s" % (
            n, t, p.flops / t / 1e9, p.grid(), p.n_kb, p._args.swap_ab, p._args.sched, p.n_total, p.box))
