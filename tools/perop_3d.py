#!/usr/bin/env python
"""Per-op CUDA-event timing of the cfg4 3D UNet forward (inference program), GPU only."""
import collections, contextlib, io, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_image_generation_b200 import ops
from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.unet_attention import UNet3DModelWithAttention
B = int(os.environ.get("B", "16"))
m = UNet3DModelWithAttention(3, base_channels=128, channel_mults=(1, 2, 4), time_emb_dim=256).cuda().eval()
prog = m.program(B, (40, 48, 40))
def run():
    ops.memset_zero(prog._arena, max(prog._arena_used, 4) * 8)
    evs = []
    for n, fn in zip(prog.op_names, prog.ops):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); evs.append((n, a, b))
    torch.cuda.synchronize()
    return [(n, a.elapsed_time(b)) for n, a, b in evs]
for _ in range(3): run()
r = run()
tot = sum(t for _, t in r)
agg = collections.defaultdict(float)
for n, t in r:
    key = "gemm" if n.startswith("gemm") else n.split(".")[-1]
    if n.startswith("gemm:mid_attn") or "softmax" in n or n.startswith("attn:"): key = "attention"
    if "norm" in n and not n.startswith("gemm"): key = "groupnorm"
    agg[key] += t
print("B=%d total %.2f ms" % (B, tot), {k: round(v, 2) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])})
plans = {("gemm:" + p.name): p for p in prog.plans}
lv = collections.defaultdict(lambda: [0.0, 0.0])
for n, t in r:
    if n.startswith("gemm") and n in plans:
        p = plans[n]
        k = "tiles %5d" % p.grid()
        lv[k][0] += t; lv[k][1] += p.flops
for k, (t, f) in sorted(lv.items()):
    print("%s: %.2f ms  %.0f TF/s" % (k, t, f / t / 1e9))
print("non-GEMM launches:")
for n, t in r:
    if not n.startswith("gemm"):
        print("  %-34s %7.3f ms" % (n, t))
