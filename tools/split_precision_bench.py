#!/usr/bin/env python
"""Cost and accuracy of the split precision mode (model.precision = "split") on BASELINE cfg4:
one graph-replayed reverse step at B volumes per GPU in both modes, the noise-prediction error of
both against the fp32 CPU oracle (B = 1), and the reference graph run eagerly with TF32 / fp32
convolutions on the same GPU (what ddpm_3d_ldm/show_model.py:254 runs)."""
import contextlib
import io
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import rel_l2, shapes_of, synthetic_state_dict  # noqa: E402
from oracle import reference_oracle as O  # noqa: E402
from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.diffusion import GaussianDiffusionLatent3D  # noqa: E402
from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.unet_attention import UNet3DModelWithAttention  # noqa: E402

LATENT = (40, 48, 40)


def timed(fn, warm=3, iters=10):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    batches = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1,4,16").split(",")]
    with contextlib.redirect_stdout(io.StringIO()):
        m = UNet3DModelWithAttention(3, base_channels=128, channel_mults=(1, 2, 4), time_emb_dim=256)
        sd = synthetic_state_dict(shapes_of(m), seed=51)
        m.load_state_dict(sd)
        m = m.cuda().eval()
        diff = GaussianDiffusionLatent3D(m, 3, timesteps=1000).cuda()
    out = {"latent": list(LATENT), "steps_ms": {}, "eps_rel_l2_vs_fp32_oracle": {}}
    g = torch.Generator().manual_seed(0)
    x1 = torch.randn(1, 3, *LATENT, generator=g)
    t1 = torch.full((1,), 500, dtype=torch.long)
    with torch.no_grad():
        want = O.unet3d_forward(sd, x1, t1)
        for prec in ("bf16", "split"):
            m.precision = prec
            out["eps_rel_l2_vs_fp32_oracle"][prec] = rel_l2(m(x1.cuda(), t1.cuda()), want)
        sdc = {k: v.cuda() for k, v in sd.items()}
        torch.backends.cudnn.benchmark = True
        for name, tf32 in (("reference_eager_tf32", True), ("reference_eager_fp32", False)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            out["eps_rel_l2_vs_fp32_oracle"][name] = rel_l2(O.unet3d_forward(sdc, x1.cuda(), t1.cuda()), want)
        for B in batches:
            x = torch.randn(B, 3, *LATENT, device="cuda")
            t = torch.full((B,), 500, device="cuda", dtype=torch.long)
            row = {}
            for prec in ("bf16", "split"):
                m.precision = prec
                row[prec] = timed(lambda: diff.p_sample(x, t))
            for name, tf32 in (("reference_eager_tf32", True), ("reference_eager_fp32", False)):
                torch.backends.cudnn.allow_tf32 = tf32
                torch.backends.cuda.matmul.allow_tf32 = tf32
                row[name] = timed(lambda: O.unet3d_forward(sdc, x, t), warm=2, iters=3)
            out["steps_ms"][f"B={B}"] = row
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
