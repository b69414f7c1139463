#!/usr/bin/env python
"""Secondary timings for the 2D / 2.5D configs of BASELINE.json (GPU only):
  cfg2  slice_cond_2d_ddpm   UNet(1, base 64, mults 1-2-4-8) B=64 1x240x240: reverse-sampling step
                             (graph-replayed) and one training step (q_sample+fwd+loss+bwd+Adam)
  cfg3  ddpm_25d_all_mod...  UNet(20 -> 4) B=32 (4+16)x192x192: one training step
Algorithmic FLOPs from SURVEY.md 8(d): cfg2 fwd 20 030 G, fwd+bwd 60 088 G; cfg3 fwd+bwd 19 292 G.
"""
import contextlib
import io
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def timeit(fn, warm=3, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    from mri_image_generation_b200.model_scripts.slice_cond_2d_ddpm.unet import UNet as U2
    from mri_image_generation_b200.model_scripts.slice_cond_2d_ddpm.diffusion import GaussianDiffusion as G2
    from mri_image_generation_b200.model_scripts.ddpm_25d_all_modalities.unet import UNet as U25
    from mri_image_generation_b200.model_scripts.ddpm_25d_all_modalities.diffusion import GaussianDiffusion as G25
    dev = "cuda"
    out = {}
    B2 = int(os.environ.get("B2", "64"))
    torch.manual_seed(0)
    m = quiet(U2, img_channels=1, base_channels=64, channel_mults=(1, 2, 4, 8), time_emb_dim=256).to(dev)
    d = quiet(G2, m, 240, channels=1, timesteps=1000).to(dev)
    # sampling: 20 graph-replayed reverse steps through the public API
    m.eval()
    with torch.no_grad():
        z = torch.rand(B2, device=dev)
        x = torch.randn(B2, 1, 240, 240, device=dev)
        ms = timeit(lambda: d.p_sample(x, torch.full((B2,), 500, device=dev, dtype=torch.long), z), 3, 5)
    out["cfg2_p_sample_ms"] = ms
    out["cfg2_p_sample_tflops"] = 20030e9 * (B2 / 64) / (ms * 1e-3) / 1e12
    # training step
    m.train()
    opt = torch.optim.Adam(m.parameters(), lr=2e-4)
    x0 = torch.randn(B2, 1, 240, 240, device=dev).clamp(-1, 1)

    def step2():
        t = torch.randint(0, 1000, (B2,), device=dev)
        opt.zero_grad(set_to_none=True)
        loss = d.p_losses(x0, t, z)
        loss.backward()
        opt.step()

    ms = timeit(step2, 3, 3)
    out["cfg2_train_step_ms"] = ms
    out["cfg2_train_tflops"] = 60088e9 * (B2 / 64) / (ms * 1e-3) / 1e12
    del m, d, opt
    torch.cuda.empty_cache()
    B3 = int(os.environ.get("B3", "32"))
    m = quiet(U25, in_channels=20, out_channels=4, base_channels=64, channel_mults=(1, 2, 4, 8),
              time_emb_dim=256).to(dev).train()
    d = quiet(G25, m, 192, channels=4, timesteps=1000).to(dev)
    opt = torch.optim.Adam(m.parameters(), lr=2e-4)
    xc = torch.randn(B3, 4, 192, 192, device=dev)
    ctx = torch.randn(B3, 16, 192, 192, device=dev)
    z = torch.rand(B3, device=dev)

    def step3():
        t = torch.randint(0, 1000, (B3,), device=dev)
        opt.zero_grad(set_to_none=True)
        loss = d.p_losses(xc, t, z, context=ctx)
        loss.backward()
        opt.step()

    ms = timeit(step3, 3, 3)
    out["cfg3_train_step_ms"] = ms
    out["cfg3_train_tflops"] = 19292e9 * (B3 / 32) / (ms * 1e-3) / 1e12
    print(json.dumps(out))


if __name__ == "__main__":
    main()
