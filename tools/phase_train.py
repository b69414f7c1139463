#!/usr/bin/env python
"""CUDA-event timing of the phases of one cfg5 training step (GPU only): where the time outside
the forward / backward launch lists goes.   B=8 python tools/phase_train.py"""
import contextlib, io, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.diffusion import GaussianDiffusionLatent3D
from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.unet_attention import UNet3DModelWithAttention
from mri_image_generation_b200.optim import Adam
B = int(os.environ.get("B", "8"))
dev = torch.device("cuda")
torch.manual_seed(0)
model = UNet3DModelWithAttention(3, base_channels=128, channel_mults=(1, 2, 4), time_emb_dim=256).to(dev).train()
with contextlib.redirect_stdout(io.StringIO()):
    diff = GaussianDiffusionLatent3D(model, 3, timesteps=1000).to(dev)
opt = Adam(model.parameters(), lr=2e-4)
z = torch.randn(B, 3, 40, 48, 40, device=dev)
names = ["zero_grad", "p_losses(q_sample+refresh+fwd+loss)", "backward", "opt.step"]
acc = [0.0] * len(names)
cpu = [0.0] * len(names)
N = 12
for it in range(6 + N):
    t = torch.randint(1, 1000, (B,), device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    c = [time.perf_counter()]
    ev[0].record(); opt.zero_grad(set_to_none=True); c.append(time.perf_counter())
    ev[1].record(); loss = diff.p_losses(z, t, cond=None, min_snr_gamma=5.0); c.append(time.perf_counter())
    ev[2].record(); loss.backward(); c.append(time.perf_counter())
    ev[3].record(); opt.step(); c.append(time.perf_counter())
    ev[4].record()
    torch.cuda.synchronize()
    if it >= 6:
        for i in range(len(names)):
            acc[i] += ev[i].elapsed_time(ev[i + 1]) / N
            cpu[i] += (c[i + 1] - c[i]) * 1e3 / N
for n, a, cc in zip(names, acc, cpu):
    print("%-40s gpu %7.3f ms   host enqueue %7.3f ms" % (n, a, cc))
print("sum gpu %.3f ms" % sum(acc))
