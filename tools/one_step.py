#!/usr/bin/env python
"""ONE eager reverse step of BASELINE cfg4 (16 volumes) inside an NVTX range, for Nsight Compute:

  python tools/one_step.py                                  # plain run first (must exit 0)
  ncu --nvtx --nvtx-include "measured_step/" --metrics dram__bytes_read.sum,dram__bytes_write.sum,\\
      gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/step.csv python tools/one_step.py
  ncu --nvtx --nvtx-include "measured_step/" --set full --import-source on -k regex:attn_flash ... python tools/one_step.py

MODE=train runs one cfg5 training step (q_sample + forward + loss + backward + Adam, 8 latents)
inside the start/end range "measured_train_step" instead (--nvtx-include "measured_train_step")."""
import contextlib
import io
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("MRI_NO_GRAPH", "1")   # eager launches: every kernel is a separate ncu result
from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.diffusion import GaussianDiffusionLatent3D  # noqa: E402
from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.unet_attention import UNet3DModelWithAttention  # noqa: E402

B = int(os.environ.get("B", "16"))
mode = os.environ.get("MODE", "sample")
torch.manual_seed(0)
m = UNet3DModelWithAttention(3, base_channels=128, channel_mults=(1, 2, 4), time_emb_dim=256).cuda()
with contextlib.redirect_stdout(io.StringIO()):
    diff = GaussianDiffusionLatent3D(m, 3, timesteps=1000).cuda()
if mode == "sample":
    m.eval()
    x = torch.randn(B, 3, 40, 48, 40, device="cuda")
    prog = m.program(B, (40, 48, 40))
    with torch.no_grad():
        x = diff._reverse_loop(prog, x, 999, 2, "ddpm", use_graph=False)
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_push("measured_step")
        x = diff._reverse_loop(prog, x, 997, 1, "ddpm", use_graph=False)
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
    print("one sampling step done; finite:", bool(torch.isfinite(x).all()))
else:
    from mri_image_generation_b200.optim import Adam
    B = int(os.environ.get("B", "8"))
    m.train()
    opt = Adam(m.parameters(), lr=2e-4)
    z = torch.randn(B, 3, 40, 48, 40, device="cuda")

    def step():
        t = torch.randint(1, 1000, (B,), device="cuda")
        opt.zero_grad(set_to_none=True)
        loss = diff.p_losses(z, t)
        loss.backward()
        opt.step()
        return loss

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    # start / end range, not push / pop: the backward launches come from the autograd thread, and
    # push / pop ranges are per thread (ncu: --nvtx-include "measured_train_step", no trailing slash)
    rid = torch.cuda.nvtx.range_start("measured_train_step")
    loss = step()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_end(rid)
    print("one training step done; loss", float(loss))
