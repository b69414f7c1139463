#!/usr/bin/env python
"""How far is the reference's OWN GPU sampling path from fp32?  ddpm_3d_ldm/show_model.py:254 runs
the UNet without autocast, i.e. cuDNN convolutions with TF32 operands (torch's default
`cudnn.allow_tf32 = True`; train.py:72 also sets float32_matmul_precision("high")).  This prints
the rel-L2 error of the noise prediction of the reference graph (oracle restatement = the same
ATen / cuDNN calls) on this GPU under (a) TF32, (b) autocast(bf16), and of the drop-in (bf16
operands, fp32 accumulation), each against the fp32 CPU oracle -- cfg4 model, full latent, B = 1."""
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
from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.unet_attention import UNet3DModelWithAttention  # noqa: E402


def main():
    torch.backends.cudnn.benchmark = True
    with contextlib.redirect_stdout(io.StringIO()):
        m = UNet3DModelWithAttention(3, base_channels=128, channel_mults=(1, 2, 4), time_emb_dim=256)
    sd = synthetic_state_dict(shapes_of(m), seed=51)
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 3, 40, 48, 40, generator=g)
    out = {}
    for t_val in (999, 500, 3):
        t = torch.full((1,), t_val, dtype=torch.long)
        with torch.no_grad():
            want = O.unet3d_forward(sd, x, t)
            sdc = {k: v.cuda() for k, v in sd.items()}
            torch.backends.cudnn.allow_tf32 = True
            torch.backends.cuda.matmul.allow_tf32 = True
            tf32 = O.unet3d_forward(sdc, x.cuda(), t.cuda())
            with torch.autocast("cuda", dtype=torch.bfloat16):
                ac = O.unet3d_forward(sdc, x.cuda(), t.cuda())
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
            fp32 = O.unet3d_forward(sdc, x.cuda(), t.cuda())
            ours = m.cuda().eval()(x.cuda(), t.cuda())
        out[f"t={t_val}"] = {"reference_gpu_fp32_no_tf32": rel_l2(fp32, want), "reference_gpu_tf32": rel_l2(tf32, want),
                             "reference_gpu_autocast_bf16": rel_l2(ac.float(), want), "drop_in_bf16": rel_l2(ours, want)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
