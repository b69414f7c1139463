"""GPU comparator ("the kernel to beat"): the reference's own PyTorch graph --
the oracle restatement, i.e. the same ATen / cuDNN / cuBLAS calls the reference modules make --
run EAGERLY on the same B200, at the bench shapes, beside the B200-native path.  Not collected by
pytest; run on a GPU box:

    python tests/gpu_comparator.py > gpurun_out/gpu_comparator.json

cfg4: one reverse step (UNet forward + the p_sample arithmetic) at 16 volumes, fp32 (TF32 as
`train.py:72` sets it) and under autocast(bf16) (`train.py:395`), `cudnn.benchmark = True`
(`train.py:71`).  cfg5: one training step at 8 latents (q_sample + autocast(bf16) forward +
min-SNR loss + backward + torch.optim.Adam).  CUDA events, 3 warm-up + 5 timed iterations.
"""
import contextlib
import io
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import shapes_of, synthetic_state_dict  # noqa: E402
from oracle import reference_oracle as O  # noqa: E402

LATENT = (3, 40, 48, 40)
T = 1000


def timed(fn, warm=3, iters=5):
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
    dev = torch.device("cuda")
    torch.backends.cudnn.benchmark = True
    torch.set_float32_matmul_precision("high")
    torch.backends.cudnn.allow_tf32 = True
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.unet_attention import UNet3DModelWithAttention
    with contextlib.redirect_stdout(io.StringIO()):
        shapes = shapes_of(UNet3DModelWithAttention(3, base_channels=128, channel_mults=(1, 2, 4), time_emb_dim=256))
    sd = {k: v.to(dev) for k, v in synthetic_state_dict(shapes, seed=51).items()}
    buf = {k: v.to(dev) for k, v in O.schedule_buffers(O.cosine_betas(T)).items()}
    out = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__,
           "cudnn": torch.backends.cudnn.version(), "flops_fwd_per_sample": 1276.4e9}

    # ---- cfg4: one reverse step at B = 16 ----------------------------------------------------
    B = 16
    x = torch.randn(B, *LATENT, device=dev)
    t = torch.full((B,), 500, device=dev, dtype=torch.long)

    def step(autocast):
        ctx = torch.autocast("cuda", dtype=torch.bfloat16) if autocast else contextlib.nullcontext()
        with torch.no_grad():
            with ctx:
                eps = O.unet3d_forward(sd, x, t)
            return O.p_sample_update(buf, x, t, eps.float(), torch.randn_like(x))

    for name, ac in (("tf32", False), ("autocast_bf16", True)):
        ms = timed(lambda: step(ac))
        out[f"cfg4_reverse_step_ms_{name}"] = ms
        out[f"cfg4_volumes_per_s_{name}"] = B / (T * ms * 1e-3)
        out[f"cfg4_tflops_{name}"] = B * 1276.4e9 / (ms * 1e-3) / 1e12

    # ---- cfg5: one training step at B = 8 -------------------------------------------------------
    B = 8
    params = {k: v.clone().requires_grad_() for k, v in sd.items()}
    opt = torch.optim.Adam(list(params.values()), lr=2e-4)
    z = torch.randn(B, *LATENT, device=dev)

    def train_step():
        tt = torch.randint(1, T, (B,), device=dev)
        noise = torch.randn_like(z)
        opt.zero_grad(set_to_none=True)
        xn = O.q_sample(buf, z, tt, noise)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            pred = O.unet3d_forward(params, xn, tt)
        loss = O.minsnr_loss(buf, pred.float(), noise, tt, 5.0)
        loss.backward()
        opt.step()

    ms = timed(train_step)
    out["cfg5_train_step_ms_autocast_bf16"] = ms
    out["cfg5_samples_per_s_autocast_bf16"] = B / (ms * 1e-3)
    out["cfg5_tflops_autocast_bf16"] = 3 * B * 1276.4e9 / (ms * 1e-3) / 1e12
    out["peak_mem_gb"] = torch.cuda.max_memory_allocated() / 2**30
    print(json.dumps(out))


if __name__ == "__main__":
    main()
