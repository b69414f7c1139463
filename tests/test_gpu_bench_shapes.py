"""Parity AT THE BENCHMARKED SHAPES (BASELINE.json configs[1..4] at the batch sizes bench.py and
tools/bench_2d.py time): plans are batch-dependent (stream-K shares, boxes that pair samples,
swap_ab box pairs), so the small-batch parity tests do not cover them.

Samples are independent through the whole UNet (per-sample GroupNorm, per-sample attention), so
the fp32 CPU oracle is run per sample -- forward on a subset (first / middle / last) for the
inference cases, forward + backward on EVERY sample (gradients accumulated) for the training
cases -- which is exact, not an approximation.

  cfg4  3D LDM reverse step,  B = 16, 3x40x48x40, 136.4 M-parameter attention UNet
  cfg5  3D LDM training step, B = 8,  same model: loss and every parameter gradient
  cfg2  2D slice-conditioned UNet forward, B = 64, 1x240x240
  cfg3  2.5D UNet training step, B = 32, (4 + 16 context)x192x192: loss and every gradient
Tolerances as in DESIGN.md section 4: eps rel-L2 <= 2e-2 (bf16 operands, fp32 accumulate, against
the fp32 oracle), gradients rel-L2 <= 5e-2 per tensor, fused updates bit-exact.
Reference graphs: ddpm_3d_ldm/unet_attention.py:157-200, ddpm_25d_all_modalities/unet.py:174-218,
slice_cond_2d_ddpm/unet.py:169-199.
"""
import contextlib
import io

import pytest
import torch

from helpers import rel_l2, shapes_of, synthetic_state_dict
from oracle import reference_oracle as O

pytestmark = pytest.mark.gpu
LATENT = (3, 40, 48, 40)
T = 1000


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


@pytest.fixture(scope="module")
def net3d():
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.unet_attention import UNet3DModelWithAttention
    m = UNet3DModelWithAttention(3, base_channels=128, channel_mults=(1, 2, 4), time_emb_dim=256)
    sd = synthetic_state_dict(shapes_of(m), seed=61)
    m.load_state_dict(sd)
    return m.cuda(), sd


def test_cfg4_reverse_step_at_batch_16(net3d):
    """One graph-replayed reverse step at the benched batch (16 volumes): eps of samples 0 / 7 / 15
    against the oracle, the fused update bit-exact given (x, eps, z), and sample k of the batch
    equal to the same sample run in a batch of 2 within bf16 noise."""
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.diffusion import GaussianDiffusionLatent3D
    m, sd = net3d
    m.eval()
    diff = quiet(GaussianDiffusionLatent3D, m, 3, timesteps=T).cuda()
    buf = O.schedule_buffers(O.cosine_betas(T))
    B = 16
    g = torch.Generator().manual_seed(14)
    x = torch.randn(B, *LATENT, generator=g)
    t = torch.randint(0, T, (B,), generator=g)
    t[0], t[B - 1] = T - 1, 0
    with torch.no_grad():
        eps = m(x.cuda(), t.cuda()).clone()
        worst = 0.0
        for k in (0, 7, B - 1):
            ref = O.unet3d_forward(sd, x[k:k + 1], t[k:k + 1])
            worst = max(worst, rel_l2(eps[k:k + 1], ref))
        # the same samples in a batch of 2 (different plans: tiles, stream-K shares)
        eps2 = m(x[[7, B - 1]].cuda(), t[[7, B - 1]].cuda())
        cross = rel_l2(eps[[7, B - 1]], eps2)
        # p_sample (second call on the program = the replayed graph) == eps + fused update
        for _ in range(2):
            torch.manual_seed(21)
            nxt = diff.p_sample(x.cuda(), t.cuda())
        torch.manual_seed(21)
        z = torch.randn(B, *LATENT, device="cuda")
        want = O.p_sample_update(buf, x, t, eps.cpu(), z.cpu())
    print(f"cfg4 B=16: worst eps rel-L2 {worst:.3e}; batch 16 vs batch 2 {cross:.3e}")
    assert worst < 2e-2, worst
    assert cross < 1e-2, cross
    assert torch.equal(nxt.cpu(), want)


def test_cfg5_training_step_at_batch_8(net3d):
    """q_sample + forward + min-SNR loss + backward at the benched batch (8 latents): loss and every
    parameter gradient against torch autograd over the oracle, accumulated sample by sample."""
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.diffusion import GaussianDiffusionLatent3D
    m, sd = net3d
    m.train()
    diff = quiet(GaussianDiffusionLatent3D, m, 3, timesteps=T).cuda()
    B = 8
    g = torch.Generator().manual_seed(15)
    x0 = torch.randn(B, *LATENT, generator=g)
    noise = torch.randn(B, *LATENT, generator=g)
    t = torch.randint(1, T, (B,), generator=g)
    for p in m.parameters():
        p.grad = None
    loss = diff.p_losses(x0.cuda(), t.cuda(), cond=None, noise=noise.cuda(), min_snr_gamma=5.0)
    loss.backward()
    buf = O.schedule_buffers(O.cosine_betas(T))
    sdg = {k: v.clone().requires_grad_() for k, v in sd.items()}
    ref_loss = 0.0
    for k in range(B):  # loss = mean_b(w_b * mse_b): per-sample terms, gradients accumulate
        sl = slice(k, k + 1)
        pred = O.unet3d_forward(sdg, O.q_sample(buf, x0[sl], t[sl], noise[sl]), t[sl])
        lk = O.minsnr_loss(buf, pred, noise[sl], t[sl], 5.0) / B
        lk.backward()
        ref_loss += lk.item()
    assert abs(loss.item() - ref_loss) <= 1e-2 * abs(ref_loss), (loss.item(), ref_loss)
    errs = sorted(((rel_l2(p.grad, sdg[n].grad), n) for n, p in m.named_parameters()), reverse=True)
    print(f"cfg5 B=8: loss {loss.item():.6f} vs oracle {ref_loss:.6f}; worst grad rel-L2 "
          f"{errs[0][0]:.3e} at {errs[0][1]}, median {errs[len(errs) // 2][0]:.3e}")
    assert errs[0][0] < 5e-2, errs[:5]


def test_cfg2_forward_at_batch_64_240():
    """2D slice-conditioned UNet (default 35.4 M parameters) at B = 64, 1x240x240."""
    from mri_image_generation_b200.model_scripts.slice_cond_2d_ddpm.unet import UNet
    m = quiet(UNet)
    sd = synthetic_state_dict(shapes_of(m), seed=62)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    B = 64
    g = torch.Generator().manual_seed(16)
    x = torch.randn(B, 1, 240, 240, generator=g)
    t = torch.randint(0, T, (B,), generator=g)
    z = torch.rand(B, generator=g)
    with torch.no_grad():
        eps = m(x.cuda(), t.cuda(), z.cuda()).clone()
        worst = 0.0
        for k in (0, 31, B - 1):
            ref = O.unet2d_forward(sd, x[k:k + 1], t[k:k + 1], z[k:k + 1])
            worst = max(worst, rel_l2(eps[k:k + 1], ref))
        eps4 = m(x[[0, 31, 62, 63]].cuda(), t[[0, 31, 62, 63]].cuda(), z[[0, 31, 62, 63]].cuda())
        cross = rel_l2(eps[[0, 31, 62, 63]], eps4)
    print(f"cfg2 B=64 240^2: worst eps rel-L2 {worst:.3e}; batch 64 vs batch 4 {cross:.3e}")
    assert torch.isfinite(eps).all()
    assert worst < 2e-2, worst
    assert cross < 1e-2, cross


def test_cfg3_training_step_at_batch_32_192_with_context():
    """2.5D UNet (4 modalities + 16 context channels) at B = 32, 192x192: MSE loss and every
    parameter gradient against autograd over the oracle, accumulated sample by sample."""
    from mri_image_generation_b200.model_scripts.ddpm_25d_all_modalities.diffusion import GaussianDiffusion
    from mri_image_generation_b200.model_scripts.ddpm_25d_all_modalities.unet import UNet
    m = quiet(UNet, in_channels=20, out_channels=4)
    sd = synthetic_state_dict(shapes_of(m), seed=63)
    m.load_state_dict(sd)
    m = m.cuda().train()
    diff = quiet(GaussianDiffusion, m, 192, channels=4, timesteps=T).cuda()
    B = 32
    g = torch.Generator().manual_seed(17)
    x0 = torch.randn(B, 4, 192, 192, generator=g)
    ctx = torch.randn(B, 16, 192, 192, generator=g)
    noise = torch.randn(B, 4, 192, 192, generator=g)
    t = torch.randint(0, T, (B,), generator=g)
    z = torch.rand(B, generator=g)
    loss = diff.p_losses(x0.cuda(), t.cuda(), z.cuda(), context=ctx.cuda(), noise=noise.cuda())
    loss.backward()
    buf = O.schedule_buffers(O.linear_betas(T), with_snr=False)
    sdg = {k: v.clone().requires_grad_() for k, v in sd.items()}
    ref_loss = 0.0
    for k in range(B):  # F.mse_loss over the batch = mean of the per-sample means
        sl = slice(k, k + 1)
        pred = O.unet2d_forward(sdg, O.q_sample(buf, x0[sl], t[sl], noise[sl]), t[sl], z[sl],
                                context=ctx[sl])
        lk = O.mse_loss(pred, noise[sl]) / B
        lk.backward()
        ref_loss += lk.item()
    assert abs(loss.item() - ref_loss) <= 1e-2 * abs(ref_loss), (loss.item(), ref_loss)
    errs = sorted(((rel_l2(p.grad, sdg[n].grad), n) for n, p in m.named_parameters()), reverse=True)
    print(f"cfg3 B=32 192^2: loss {loss.item():.6f} vs oracle {ref_loss:.6f}; worst grad rel-L2 "
          f"{errs[0][0]:.3e} at {errs[0][1]}, median {errs[len(errs) // 2][0]:.3e}")
    assert errs[0][0] < 5e-2, errs[:5]
