"""BASELINE.json configs[3] and [4] at their real per-sample size: the base-128 attention UNet
(136.4 M parameters) on 3x40x48x40 latents, T = 1000.

* cfg4 (sampling): a window of the class's own reverse loop (`sample_from`, the graph-replayed
  path `sample` uses) against a step-by-step run with the fp32 CPU oracle teacher-forced at every
  step -- noise prediction rel-L2 <= 2e-2, fused update bit-exact given (x, eps, z) -- at the high-noise end
  (t = 999.., where sqrt_recip_alphas = 31.6 amplifies eps errors) and at the low-noise end.
* cfg5 (training): q_sample + forward + min-SNR loss + backward at the full latent size; loss and
  every parameter gradient against torch autograd over the oracle (rel-L2 <= 5e-2 per tensor).
  The batch is 2 instead of 8 to keep the CPU oracle at ~20 s; per-sample shapes, tiles and
  launch lists are the ones the bench times.
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
def net():
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.unet_attention import UNet3DModelWithAttention
    m = UNet3DModelWithAttention(3, base_channels=128, channel_mults=(1, 2, 4), time_emb_dim=256)
    assert sum(p.numel() for p in m.parameters()) == 136_392_835
    sd = synthetic_state_dict(shapes_of(m), seed=51)
    m.load_state_dict(sd)
    return m.cuda(), sd


def test_cfg4_full_size_reverse_window_teacher_forced(net):
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.diffusion import GaussianDiffusionLatent3D
    m, sd = net
    m.eval()
    diff = quiet(GaussianDiffusionLatent3D, m, 3, timesteps=T).cuda()
    buf = O.schedule_buffers(O.cosine_betas(T))
    B = 2
    g = torch.Generator().manual_seed(4)
    # ---- high-noise end: the first steps of sample() (t = 999, 998, 997) ----------------------
    x_T = torch.randn(B, *LATENT, generator=g).cuda()
    img, worst = x_T, 0.0
    torch.manual_seed(9)
    for i in (999, 998, 997):
        t = torch.full((B,), i, device="cuda", dtype=torch.long)
        with torch.no_grad():
            eps = m(img, t)
            ref = O.unet3d_forward(sd, img.cpu(), t.cpu())
        worst = max(worst, rel_l2(eps, ref))
        z = torch.randn_like(img)
        nxt = diff._p_update(img, t, eps, z)
        assert torch.equal(nxt.cpu(), O.p_sample_update(buf, img.cpu(), t.cpu(), eps.cpu(), z.cpu())), i
        img = nxt
    assert torch.isfinite(img).all()
    # ---- low-noise end through the class's own loop: sample_from(x_3, 3) runs t = 3, 2, 1, 0 --
    x3 = 0.3 * torch.randn(B, *LATENT, generator=g).cuda()
    torch.manual_seed(10)
    a = diff.sample_from(x3, 3)
    torch.manual_seed(10)
    img = x3
    for i in (3, 2, 1, 0):
        t = torch.full((B,), i, device="cuda", dtype=torch.long)
        with torch.no_grad():
            eps = m(img, t)
            ref = O.unet3d_forward(sd, img.cpu(), t.cpu())
        worst = max(worst, rel_l2(eps, ref))
        z = torch.randn_like(img)   # drawn for every t including 0 (diffusion.py:118-125)
        img = diff._p_update(img, t, eps, z)
    print(f"cfg4 full size: teacher-forced worst eps rel-L2 {worst:.3e}; loop vs step-by-step {rel_l2(a, img):.3e}")
    assert worst < 2e-2, worst
    assert rel_l2(a, img) < 5e-3


def test_cfg5_full_size_training_step_vs_oracle(net):
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.diffusion import GaussianDiffusionLatent3D
    m, sd = net
    m.train()
    diff = quiet(GaussianDiffusionLatent3D, m, 3, timesteps=T).cuda()
    B = 2
    g = torch.Generator().manual_seed(5)
    x0 = torch.randn(B, *LATENT, generator=g)
    noise = torch.randn(B, *LATENT, generator=g)
    t = torch.randint(1, T, (B,), generator=g)
    for p in m.parameters():
        p.grad = None
    loss = diff.p_losses(x0.cuda(), t.cuda(), cond=None, noise=noise.cuda(), min_snr_gamma=5.0)
    loss.backward()
    buf = O.schedule_buffers(O.cosine_betas(T))
    sdg = {k: v.clone().requires_grad_() for k, v in sd.items()}
    pred = O.unet3d_forward(sdg, O.q_sample(buf, x0, t, noise), t)
    ref_loss = O.minsnr_loss(buf, pred, noise, t, 5.0)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-2 * abs(ref_loss.item()), (loss.item(), ref_loss.item())
    errs = sorted(((rel_l2(p.grad, sdg[n].grad), n) for n, p in m.named_parameters()), reverse=True)
    print(f"cfg5 full size: loss {loss.item():.6f} vs oracle {ref_loss.item():.6f}; worst grad rel-L2 "
          f"{errs[0][0]:.3e} at {errs[0][1]}, median {errs[len(errs) // 2][0]:.3e}")
    assert errs[0][0] < 5e-2, errs[:5]
