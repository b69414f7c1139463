"""BASELINE.json configs[0], end to end at its real size: the slice-conditioned 2D UNet in its
default architecture (base 64, mults 1-2-4-8, 35.4 M parameters), batch 4 of 1x128x128 slices:
one training step (q_sample + forward + min-SNR loss + backward + Adam) followed by 50-step reverse
sampling through GaussianDiffusion.sample(), each compared with the fp32 CPU oracle
(inputs, seeds and the comparison protocol).

Tolerances: loss 1e-2 relative; parameter gradients rel-L2 <= 5e-2 per tensor (bf16 operands,
fp32 accumulation vs an fp32 oracle); noise prediction rel-L2 <= 2e-2 at EVERY step of the
trajectory (teacher-forced); the fused reverse update bit-exact given (x, eps, z); timestep schedule
49..0 bit-exact."""
import contextlib
import io

import pytest
import torch

from helpers import rel_l2, shapes_of, synthetic_state_dict
from oracle import reference_oracle as O

pytestmark = pytest.mark.gpu


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def test_cfg1_train_step_then_50_step_sampling():
    from mri_image_generation_b200.model_scripts.slice_cond_2d_ddpm.diffusion import GaussianDiffusion
    from mri_image_generation_b200.model_scripts.slice_cond_2d_ddpm.unet import UNet
    from mri_image_generation_b200.optim import Adam
    B, HW = 4, 128
    m = quiet(UNet, img_channels=1)       # reference defaults (slice_cond_2d_ddpm/model.py:104-110)
    assert sum(p.numel() for p in m.parameters()) == 35_377_985
    sd = synthetic_state_dict(shapes_of(m), seed=41)
    m.load_state_dict(sd)
    m = m.cuda().train()
    g = torch.Generator().manual_seed(1234)
    x0 = (torch.randn(B, 1, HW, HW, generator=g) * 0.5).clamp(-1, 1)
    z_pos = torch.rand(B, generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    noise = torch.randn(B, 1, HW, HW, generator=g)

    # ---- one training step ------------------------------------------------------------------
    diff = quiet(GaussianDiffusion, m, HW, channels=1, timesteps=1000).cuda()
    opt = Adam(m.parameters(), lr=2e-4)
    opt.zero_grad(set_to_none=True)
    loss = diff.p_losses(x0.cuda(), t.cuda(), z_pos.cuda(), noise=noise.cuda())
    loss.backward()
    buf = O.schedule_buffers(O.linear_betas(1000))
    sdg = {k: v.clone().requires_grad_() for k, v in sd.items()}
    pred = O.unet2d_forward(sdg, O.q_sample(buf, x0, t, noise), t, z_pos)
    ref_loss = O.minsnr_loss({"snr": buf["snr"]}, pred, noise, t, 5.0)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-2 * abs(ref_loss.item()), (loss.item(), ref_loss.item())
    errs = sorted(((rel_l2(p.grad, sdg[n].grad), n) for n, p in m.named_parameters()), reverse=True)
    print(f"cfg1 train step: loss {loss.item():.6f} vs oracle {ref_loss.item():.6f}; worst grad rel-L2 "
          f"{errs[0][0]:.3e} at {errs[0][1]}, median {errs[len(errs) // 2][0]:.3e}")
    assert errs[0][0] < 5e-2, errs[:5]
    before = {n: p.detach().clone() for n, p in m.named_parameters()}
    opt.step()
    # Adam's first step moves every parameter by lr * g / (|g| + eps): +-2e-4 wherever g != 0
    moved = torch.cat([(p.detach() - before[n]).abs().flatten() for n, p in m.named_parameters()])
    assert 1.9e-4 < moved.median().item() < 2.1e-4 and moved.max().item() <= 2.002e-4  # + fp32 rounding of p - lr * x near |p| ~ 1
    sd1 = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}

    # ---- 50-step reverse sampling through the class's own loop ------------------------------
    m.eval()
    T = 50
    diff50 = quiet(GaussianDiffusion, m, HW, channels=1, timesteps=T).cuda()
    torch.manual_seed(77)
    img_a = diff50.sample(batch_size=B, z_pos=z_pos.cuda())
    assert img_a.shape == (B, 1, HW, HW) and torch.isfinite(img_a).all()
    # the same trajectory step by step (same seed -> same noise stream), teacher-forcing the
    # oracle (with the UPDATED weights) at every step
    torch.manual_seed(77)
    img = torch.randn(B, 1, HW, HW, device="cuda")
    buf50 = O.schedule_buffers(O.linear_betas(T))
    zc = z_pos.cuda()
    worst, steps = 0.0, []
    for i in reversed(range(T)):
        tt = torch.full((B,), i, device="cuda", dtype=torch.long)
        steps.append(i)
        with torch.no_grad():
            eps = m(img, tt, zc)
            ref_eps = O.unet2d_forward(sd1, img.cpu(), tt.cpu(), z_pos)
        worst = max(worst, rel_l2(eps, ref_eps))
        z = torch.randn_like(img)
        nxt = diff50._p_update(img, tt, eps, z)
        assert torch.equal(nxt.cpu(), O.p_sample_update(buf50, img.cpu(), tt.cpu(), eps.cpu(), z.cpu())), i
        img = nxt
    assert steps == list(range(T - 1, -1, -1))
    print(f"cfg1 sampling: teacher-forced worst eps rel-L2 over {T} steps {worst:.3e}; "
          f"loop vs step-by-step rel-L2 {rel_l2(img_a, img):.3e}")
    assert worst < 2e-2, worst
    assert rel_l2(img_a, img) < 5e-3
