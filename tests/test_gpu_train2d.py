"""Training step of the 2D slice-conditioned UNet and the 2.5D UNet (with context) on the B200
kernels vs torch autograd on the CPU oracle: loss and every parameter gradient.
Tolerances as in test_gpu_train3d.py."""
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


def oracle_grads(sd, xn, t, z, noise, ctx, loss_fn, device="cpu", autocast=False):
    sd = {k: v.clone().to(device).requires_grad_() for k, v in sd.items()}
    args = [a.to(device) for a in (xn, t, z)]
    kw = {} if ctx is None else {"context": ctx.to(device)}
    c = torch.autocast("cuda", dtype=torch.bfloat16) if autocast else contextlib.nullcontext()
    with c:
        pred = O.unet2d_forward(sd, *args, **kw)
    loss = loss_fn(pred.float(), noise.to(device), t.to(device))
    loss.backward()
    return loss.item(), {k: v.grad.detach().cpu() for k, v in sd.items()}


def compare(m, ref, auto, tag):
    worst, wn, bad = 0.0, "", []
    for name, p in m.named_parameters():
        assert p.grad is not None, name
        e, ea = rel_l2(p.grad, ref[name]), rel_l2(auto[name], ref[name])
        if e > worst:
            worst, wn = e, name
        if e > max(5e-2, 1.5 * ea):
            bad.append((name, e, ea))
    print(f"{tag}: worst grad rel-L2 {worst:.3e} at {wn}")
    assert not bad, bad[:10]


def test_train_step_2d_minsnr():
    from mri_image_generation_b200.model_scripts.slice_cond_2d_ddpm.diffusion import GaussianDiffusion
    from mri_image_generation_b200.model_scripts.slice_cond_2d_ddpm.unet import UNet
    m = quiet(UNet, img_channels=1, base_channels=64, time_emb_dim=64)
    sd = synthetic_state_dict(shapes_of(m), seed=31)
    m.load_state_dict(sd)
    m = m.cuda().train()
    T = 100
    diff = quiet(GaussianDiffusion, m, 32, channels=1, timesteps=T).cuda()
    g = torch.Generator().manual_seed(9)
    x0 = torch.randn(2, 1, 32, 32, generator=g)
    noise = torch.randn(2, 1, 32, 32, generator=g)
    t = torch.randint(0, T, (2,), generator=g)
    z = torch.rand(2, generator=g)
    # the training script calls diffusion.p_losses(x, t, z_pos) (model.py:164): cond = z_pos
    loss = diff.p_losses(x0.cuda(), t.cuda(), z.cuda(), noise=noise.cuda())
    loss.backward()
    buf = O.schedule_buffers(O.linear_betas(T))
    xn = O.q_sample(buf, x0, t, noise)
    lf = lambda pred, nz, tt: O.minsnr_loss({"snr": buf["snr"].to(pred.device)}, pred, nz, tt, 5.0)
    ref_loss, ref = oracle_grads(sd, xn, t, z, noise, None, lf)
    _, auto = oracle_grads(sd, xn, t, z, noise, None, lf, device="cuda", autocast=True)
    assert abs(loss.item() - ref_loss) <= 1e-2 * abs(ref_loss), (loss.item(), ref_loss)
    compare(m, ref, auto, "2D")


def test_train_step_25d_mse_with_context():
    from mri_image_generation_b200.model_scripts.ddpm_25d_all_modalities.diffusion import GaussianDiffusion
    from mri_image_generation_b200.model_scripts.ddpm_25d_all_modalities.unet import UNet
    m = quiet(UNet, in_channels=20, out_channels=4, base_channels=64, time_emb_dim=64)
    sd = synthetic_state_dict(shapes_of(m), seed=32)
    m.load_state_dict(sd)
    m = m.cuda().train()
    T = 100
    diff = quiet(GaussianDiffusion, m, 32, channels=4, timesteps=T).cuda()
    g = torch.Generator().manual_seed(10)
    x0 = torch.randn(2, 4, 32, 32, generator=g)
    ctx = torch.randn(2, 16, 32, 32, generator=g)
    noise = torch.randn(2, 4, 32, 32, generator=g)
    t = torch.randint(0, T, (2,), generator=g)
    z = torch.rand(2, generator=g)
    loss = diff.p_losses(x0.cuda(), t.cuda(), z.cuda(), context=ctx.cuda(), noise=noise.cuda())
    loss.backward()
    buf = O.schedule_buffers(O.linear_betas(T), with_snr=False)
    xn = O.q_sample(buf, x0, t, noise)
    lf = lambda pred, nz, tt: O.mse_loss(pred, nz)
    ref_loss, ref = oracle_grads(sd, xn, t, z, noise, ctx, lf)
    _, auto = oracle_grads(sd, xn, t, z, noise, ctx, lf, device="cuda", autocast=True)
    assert abs(loss.item() - ref_loss) <= 1e-2 * abs(ref_loss), (loss.item(), ref_loss)
    compare(m, ref, auto, "2.5D")
