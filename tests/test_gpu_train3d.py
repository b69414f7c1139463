"""Training step of the 3D latent UNet on the B200 kernels vs torch autograd on the CPU oracle:
loss and every parameter gradient (p_losses with min-SNR weighting, ddpm_3d_ldm/train.py:391-398).

Tolerance: bf16 operands/activations/gradient tensors with fp32 accumulation; per-parameter
gradient rel-L2 <= 5e-2 against the fp32 oracle (or 1.5x the error torch autocast-bf16 makes on
this GPU, whichever is larger), loss within 1e-2 relative."""
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


def oracle_grads(sd, x0, t, noise, T, device="cpu", autocast=False):
    sd = {k: v.clone().to(device).requires_grad_() for k, v in sd.items()}
    buf = {k: v.to(device) for k, v in O.schedule_buffers(O.cosine_betas(T)).items()}
    x0, t, noise = x0.to(device), t.to(device), noise.to(device)
    xn = O.q_sample(buf, x0, t, noise)
    ctx = torch.autocast("cuda", dtype=torch.bfloat16) if autocast else contextlib.nullcontext()
    with ctx:
        pred = O.unet3d_forward(sd, xn, t)
    loss = O.minsnr_loss(buf, pred.float(), noise, t, 5.0) if device == "cpu" else \
        O.minsnr_loss({"snr": buf["snr"]}, pred.float(), noise, t, 5.0)
    loss.backward()
    return loss.item(), {k: v.grad.detach().cpu() for k, v in sd.items()}


@pytest.mark.parametrize("cls_name,kw,B,sp", [
    ("UNet3DModelWithAttention", dict(in_channels=3, base_channels=64, time_emb_dim=64), 2, (8, 8, 8)),
    ("UNet3DModel", dict(in_channels=3, base_channels=64, time_emb_dim=64), 1, (8, 12, 8)),
])
def test_train_step_gradients_vs_oracle(cls_name, kw, B, sp):
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm import unet, unet_attention
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.diffusion import GaussianDiffusionLatent3D
    cls = getattr(unet_attention, cls_name, None) or getattr(unet, cls_name)
    m = cls(**kw)
    sd = synthetic_state_dict(shapes_of(m), seed=21)
    m.load_state_dict(sd)
    m = m.cuda().train()
    T = 100
    diff = quiet(GaussianDiffusionLatent3D, m, 3, timesteps=T).cuda()
    g = torch.Generator().manual_seed(8)
    x0 = torch.randn(B, 3, *sp, generator=g)
    noise = torch.randn(B, 3, *sp, generator=g)
    t = torch.randint(1, T, (B,), generator=g)

    loss = diff.p_losses(x0.cuda(), t.cuda(), cond=None, noise=noise.cuda(), min_snr_gamma=5.0)
    loss.backward()
    ref_loss, ref = oracle_grads(sd, x0, t, noise, T)
    _, auto = oracle_grads(sd, x0, t, noise, T, device="cuda", autocast=True)
    assert abs(loss.item() - ref_loss) <= 1e-2 * abs(ref_loss), (loss.item(), ref_loss)
    worst, worst_name = 0.0, ""
    bad = []
    for name, p in m.named_parameters():
        assert p.grad is not None, name
        e = rel_l2(p.grad, ref[name])
        ea = rel_l2(auto[name], ref[name])
        if e > worst:
            worst, worst_name = e, name
        if e > max(5e-2, 1.5 * ea):
            bad.append((name, e, ea))
    print(f"{cls_name}: loss {loss.item():.6f} (oracle {ref_loss:.6f}); worst grad rel-L2 {worst:.3e} at {worst_name}")
    assert not bad, bad[:10]

    # a second step re-uses the cached training program (static buffers) and still matches
    for p in m.parameters():
        p.grad = None
    loss2 = diff.p_losses(x0.cuda(), t.cuda(), cond=None, noise=noise.cuda(), min_snr_gamma=5.0)
    loss2.backward()
    assert abs(loss2.item() - loss.item()) <= 1e-4 * abs(loss.item())
    assert rel_l2(m.in_conv.weight.grad, ref["in_conv.weight"]) < max(5e-2, 1.5 * rel_l2(auto["in_conv.weight"], ref["in_conv.weight"]))


def test_optimizer_step_changes_output_and_repacks_weights():
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.unet import UNet3DModel
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.diffusion import GaussianDiffusionLatent3D
    torch.manual_seed(0)
    m = UNet3DModel(3, base_channels=64, time_emb_dim=64).cuda().train()
    diff = quiet(GaussianDiffusionLatent3D, m, 3, timesteps=50).cuda()
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)
    x0 = torch.randn(2, 3, 8, 8, 8, device="cuda")
    t = torch.tensor([3, 40], device="cuda")
    noise = torch.randn_like(x0)
    with torch.no_grad():
        before = diff.p_losses(x0, t, noise=noise).item()  # inference program, packed once
    losses = []
    for _ in range(6):
        opt.zero_grad()
        loss = diff.p_losses(x0, t, noise=noise)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    print("losses", before, losses)
    assert abs(losses[0] - before) <= 1e-3 * abs(before)
    assert losses[-1] < losses[0]  # same batch, small Adam steps: the loss must go down
    # both programs must have re-packed the updated weights: compare with the oracle on them
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    buf = O.schedule_buffers(O.cosine_betas(50))
    xn = O.q_sample(buf, x0.cpu(), t.cpu(), noise.cpu())
    with torch.no_grad():
        want = O.minsnr_loss(buf, O.unet3d_forward(sd, xn, t.cpu()), noise.cpu(), t.cpu(), 5.0).item()
        got_eval = diff.p_losses(x0, t, noise=noise).item()
    got_train = diff.p_losses(x0, t, noise=noise).item()
    assert abs(got_eval - want) <= 1e-2 * abs(want), (got_eval, want)
    assert abs(got_train - want) <= 1e-2 * abs(want), (got_train, want)


def test_graph_replayed_training_step_equals_eager(monkeypatch):
    """From the third step on the forward / backward launch lists run as replayed CUDA graphs
    (engine.UNetProgram._replay): gradients must equal those of the eager launch lists."""
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.unet_attention import UNet3DModelWithAttention
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.diffusion import GaussianDiffusionLatent3D

    def grads(no_graph: bool):
        monkeypatch.setenv("MRI_NO_GRAPH", "1" if no_graph else "0")
        torch.manual_seed(0)
        m = UNet3DModelWithAttention(3, base_channels=64, time_emb_dim=64).cuda().train()
        diff = quiet(GaussianDiffusionLatent3D, m, 3, timesteps=50).cuda()
        g = torch.Generator(device="cuda").manual_seed(5)
        out = None
        for it in range(4):
            x0 = torch.randn(2, 3, 8, 8, 8, device="cuda", generator=g)
            noise = torch.randn(2, 3, 8, 8, 8, device="cuda", generator=g)
            t = torch.tensor([3 + it, 40 - it], device="cuda")
            for p in m.parameters():
                p.grad = None
            diff.p_losses(x0, t, noise=noise).backward()
            out = {n: p.grad.detach().clone() for n, p in m.named_parameters()}
        prog = m.program(2, (8, 8, 8), training=True)
        return out, prog

    ge, pe = grads(True)
    gg, pg = grads(False)
    assert pe._rt_graphs.get("bwd", [0, None])[1] is None
    assert pg._rt_graphs["fwd"][1] is not None and pg._rt_graphs["bwd"][1] is not None
    # wgrad accumulates with fp32 red.global.add (arrival order varies run to run): equality up to
    # that summation noise, far below any addressing / ordering mistake a broken capture would show
    worst = max(rel_l2(gg[n], ge[n]) for n in ge)
    print(f"graph vs eager: worst gradient rel-L2 {worst:.2e}")
    assert worst < 1e-4, worst
