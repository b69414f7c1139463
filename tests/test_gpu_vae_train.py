"""VAE3D training on the B200 engine (stage 1 of ddpm_3d_ldm/train.py:258-300) vs torch autograd
through the fp32 CPU oracle (oracle.vae3d_encode / vae3d_decode on the same state_dict).

The loss is the script's: L1(recon, x) + 1e-4 * KL(mu, logvar) with z = mu + eps * exp(logvar / 2)
(train.py:270-274, vae.py:106-118); the reparameterisation noise is drawn once and fed to both
sides.  Tolerances: outputs rel-L2 <= 3e-2 (as tests/test_gpu_vae3d.py); every parameter gradient
and the latent gradient rel-L2 <= max(5e-2, 1.5 x the error the reference's own graph makes on the
same GPU under autocast(bf16), which is how the script runs it: train.py:269) against fp32 autograd
-- bf16 operands / fp32 accumulation through ~28 convolutions forward and backward and an L1 loss
whose gradient is a sign.  Measured values: profiles/r04b_vae_train_parity.txt."""
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_l2
from oracle import reference_oracle as O

pytestmark = pytest.mark.gpu
KL_W = 1e-4


def _loss(recon, x, mu, logvar):
    kl = -0.5 * torch.mean(1 + logvar - mu.pow(2) - logvar.exp())
    return F.l1_loss(recon, x) + KL_W * kl


def _oracle_grads(sd, x, eps, autocast=False):
    """fp32 on the CPU (the oracle), or the same graph on the GPU under autocast(bf16)."""
    if autocast:
        sd, x, eps = {k: v.cuda() for k, v in sd.items()}, x.cuda(), eps.cuda()
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    with torch.autocast(device_type="cuda", dtype=torch.bfloat16, enabled=autocast):
        mu, logvar = O.vae3d_encode(sd, x)
        z = mu.float() + eps * torch.exp(0.5 * logvar.float())
        z.retain_grad()
        recon = O.vae3d_decode(sd, z)
        loss = _loss(recon.float(), x, mu.float(), logvar.float())
    loss.backward()
    return loss.item(), recon.detach(), mu.detach(), logvar.detach(), z.grad, {k: v.grad for k, v in sd.items()}


@pytest.mark.parametrize("cfg", [
    dict(base=32, num_down=3, latent=16, shape=(2, 4, 16, 24, 16)),   # train.py:43-46 (32-channel level padded)
    dict(base=32, num_down=3, latent=3, shape=(1, 4, 24, 16, 32)),    # BASELINE cfg4's VAE3D(4, 32, 3, 3)
    dict(base=64, num_down=2, latent=8, shape=(2, 4, 16, 16, 16)),    # no padded level
])
def test_vae_training_step_matches_fp32_autograd(cfg):
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.vae import VAE3D
    torch.manual_seed(0)
    m = VAE3D(4, cfg["base"], cfg["num_down"], cfg["latent"])
    with torch.no_grad():   # lift the biases / norm parameters off their init so their gradients matter
        for k, p in m.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.cuda().train()
    x = torch.randn(*cfg["shape"]).clamp_(-1, 1)
    f = 2 ** (cfg["num_down"] - 1)
    eps = torch.randn(cfg["shape"][0], cfg["latent"], *[s // f for s in cfg["shape"][2:]])
    want_loss, want_rec, want_mu, want_lv, want_dz, want_g = _oracle_grads(sd, x, eps)
    _, _, _, _, ac_dz, ac_g = _oracle_grads(sd, x, eps, autocast=True)
    ac_err = {k: rel_l2(ac_g[k], want_g[k]) for k in want_g}
    ac_worst = max(ac_err.items(), key=lambda kv: kv[1])
    print(f"reference graph under autocast(bf16) on this GPU: dz {rel_l2(ac_dz, want_dz):.3e}  worst grad "
          f"{ac_worst[0]} {ac_worst[1]:.3e}")
    tol_g, tol_z = max(5e-2, 1.5 * ac_worst[1]), max(5e-2, 1.5 * rel_l2(ac_dz, want_dz))

    for rep in range(4):      # eager, eager, graph capture, graph replay
        m.zero_grad(set_to_none=True)
        xc = x.cuda()
        mu, logvar = m.encode(xc)
        z = mu + eps.cuda() * torch.exp(0.5 * logvar)
        z.retain_grad()
        recon = m.decode(z)
        loss = _loss(recon, xc, mu, logvar)
        loss.backward()
        assert rel_l2(mu, want_mu) < 3e-2 and rel_l2(logvar, want_lv) < 3e-2 and rel_l2(recon, want_rec) < 3e-2
        assert abs(loss.item() - want_loss) < 2e-2 * abs(want_loss)
        e_dz = rel_l2(z.grad, want_dz)
        worst = ("", 0.0)
        for k, p in m.named_parameters():
            assert p.grad is not None, k
            e = rel_l2(p.grad, want_g[k])
            if e > worst[1]:
                worst = (k, e)
        print(f"rep {rep}: loss {loss.item():.5f} vs {want_loss:.5f}  dz {e_dz:.3e}  worst grad {worst[0]} {worst[1]:.3e}")
        assert e_dz < tol_z, (e_dz, tol_z)
        assert worst[1] < tol_g, (worst, tol_g)


def test_vae_training_loop_like_the_script():
    """train.py:258-300 verbatim in miniature: autocast(bf16) + GradScaler + Adam(lr 1e-4) on vae(x),
    beside the reference's graph trained eagerly the same way from the same weights with the same
    reparameterisation noise: the two loss trajectories stay together.  Then the frozen-VAE use of
    stage 2 (requires_grad False -> inference programs) sees the trained weights."""
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.vae import VAE3D
    torch.manual_seed(1)
    vae = VAE3D(in_channels=4, base_channels=32, num_down=3, latent_channels=16).cuda()
    ref_sd = {k: v.detach().clone().requires_grad_(True) for k, v in vae.state_dict().items()}
    x = torch.randn(1, 4, 16, 16, 16, device="cuda").clamp_(-1, 1)
    steps = 16

    def run(params, forward):
        opt = torch.optim.Adam(params, lr=1e-4)
        scaler = torch.amp.GradScaler("cuda", enabled=True)
        torch.manual_seed(7)          # the same randn_like draws on both sides
        losses = []
        for _ in range(steps):
            opt.zero_grad()
            with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
                recon, mu, logvar = forward()
                loss = _loss(recon.float(), x, mu.float(), logvar.float())
            scaler.scale(loss).backward()
            scaler.step(opt)
            scaler.update()
            losses.append(loss.item())
        return losses

    def ref_forward():
        mu, logvar = O.vae3d_encode(ref_sd, x)
        std = torch.exp(0.5 * logvar)
        return O.vae3d_decode(ref_sd, mu + torch.randn_like(std) * std), mu, logvar

    vae.train()
    ours = run(list(vae.parameters()), lambda: vae(x))
    ref = run(list(ref_sd.values()), ref_forward)
    print("drop-in  ", [round(v, 4) for v in ours])
    print("reference", [round(v, 4) for v in ref])
    assert all(torch.isfinite(torch.tensor(ours))) and ours[-1] < ours[0]
    for a, b in zip(ours, ref):
        assert abs(a - b) < 3e-2 * abs(b), (ours, ref)
    for p in vae.parameters():
        p.requires_grad = False
    vae.eval()
    z = vae.encode_to_latent(x)
    sd = {k: v.detach().cpu().clone() for k, v in vae.state_dict().items()}
    want, _ = O.vae3d_encode(sd, x.cpu())
    assert rel_l2(z, want) < 3e-2     # the inference program sees the trained weights
