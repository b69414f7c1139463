"""Split precision mode (model.precision = "split", split_engine.py): parity with the reference's
fp32 / TF32 sampling path (ddpm_3d_ldm/show_model.py:254 runs the UNet without autocast).

Tolerance: noise prediction rel-L2 <= 3e-4 against the fp32 CPU oracle -- the reference's own graph
with TF32 convolutions on the same GPU is at 1.1e-3 (profiles/r04g_reference_precision_modes.json),
the default bf16 path at 1.2e-2.  Measured values: profiles/r04i_split_precision.txt."""
import contextlib
import io

import pytest
import torch
import torch.nn.functional as F

from helpers import rel_l2, shapes_of, synthetic_state_dict
from oracle import reference_oracle as O

pytestmark = pytest.mark.gpu
TOL = 3e-4


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def test_split_kernels():
    from mri_image_generation_b200 import ops
    g = torch.Generator().manual_seed(0)
    B, S, C, G = 2, 200, 128, 8
    x = torch.randn(B, S, C, generator=g) * 3 + 0.5
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    rb = torch.randn(B, C, generator=g)
    xd = x.double().view(B, S, G, C // G)
    stats = torch.stack([xd.sum((1, 3)), (xd * xd).sum((1, 3))], -1).contiguous()      # [B, G, 2] fp64
    y = torch.zeros(B, S, 3 * C, dtype=torch.bfloat16, device="cuda")
    ops.gn_split(x.cuda(), y, stats.cuda(), gamma.cuda(), beta.cuda(), B, S, C, G, C // G, 1e-5, True,
                 rowbias=rb.cuda(), rowbias_ld=C)
    want = F.silu(F.group_norm(x.permute(0, 2, 1), G, gamma, beta, 1e-5)).permute(0, 2, 1) + rb[:, None, :]
    y = y.float().cpu()
    assert torch.equal(y[..., :C], y[..., 2 * C:])
    assert rel_l2(y[..., :C] + y[..., C:2 * C], want) < 2e-5
    assert rel_l2(y[..., :C], want) > 1e-3          # hi alone is bf16
    y2 = torch.zeros(B, S, 3 * C, dtype=torch.bfloat16, device="cuda")
    ops.gn_split(x.cuda(), y2, None, None, None, B, S, C, 1, 1, 0.0, False)
    y2 = y2.float().cpu()
    assert rel_l2(y2[..., :C] + y2[..., C:2 * C], x) < 1e-5
    # split3, both patterns, strided source with an inner dimension and a segment longer than the width
    src = torch.randn(6, 3 * 10, generator=g)
    dst = torch.zeros(6, 3, 3 * 16, dtype=torch.bfloat16, device="cuda")
    ops.split3(src.cuda(), dst, 6, 3, 10, 30, 10, 3 * 48, 48, 16, 1)
    d = dst.float().cpu()
    v = src.view(6, 3, 10)
    hi = v.to(torch.bfloat16).float()
    lo = (v - hi).to(torch.bfloat16).float()
    assert torch.equal(d[..., 0:10], hi) and torch.equal(d[..., 16:26], hi) and torch.equal(d[..., 32:42], lo)
    assert not d[..., 10:16].any() and not d[..., 42:].any()
    # softmax
    Sm = torch.randn(20, 72, generator=g) * 4
    P3 = torch.zeros(20, 3 * 128, dtype=torch.bfloat16, device="cuda")
    ops.softmax_rows_split(Sm.cuda(), P3, 20, 70, 72, 3 * 128, 128, 0.3)
    p = P3.float().cpu()
    want = torch.softmax(Sm[:, :70] * 0.3, -1)
    assert rel_l2(p[:, :70] + p[:, 128:198], want) < 1e-5 and torch.equal(p[:, :70], p[:, 256:326])
    assert not p[:, 70:128].any()


@pytest.mark.parametrize("attn", [True, False])
def test_small_unet_split_vs_fp32_oracle(attn):
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.unet import UNet3DModel
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.unet_attention import UNet3DModelWithAttention
    cls = UNet3DModelWithAttention if attn else UNet3DModel
    m = quiet(cls, 3, base_channels=64, time_emb_dim=64)
    sd = synthetic_state_dict(shapes_of(m), seed=61)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 3, 8, 12, 8, generator=g)
    t = torch.tensor([49, 3])
    with torch.no_grad():
        want = O.unet3d_forward(sd, x, t)     # the plain model only lacks the mid_attn keys
        bf16 = m(x.cuda(), t.cuda()).clone()
        m.precision = "split"
        got = m(x.cuda(), t.cuda()).clone()
        again = m(x.cuda(), t.cuda())
    e_split, e_bf16 = rel_l2(got, want), rel_l2(bf16, want)
    print(f"attn={attn}: split {e_split:.3e}  bf16 {e_bf16:.3e}")
    assert torch.equal(got, again)
    assert e_split < TOL and e_bf16 > 10 * e_split


def test_cfg4_model_split_vs_fp32_oracle():
    """BASELINE cfg4's model at the full 3 x 40 x 48 x 40 latent, B = 1."""
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.unet_attention import UNet3DModelWithAttention
    m = quiet(UNet3DModelWithAttention, 3, base_channels=128, channel_mults=(1, 2, 4), time_emb_dim=256)
    sd = synthetic_state_dict(shapes_of(m), seed=51)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    m.precision = "split"
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 3, 40, 48, 40, generator=g)
    for t_val in (999, 3):
        t = torch.full((1,), t_val, dtype=torch.long)
        with torch.no_grad():
            want = O.unet3d_forward(sd, x, t)
            got = m(x.cuda(), t.cuda())
        e = rel_l2(got, want)
        print(f"cfg4 t={t_val}: split {e:.3e}")
        assert e < TOL


def test_sampling_in_split_precision_follows_the_fp32_reference_trajectory():
    """sample() in split mode (graph-replayed) against the oracle's loop fed with the same x_T and
    the same per-step noise (torch's generator: the in-kernel draws are torch.randn's, bit for bit)."""
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.diffusion import GaussianDiffusionLatent3D
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.unet_attention import UNet3DModelWithAttention
    m = quiet(UNet3DModelWithAttention, 3, base_channels=64, time_emb_dim=64)
    sd = synthetic_state_dict(shapes_of(m), seed=62)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    T, shape = 12, (2, 3, 8, 8, 8)
    diff = quiet(GaussianDiffusionLatent3D, m, 3, timesteps=T).cuda()
    buf = O.schedule_buffers(O.cosine_betas(T))
    torch.manual_seed(11)
    x_T = torch.randn(shape, device="cuda")
    noises = [torch.randn(shape, device="cuda").cpu() for _ in range(T)]
    want = O.sample_loop(buf, lambda x, t: O.unet3d_forward(sd, x, t), x_T.cpu(), noises, T)
    errs = {}
    for prec in ("bf16", "split"):
        m.precision = prec
        torch.manual_seed(11)
        got = diff.sample(2, (8, 8, 8))
        errs[prec] = rel_l2(got, want)
    print("trajectory error after", T, "steps:", errs)
    assert errs["split"] < 1e-3 and errs["bf16"] > 5 * errs["split"]
    m.precision = "split"
    with pytest.raises(Exception):
        m.train()
        m(torch.randn(shape, device="cuda"), torch.zeros(2, dtype=torch.long, device="cuda"))


@pytest.mark.parametrize("kind", ["2d", "25d"])
def test_unet2d_split_vs_fp32_oracle(kind):
    """slice_cond_2d_ddpm / ddpm_25d_all_modalities UNets (post-norm blocks, slice conditioning, context)."""
    if kind == "2d":
        from mri_image_generation_b200.model_scripts.slice_cond_2d_ddpm.unet import UNet
        m = quiet(UNet, img_channels=1, base_channels=64, time_emb_dim=64)
        cx, cc = 1, 0
    else:
        from mri_image_generation_b200.model_scripts.ddpm_25d_all_modalities.unet import UNet
        m = quiet(UNet, in_channels=20, out_channels=4, base_channels=64, time_emb_dim=64)
        cx, cc = 4, 16
    sd = synthetic_state_dict(shapes_of(m), seed=63)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(3, cx, 32, 48, generator=g)
    ctx = torch.randn(3, cc, 32, 48, generator=g) if cc else None
    t = torch.randint(0, 1000, (3,), generator=g)
    z = torch.rand(3, generator=g)
    kw = {} if ctx is None else {"context": ctx}
    kwc = {} if ctx is None else {"context": ctx.cuda()}
    with torch.no_grad():
        want = O.unet2d_forward(sd, x, t, z, **kw)
        bf16 = m(x.cuda(), t.cuda(), z.cuda(), **kwc).clone()
        m.precision = "split"
        got = m(x.cuda(), t.cuda(), z.cuda(), **kwc).clone()
    e_split, e_bf16 = rel_l2(got, want), rel_l2(bf16, want)
    print(f"{kind}: split {e_split:.3e}  bf16 {e_bf16:.3e}")
    assert e_split < TOL and e_bf16 > 10 * e_split


def test_2d_sampling_in_split_precision():
    from mri_image_generation_b200.model_scripts.slice_cond_2d_ddpm.diffusion import GaussianDiffusion
    from mri_image_generation_b200.model_scripts.slice_cond_2d_ddpm.unet import UNet
    m = quiet(UNet, img_channels=1, base_channels=64, time_emb_dim=64)
    sd = synthetic_state_dict(shapes_of(m), seed=64)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    m.precision = "split"
    T, shape = 10, (2, 1, 32, 32)
    diff = quiet(GaussianDiffusion, m, 32, channels=1, timesteps=T).cuda()
    buf = O.schedule_buffers(O.linear_betas(T))
    z = torch.tensor([0.25, 0.75])
    torch.manual_seed(12)
    x_T = torch.randn(shape, device="cuda")
    noises = [torch.randn(shape, device="cuda").cpu() for _ in range(T)]
    want = O.sample_loop(buf, lambda x, t: O.unet2d_forward(sd, x, t, z), x_T.cpu(), noises, T)
    torch.manual_seed(12)
    got = diff.sample(2, z.cuda())
    e = rel_l2(got, want)
    print("2D trajectory error after", T, "steps:", e)
    assert e < 1e-3


@pytest.mark.parametrize("cfg", [dict(base=32, num_down=3, latent=3, shape=(2, 4, 16, 24, 16)),
                                 dict(base=32, num_down=3, latent=16, shape=(1, 4, 16, 16, 16)),
                                 dict(base=64, num_down=2, latent=8, shape=(1, 4, 8, 16, 16))])
def test_vae_split_vs_fp32_oracle(cfg):
    """VAE3D encode / decode (show_model.py:255 decodes without autocast), padded 32-channel level included."""
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.vae import VAE3D
    torch.manual_seed(3)
    m = VAE3D(4, cfg["base"], cfg["num_down"], cfg["latent"])
    with torch.no_grad():
        for p_ in m.parameters():
            if p_.dim() == 1:
                p_.add_(0.1 * torch.randn_like(p_))
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.cuda().eval()
    x = torch.randn(*cfg["shape"]).clamp_(-1, 1)
    with torch.no_grad():
        want_mu, want_lv = O.vae3d_encode(sd, x)
        want_rec = O.vae3d_decode(sd, want_mu)
        mu_b, _ = m.encode(x.cuda())
        rec_b = m.decode_from_latent(want_mu.cuda())
        m.precision = "split"
        mu, lv = m.encode(x.cuda())
        rec = m.decode_from_latent(want_mu.cuda())
    errs = dict(mu=rel_l2(mu, want_mu), logvar=rel_l2(lv, want_lv), recon=rel_l2(rec, want_rec),
                mu_bf16=rel_l2(mu_b, want_mu), recon_bf16=rel_l2(rec_b, want_rec))
    print({k: f"{v:.2e}" for k, v in errs.items()})
    assert errs["mu"] < TOL and errs["logvar"] < TOL and errs["recon"] < TOL
    assert errs["recon_bf16"] > 10 * errs["recon"]
