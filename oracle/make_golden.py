"""Generate tests/golden/*.pt from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py            # needs /root/reference (read-only, never copied)

The reference modules are imported from /root/reference and executed on CPU with seeded inputs
and synthetic seeded weights (see synthetic_state_dict); inputs and outputs are stored as small
fixtures.  The GPU box has no /root/reference: tests there read only the committed fixtures.
This script also asserts that oracle/reference_oracle.py reproduces every fixture bit-for-bit
(same torch build, same ops) before writing them.
"""
from __future__ import annotations

import contextlib
import hashlib
import io
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import reference_oracle as O  # noqa: E402


def synthetic_state_dict(shapes, seed: int):
    """Deterministic weights independent of nn init: N(0, 0.05) everywhere, norm weights ~ 1."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in shapes:
        w = torch.randn(*shp, generator=g) * 0.05
        if ("norm" in k) and k.endswith(".weight"):
            w = w + 1.0
        sd[k] = w
    return sd


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().contiguous().numpy().tobytes()).hexdigest()[:16]


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def vae_cases():
    """tests/golden/vae3d.pt: encode / decode of the unmodified reference VAE3D (vae.py)."""
    from model_scripts.ddpm_3d_ldm.vae import VAE3D
    cases = {}
    for name, kw, vol, seed in [("vae_b32", dict(in_channels=4, base_channels=32, num_down=3, latent_channels=3), (16, 16, 16), 301),
                                ("vae_b64_l8", dict(in_channels=4, base_channels=64, num_down=2, latent_channels=8), (8, 12, 8), 302)]:
        m = VAE3D(**kw).eval()
        shapes = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
        sd = synthetic_state_dict(shapes, seed)
        m.load_state_dict(sd, strict=True)
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(2, kw["in_channels"], *vol, generator=g)
        with torch.no_grad():
            mu, logvar = m.encode(x)
            rec = m.decode(mu)
            lat = m.encode_to_latent(x)
            mu_o, lv_o = O.vae3d_encode(sd, x)
            rec_o = O.vae3d_decode(sd, mu)
        assert torch.equal(mu, mu_o) and torch.equal(logvar, lv_o) and torch.equal(rec, rec_o), name
        assert torch.equal(lat, mu)
        cases[name] = {"kwargs": kw, "shapes": shapes, "seed": seed, "x": x, "mu": mu, "logvar": logvar,
                       "recon_of_mu": rec}
        print(f"{name}: {len(shapes)} keys, mu {tuple(mu.shape)} sha {sha(mu)}, recon {tuple(rec.shape)} "
              f"sha {sha(rec)}  (oracle == reference)")
    torch.save(cases, os.path.join(GOLD, "vae3d.pt"))


def main():
    sys.path.insert(0, "/root/reference")
    if "--only=vae" in sys.argv:
        os.makedirs(GOLD, exist_ok=True)
        torch.set_num_threads(4)
        vae_cases()
        return
    from model_scripts.ddpm_3d_ldm.diffusion import GaussianDiffusionLatent3D
    from model_scripts.ddpm_3d_ldm.unet import UNet3DModel
    from model_scripts.ddpm_3d_ldm.unet_attention import UNet3DModelWithAttention
    from model_scripts.ddpm_25d_all_modalities.diffusion import GaussianDiffusion as GD25
    from model_scripts.ddpm_25d_all_modalities.unet import UNet as UNet25
    from model_scripts.slice_cond_2d_ddpm.diffusion import GaussianDiffusion as GD2
    from model_scripts.slice_cond_2d_ddpm.unet import UNet as UNet2

    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(4)

    # ---------------------------------------------------------------- UNet forwards
    def unet_case(name, ctor, kwargs, inputs, oracle_fn, seed):
        m = quiet(ctor, **kwargs).eval()
        shapes = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
        sd = synthetic_state_dict(shapes, seed)
        m.load_state_dict(sd, strict=True)
        with torch.no_grad():
            out = m(*inputs["args"], **inputs.get("kwargs", {}))
            mine = oracle_fn(sd, *inputs["args"], **inputs.get("kwargs", {}))
        assert torch.equal(out, mine), f"{name}: oracle differs from the reference " \
            f"(max abs {(out - mine).abs().max().item():.3e})"
        torch.save({"kwargs": kwargs, "shapes": shapes, "seed": seed, "args": inputs["args"],
                    "call_kwargs": inputs.get("kwargs", {}), "out": out},
                   os.path.join(GOLD, f"{name}.pt"))
        print(f"{name}: {len(shapes)} keys, out {tuple(out.shape)} sha {sha(out)}  (oracle == reference)")

    g = torch.Generator().manual_seed(11)
    x3 = torch.randn(2, 3, 8, 8, 8, generator=g)
    t3 = torch.tensor([3, 977])
    unet_case("unet3d_attn", UNet3DModelWithAttention,
              dict(in_channels=3, base_channels=16, channel_mults=(1, 2, 4), time_emb_dim=32,
                   groups=8, num_heads=4), {"args": (x3, t3)}, O.unet3d_forward, 101)
    unet_case("unet3d", UNet3DModel,
              dict(in_channels=3, base_channels=16, channel_mults=(1, 2, 4), time_emb_dim=32, groups=8),
              {"args": (x3, t3)}, O.unet3d_forward, 102)
    x2 = torch.randn(2, 1, 16, 16, generator=g)
    z2 = torch.rand(2, generator=g)
    unet_case("unet2d", UNet2, dict(img_channels=1, base_channels=16, channel_mults=(1, 2, 4, 8),
                                    time_emb_dim=32),
              {"args": (x2, t3, z2)}, O.unet2d_forward, 103)
    x25 = torch.randn(2, 4, 16, 16, generator=g)
    c25 = torch.randn(2, 16, 16, 16, generator=g)
    unet_case("unet25d", UNet25, dict(in_channels=20, out_channels=4, base_channels=16,
                                      channel_mults=(1, 2, 4, 8), time_emb_dim=32),
              {"args": (x25, t3, z2), "kwargs": {"context": c25}}, O.unet2d_forward, 104)

    # ---------------------------------------------------------------- schedules (bit-exact)
    class Stub(torch.nn.Module):
        def forward(self, x, t, *a, **k):
            return 0.1 * x + 0.01 * t.float().view(-1, *([1] * (x.dim() - 1)))

    sched = {}
    for name, mk in [("linear_1000", lambda: quiet(GD2, Stub(), 16, channels=1, timesteps=1000)),
                     ("linear_50", lambda: quiet(GD2, Stub(), 16, channels=1, timesteps=50)),
                     ("linear25_1000", lambda: quiet(GD25, Stub(), 16, channels=4, timesteps=1000)),
                     ("cosine_1000", lambda: quiet(GaussianDiffusionLatent3D, Stub(), 3, timesteps=1000)),
                     ("cosine_400", lambda: quiet(GaussianDiffusionLatent3D, Stub(), 3, timesteps=400))]:
        d = mk()
        bufs = {k: v.clone() for k, v in d.state_dict().items() if not k.startswith("model.")}
        T = d.timesteps
        betas = O.cosine_betas(T) if name.startswith("cosine") else O.linear_betas(T)
        mine = O.schedule_buffers(betas, with_snr=("snr" in bufs))
        assert list(mine.keys()) == list(bufs.keys()), (name, list(mine), list(bufs))
        for k in bufs:
            assert torch.equal(mine[k], bufs[k]), f"schedule {name}.{k} differs"
        sched[name] = {"T": T, "sha": {k: sha(v) for k, v in bufs.items()},
                       "first_last": {k: (v[0].item(), v[-1].item()) for k, v in bufs.items()}}
        print(f"schedule {name}: {len(bufs)} buffers bit-exact vs oracle")
    # index schedules (train.py:391, :446-448; diffusion.py:138)
    sched["val_grid_400"] = torch.linspace(1, 399, steps=8).long()
    sched["val_grid_1000"] = torch.linspace(1, 999, steps=8).long()
    sched["sin_emb_256"] = O.sinusoidal(torch.tensor([0, 1, 500, 999]), 256)
    torch.save(sched, os.path.join(GOLD, "schedules.pt"))

    # ---------------------------------------------------------------- diffusion arithmetic
    d3 = quiet(GaussianDiffusionLatent3D, Stub(), 3, timesteps=1000)
    d2 = quiet(GD2, Stub(), 16, channels=1, timesteps=1000)
    buf3 = {k: v for k, v in d3.state_dict().items() if not k.startswith("model.")}
    buf2 = {k: v for k, v in d2.state_dict().items() if not k.startswith("model.")}
    g = torch.Generator().manual_seed(21)
    x = torch.randn(4, 3, 6, 8, 6, generator=g)
    noise = torch.randn(4, 3, 6, 8, 6, generator=g)
    t = torch.tensor([0, 1, 500, 999])
    diff = {"x": x, "noise": noise, "t": t}
    diff["q_sample_3d"] = d3.q_sample(x, t, noise=noise)
    assert torch.equal(diff["q_sample_3d"], O.q_sample(buf3, x, t, noise))
    torch.manual_seed(77)
    diff["p_sample_3d"] = d3.p_sample(x, t)
    torch.manual_seed(77)
    z = torch.randn_like(x)
    diff["p_sample_noise"] = z
    eps = Stub()(x, t)
    diff["eps"] = eps
    assert torch.equal(diff["p_sample_3d"], O.p_sample_update(buf3, x, t, eps, z))
    diff["p_losses_3d"] = d3.p_losses(x, t, noise=noise, min_snr_gamma=5.0)
    xn = O.q_sample(buf3, x, t, noise)
    assert torch.equal(diff["p_losses_3d"], O.minsnr_loss(buf3, Stub()(xn, t), noise, t, 5.0))
    t_prev = torch.tensor([0, 0, 499, 998])
    diff["t_prev"] = t_prev
    diff["ddim_3d"] = d3.p_sample_ddim(x, t, t_prev)
    assert torch.equal(diff["ddim_3d"], O.ddim_update(buf3, x, t, t_prev, eps))
    torch.manual_seed(5)
    diff["sample_from_3d"] = d3.sample_from(x, 9)
    torch.manual_seed(5)
    img = x
    for i in reversed(range(10)):
        ti = torch.full((4,), i, dtype=torch.long)
        img = O.p_sample_update(buf3, img, ti, Stub()(img, ti), torch.randn_like(img))
    assert torch.equal(diff["sample_from_3d"], img)
    diff["sample_from_ddim_3d"] = d3.sample_from_ddim(x, 9)
    # 2D: q_sample / p_sample with z_pos (model stub ignores it)
    x2d = torch.randn(4, 1, 16, 16, generator=g)
    n2d = torch.randn(4, 1, 16, 16, generator=g)
    diff["x2d"], diff["n2d"] = x2d, n2d
    diff["q_sample_2d"] = d2.q_sample(x2d, t, noise=n2d)
    assert torch.equal(diff["q_sample_2d"], O.q_sample(buf2, x2d, t, n2d))
    torch.manual_seed(78)
    diff["p_sample_2d"] = d2.p_sample(x2d, t, torch.rand(4))
    torch.manual_seed(78)
    _ = torch.rand(4)
    z2d = torch.randn_like(x2d)
    diff["p_sample_noise_2d"] = z2d
    assert torch.equal(diff["p_sample_2d"], O.p_sample_update(buf2, x2d, t, Stub()(x2d, t), z2d))
    torch.save(diff, os.path.join(GOLD, "diffusion.pt"))
    print("diffusion arithmetic fixtures written (oracle == reference, bit-exact)")
    vae_cases()
    for f in sorted(os.listdir(GOLD)):
        print(f"  {f}: {os.path.getsize(os.path.join(GOLD, f)) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
