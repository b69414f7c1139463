"""In-kernel Philox4x32-10 + Box-Muller (csrc/diffusion_ops.cu) against torch's own CUDA generator:
the noise the reference draws with torch.randn / torch.randn_like (ddpm_3d_ldm/diffusion.py:75,
86,122,131; slice_cond_2d_ddpm/diffusion.py:67,93,129,140) must come out BIT-IDENTICAL for the
same torch.manual_seed, and torch's generator must end up at the same offset, so that every later
draw matches as well."""
import contextlib
import io

import pytest
import torch

from helpers import shapes_of, synthetic_state_dict
from oracle import reference_oracle as O

pytestmark = pytest.mark.gpu


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def _gen():
    return torch.cuda.default_generators[torch.cuda.current_device()]


@pytest.mark.parametrize("numel", [1, 3, 4, 255, 1024, 1025, 230_400, 303_104 * 4, 303_104 * 4 + 5,
                                   3_686_400, 5_000_001])
@pytest.mark.parametrize("seed", [0, 1234, 2 ** 63 + 12345])
def test_randn_matches_torch_bit_exactly(numel, seed):
    from mri_image_generation_b200 import ops
    dev = torch.device("cuda", torch.cuda.current_device())
    rng = ops.DeviceRng(dev)
    torch.manual_seed(seed)
    burn = torch.randn(7, device=dev)          # a non-zero starting offset
    off0 = _gen().get_offset()
    want = torch.randn(numel, device=dev)
    off1 = _gen().get_offset()
    torch.manual_seed(seed)
    torch.randn(7, device=dev)
    got = torch.empty(numel, device=dev)
    assert rng.load() == off0
    ops.randn(got, rng)
    inc = ops.randn_offset_increment(numel)
    assert off0 + inc == off1, (off0, inc, off1)
    assert torch.equal(got, want), (got - want).abs().max().item()
    del burn


def test_diffusion_draws_equal_the_reference_call_sites():
    """q_sample(noise=None), p_sample and sample() consume torch's generator exactly like
    `noise = torch.randn_like(x)` at the reference's call sites."""
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.diffusion import GaussianDiffusionLatent3D
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.unet import UNet3DModel
    m = UNet3DModel(3, base_channels=64, time_emb_dim=64)
    sd = synthetic_state_dict(shapes_of(m), seed=71)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    T = 20
    diff = quiet(GaussianDiffusionLatent3D, m, 3, timesteps=T).cuda()
    buf = O.schedule_buffers(O.cosine_betas(T))
    g = torch.Generator().manual_seed(3)
    x0 = torch.randn(3, 3, 8, 8, 8, generator=g).cuda()
    t = torch.tensor([0, 7, T - 1], device="cuda")
    # ---- q_sample with its own noise -------------------------------------------------------
    torch.manual_seed(5)
    got = diff.q_sample(x0, t)
    after = torch.randn(5, device="cuda")
    torch.manual_seed(5)
    noise = torch.randn_like(x0)
    assert torch.equal(got.cpu(), O.q_sample(buf, x0.cpu(), t.cpu(), noise.cpu()))
    assert torch.equal(after, torch.randn(5, device="cuda"))   # generator left where torch leaves it
    # ---- sample(): x_T = randn(shape), then z = randn_like(x) at EVERY step, t = 0 included ----
    torch.manual_seed(6)
    vol = diff.sample(2, (8, 8, 8))
    after = torch.randn(5, device="cuda")
    torch.manual_seed(6)
    img = torch.randn(2, 3, 8, 8, 8, device="cuda")
    with torch.no_grad():
        for i in reversed(range(T)):
            tt = torch.full((2,), i, device="cuda", dtype=torch.long)
            eps = m(img, tt)
            z = torch.randn_like(img)
            img = diff._p_update(img, tt, eps, z)
    assert torch.equal(after, torch.randn(5, device="cuda"))
    assert torch.equal(vol, img), (vol - img).abs().max().item()


def test_strided_ddim_graph_loop_equals_step_by_step():
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.diffusion import GaussianDiffusionLatent3D
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.unet import UNet3DModel
    m = UNet3DModel(3, base_channels=64, time_emb_dim=64)
    sd = synthetic_state_dict(shapes_of(m), seed=72)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    T = 50
    diff = quiet(GaussianDiffusionLatent3D, m, 3, timesteps=T).cuda()
    buf = O.schedule_buffers(O.cosine_betas(T))
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 3, 8, 8, 8, generator=g).cuda()
    for stride in (7, 10):
        got = diff.sample_from_ddim(x, T - 1, stride=stride)
        img = x
        ts = list(range(T - 1, 0, -stride))
        with torch.no_grad():
            for i in ts:
                t = torch.full((2,), i, device="cuda", dtype=torch.long)
                tp = torch.full((2,), max(i - stride, 0), device="cuda", dtype=torch.long)
                eps = m(img, t)
                want = O.ddim_update(buf, img.cpu(), t.cpu(), tp.cpu(), eps.cpu())
                img = diff._ddim_update(img, t, tp, eps)
                assert torch.equal(img.cpu(), want), (stride, i)
        assert len(ts) == -(-(T - 1) // stride)
        assert torch.equal(got, img), (stride, (got - img).abs().max().item())
    vol = diff.sample_ddim(2, 8, num_steps=5)
    assert vol.shape == (2, 3, 8, 8, 8) and torch.isfinite(vol).all()


@pytest.mark.parametrize("kind", ["3d", "2d", "25d"])
def test_fused_gather_step_equals_unfused_bit_exactly(kind, monkeypatch):
    """The sampling loop finishes out_conv and applies the update in ONE kernel
    (mri_tap_gather_step): same tap order, same bf16 rounding of eps, same Philox draws ->
    bit-identical to the unfused launches (tap gather, then mri_ddpm_step_rng / mri_ddim_step)."""
    if kind == "3d":
        from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.diffusion import GaussianDiffusionLatent3D
        from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.unet import UNet3DModel
        m = UNet3DModel(3, base_channels=64, time_emb_dim=64)
        mk = lambda mm: quiet(GaussianDiffusionLatent3D, mm, 3, timesteps=12).cuda()
        run = lambda d: d.sample(2, (8, 12, 8))
    elif kind == "2d":
        from mri_image_generation_b200.model_scripts.slice_cond_2d_ddpm.diffusion import GaussianDiffusion
        from mri_image_generation_b200.model_scripts.slice_cond_2d_ddpm.unet import UNet
        m = quiet(UNet, img_channels=1, base_channels=64, time_emb_dim=64)
        mk = lambda mm: quiet(GaussianDiffusion, mm, 40, channels=1, timesteps=12).cuda()
        run = lambda d: d.sample(3, z_pos=0.3)
    else:
        from mri_image_generation_b200.model_scripts.ddpm_25d_all_modalities.diffusion import GaussianDiffusion
        from mri_image_generation_b200.model_scripts.ddpm_25d_all_modalities.unet import UNet
        m = quiet(UNet, in_channels=20, out_channels=4, base_channels=64, time_emb_dim=64)
        mk = lambda mm: quiet(GaussianDiffusion, mm, 24, channels=4, timesteps=12).cuda()
        ctx = torch.randn(2, 16, 24, 24, generator=torch.Generator().manual_seed(1)).cuda()
        run = lambda d: d.sample(2, z_pos=0.6, context=ctx)
    sd = synthetic_state_dict(shapes_of(m), seed=91)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    outs = []
    for fused in ("1", "0"):
        monkeypatch.setenv("MRI_FUSED_STEP", fused)
        m.__dict__.pop("_mri_programs", None)      # fresh program: fresh step graph
        diff = mk(m)
        torch.manual_seed(17)
        outs.append(run(diff).clone())
        prog = next(iter(m._programs().values()))
        assert (prog.fused_head is not None)
    assert torch.isfinite(outs[0]).all()
    assert torch.equal(outs[0], outs[1]), (outs[0] - outs[1]).abs().max().item()
    if kind == "3d":   # DDIM (strided) through the fused kernel as well
        x = torch.randn(2, 3, 8, 12, 8, generator=torch.Generator().manual_seed(2)).cuda()
        res = []
        for fused in ("1", "0"):
            monkeypatch.setenv("MRI_FUSED_STEP", fused)
            m.__dict__.pop("_mri_programs", None)
            res.append(mk(m).sample_from_ddim(x, 11, stride=3).clone())
        assert torch.equal(res[0], res[1])
