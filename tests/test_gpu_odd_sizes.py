"""Image sizes the down-sampling factor does not divide: the 2D / 2.5D UpBlock resizes the
up-sampled tensor to the skip's size with F.interpolate(mode="bilinear", align_corners=False)
(slice_cond_2d_ddpm/unet.py:95-99, ddpm_25d_all_modalities/unet.py:95-99).  On the B200 path:
odd levels are zero-padded to an even size in front of the stride-2 convolution and cropped back
(mri_copy_window_nhwc), the resize and its adjoint are mri_resize_bilinear_nhwc[_bwd].
Forward vs the fp32 CPU oracle (rel-L2 <= 2e-2), every gradient as in test_gpu_train2d.py."""
import pytest
import torch

from helpers import rel_l2, shapes_of, synthetic_state_dict
from oracle import reference_oracle as O
from test_gpu_train2d import compare, oracle_grads, quiet

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("size", [(36, 44), (30, 25), (100, 100), (17, 23)])
def test_resize_kernels_match_torch(size):
    """The two resample kernels alone against F.interpolate and its autograd."""
    from mri_image_generation_b200 import ops
    H, W = size
    g = torch.Generator().manual_seed(H * 100 + W)
    for (sh, sw) in [(H - 1, W - 1), (H - 1, W), (2 * (H // 2), 2 * (W // 2)), (H // 2, W // 2 + 3)]:
        x = torch.randn(3, sh, sw, 16, generator=g).to(torch.bfloat16)
        y = torch.zeros(3, H, W, 16, dtype=torch.bfloat16, device="cuda")
        ops.resize_bilinear(x.cuda(), y)
        xr = x.float().permute(0, 3, 1, 2).requires_grad_()
        want = torch.nn.functional.interpolate(xr, size=(H, W), mode="bilinear", align_corners=False)
        assert rel_l2(y.float().permute(0, 3, 1, 2), want) < 4e-3, (size, sh, sw)
        dy = torch.randn(3, H, W, 16, generator=g).to(torch.bfloat16)
        want.backward(dy.float().permute(0, 3, 1, 2))
        dx = torch.zeros(3, sh, sw, 16, dtype=torch.bfloat16, device="cuda")
        ops.resize_bilinear_bwd(dy.cuda(), dx)
        assert rel_l2(dx.float().permute(0, 3, 1, 2), xr.grad) < 4e-3, (size, sh, sw)
        add = torch.randn(3, sh, sw, 16, generator=g).to(torch.bfloat16).cuda()
        dx2 = add.clone()
        ops.resize_bilinear_bwd(dy.cuda(), dx2, add=dx2)
        assert rel_l2(dx2.float(), dx.float() + add.float()) < 6e-3


def test_copy_window():
    from mri_image_generation_b200 import ops
    src = torch.randn(2, 9, 11, 8).to(torch.bfloat16).cuda()
    dst = torch.zeros(2, 10, 12, 8, dtype=torch.bfloat16, device="cuda")
    ops.copy_window(src, dst, (0, 0), (0, 0), (9, 11))
    assert torch.equal(dst[:, :9, :11], src) and not dst[:, 9:].any() and not dst[:, :, 11:].any()
    back = torch.ones(2, 9, 11, 8, dtype=torch.bfloat16, device="cuda")
    ops.copy_window(dst, back, (1, 2), (0, 1), (4, 5), add=back)
    want = torch.ones_like(back)
    want[:, 0:4, 1:6] += dst[:, 1:5, 2:7]
    assert torch.equal(back, want)


@pytest.mark.parametrize("size,batch", [((36, 44), 2), ((30, 25), 3), ((100, 100), 1)])
def test_unet2d_forward_at_sizes_not_divisible_by_8(size, batch):
    from mri_image_generation_b200.model_scripts.slice_cond_2d_ddpm.unet import UNet
    m = quiet(UNet, img_channels=1, base_channels=64, time_emb_dim=64)
    sd = synthetic_state_dict(shapes_of(m), seed=41)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(batch, 1, *size, generator=g)
    t = torch.randint(0, 1000, (batch,), generator=g)
    z = torch.rand(batch, generator=g)
    with torch.no_grad():
        got = m(x.cuda(), t.cuda(), z.cuda())
        want = O.unet2d_forward(sd, x, t, z)
    assert got.shape == want.shape == x.shape
    e = rel_l2(got, want)
    print(f"{size} B={batch}: eps rel-L2 {e:.3e}")
    assert e < 2e-2


def test_sampling_loop_at_an_odd_size():
    """sample() through the fused reverse step at 30 x 30 -- levels 30, 15, 7, 3: two odd levels and
    two resizes (the reference's GaussianDiffusion samples squares: diffusion.py:144-149) -- graph
    replay == the eager steps."""
    from mri_image_generation_b200.model_scripts.slice_cond_2d_ddpm.diffusion import GaussianDiffusion
    from mri_image_generation_b200.model_scripts.slice_cond_2d_ddpm.unet import UNet
    m = quiet(UNet, img_channels=1, base_channels=64, time_emb_dim=64)
    m.load_state_dict(synthetic_state_dict(shapes_of(m), seed=42))
    m = m.cuda().eval()
    diff = quiet(GaussianDiffusion, m, 30, channels=1, timesteps=8).cuda()
    z = torch.rand(2, device="cuda")
    torch.manual_seed(5)
    a = diff.sample(2, z)
    torch.manual_seed(5)
    x = torch.randn(2, 1, 30, 30, device="cuda")
    for i in reversed(range(8)):
        x = diff.p_sample(x, torch.full((2,), i, device="cuda", dtype=torch.long), z)
    assert a.shape == (2, 1, 30, 30) and torch.isfinite(a).all()
    assert torch.equal(a, x)


@pytest.mark.parametrize("size", [(36, 44), (30, 25)])
def test_train_step_25d_with_context_at_odd_sizes(size):
    from mri_image_generation_b200.model_scripts.ddpm_25d_all_modalities.diffusion import GaussianDiffusion
    from mri_image_generation_b200.model_scripts.ddpm_25d_all_modalities.unet import UNet
    m = quiet(UNet, in_channels=20, out_channels=4, base_channels=64, time_emb_dim=64)
    sd = synthetic_state_dict(shapes_of(m), seed=43)
    m.load_state_dict(sd)
    m = m.cuda().train()
    T = 100
    diff = quiet(GaussianDiffusion, m, size[0], channels=4, timesteps=T).cuda()   # image_size only sizes sample()
    g = torch.Generator().manual_seed(11)
    x0 = torch.randn(2, 4, *size, generator=g)
    noise = torch.randn(2, 4, *size, generator=g)
    ctx = torch.randn(2, 16, *size, generator=g)
    t = torch.randint(0, T, (2,), generator=g)
    z = torch.rand(2, generator=g)
    buf = O.schedule_buffers(O.linear_betas(T), with_snr=False)
    xn = O.q_sample(buf, x0, t, noise)
    lf = lambda pred, nz, tt: O.mse_loss(pred, nz)
    ref_loss, ref = oracle_grads(sd, xn, t, z, noise, ctx, lf)
    _, auto = oracle_grads(sd, xn, t, z, noise, ctx, lf, device="cuda", autocast=True)
    for rep in range(4):     # eager, eager, capture, replay
        m.zero_grad(set_to_none=True)
        loss = diff.p_losses(x0.cuda(), t.cuda(), z.cuda(), context=ctx.cuda(), noise=noise.cuda())
        loss.backward()
        assert abs(loss.item() - ref_loss) <= 1e-2 * abs(ref_loss), (loss.item(), ref_loss)
        compare(m, ref, auto, f"2.5D {size} rep {rep}")
