"""mri_image_generation_b200.optim.Adam (one-launch mri_adam_step) vs torch.optim.Adam on the same
parameters / gradients: fp32 arithmetic in a different association order -> rtol 2e-6 per step."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def make_params(seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(128, 64, 3, 3, 3), (128,), (7, 5), (1,), (300, 33)]
    return [torch.nn.Parameter(torch.randn(*s, generator=g).cuda()) for s in shapes]


@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_adam_matches_torch(wd):
    from mri_image_generation_b200.optim import Adam
    pa, pb = make_params(1), make_params(1)
    oa = Adam(pa, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    ob = torch.optim.Adam(pb, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    g = torch.Generator().manual_seed(2)
    for step in range(5):
        v0 = [p._version for p in pa]
        for x, y in zip(pa, pb):
            gr = torch.randn(*x.shape, generator=g).cuda()
            x.grad, y.grad = gr.clone(), gr.clone()
        oa.step()
        ob.step()
        assert all(p._version > v for p, v in zip(pa, v0)), "parameter versions must advance"
        for x, y in zip(pa, pb):
            assert torch.allclose(x, y, rtol=2e-6, atol=1e-7), (step, (x - y).abs().max().item())
    sa, sb = oa.state_dict(), ob.state_dict()
    assert sa["state"].keys() == sb["state"].keys()
    for k in sa["state"]:
        assert set(sa["state"][k]) == {"step", "exp_avg", "exp_avg_sq"}
        for key in ("exp_avg", "exp_avg_sq"):
            a, b = sa["state"][k][key], sb["state"][k][key]
            assert torch.allclose(a, b, rtol=1e-4, atol=1e-6), (k, key, (a - b).abs().max().item())
        assert float(sa["state"][k]["step"]) == float(sb["state"][k]["step"]) == 5.0


def test_adam_with_grad_scaler_skips_on_inf_and_unscales():
    from mri_image_generation_b200.optim import Adam
    pa, pb = make_params(3), make_params(3)
    oa = Adam(pa, lr=1e-3)
    ob = torch.optim.Adam(pb, lr=1e-3)
    sa = torch.amp.GradScaler("cuda", init_scale=1024.0)
    sb = torch.amp.GradScaler("cuda", init_scale=1024.0)
    for it in range(3):
        la = sum((p * p).sum() for p in pa)
        lb = sum((p * p).sum() for p in pb)
        oa.zero_grad(set_to_none=True)
        ob.zero_grad(set_to_none=True)
        sa.scale(la).backward()
        sb.scale(lb).backward()
        if it == 1:  # poison one gradient: both scalers must skip the step
            pa[0].grad[0, 0, 0, 0, 0] = float("inf")
            pb[0].grad[0, 0, 0, 0, 0] = float("inf")
        before = [p.detach().clone() for p in pa]
        sa.step(oa)
        sb.step(ob)
        sa.update()
        sb.update()
        if it == 1:
            assert all(torch.equal(a, b) for a, b in zip(before, pa))
        for x, y in zip(pa, pb):
            assert torch.allclose(x, y, rtol=5e-6, atol=1e-7), it
    assert sa.get_scale() == sb.get_scale()


def test_adam_resume_from_state_dict_uses_the_loaded_moments():
    """load_state_dict installs new moment tensors: the cached segment table must follow them
    (it used to keep pointing at the freed ones), and the step counter must continue."""
    import copy
    from mri_image_generation_b200.optim import Adam
    pa, pb = make_params(5), make_params(5)
    oa = Adam(pa, lr=1e-3)
    ob = torch.optim.Adam(pb, lr=1e-3)
    g = torch.Generator().manual_seed(6)

    def give_grads(ps_a, ps_b):
        for x, y in zip(ps_a, ps_b):
            gr = torch.randn(*x.shape, generator=g).cuda()
            if x.grad is None:
                x.grad, y.grad = gr.clone(), gr.clone()
            else:  # same .grad storage as before: the table's (param, grad) pointers do not change
                x.grad.copy_(gr)
                y.grad.copy_(gr)

    for _ in range(3):
        give_grads(pa, pb)
        oa.step()
        ob.step()
    saved = copy.deepcopy(oa.state_dict())
    for _ in range(2):  # move on, then roll the optimizer state back
        give_grads(pa, pb)
        oa.step()
    with torch.no_grad():
        for x, y in zip(pa, pb):
            x.copy_(y)
    oa.load_state_dict(saved)
    for _ in range(2):
        give_grads(pa, pb)
        oa.step()
        ob.step()
        for x, y in zip(pa, pb):
            assert torch.allclose(x, y, rtol=5e-6, atol=1e-7), (x - y).abs().max().item()
    assert float(oa.state_dict()["state"][0]["step"]) == 5.0


def test_adam_late_parameter_does_not_reset_the_group_step():
    """A parameter that gets its first gradient later (and is first in the list) must not alias
    a fresh step = 0 onto the others."""
    from mri_image_generation_b200.optim import Adam
    pa = make_params(7)
    oa = Adam(pa, lr=1e-3)
    for _ in range(3):
        for p in pa[1:]:
            p.grad = torch.ones_like(p)
        oa.step()
    for p in pa:
        p.grad = torch.ones_like(p)
    oa.step()
    assert float(oa.state[pa[1]]["step"]) == 4.0


def test_weight_repack_after_a_step_is_pure_gather():
    """After an optimizer step every derived weight buffer (packed bf16 matrices, the concatenated
    block projections, padded biases) is refreshed by ONE mri_gather_pack launch: no buffer is left
    on the torch-expression fallback, and the refreshed buffers equal a fresh packing."""
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.unet_attention import UNet3DModelWithAttention
    torch.manual_seed(3)
    m = UNet3DModelWithAttention(3, base_channels=64, time_emb_dim=64).cuda().train()
    prog = m.program(2, (8, 8, 8), training=True)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(torch.randn_like(p) * 0.01)
    prog.do_refresh()
    info = prog.gather_info
    assert info["fallback"] == 0, info
    for buf, make in prog._packed:
        assert torch.equal(buf, make().reshape(buf.shape).to(buf.dtype)), "stale packed buffer"
