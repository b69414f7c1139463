"""HBM-bound kernels vs torch on the same inputs (GPU): GroupNorm statistics / apply(+SiLU,
+rowbias, +residual), patch matrix, row softmax, linear, sinusoidal embedding, layout edges."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import reference_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from mri_image_generation_b200 import _lib, ops
    _lib.require_device()
    return ops


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("N,sp,C,groups,comb", [(2, (6, 8, 10), 128, 8, 1), (3, (30, 30), 64, 8, 1),
                                                  (2, (5, 7, 3), 256, 4, 2), (1, (40, 48, 40), 128, 8, 1),
                                                  (2, (9, 9), 768, 8, 1)])
def test_gn_stats_and_apply(ops, N, sp, C, groups, comb):
    torch.manual_seed(0)
    dev = "cuda"
    S = math.prod(sp)
    x = (torch.randn(N, S, C, device=dev) * 2 + 0.5).to(torch.bfloat16)
    gamma = torch.randn(C, device=dev)
    beta = torch.randn(C, device=dev)
    rb = torch.randn(N, C + 16, device=dev)
    res = torch.randn(N, S, C, device=dev).to(torch.bfloat16)
    fine = groups * comb
    stats = torch.zeros(N, fine, 2, device=dev, dtype=torch.float64)
    ops.gn_stats(x, stats, N, S, C, C // fine)
    xr = x.float().reshape(N, S, fine, C // fine)
    assert rel(stats[:, :, 0], xr.sum((1, 3))) < 1e-5
    assert rel(stats[:, :, 1], (xr ** 2).sum((1, 3))) < 1e-5
    for silu, use_rb, use_res in [(True, False, False), (True, True, True), (False, False, True)]:
        y = torch.empty_like(x)
        ops.gn_apply(x, y, stats, gamma, beta, N, S, C, groups, C // fine, 1e-5, silu,
                     rowbias=rb if use_rb else None, rowbias_ld=C + 16,
                     residual=res if use_res else None)
        ref = F.group_norm(x.float().permute(0, 2, 1), groups, gamma, beta, 1e-5).permute(0, 2, 1)
        if silu:
            ref = F.silu(ref)
        if use_rb:
            ref = ref + rb[:, None, :C]
        if use_res:
            ref = ref + res.float()
        assert rel(y, ref) < 4e-3  # bf16 output rounding


def test_im2col_matches_unfold(ops):
    torch.manual_seed(1)
    dev = "cuda"
    # 3D, cin 3
    x = torch.randn(2, 3, 6, 8, 10, device=dev)
    kpad = 128
    col = torch.full((2 * 480, kpad), 7.0, device=dev, dtype=torch.bfloat16)
    ops.im2col(x, col, 2, 3, 6, 8, 10, 3, 3, kpad)
    xp = F.pad(x, (1, 1, 1, 1, 1, 1))
    ref = torch.zeros(2, 6, 8, 10, 27, 3, device=dev)
    for kd in range(3):
        for kh in range(3):
            for kw in range(3):
                ref[..., (kd * 3 + kh) * 3 + kw, :] = xp[:, :, kd:kd + 6, kh:kh + 8, kw:kw + 10].permute(0, 2, 3, 4, 1)
    assert torch.equal(col[:, :81].float(), ref.reshape(-1, 81).to(torch.bfloat16).float())
    assert (col[:, 81:] == 0).all()
    # 2D with context channels (4 + 16)
    a, c = torch.randn(2, 4, 12, 16, device=dev), torch.randn(2, 16, 12, 16, device=dev)
    col = torch.empty(2 * 192, 192, device=dev, dtype=torch.bfloat16)
    ops.im2col(a, col, 2, 4, 1, 12, 16, 3, 2, 192, src2=c, cin2=16)
    xp = F.pad(torch.cat([a, c], 1), (1, 1, 1, 1))
    ref = torch.zeros(2, 12, 16, 9, 20, device=dev)
    for kh in range(3):
        for kw in range(3):
            ref[..., kh * 3 + kw, :] = xp[:, :, kh:kh + 12, kw:kw + 16].permute(0, 2, 3, 1)
    assert torch.equal(col[:, :180].float(), ref.reshape(-1, 180).to(torch.bfloat16).float())


@pytest.mark.parametrize("rows,cols", [(64, 8), (100, 200), (4800, 1200), (7, 2000)])
def test_softmax_rows(ops, rows, cols):
    torch.manual_seed(2)
    ld = (cols + 7) // 8 * 8
    S = torch.randn(rows, ld, device="cuda") * 3
    P = torch.empty(rows, ld, device="cuda", dtype=torch.bfloat16)
    ops.softmax_rows(S, P, rows, cols, ld, ld, 0.125)
    ref = torch.softmax(S[:, :cols] * 0.125, -1)
    assert rel(P[:, :cols], ref) < 4e-3
    assert (P[:, cols:] == 0).all()


def test_linear_and_sinusoidal(ops):
    torch.manual_seed(3)
    dev = "cuda"
    for B in (1, 4, 19):
        x = torch.randn(B, 256, device=dev)
        W = torch.randn(1024, 256, device=dev) * 0.05
        b = torch.randn(1024, device=dev)
        add = torch.randn(B, 1024, device=dev)
        y = torch.empty(B, 1024, device=dev)
        ops.linear(x, W, b, y, act=1, addend=add)
        ref = F.silu(F.linear(x.double(), W.double(), b.double())).float() + add
        assert rel(y, ref) < 1e-5
    t = torch.tensor([0, 1, 500, 999], device=dev)
    out = torch.empty(4, 256, device=dev)
    ops.sinusoidal(t, ops.sinusoidal_freqs(256, dev), out)
    ref = O.sinusoidal(t.cpu(), 256)
    assert (out.cpu() - ref).abs().max() < 2e-6  # only sinf/cosf implementations differ (ulps)


def test_layout_edges(ops):
    x = torch.randn(2, 3, 5, 6, 7, device="cuda")
    nh = torch.empty(2, 210, 16, device="cuda", dtype=torch.bfloat16)
    ops.nchw_to_nhwc(x, nh, 2, 210, 3, 16)
    assert torch.equal(nh[:, :, :3].float(), x.reshape(2, 3, 210).permute(0, 2, 1).to(torch.bfloat16).float())
    assert (nh[:, :, 3:] == 0).all()
    back = torch.empty(2, 3, 210, device="cuda")
    ops.nhwc_to_nchw(nh, back, 2, 210, 3, 16)
    assert torch.equal(back, x.reshape(2, 3, 210).to(torch.bfloat16).float())
