"""Split precision mode, host side (no GPU): the widened operands [hi | lo | hi] x [w_hi | w_hi | w_lo]
through the ordinary convolution plans (CPU emulation of the tables, plan.simulate) reproduce the
fp32 convolution to ~1e-5, where the plain bf16 operands give ~3e-3."""
import pytest
import torch
import torch.nn.functional as F

from mri_image_generation_b200 import plan as P
from mri_image_generation_b200 import split_engine as SE


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


@pytest.mark.parametrize("nd,sp", [(3, (4, 6, 8)), (2, (8, 12))])
def test_widened_conv_matches_fp32(nd, sp):
    torch.manual_seed(0)
    N, C, Co = 2, 64, 128
    conv = F.conv3d if nd == 3 else F.conv2d
    x = torch.randn(N, C, *sp)
    w = torch.randn(Co, C, *([3] * nd)) * 0.05
    b = torch.randn(Co)
    want = conv(x, w, b, padding=1)
    x_cl = x.permute(0, *range(2, 2 + nd), 1).contiguous()                 # [N, *sp, C] fp32
    x3 = SE.widen_activation(x_cl)                                          # [N, *sp, 3C] bf16
    w3 = SE.widen_weight(w)                                                 # [Co, 3C, *k] fp32, bf16-exact
    assert torch.equal(w3.to(torch.bfloat16).float(), w3)
    y = torch.zeros(N, *sp, Co, dtype=torch.float32)
    P.conv_plan([P.ConvSource(x3)], P.pack_conv_weight(w3), y, 3, bias=b, out_f32=True).simulate()
    got = y.permute(0, nd + 1, *range(1, nd + 1))
    e_split = rel(got, want)
    # plain bf16 operands for comparison
    yb = torch.zeros(N, *sp, Co, dtype=torch.bfloat16)
    P.conv_plan([P.ConvSource(x_cl.to(torch.bfloat16))], P.pack_conv_weight(w), yb, 3, bias=b).simulate()
    e_bf16 = rel(yb.float().permute(0, nd + 1, *range(1, nd + 1)), want)
    print(f"split {e_split:.2e}  bf16 {e_bf16:.2e}")
    assert e_split < 2e-5 and e_bf16 > 1e-3


def test_widened_down_and_up_conv_match_fp32():
    torch.manual_seed(1)
    N, C, Co, sp = 1, 64, 64, (4, 8, 8)
    x = torch.randn(N, C, *sp)
    x3 = SE.widen_activation(x.permute(0, 2, 3, 4, 1).contiguous())
    w = torch.randn(Co, C, 4, 4, 4) * 0.05
    want = F.conv3d(x, w, None, stride=2, padding=1)
    y = torch.zeros(N, *[s // 2 for s in sp], Co, dtype=torch.float32)
    P.down_conv_plan(x3, P.pack_conv_weight(SE.widen_weight(w)), y, out_f32=True).simulate()
    assert rel(y.permute(0, 4, 1, 2, 3), want) < 2e-5
    wt = torch.randn(C, Co, 4, 4, 4) * 0.05
    want = F.conv_transpose3d(x, wt, None, stride=2, padding=1)
    y = torch.zeros(N, *[s * 2 for s in sp], Co, dtype=torch.float32)
    P.up_conv_plan(x3, P.pack_convT_weight(SE.widen_weight(wt, transposed=True)), y, out_f32=True).simulate()
    assert rel(y.permute(0, 4, 1, 2, 3), want) < 2e-5


def test_identity_and_concat_weights():
    torch.manual_seed(2)
    w = torch.randn(8, 6, 3, 3, 3)
    w3 = SE.widen_weight(w, splits=[2, 4])
    assert w3.shape == (8, 18, 3, 3, 3)
    hi = w.to(torch.bfloat16).float()
    lo = (w - hi).to(torch.bfloat16).float()
    # per source: [hi | hi | lo] of that source's channels
    assert torch.equal(w3[:, 0:2], hi[:, 0:2]) and torch.equal(w3[:, 2:4], hi[:, 0:2]) and torch.equal(w3[:, 4:6], lo[:, 0:2])
    assert torch.equal(w3[:, 6:10], hi[:, 2:6]) and torch.equal(w3[:, 14:18], lo[:, 2:6])
    eye = SE.identity_weight(4)
    assert eye.shape == (4, 12) and torch.equal(eye[:, :4], torch.eye(4)) and torch.equal(eye[:, 4:8], torch.eye(4))
    assert not eye[:, 8:].any()


def test_widened_attention_products_match_fp32():
    """q k^T and P v of AttentionBlock3D (unet_attention.py:49-51) with widened operands
    ([hi | lo | hi] x [hi | hi | lo] along the contraction) through matrix_plan's CPU emulation."""
    torch.manual_seed(3)
    B, heads, n, d = 1, 2, 100, 64
    npad = 128
    q, k, v = (torch.randn(B, n, heads, d) for _ in range(3))
    want_S = torch.einsum("bnhd,bmhd->bhnm", q, k)

    def widen(x, pattern):            # along the last axis
        hi = x.to(torch.bfloat16)
        lo = (x - hi.float()).to(torch.bfloat16)
        return torch.cat([hi, lo, hi] if pattern == 0 else [hi, hi, lo], dim=-1).contiguous()

    q3, k3 = widen(q, 0), widen(k, 1)                     # [B, n, heads, 3d]
    S = torch.zeros(B, heads, n, npad)
    ld = heads * 3 * d
    qa = P.TView(q3, (3 * d, n, heads, B, 1), (1, ld, 3 * d, n * ld, B * n * ld))
    kb = P.TView(k3, (3 * d, n, heads, B), (1, ld, 3 * d, n * ld))
    so = P.TView(S, (npad, n, heads, B, 1), (1, npad, n * npad, heads * n * npad, B * heads * n * npad))
    P.matrix_plan(qa, (128, 1, 1, 1), kb, so, K=3 * d, n_total=npad, block_n=128, ext=(n, heads, B, 1),
                  tiles=(-(-n // 128), heads, B, 1), bz_sel=(3, 4), out_f32=True).simulate()
    assert rel(S[..., :n], want_S) < 2e-5
    # P v with widened probabilities [hi | lo | hi] (segments of npad keys) and values [hi | hi | lo]
    Pm = torch.softmax(want_S * d ** -0.5, -1)
    want_O = torch.einsum("bhnm,bmhd->bnhd", Pm, v)
    P3 = torch.zeros(B, heads, n, 3 * npad, dtype=torch.bfloat16)
    hi = Pm.to(torch.bfloat16)
    P3[..., :n], P3[..., npad:npad + n], P3[..., 2 * npad:2 * npad + n] = hi, (Pm - hi.float()).to(torch.bfloat16), hi
    vT = v.permute(0, 2, 3, 1).reshape(B, heads * d, n)   # [B, C, n]: keys contiguous
    vT3 = torch.zeros(B, heads * d, 3 * npad, dtype=torch.bfloat16)
    vh = vT.to(torch.bfloat16)
    vT3[..., :n], vT3[..., npad:npad + n], vT3[..., 2 * npad:2 * npad + n] = vh, vh, (vT - vh.float()).to(torch.bfloat16)
    C_ = heads * d
    O = torch.zeros(B, n, C_)
    pa = P.TView(P3, (3 * npad, n, heads, B, 1), (1, 3 * npad, n * 3 * npad, heads * n * 3 * npad, B * heads * n * 3 * npad))
    vb = P.TView(vT3, (3 * npad, d, heads, B), (1, 3 * npad, d * 3 * npad, C_ * 3 * npad))
    oo = P.TView(O, (d, n, heads, B, 1), (1, C_, d, n * C_, B * n * C_))
    P.matrix_plan(pa, (128, 1, 1, 1), vb, oo, K=3 * npad, n_total=d, block_n=64, ext=(n, heads, B, 1),
                  tiles=(-(-n // 128), heads, B, 1), bz_sel=(3, 4), out_f32=True).simulate()
    assert rel(O.view(B, n, heads, d), want_O) < 2e-5
