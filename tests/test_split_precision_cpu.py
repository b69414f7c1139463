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
