"""tcgen05 implicit-GEMM kernel vs torch convolutions on the same bf16-rounded operands (GPU).

Tolerance: outputs are bf16 (8 mantissa bits) of an fp32-accumulated sum -> rel-L2 <= 4e-3
against the fp32 result of the same rounded operands; GroupNorm partial sums (fp32 atomics of
the un-rounded accumulators) rel <= 1e-3."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = 4e-3


def nhwc(x):
    nd = x.dim() - 2
    return x.permute(0, *range(2, 2 + nd), 1).contiguous().to(torch.bfloat16)


def nchw(y):
    nd = y.dim() - 2
    return y.permute(0, nd + 1, *range(1, nd + 1)).float()


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def conv_fn(nd):
    return F.conv3d if nd == 3 else F.conv2d


@pytest.fixture(scope="module")
def P():
    from mri_image_generation_b200 import _lib, plan
    _lib.require_device()
    return plan


CASES = [
    # nd, spatial, N, [Cin sources], Cout, k
    (3, (8, 12, 10), 2, [128], 128, 3),
    (3, (40, 48, 40), 1, [128], 128, 3),
    (3, (10, 12, 10), 2, [512], 512, 3),
    (3, (20, 24, 20), 1, [256, 256], 256, 3),
    (2, (32, 32), 4, [64], 64, 3),
    (2, (30, 30), 3, [192], 128, 3),
    (3, (10, 12, 10), 2, [512], 1024, 1),
    (3, (8, 8, 8), 1, [128], 16, 3),
    (3, (10, 12, 10), 8, [128], 256, 3),   # 128-row boxes spanning all 8 samples (2,4,2,8)
    (3, (6, 6, 6), 5, [64], 128, 3),        # odd box count per class, ragged boxes
    (2, (24, 24), 3, [64, 64], 384, 3),     # cpg = 48 (not a power of two)
]


@pytest.mark.parametrize("nd,sp,N,cins,cout,k,xr", [c + (None,) for c in CASES] + [
    (3, (12, 32, 16), 2, [128], 128, 3, 2),   # 10 x 18 halo tiles shared by the nine (kh, kw) taps
    (2, (48, 24), 3, [64, 64], 128, 3, 2),
    (3, (12, 32, 16), 2, [128], 128, 3, 0),   # no tap sharing at all
])
def test_conv_stride1(P, nd, sp, N, cins, cout, k, xr):
    torch.manual_seed(1)
    dev = "cuda"
    xs = [torch.randn(N, c, *sp, device=dev) for c in cins]
    acts = [nhwc(x) for x in xs]
    cin = sum(cins)
    w = torch.randn(cout, cin, *([k] * nd), device=dev) / (cin * k ** nd) ** 0.5
    bias = torch.randn(cout, device=dev)
    rb = torch.randn(N, cout + 8, device=dev)
    res = nhwc(torch.randn(N, cout, *sp, device=dev))
    xin = torch.cat([nchw(a) for a in acts], 1)
    ref = conv_fn(nd)(xin, w.to(torch.bfloat16).float(), bias, padding=k // 2)
    ref = ref + rb[:, :cout].reshape(N, cout, *[1] * nd) + nchw(res)
    wm = P.pack_conv_weight(w, splits=cins)
    y = torch.zeros(N, *sp, cout, dtype=torch.bfloat16, device=dev)
    groups = 8 if cout % 64 == 0 else 2
    stats = torch.zeros(N, groups, 2, device=dev, dtype=torch.float64)
    pl = P.conv_plan([P.ConvSource(a) for a in acts], wm, y, k, bias=bias, rowbias=rb,
                     rowbias_ld=cout + 8, residual=res, stats=stats, stats_cpg=cout // groups, xreuse=xr)
    if xr is not None:
        assert pl.xreuse == xr, (pl.xreuse, pl.box)
    pl.materialize(dev)
    pl.launch()
    torch.cuda.synchronize()
    assert rel(nchw(y), ref) < TOL
    r = ref.reshape(N, groups, -1)
    assert rel(stats[:, :, 0], r.sum(-1)) < 1e-3 or (stats[:, :, 0] - r.sum(-1)).abs().max() < 0.05 * r.abs().sum(-1).max() ** 0.5
    assert rel(stats[:, :, 1], (r ** 2).sum(-1)) < 1e-3


@pytest.mark.parametrize("nd,sp,N,cin,cout", [(3, (8, 12, 8), 2, 128, 256), (2, (32, 48), 2, 64, 64),
                                                (3, (40, 48, 40), 1, 128, 256)])
def test_down_conv(P, nd, sp, N, cin, cout):
    torch.manual_seed(2)
    dev = "cuda"
    a = nhwc(torch.randn(N, cin, *sp, device=dev))
    w = torch.randn(cout, cin, *([4] * nd), device=dev) / (cin * 4 ** nd) ** 0.5
    bias = torch.randn(cout, device=dev)
    ref = conv_fn(nd)(nchw(a), w.to(torch.bfloat16).float(), bias, stride=2, padding=1)
    y = torch.zeros(N, *[s // 2 for s in sp], cout, dtype=torch.bfloat16, device=dev)
    stats = torch.zeros(N, 8, 2, device=dev, dtype=torch.float64)
    pl = P.down_conv_plan(a, P.pack_conv_weight(w), y, bias=bias, stats=stats, stats_cpg=cout // 8)
    pl.materialize(dev)
    pl.launch()
    torch.cuda.synchronize()
    assert rel(nchw(y), ref) < TOL
    assert rel(stats[:, :, 1], (ref.reshape(N, 8, -1) ** 2).sum(-1)) < 1e-3


@pytest.mark.parametrize("nd,sp,N,cin,cout", [(3, (4, 6, 4), 2, 128, 64), (2, (16, 24), 2, 128, 64),
                                                (3, (20, 24, 20), 1, 256, 128)])
def test_up_conv(P, nd, sp, N, cin, cout):
    torch.manual_seed(3)
    dev = "cuda"
    a = nhwc(torch.randn(N, cin, *sp, device=dev))
    w = torch.randn(cin, cout, *([4] * nd), device=dev) / (cin * 2 ** nd) ** 0.5
    bias = torch.randn(cout, device=dev)
    ct = F.conv_transpose3d if nd == 3 else F.conv_transpose2d
    ref = ct(nchw(a), w.to(torch.bfloat16).float(), bias, stride=2, padding=1)
    y = torch.zeros(N, *[s * 2 for s in sp], cout, dtype=torch.bfloat16, device=dev)
    stats = torch.zeros(N, 8, 2, device=dev, dtype=torch.float64)
    pl = P.up_conv_plan(a, P.pack_convT_weight(w), y, bias=bias, stats=stats, stats_cpg=cout // 8)
    pl.materialize(dev)
    pl.launch()
    torch.cuda.synchronize()
    assert rel(nchw(y), ref) < TOL
    assert rel(stats[:, :, 1], (ref.reshape(N, 8, -1) ** 2).sum(-1)) < 1e-3


def test_matrix_fp32_out(P):
    """Batched Q K^T with fp32 output (the attention logits GEMM)."""
    torch.manual_seed(4)
    dev = "cuda"
    B, heads, n, d = 2, 4, 200, 128
    C = heads * d
    qk = torch.randn(B, n, 2 * C, device=dev).to(torch.bfloat16)
    npad = (n + 7) // 8 * 8
    S = torch.zeros(B, heads, n, npad, device=dev)
    a = P.TView(qk, (d, n, heads, B, 1), (1, 2 * C, d, n * 2 * C, B * n * 2 * C))
    b = P.TView(qk, (d, n, heads, B), (1, 2 * C, d, n * 2 * C), offset=C)
    o = P.TView(S, (npad, n, heads, B, 1), (1, npad, n * npad, heads * n * npad, B * heads * n * npad))
    pl = P.matrix_plan(a, (128, 1, 1, 1), b, o, K=d, n_total=npad, block_n=128,
                       ext=(n, heads, B, 1), tiles=(-(-n // 128), heads, B, 1), bz_sel=(3, 4),
                       out_f32=True)
    pl.materialize(dev)
    pl.launch()
    torch.cuda.synchronize()
    q = qk[:, :, :C].float().reshape(B, n, heads, d).permute(0, 2, 1, 3)
    k = qk[:, :, C:].float().reshape(B, n, heads, d).permute(0, 2, 1, 3)
    ref = q @ k.transpose(-1, -2)
    assert rel(S[..., :n], ref) < 1e-5


# ------------------------------------------------------------------------------------------
# backward: weight gradients (MN-major tcgen05 GEMM) and data gradients (fprop kernel on the
# adjoint plan) vs torch autograd on the same bf16-rounded operands
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nd,sp,N", [(3, (8, 12, 10), 2), (2, (30, 30), 3), (3, (10, 12, 10), 1),
                                     (3, (12, 32, 16), 2), (2, (48, 24), 3),   # 8 x 16 boxes: wgrad xgroup mode
                                     (3, (5, 16, 8), 1)])                        # ... with an odd number of boxes
def test_wgrad_and_dgrad_stride1_concat_skipfold(P, nd, sp, N):
    torch.manual_seed(5)
    dev = "cuda"
    C1, C2, Co = 64, 128, 128
    conv = conv_fn(nd)
    a1, a2 = nhwc(torch.randn(N, C1, *sp, device=dev)), nhwc(torch.randn(N, C2, *sp, device=dev))
    w = (torch.randn(Co, C1 + C2, *([3] * nd), device=dev) * 0.05).requires_grad_()
    ws = (torch.randn(Co, C1, *([1] * nd), device=dev) * 0.1).requires_grad_()
    x1, x2 = nchw(a1).requires_grad_(), nchw(a2).requires_grad_()
    out = conv(torch.cat([x1, x2], 1), w, None, padding=1) + conv(x1, ws)
    dy = nhwc(torch.randn(N, Co, *sp, device=dev))
    out.backward(nchw(dy))
    wm = P.pack_conv_weight(w.detach(), splits=[C1, C2], extra=[ws.detach()])
    y = torch.zeros(N, *sp, Co, dtype=torch.bfloat16, device=dev)
    fwd = P.conv_plan([P.ConvSource(a1), P.ConvSource(a2), P.ConvSource(a1, taps=False)], wm, y, 3)
    dw = torch.zeros(1, Co, wm.shape[1], device=dev)
    wg = P.WgradPlan(fwd, dy, dw, Co)
    wg.materialize(dev)
    wg.launch()
    torch.cuda.synchronize()
    g, ex = P.unpack_conv_wgrad(dw[0], w.shape, [C1, C2], [ws.shape])
    assert rel(g, w.grad) < 1e-4
    assert rel(ex[0], ws.grad) < 1e-4
    # data gradient of source 2 through the adjoint plan (flipped, transposed weights)
    wT = w.detach()[:, C1:].transpose(0, 1).flip(*range(2, 2 + nd)).to(torch.bfloat16).float()
    gx = torch.zeros(N, *sp, C2, dtype=torch.bfloat16, device=dev)
    pl = P.conv_plan([P.ConvSource(dy)], P.pack_conv_weight(wT), gx, 3)
    pl.materialize(dev)
    pl.launch()
    torch.cuda.synchronize()
    ref = conv(nchw(dy), wT, None, padding=1)
    assert rel(nchw(gx), ref) < TOL


@pytest.mark.parametrize("nd,sp,N", [(3, (8, 12, 8), 2), (2, (32, 48), 2)])
def test_wgrad_and_dgrad_strided(P, nd, sp, N):
    torch.manual_seed(6)
    dev = "cuda"
    C, Cd = 128, 256
    conv = conv_fn(nd)
    convT = F.conv_transpose3d if nd == 3 else F.conv_transpose2d
    a = nhwc(torch.randn(N, C, *sp, device=dev))
    # ---- down conv
    wd = (torch.randn(Cd, C, *([4] * nd), device=dev) * 0.05).to(torch.bfloat16).float().requires_grad_()
    xd = nchw(a).requires_grad_()
    od = conv(xd, wd, None, stride=2, padding=1)
    dyd = nhwc(torch.randn(*od.shape, device=dev))
    od.backward(nchw(dyd))
    yd = torch.zeros(N, *[s // 2 for s in sp], Cd, dtype=torch.bfloat16, device=dev)
    fwd = P.down_conv_plan(a, P.pack_conv_weight(wd.detach()), yd)
    dw = torch.zeros(1, Cd, fwd.n_kb * 64, device=dev)
    wg = P.WgradPlan(fwd, dyd, dw, Cd)
    wg.materialize(dev)
    wg.launch()
    gx = torch.zeros(N, *sp, C, dtype=torch.bfloat16, device=dev)
    pl = P.up_conv_plan(dyd, P.pack_convT_weight(wd.detach()), gx)
    pl.materialize(dev)
    pl.launch()
    torch.cuda.synchronize()
    assert rel(P.unpack_conv_wgrad(dw[0], wd.shape)[0], wd.grad) < 1e-4
    assert rel(nchw(gx), xd.grad) < TOL
    # ---- up conv
    wu = (torch.randn(C, Cd, *([4] * nd), device=dev) * 0.05).to(torch.bfloat16).float().requires_grad_()
    xu = nchw(a).requires_grad_()
    ou = convT(xu, wu, None, stride=2, padding=1)
    dyu = nhwc(torch.randn(*ou.shape, device=dev))
    ou.backward(nchw(dyu))
    yu = torch.zeros(N, *[s * 2 for s in sp], Cd, dtype=torch.bfloat16, device=dev)
    fwd = P.up_conv_plan(a, P.pack_convT_weight(wu.detach()), yu)
    dw = torch.zeros(fwd.n_class, Cd, fwd.n_kb * 64, device=dev)
    wg = P.WgradPlan(fwd, dyu, dw, Cd)
    wg.materialize(dev)
    wg.launch()
    gx = torch.zeros(N, *sp, C, dtype=torch.bfloat16, device=dev)
    res = nhwc(torch.randn(N, C, *sp, device=dev))
    pl = P.down_conv_plan(dyu, P.pack_conv_weight(wu.detach()), gx, residual=res)
    pl.materialize(dev)
    pl.launch()
    torch.cuda.synchronize()
    assert rel(P.unpack_convT_wgrad(dw, wu.shape), wu.grad) < 1e-4
    assert rel(nchw(gx), xu.grad + nchw(res)) < TOL


def test_box_order_does_not_change_results(P):
    """Depth-first box enumeration (tile_fast_dim = 2) only changes WHICH CTA computes a tile and
    when: outputs are bit-identical, statistics equal to fp64 rounding."""
    torch.manual_seed(11)
    dev = "cuda"
    N, C, sp = 2, 128, (12, 32, 16)
    a = nhwc(torch.randn(N, C, *sp, device=dev))
    w = torch.randn(C, C, 3, 3, 3, device=dev) / (C * 27) ** 0.5
    wm = P.pack_conv_weight(w)
    outs = []
    for order in (0, 2, 1):
        y = torch.zeros(N, *sp, C, dtype=torch.bfloat16, device=dev)
        stats = torch.zeros(N, 8, 2, device=dev, dtype=torch.float64)
        pl = P.conv_plan([P.ConvSource(a)], wm, y, 3, stats=stats, stats_cpg=C // 8)
        assert pl.tile_fast_dim == 2          # the default for 3-D 3x3x3 convolutions
        pl.tile_fast_dim = order
        pl.sched = 0   # whole K loops per CTA: with stream-K the split points depend on the tile index
        pl.materialize(dev)
        pl.launch()
        torch.cuda.synchronize()
        outs.append((y, stats))
    for y, st in outs[1:]:
        assert torch.equal(y, outs[0][0])
        # per-thread fp32 partial sums now span the (up to two) boxes of a tile before they enter the
        # fp64 accumulators, and which boxes share a tile depends on the box order: fp32 rounding
        assert torch.allclose(st, outs[0][1], rtol=2e-6, atol=1e-6)


@pytest.mark.parametrize("n,d,heads", [(1200, 128, 2), (700, 64, 3), (100, 128, 1)])
def test_attn_flash_kernel_with_growing_logits(n, d, heads):
    """mri_attn_flash_launch alone against fp32 torch attention.  The keys grow in norm along the
    sequence, so that the running row maximum keeps rising across key tiles and the lazy
    reference-maximum update (rescaling l and the O accumulator in TMEM) is exercised, not just
    the no-rescale fast path."""
    import ctypes as C
    from helpers import rel_l2
    from mri_image_generation_b200 import _lib, plan as PL
    B, Cc = 2, heads * d
    npad = (n + 7) // 8 * 8
    g = torch.Generator().manual_seed(n + d)
    q = torch.randn(B, n, Cc, generator=g)
    ramp = torch.linspace(0.2, 6.0, n).view(1, n, 1)        # later keys are much "louder"
    k = torch.randn(B, n, Cc, generator=g) * ramp
    v = torch.randn(B, n, Cc, generator=g)
    qkv = torch.cat([q, k, v], dim=-1).to(torch.bfloat16).cuda().contiguous()     # [B, n, 3C]
    vT = torch.zeros(B, Cc, npad, dtype=torch.bfloat16, device="cuda")
    vT[:, :, :n] = qkv[:, :, 2 * Cc:].transpose(1, 2)
    out = torch.zeros(B, n, Cc, dtype=torch.bfloat16, device="cuda")
    maps = PL.encode_maps([
        PL.MapSpec(PL.TView(qkv, (3 * Cc, n, B), (1, 3 * Cc, n * 3 * Cc)), (64, 128, 1), 3),
        PL.MapSpec(PL.TView(vT, (npad, Cc, B), (1, npad, Cc * npad)), (64, d, 1), 3)], "cuda")
    a = _lib.MriAttnArgs()
    a.qk_map, a.vt_map, a.out = maps.data_ptr(), maps.data_ptr() + 128, out.data_ptr()
    a.batch, a.heads, a.n, a.d, a.C = B, heads, n, d, Cc
    a.k_col0, a.v_row0, a.ld_out = Cc, 0, Cc
    a.scale = float(d) ** -0.5
    _lib.check(_lib.load().mri_attn_flash_launch(C.byref(a), _lib.current_stream_ptr()), "attn")
    torch.cuda.synchronize()
    qf, kf, vf = (t.float().view(B, n, heads, d).transpose(1, 2) for t in
                  (qkv[..., :Cc], qkv[..., Cc:2 * Cc], qkv[..., 2 * Cc:]))
    logits = (qf @ kf.transpose(-1, -2)) * a.scale
    # the row maximum really does climb by far more than 8 / log2(e) between the first and last tile
    if n > 256:
        assert (logits[..., 128:].amax(-1) - logits[..., :128].amax(-1)).median().item() > 8.0
    want = (torch.softmax(logits, -1) @ vf).transpose(1, 2).reshape(B, n, Cc)
    err = rel_l2(out, want)
    print(f"attn_flash n={n} d={d} heads={heads}: rel-L2 {err:.3e}")
    assert err < 1e-2, err


@pytest.mark.parametrize("sp,cin,cout", [((8, 12, 8), 3, 128), ((5, 7, 9), 3, 64), ((40, 48, 40), 3, 128),
                                         ((37, 50), 1, 64), ((128, 128), 1, 64), ((16, 24), 4, 128)])
def test_thin_in_conv_kernel(sp, cin, cout):
    """mri_thin_in_conv (patch matrix built in shared memory, one tcgen05 launch) against torch's
    convolution on the same bf16-rounded operands, and its GroupNorm partial sums against the
    sums of the stored output; spatial sizes that are not multiples of the 128-position tile."""
    from helpers import rel_l2
    from mri_image_generation_b200 import ops, plan as PL
    nd = len(sp)
    B = 2
    g = torch.Generator().manual_seed(sum(sp) + cin + cout)
    x = torch.randn(B, cin, *sp, generator=g).cuda()
    w = (torch.randn(cout, cin, *([3] * nd), generator=g) * 0.2).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    wp = torch.zeros(cout, 4, *([3] * nd), device="cuda")
    wp[:, :cin] = w
    packed = PL.pack_conv_weight(wp)
    w128 = torch.zeros(cout, 128, dtype=torch.bfloat16, device="cuda")
    w128[:, :packed.shape[1]] = packed
    S = 1
    for e in sp:
        S *= e
    y = torch.zeros(B, S, cout, dtype=torch.bfloat16, device="cuda")
    stats = torch.zeros(B, 8, 2, dtype=torch.float64, device="cuda")
    sp3 = (1,) * (3 - nd) + tuple(sp)
    ops.thin_in_conv(x, w128, bias, y, stats, B, cin, sp3[0], sp3[1], sp3[2], nd, cout)
    torch.cuda.synchronize()
    conv = F.conv3d if nd == 3 else F.conv2d
    want = conv(x.bfloat16().float(), w.bfloat16().float(), bias, padding=1)       # [B, cout, *sp]
    want = want.reshape(B, cout, S).transpose(1, 2)
    err = rel_l2(y, want)
    yf = y.double().view(B, S, 8, cout // 8)
    # the sums are taken over the fp32 values BEFORE the bf16 rounding of the stored output
    wf = want.double().reshape(B, S, 8, cout // 8)
    assert torch.allclose(stats[..., 0], wf.sum((1, 3)), rtol=1e-4, atol=1e-2), (stats[0, :, 0], wf.sum((1, 3))[0])
    assert torch.allclose(stats[..., 1], (wf * wf).sum((1, 3)), rtol=1e-4, atol=1e-2)
    print(f"thin_in_conv {sp} {cin}->{cout}: rel-L2 {err:.3e}")
    assert err < 5e-3, err
    del yf
