"""GemmPlan tables executed by the CPU emulator (plan.simulate) vs torch convolutions: checks
taps, parity views, weight packing, virtual concat, folded 1x1 skip, residual, time-embedding
bias and the GroupNorm partial sums -- the addressing logic of the tensor-core kernel."""
import pytest
import torch
import torch.nn.functional as F

from mri_image_generation_b200 import plan as P

TOL = 4e-3  # bf16 output rounding


def nhwc(x):
    nd = x.dim() - 2
    return x.permute(0, *range(2, 2 + nd), 1).contiguous().to(torch.bfloat16)


def nchw(y):
    nd = y.dim() - 2
    return y.permute(0, nd + 1, *range(1, nd + 1)).float()


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


@pytest.mark.parametrize("nd,sp", [(3, (6, 8, 10)), (2, (12, 20)), (3, (4, 4, 6)), (3, (2, 16, 8)), (2, (16, 8))])
def test_conv_concat_skipfold_residual(nd, sp):
    torch.manual_seed(0)
    N, C1, C2, Co = 2, 64, 128, 64
    conv = F.conv3d if nd == 3 else F.conv2d
    a1, a2 = nhwc(torch.randn(N, C1, *sp)), nhwc(torch.randn(N, C2, *sp))
    w = torch.randn(Co, C1 + C2, *([3] * nd)) * 0.05
    ws = torch.randn(Co, C1, *([1] * nd)) * 0.1
    b = torch.randn(Co)
    rb = torch.randn(N, Co + 5)
    res = nhwc(torch.randn(N, Co, *sp))
    ref = conv(torch.cat([nchw(a1), nchw(a2)], 1), w.to(torch.bfloat16).float(), b, padding=1)
    ref = ref + conv(nchw(a1), ws.to(torch.bfloat16).float()) + rb[:, :Co].reshape(N, Co, *[1] * nd) + nchw(res)
    wm = P.pack_conv_weight(w, splits=[C1, C2], extra=[ws])
    y = torch.zeros(N, *sp, Co, dtype=torch.bfloat16)
    stats = torch.zeros(N, 8, 2, dtype=torch.float64)
    pl = P.conv_plan([P.ConvSource(a1), P.ConvSource(a2), P.ConvSource(a1, taps=False)], wm, y, 3,
                     bias=b, rowbias=rb, rowbias_ld=Co + 5, residual=res, stats=stats, stats_cpg=8)
    if sp[-1] == 8:  # boxes of 8 x 16 positions: the kw taps share one 10-wide activation tile
        assert pl.xreuse and pl.box[0] == 8 and pl.a_maps[0].box[1] == 10 and pl.a_maps_std[0].box[1] == 8
        lead = pl.ktable[0, :, 7]
        if pl.xreuse == 2:  # 8 x 16 boxes: the nine (kh, kw) taps share one 10 x 18 tile
            assert pl.a_maps[0].box[2] == 18
            assert lead.sum() == (pl.n_kb - 1) // 9 + 1     # one leader per (kd, slab) + the 1x1 slab
        else:
            assert lead.sum() == pl.n_kb - 2 * (pl.n_kb - 1) // 3  # one leader per (kd, kh, slab) + the 1x1 slab
    pl.simulate()
    assert rel(nchw(y), ref) < TOL
    r = ref.reshape(N, 8, -1)
    assert rel(stats[:, :, 0], r.sum(-1)) < 1e-4
    assert rel(stats[:, :, 1], (r ** 2).sum(-1)) < 1e-4


@pytest.mark.parametrize("nd,sp", [(3, (4, 8, 6)), (2, (12, 20))])
def test_down_and_up(nd, sp):
    torch.manual_seed(1)
    N, C, Co = 2, 64, 128
    a = nhwc(torch.randn(N, C, *sp))
    conv = F.conv3d if nd == 3 else F.conv2d
    convT = F.conv_transpose3d if nd == 3 else F.conv_transpose2d
    w = torch.randn(Co, C, *([4] * nd)) * 0.05
    b = torch.randn(Co)
    ref = conv(nchw(a), w.to(torch.bfloat16).float(), b, stride=2, padding=1)
    y = torch.zeros(N, *[s // 2 for s in sp], Co, dtype=torch.bfloat16)
    P.down_conv_plan(a, P.pack_conv_weight(w), y, bias=b).simulate()
    assert rel(nchw(y), ref) < TOL
    wt = torch.randn(C, Co, *([4] * nd)) * 0.05
    ref = convT(nchw(a), wt.to(torch.bfloat16).float(), b, stride=2, padding=1)
    y = torch.zeros(N, *[s * 2 for s in sp], Co, dtype=torch.bfloat16)
    st = torch.zeros(N, 8, 2, dtype=torch.float64)
    P.up_conv_plan(a, P.pack_convT_weight(wt), y, bias=b, stats=st, stats_cpg=16).simulate()
    assert rel(nchw(y), ref) < TOL
    assert rel(st[:, :, 0], ref.reshape(N, 8, -1).sum(-1)) < 1e-4


def test_choose_box():
    assert P.choose_box((40, 48, 40, 1))[3] == 1
    b = P.choose_box((40, 48, 40, 1))
    assert b[0] * b[1] * b[2] == 128
    assert P.choose_box((10, 12, 10, 4)) == (10, 12, 1, 1)
    b = P.choose_box((240, 240, 64, 1))
    assert b[0] * b[1] * b[2] * b[3] == 128 and 240 % b[0] == 0 and 240 % b[1] == 0
    b = P.choose_box((2, 2, 2, 2))
    assert b == (2, 2, 2, 2)


def test_schedule_choice_and_smem_budget():
    """Stream-K is chosen when whole tiles would under-fill the last wave of a 148-CTA grid; the
    library's stage count keeps ring + staging inside 227 KB of shared memory."""
    y = torch.zeros(1, 8, 8, 8, 128, dtype=torch.bfloat16)
    a = torch.zeros(1, 8, 8, 8, 64, dtype=torch.bfloat16)
    pl = P.conv_plan([P.ConvSource(a)], torch.zeros(128, 27 * 64, dtype=torch.bfloat16), y, 3)
    assert pl.pick_stages() == 0            # library decides
    assert pl.pick_swap() and pl.grid() == 2 and pl.pick_sched(148) == 1   # 2 box pairs x 27 k-steps
    pl.ktable = pl.ktable[:, :4]
    assert pl.pick_sched(148) == 0          # short K loops are never split
    from mri_image_generation_b200 import _lib
    lib = _lib.load()
    for bn in (16, 64, 128, 256):
        assert lib.mri_gemm_smem_bytes(bn, 0, 0) <= 227 * 1024
        assert lib.mri_gemm_smem_bytes(bn, 0, 2) == 2 * (128 * 128 + bn * 128) + 2 * 16384 + 1024
    assert lib.mri_gemm_smem_bytes(128, 1, 0) == 4 * 3 * 16384 + 2 * 16384 + 1024


def test_wgrad_runs_split_the_ktable_into_triples_and_singles():
    """WgradPlan.runs(): forward plans that share activation tiles list their taps as kw triples;
    whole runs of triples go to the kernel's xgroup mode, folded 1x1 slabs to the plain mode."""
    a1, a2 = torch.zeros(2, 16, 32, 8, 128, dtype=torch.bfloat16), torch.zeros(2, 16, 32, 8, 64, dtype=torch.bfloat16)
    wm = P.pack_conv_weight(torch.zeros(128, 192, 3, 3, 3), splits=[128, 64], extra=[torch.zeros(128, 64, 1, 1, 1)])
    y = torch.zeros(2, 16, 32, 8, 128, dtype=torch.bfloat16)
    pl = P.conv_plan([P.ConvSource(a1), P.ConvSource(a2), P.ConvSource(a2, taps=False)], wm, y, 3)
    assert pl.xreuse == 1 and tuple(pl.box) == (8, 16, 1, 1) and pl.tile_fast_dim == 2
    wg = P.WgradPlan(pl, torch.zeros_like(y), torch.zeros(1, 128, pl.n_kb * 64), 128)
    runs = wg.runs()
    assert runs == [(0, 27 * 3, True), (81, 1, False)]
    kt = pl.ktable[0]
    for i in range(0, 81, 3):   # a triple: same source / slab / (kd, kh), kw = -1, 0, +1, one leader
        assert list(kt[i:i + 3, 2]) == [-1, 0, 1] and list(kt[i:i + 3, 7]) == [1, 0, 0]
        assert (kt[i:i + 3, [0, 1, 3, 4]] == kt[i, [0, 1, 3, 4]]).all()
    # boxes that are not 8 x 16 positions, or plans without tile sharing: one plain launch
    y2 = torch.zeros(2, 10, 12, 10, 128, dtype=torch.bfloat16)
    a3 = torch.zeros(2, 10, 12, 10, 128, dtype=torch.bfloat16)
    pl2 = P.conv_plan([P.ConvSource(a3)], P.pack_conv_weight(torch.zeros(128, 128, 3, 3, 3)), y2, 3)
    wg2 = P.WgradPlan(pl2, torch.zeros_like(y2), torch.zeros(1, 128, pl2.n_kb * 64), 128)
    assert wg2.runs() == [(0, pl2.n_kb, False)]
    # the grid of every launch fills whole waves of the 148 SMs
    for base, mt in ((14, 4800), (9, 4800), (54, 600), (216, 75)):
        s = P.pick_wgrad_splits(base, mt)
        assert 1 <= s <= mt and (base * s) % 148 <= 148 and -(-(base * s) // 148) * 148 - base * s < 0.1 * base * s + 148
