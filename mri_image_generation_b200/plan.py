"""Host-side description of one tensor-core implicit-GEMM launch (MriGemmArgs, include/mri_b200.h).

A :class:`GemmPlan` is plain host data: tensor views, TMA box shapes and the k-table that
turns a convolution (stride 1, stride 2, transposed, channel-concatenated inputs, folded 1x1
skip) into "load a shifted box of 64 channels, multiply by a 64-wide slab of packed weights".
`materialize()` encodes the TMA descriptors and uploads the tables (needs the GPU);
`simulate()` executes exactly the same tables with torch on the CPU and exists so that the
addressing logic (taps, parity views, weight packing, clipping) is testable without a GPU --
it is test infrastructure, never part of the product path.

Reference semantics being reproduced: nn.Conv2d/3d, nn.ConvTranspose2d/3d and torch.cat as used
in model_scripts/*/unet.py and ddpm_3d_ldm/unet_attention.py (file:line in include/mri_b200.h).
"""
from __future__ import annotations

import ctypes as C
import itertools
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

BLOCK_M = 128
BLOCK_K = 64


def _esize(dt: torch.dtype) -> int:
    return {torch.bfloat16: 2, torch.float32: 4}[dt]


@dataclass
class TView:
    """A strided view (inner dimension first, element strides) over a flat torch tensor."""
    base: torch.Tensor
    dims: Tuple[int, ...]
    strides: Tuple[int, ...]
    offset: int = 0

    def __post_init__(self):
        assert len(self.dims) == len(self.strides) and self.strides[0] == 1
        assert self.base.is_contiguous()

    @property
    def esize(self) -> int:
        return _esize(self.base.dtype)

    def ptr(self) -> int:
        return self.base.data_ptr() + self.offset * self.esize


@dataclass
class MapSpec:
    view: TView
    box: Tuple[int, ...]
    swizzle: int  # 0 none, 1 32B, 2 64B, 3 128B


def pad_rank(view: TView, box: Sequence[int], rank: int) -> Tuple[TView, Tuple[int, ...]]:
    """Append unit dims so that the view/box have exactly `rank` dims."""
    dims, strides, box = list(view.dims), list(view.strides), list(box)
    while len(dims) < rank:
        # a unit dim's stride is irrelevant but must be a multiple of 16 bytes: reuse the
        # largest extent so far
        big = max(d * s for d, s in zip(dims, strides))
        big = (big + 7) // 8 * 8
        dims.append(1)
        strides.append(big)
        box.append(1)
    return TView(view.base, tuple(dims), tuple(strides), view.offset), tuple(box)


def encode_maps(specs: Sequence[MapSpec], device) -> torch.Tensor:
    """Encode CUtensorMaps on the host and upload them (uint8 [n, 128], 128-byte aligned rows)."""
    lib = _lib.load()
    host = np.zeros((len(specs), 128), dtype=np.uint8)
    for i, sp in enumerate(specs):
        v = sp.view
        rank = len(v.dims)
        es = v.esize
        dims = (C.c_uint64 * rank)(*v.dims)
        strides = (C.c_uint64 * max(rank - 1, 1))(*[s * es for s in v.strides[1:]] or [0])
        box = (C.c_uint32 * rank)(*sp.box)
        buf = (C.c_uint8 * 128)()
        rc = lib.mri_tmap_encode(C.addressof(buf), v.ptr(), 0 if v.base.dtype == torch.bfloat16 else 1,
                                 rank, dims, strides, box, sp.swizzle)
        _lib.check(rc, "mri_tmap_encode")
        host[i] = np.frombuffer(buf, dtype=np.uint8)
    return torch.from_numpy(host).to(device)


class _Workspace:
    """Stream-K scratch of one device: fp32 partial tiles + self-resetting flags."""

    def __init__(self, device):
        n = C.c_int(0)
        nbytes = _lib.load().mri_gemm_workspace_bytes(C.byref(n))
        if nbytes <= 0:
            _lib.check(nbytes, "mri_gemm_workspace_bytes")
        self.n_ctas = n.value
        self.buf = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        self.partials_ptr = self.buf.data_ptr()
        self.flags_ptr = self.buf.data_ptr() + nbytes - 1024


_WORKSPACES = {}


def gemm_workspace(device) -> _Workspace:
    dev = torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    ws = _WORKSPACES.get(idx)
    if ws is None:
        with torch.cuda.device(idx):
            ws = _Workspace(torch.device("cuda", idx))
        _WORKSPACES[idx] = ws
    return ws


def choose_box(ext: Sequence[int], prefer_unit: Sequence[int] = ()) -> Tuple[int, int, int, int]:
    """Pick box extents (product <= 128) over up to four output dims maximising tile fill.

    `prefer_unit`: indices of dims (e.g. the sample dim) that should stay 1 unless that costs
    efficiency (keeps the GroupNorm-statistics epilogue on its warp-reduced fast path)."""
    ext = list(ext) + [1] * (4 - len(ext))

    def cands(e):
        c = {d for d in range(1, min(e, BLOCK_M) + 1) if e % d == 0}
        c |= {p for p in (1, 2, 4, 8, 16, 32, 64, 128) if p <= e}
        c.add(min(e, BLOCK_M))
        return sorted(c)

    best, best_key = None, None
    total = ext[0] * ext[1] * ext[2] * ext[3]
    for b in itertools.product(*[cands(e) for e in ext]):
        rows = b[0] * b[1] * b[2] * b[3]
        if rows > BLOCK_M:
            continue
        tiles = 1
        for e, bb in zip(ext, b):
            tiles *= -(-e // bb)
        eff = total / (tiles * BLOCK_M)
        unit_pen = sum(1 for i in prefer_unit if b[i] != 1)
        # a box that spans several samples is only cheap for the statistics epilogue when every
        # 16-position chunk of it lies inside one sample (positions of a sample are contiguous
        # in the box): otherwise demand a clear efficiency advantage
        score = eff
        for i in prefer_unit:
            if b[i] != 1 and (rows // b[i]) % 16 != 0:
                score = eff * 0.75
        key = (round(score, 6), -unit_pen, b[0] == 8, b[0], b[1], b[2])
        if best_key is None or key > best_key:
            best, best_key = b, key
    return best


@dataclass
class GemmPlan:
    a_maps: List[MapSpec]
    b_map: Optional[MapSpec]
    o_maps: List[MapSpec]
    ktable: np.ndarray  # int32 [n_class, n_kb, 8]
    tiles: Tuple[int, int, int, int]
    box: Tuple[int, int, int, int]
    ext: Tuple[int, int, int, int]
    block_n: int
    n_total: int
    r_maps: Optional[List[MapSpec]] = None
    bz_sel: Tuple[int, int] = (0, 0)
    sample_dim: int = 0
    out_f32: bool = False
    bias: Optional[torch.Tensor] = None       # fp32 [>= n_tiles_n*block_n]
    bias_m: Optional[torch.Tensor] = None     # fp32 [ext[0]]
    rowbias: Optional[torch.Tensor] = None    # fp32 view [samples, >= n]; row stride = rowbias_ld
    rowbias_ld: int = 0
    stats: Optional[torch.Tensor] = None      # fp64 [samples, stats_ld, 2]
    stats_ld: int = 0
    stats_cpg: int = 0
    stages: int = 0
    sched: Optional[int] = None
    swap_ab: Optional[bool] = None
    tile_fast_dim: int = 0                     # boxes are enumerated along this x dim first (see mri_b200.h)
    xreuse: int = 0                            # 1: three kw taps share a 10-wide tile; 2: nine (kh, kw) taps share a 10 x 18 tile
    a_maps_std: Optional[List[MapSpec]] = None  # xreuse: the same inputs with the standard boxes
    trace: Optional[torch.Tensor] = None     # int64 [grid, 8]: per-CTA timestamps (profiling only)
    name: str = ""
    flops: int = 0
    _args: Optional[_lib.MriGemmArgs] = field(default=None, repr=False)
    _keep: list = field(default_factory=list, repr=False)

    @property
    def n_class(self) -> int:
        return self.ktable.shape[0]

    @property
    def n_kb(self) -> int:
        return self.ktable.shape[1]

    @property
    def n_tiles_n(self) -> int:
        return -(-self.n_total // self.block_n)

    def pick_swap(self) -> bool:
        """Weights as the M = 128 MMA operand, two boxes of positions as N = 256 (gemm_tc.cu):
        every bf16 convolution whose output channels fill whole 128-channel tiles."""
        if self.swap_ab is not None:
            return self.swap_ab
        env = os.environ.get("MRI_GEMM_SWAP")  # tuning experiments only
        if env:
            return bool(int(env)) and self.block_n == 128 and self.n_total % 64 == 0
        return (self.block_n == 128 and self.n_total % 64 == 0 and not self.out_f32
                and self.bias_m is None and max(self.bz_sel) <= 1)

    def grid(self) -> int:
        """Number of tiles (work items of n_kb k-steps each)."""
        t = self.tiles
        boxes = t[0] * t[1] * t[2] * t[3]
        groups = -(-boxes // 2) if self.pick_swap() else boxes
        return groups * self.n_tiles_n * self.n_class

    def pick_stages(self) -> int:
        """Requested TMA ring depth: 0 lets the library take what shared memory allows."""
        if self.stages:
            return self.stages
        env = os.environ.get("MRI_GEMM_STAGES")  # tuning experiments only
        return int(env) if env else 0

    def pick_sched(self, n_sms: int) -> int:
        """1 = stream-K (split K loops across neighbouring CTAs) when whole tiles would leave
        the last wave of the persistent grid under-filled; 0 = tile-aligned ranges."""
        if self.sched is not None:
            return self.sched
        env = os.environ.get("MRI_GEMM_SCHED")  # tuning experiments only
        if env:
            return int(env)
        tiles = self.grid()
        if self.n_kb < 8:
            return 0
        waves = tiles / n_sms
        return 1 if -(-tiles // n_sms) / waves > 1.02 else 0

    # ------------------------------------------------------------------ GPU
    def materialize(self, device) -> None:
        maps = list(self.a_maps) + [self.b_map] + list(self.o_maps) + list(self.r_maps or [])
        blob = encode_maps(maps, device)
        na, no = len(self.a_maps), len(self.o_maps)
        base = blob.data_ptr()
        kt = torch.from_numpy(np.ascontiguousarray(self.ktable, dtype=np.int32)).to(device)
        self._keep = [blob, kt]
        a = _lib.MriGemmArgs()
        a.a_maps = base
        a.b_map = base + 128 * na
        a.o_maps = base + 128 * (na + 1)
        if self.r_maps:
            # the residual is read with plain loads: base pointer + per-class offset + strides
            if len(self.r_maps) > 8:
                raise _lib.MriError("residual input supports at most 8 output classes")
            v0 = self.r_maps[0].view
            assert v0.base.dtype == torch.bfloat16
            a.r_maps = base + 128 * (na + 1 + no)
            a.r_base = v0.base.data_ptr()
            for ci, m in enumerate(self.r_maps):
                assert m.view.base is v0.base and m.view.strides == v0.strides
                a.r_cls_off[ci] = m.view.offset
            for i in range(4):
                a.r_stride[i] = v0.strides[1 + i]
        else:
            a.r_maps = None
            a.r_base = None
        a.ktable = kt.data_ptr()
        a.n_kb, a.n_class = self.n_kb, self.n_class
        for i in range(4):
            a.tiles[i], a.box[i], a.ext[i] = self.tiles[i], self.box[i], self.ext[i]
        a.n_tiles_n, a.block_n, a.n_total = self.n_tiles_n, self.block_n, self.n_total
        a.bz_sel[0], a.bz_sel[1] = self.bz_sel
        a.sample_dim = self.sample_dim
        a.out_f32 = 1 if self.out_f32 else 0
        a.bias = self.bias.data_ptr() if self.bias is not None else None
        a.bias_m = self.bias_m.data_ptr() if self.bias_m is not None else None
        a.rowbias = self.rowbias.data_ptr() if self.rowbias is not None else None
        a.rowbias_ld = self.rowbias_ld
        a.stats = self.stats.data_ptr() if self.stats is not None else None
        a.stats_ld, a.stats_cpg = self.stats_ld, self.stats_cpg
        a.stages = self.pick_stages()
        ws = gemm_workspace(device)
        a.sk_partials, a.sk_flags, a.sk_ctas = ws.partials_ptr, ws.flags_ptr, ws.n_ctas
        a.sched = self.pick_sched(ws.n_ctas)
        a.swap_ab = 1 if self.pick_swap() else 0
        a.xreuse = int(self.xreuse)
        a.tile_fast_dim = int(self.tile_fast_dim)
        if self.xreuse and not a.swap_ab:
            raise _lib.MriError("xreuse plans need the swap_ab tile shape")
        a.trace = self.trace.data_ptr() if self.trace is not None else None
        self._args = a

    def launch(self, stream: Optional[int] = None) -> None:
        if self._args is None:
            raise _lib.MriError("GemmPlan.launch before materialize()")
        rc = _lib.load().mri_gemm_launch(C.byref(self._args),
                                         stream if stream is not None else _lib.current_stream_ptr())
        _lib.check(rc, f"mri_gemm_launch[{self.name}]")

    # ------------------------------------------------------------------ CPU emulation (tests)
    def simulate(self) -> None:
        """Execute the plan with torch on the CPU (same tables, TMA zero-fill / clipping rules)."""
        nt_n = self.n_tiles_n
        rows_in_box = int(np.prod(self.box))
        for cls in range(self.n_class):
            for tix in itertools.product(*[range(t) for t in reversed(self.tiles)]):
                tix = tuple(reversed(tix))
                org = [tix[i] * self.box[i] for i in range(4)]
                # row validity / sample per row
                rl = _row_coords(self.box)
                valid = torch.ones(rows_in_box, dtype=torch.bool)
                for i in range(4):
                    valid &= (org[i] + rl[i]) < self.ext[i]
                sample = (org[self.sample_dim - 1] + rl[self.sample_dim - 1]) if self.sample_dim else \
                    torch.zeros(rows_in_box, dtype=torch.long)
                for nt in range(nt_n):
                    n0 = nt * self.block_n
                    acc = torch.zeros(rows_in_box, self.block_n, dtype=torch.float32)
                    sel = lambda s: 0 if s == 0 else (cls if s == 1 else tix[s - 2])
                    bz = (sel(self.bz_sel[0]), sel(self.bz_sel[1]))
                    for kb in range(self.n_kb):
                        e = self.ktable[cls, kb]
                        am = self.a_maps[int(e[0])]
                        coords = [int(e[1])] + [org[i] + int(e[2 + i]) for i in range(4)]
                        a = _load_box(am.view, coords, (BLOCK_K,) + tuple(self.box)).reshape(rows_in_box, BLOCK_K)
                        b = _load_box(self.b_map.view, [int(e[6]), n0, bz[0], bz[1]],
                                      self.b_map.box).reshape(self.block_n, BLOCK_K)
                        acc += a @ b.t()
                    ncols = torch.arange(n0, n0 + self.block_n)
                    colok = ncols < self.n_total
                    if self.bias is not None:
                        acc += torch.where(colok, self.bias[ncols.clamp(max=self.bias.numel() - 1)],
                                           torch.zeros(()))[None, :]
                    if self.bias_m is not None:
                        x1 = (org[0] + rl[0]).clamp(max=self.bias_m.numel() - 1)
                        acc += torch.where(valid, self.bias_m[x1], torch.zeros(()))[:, None]
                    if self.rowbias is not None:
                        if self.rowbias_ld == 0:      # one row for every sample (a second bias vector)
                            rb = self.rowbias.reshape(1, -1)
                        else:
                            rb = self.rowbias.reshape(-1, self.rowbias_ld) if self.rowbias.dim() == 1 \
                                else self.rowbias
                        smp = sample.clamp(max=rb.shape[0] - 1)
                        add = rb[smp][:, ncols.clamp(max=rb.shape[1] - 1)]
                        acc += torch.where(valid[:, None] & colok[None, :], add, torch.zeros(()))
                    om = self.o_maps[cls]
                    ocoords = [n0] + org
                    if self.r_maps is not None:
                        rm = self.r_maps[cls]
                        res = _load_box(rm.view, ocoords, _full_box(rm, self.block_n))
                        acc += res.reshape(rows_in_box, self.block_n)
                    if self.stats is not None:
                        for g0 in range(0, self.block_n, self.stats_cpg):
                            if n0 + g0 >= self.n_total:
                                break
                            blk = acc[:, g0:g0 + self.stats_cpg]
                            g = (n0 + g0) // self.stats_cpg
                            for smp in torch.unique(sample[valid]).tolist():
                                m = valid & (sample == smp)
                                self.stats[smp, g, 0] += blk[m].sum()
                                self.stats[smp, g, 1] += (blk[m] ** 2).sum()
                    _store_box(om.view, ocoords, _full_box(om, self.block_n),
                               acc.reshape(*reversed(self.box), self.block_n), self.n_total)


def _full_box(m: MapSpec, block_n: int) -> Tuple[int, ...]:
    # the kernel issues block_n/chunk stores of box[0] columns each; emulate them as one box
    return (block_n,) + tuple(m.box[1:])


def _row_coords(box):
    r = torch.arange(int(np.prod(box)))
    out = []
    for b in box:
        out.append(r % b)
        r = r // b
    return out


def _box_index(view: TView, coords, box):
    rank = len(view.dims)
    idx = None
    mask = None
    shape_tail = []
    for i in range(rank):
        ar = coords[i] + torch.arange(box[i])
        ok = (ar >= 0) & (ar < view.dims[i])
        shp = [1] * rank
        shp[rank - 1 - i] = box[i]
        term = (ar.clamp(0, view.dims[i] - 1) * view.strides[i]).reshape(shp)
        okr = ok.reshape(shp)
        idx = term if idx is None else idx + term
        mask = okr if mask is None else mask & okr
    return idx + view.offset, mask


def _load_box(view: TView, coords, box) -> torch.Tensor:
    idx, mask = _box_index(view, coords, box)
    flat = view.base.reshape(-1)
    vals = flat[idx.reshape(-1)].float().reshape(idx.shape)
    return torch.where(mask, vals, torch.zeros(()))


def _store_box(view: TView, coords, box, vals: torch.Tensor, n_total: int) -> None:
    idx, mask = _box_index(view, coords, box)
    # columns at or beyond n_total are not stored by the kernel when a whole chunk is outside;
    # inside a chunk TMA clips against the tensor extent (dims[0])
    flat = view.base.reshape(-1)
    sel = mask.reshape(-1)
    flat[idx.reshape(-1)[sel]] = vals.reshape(-1)[sel].to(flat.dtype)


# ======================================================================================
# Convolution -> GemmPlan builders.  Activations are channels-last bf16 tensors of shape
# [N, *spatial, C] (spatial = (H, W) or (D, H, W)).
# ======================================================================================

def _act_view(x: torch.Tensor, ndim: int) -> TView:
    """[N, (D,) H, W, C] contiguous -> rank-5 view (C, W, H, D|N, N|1)."""
    shp = list(x.shape)
    C_ = shp[-1]
    sp = shp[1:-1]
    assert len(sp) == ndim
    dims, strides = [C_], [1]
    s = C_
    for e in reversed(sp):
        dims.append(e)
        strides.append(s)
        s *= e
    dims.append(shp[0])
    strides.append(s)
    v = TView(x, tuple(dims), tuple(strides))
    v, _ = pad_rank(v, [1] * len(dims), 5)
    return v


def _parity_view(x: torch.Tensor, ndim: int, par: Sequence[int]) -> TView:
    """Every second position along each spatial dim, starting at par (w first): the stride-2
    sub-lattice of a channels-last tensor, as a rank-5 view."""
    shp = list(x.shape)
    C_ = shp[-1]
    sp = shp[1:-1]
    dims, strides = [C_], [1]
    s = C_
    off = 0
    for i, e in enumerate(reversed(sp)):  # w, h, (d)
        assert e % 2 == 0, "stride-2 paths need even spatial sizes"
        dims.append(e // 2)
        strides.append(2 * s)
        off += par[i] * s
        s *= e
    dims.append(shp[0])
    strides.append(s)
    v = TView(x, tuple(dims), tuple(strides), off)
    v, _ = pad_rank(v, [1] * len(dims), 5)
    return v


def _out_swizzle(chunk_bytes: int) -> int:
    return 3 if chunk_bytes == 128 else 0


def _out_chunk(block_n: int, out_f32: bool) -> int:
    full = 32 if out_f32 else 64
    return min(block_n, full)


def pick_block_n(cout_pad: int) -> int:
    if cout_pad % 64 == 0:
        return 128  # Cout = 64 rides in half of a 128-channel tile (zero-filled weight rows)
    for bn in (32, 16):
        if cout_pad % bn == 0:
            return bn
    raise ValueError(f"Cout {cout_pad} must be a multiple of 16")


@dataclass
class ConvSource:
    """One input of a convolution: tensor [N, *sp, C]; `taps` False = centre tap only
    (a folded 1x1 skip convolution riding on a k^d convolution's K loop)."""
    x: torch.Tensor
    taps: bool = True


def _spatial_ext(y: torch.Tensor, ndim: int):
    sp = list(y.shape[1:-1])
    ext = list(reversed(sp)) + [y.shape[0]]
    sample_dim = ndim + 1
    while len(ext) < 4:
        ext.append(1)
    return tuple(ext), sample_dim


def conv_plan(sources: Sequence[ConvSource], wmat: torch.Tensor, y: torch.Tensor, ksize: int,
              *, bias=None, rowbias=None, rowbias_ld=0, residual: Optional[torch.Tensor] = None,
              stats=None, stats_cpg=0, block_n: Optional[int] = None, stages=0, name="",
              xreuse=None, out_f32: bool = False) -> GemmPlan:
    """Stride-1, pad k//2 convolution over the channel-concatenation of `sources`.

    wmat: packed weights [Cout_pad, K] bf16, K = sum over sources of (taps * C_i), see
    pack_conv_weight().  y: [N, *sp, Cout_pad] bf16 (fp32 with out_f32: the split precision mode,
    plain tiles only)."""
    assert y.dtype == (torch.float32 if out_f32 else torch.bfloat16)
    assert not (out_f32 and residual is not None)
    ndim = y.dim() - 2
    ext, sample_dim = _spatial_ext(y, ndim)
    box = choose_box(ext, prefer_unit=(sample_dim - 1,))
    cout_pad = y.shape[-1]
    bn = block_n or pick_block_n(cout_pad)
    p = ksize // 2
    # "xreuse": with boxes of 8 x 16 positions the three kw taps of a (kd, kh, channel slab) read
    # ONE activation tile loaded 10 positions wide -> a third of the activation traffic from L2
    env_x = os.environ.get("MRI_GEMM_XREUSE")   # tuning experiments: 0 / 1 / 2
    # default 1.  Level 2 (10 x 18 tile, nine taps) is correct but SLOWER on B200 (1165 vs 1206
    # TFLOP/s at the cfg4 top level): the second box of an 18-line tile breaks the uniform group
    # stride of a 256-row operand, so every k-step needs two N = 128 instructions, which read the
    # weight operand twice -- 128 B/clk of shared-memory operand traffic, the whole budget.
    if xreuse is None:
        level = int(env_x) if env_x in ("0", "1", "2") else 1
    elif isinstance(xreuse, bool):
        level = 1 if xreuse else 0
    else:
        level = int(xreuse)
    use_x = bool(level >= 1 and ksize == 3 and bn == 128 and cout_pad % 64 == 0 and box[0] == 8
                 and int(np.prod(box)) == BLOCK_M and os.environ.get("MRI_GEMM_SWAP", "1") != "0"
                 and not out_f32)
    # level 2: boxes of exactly 8 x 16 x 1 x 1 positions load a 10 x 18 tile once per (kd, channel
    # slab) and all nine (kh, kw) taps read shifted views of it
    use_xy = bool(use_x and level >= 2 and box[1] == 16)
    a_maps, a_std, rows = [], [], []
    bk = 0
    for si, src in enumerate(sources):
        Ci = src.x.shape[-1]
        assert Ci % BLOCK_K == 0, f"source channels {Ci} must be a multiple of 64"
        assert list(src.x.shape[:-1]) == list(y.shape[:-1])
        a_std.append(MapSpec(_act_view(src.x, ndim), (BLOCK_K,) + box, 3))
        xbox = (BLOCK_K, 10, 18) + tuple(box[2:]) if use_xy else (BLOCK_K, 10) + tuple(box[1:])
        a_maps.append(MapSpec(_act_view(src.x, ndim), xbox, 3) if use_x else a_std[-1])
        if not src.taps:  # centre tap only (folded 1x1 skip): its own group
            for c0 in range(0, Ci, BLOCK_K):
                rows.append([si, c0, 0, 0, 0, 0, bk, 1])
                bk += BLOCK_K
        elif use_xy:
            for ot in itertools.product(range(ksize), repeat=ndim - 2):  # (kd,) / ()
                for c0 in range(0, Ci, BLOCK_K):
                    for kh in range(ksize):
                        for kw in range(ksize):
                            tap = tuple(ot) + (kh, kw)
                            tap_index = 0
                            for t_ in tap:
                                tap_index = tap_index * ksize + t_
                            offs = [k - p for k in reversed(tap)] + [0] * (4 - ndim)
                            rows.append([si, c0] + offs + [bk + tap_index * Ci + c0,
                                                           1 if (kh == 0 and kw == 0) else 0])
            bk += (ksize ** ndim) * Ci
        elif use_x:
            for ot in itertools.product(range(ksize), repeat=ndim - 1):  # (kd, kh) / (kh,)
                for c0 in range(0, Ci, BLOCK_K):
                    for kw in range(ksize):
                        tap = tuple(ot) + (kw,)
                        tap_index = 0
                        for t_ in tap:
                            tap_index = tap_index * ksize + t_
                        offs = [k - p for k in reversed(tap)] + [0] * (4 - ndim)
                        rows.append([si, c0] + offs + [bk + tap_index * Ci + c0, 1 if kw == 0 else 0])
            bk += (ksize ** ndim) * Ci
        else:
            for tap in itertools.product(range(ksize), repeat=ndim):
                # tap = (kd, kh, kw) / (kh, kw); x1 = w
                offs = [k - p for k in reversed(tap)] + [0] * (4 - ndim)
                for c0 in range(0, Ci, BLOCK_K):
                    rows.append([si, c0] + offs + [bk, 0])
                    bk += BLOCK_K
    assert wmat.shape == (cout_pad, bk), (wmat.shape, cout_pad, bk)
    kt = np.asarray(rows, dtype=np.int32)[None]
    bview = TView(wmat, (bk, cout_pad, 1, 1), (1, bk, _rup8(bk * cout_pad), _rup8(bk * cout_pad)))
    chunk = _out_chunk(bn, out_f32)
    omap = MapSpec(_act_view(y, ndim), (chunk,) + box, _out_swizzle(chunk * (4 if out_f32 else 2)))
    rmaps = None
    if residual is not None:
        assert residual.shape == y.shape
        rmaps = [MapSpec(_act_view(residual, ndim), (chunk,) + box, _out_swizzle(chunk * 2))]
    tiles = tuple(-(-e // b) for e, b in zip(ext, box))
    m_rows = int(np.prod(ext))
    return GemmPlan(a_maps=a_maps, b_map=MapSpec(bview, (BLOCK_K, bn, 1, 1), 3), o_maps=[omap],
                    r_maps=rmaps, ktable=kt, tiles=tiles, box=box, ext=ext, block_n=bn, out_f32=out_f32,
                    n_total=cout_pad, sample_dim=sample_dim, bias=bias, rowbias=rowbias,
                    rowbias_ld=rowbias_ld, stats=stats, stats_ld=(stats.shape[1] if stats is not None else 0),
                    stats_cpg=stats_cpg, stages=stages, name=name, flops=2 * m_rows * cout_pad * bk,
                    xreuse=(2 if use_xy else 1) if use_x else 0, a_maps_std=a_std if use_x else None,
                    tile_fast_dim=_depth_first(ndim, ksize, tiles))


def _depth_first(ndim: int, ksize: int, tiles) -> int:
    """3-D convolutions with a depth extent in their filter walk the boxes depth-first (x3)."""
    env = os.environ.get("MRI_GEMM_TILE_ORDER")  # tuning experiments: force a dim
    if env is not None:
        return int(env)
    return 2 if (ndim == 3 and ksize > 1 and tiles[2] > 1) else 0


def _rup8(n: int) -> int:
    return (n + 7) // 8 * 8


def down_conv_plan(x: torch.Tensor, wmat: torch.Tensor, y: torch.Tensor, *, bias=None, stats=None,
                   stats_cpg=0, block_n=None, stages=0, name="",
                   residual: Optional[torch.Tensor] = None, out_f32: bool = False) -> GemmPlan:
    """Conv k=4, stride 2, pad 1 (slice_cond_2d_ddpm/unet.py:70, unet_attention.py:123): input
    index 2*o - 1 + k = 2*(o + a) + par with (par, a) = (1,-1), (0,0), (1,0), (0,1) for k = 0..3,
    i.e. tap k reads the parity-`par` sub-lattice shifted by a."""
    ndim = y.dim() - 2
    ext, sample_dim = _spatial_ext(y, ndim)
    box = choose_box(ext, prefer_unit=(sample_dim - 1,))
    Ci, cout_pad = x.shape[-1], y.shape[-1]
    assert Ci % BLOCK_K == 0
    bn = block_n or pick_block_n(cout_pad)
    PA = [(1, -1), (0, 0), (1, 0), (0, 1)]
    pars = list(itertools.product((0, 1), repeat=ndim))  # (pw, ph, pd) index: w fastest
    a_maps = [MapSpec(_parity_view(x, ndim, par), (BLOCK_K,) + box, 3) for par in pars]
    par_index = {par: i for i, par in enumerate(pars)}
    rows, bk = [], 0
    for tap in itertools.product(range(4), repeat=ndim):  # (kd, kh, kw)
        kk = list(reversed(tap))  # w, h, d
        par = tuple(PA[k][0] for k in kk)
        offs = [PA[k][1] for k in kk] + [0] * (4 - ndim)
        for c0 in range(0, Ci, BLOCK_K):
            rows.append([par_index[par], c0] + offs + [bk, 0])
            bk += BLOCK_K
    assert wmat.shape == (cout_pad, bk)
    kt = np.asarray(rows, dtype=np.int32)[None]
    bview = TView(wmat, (bk, cout_pad, 1, 1), (1, bk, _rup8(bk * cout_pad), _rup8(bk * cout_pad)))
    assert y.dtype == (torch.float32 if out_f32 else torch.bfloat16) and not (out_f32 and residual is not None)
    chunk = _out_chunk(bn, out_f32)
    omap = MapSpec(_act_view(y, ndim), (chunk,) + box, _out_swizzle(chunk * (4 if out_f32 else 2)))
    tiles = tuple(-(-e // b) for e, b in zip(ext, box))
    rmaps = None
    if residual is not None:
        assert residual.shape == y.shape
        rmaps = [MapSpec(_act_view(residual, ndim), (chunk,) + box, _out_swizzle(chunk * 2))]
    return GemmPlan(a_maps=a_maps, b_map=MapSpec(bview, (BLOCK_K, bn, 1, 1), 3), o_maps=[omap], out_f32=out_f32,
                    r_maps=rmaps, ktable=kt, tiles=tiles, box=box, ext=ext, block_n=bn, n_total=cout_pad,
                    sample_dim=sample_dim, bias=bias, stats=stats,
                    stats_ld=(stats.shape[1] if stats is not None else 0), stats_cpg=stats_cpg,
                    stages=stages, name=name, flops=2 * int(np.prod(ext)) * cout_pad * bk)


# output parity rho -> [(k, input shift)] for ConvTranspose k=4, s=2, p=1:  y = 2*i - 1 + k
_CT_TAPS = {0: [(1, 0), (3, -1)], 1: [(0, 1), (2, 0)]}


def up_conv_plan(x: torch.Tensor, wmat: torch.Tensor, y: torch.Tensor, *, bias=None, stats=None,
                 stats_cpg=0, block_n=None, stages=0, name="",
                 residual: Optional[torch.Tensor] = None, out_f32: bool = False) -> GemmPlan:
    """ConvTranspose k=4, stride 2, pad 1 (slice_cond_2d_ddpm/unet.py:89, unet_attention.py:142)
    as 2^d output-parity classes, each a 2^d-tap stride-1 convolution of the input written to
    the parity sub-lattice of the output.  wmat: [n_class, Cout_pad, 2^d * Cin]."""
    ndim = y.dim() - 2
    in_ext, sample_dim = _spatial_ext(x, ndim)
    box = choose_box(in_ext, prefer_unit=(sample_dim - 1,))
    Ci, cout_pad = x.shape[-1], y.shape[-1]
    assert Ci % BLOCK_K == 0
    for a, b in zip(x.shape[1:-1], y.shape[1:-1]):
        assert b == 2 * a
    bn = block_n or pick_block_n(cout_pad)
    classes = list(itertools.product((0, 1), repeat=ndim))  # (rw, rh, rd), w fastest
    a_maps = [MapSpec(_act_view(x, ndim), (BLOCK_K,) + box, 3)]
    kts = []
    K = (2 ** ndim) * Ci
    for rho in classes:
        rows, bk = [], 0
        for combo in itertools.product((0, 1), repeat=ndim):  # tap choice per dim (w fastest)
            offs = [_CT_TAPS[rho[i]][combo[i]][1] for i in range(ndim)] + [0] * (4 - ndim)
            for c0 in range(0, Ci, BLOCK_K):
                rows.append([0, c0] + offs + [bk, 0])
                bk += BLOCK_K
        assert bk == K
        kts.append(rows)
    kt = np.asarray(kts, dtype=np.int32)
    ncls = len(classes)
    assert wmat.shape == (ncls, cout_pad, K), (wmat.shape, (ncls, cout_pad, K))
    bview = TView(wmat, (K, cout_pad, ncls, 1), (1, K, K * cout_pad, _rup8(K * cout_pad * ncls)))
    assert y.dtype == (torch.float32 if out_f32 else torch.bfloat16) and not (out_f32 and residual is not None)
    chunk = _out_chunk(bn, out_f32)
    o_maps = [MapSpec(_parity_view(y, ndim, rho), (chunk,) + box, _out_swizzle(chunk * (4 if out_f32 else 2)))
              for rho in classes]
    tiles = tuple(-(-e // b) for e, b in zip(in_ext, box))
    rmaps = None
    if residual is not None:
        assert residual.shape == y.shape
        rmaps = [MapSpec(_parity_view(residual, ndim, rho), (chunk,) + box, _out_swizzle(chunk * 2))
                 for rho in classes]
    return GemmPlan(a_maps=a_maps, b_map=MapSpec(bview, (BLOCK_K, bn, 1, 1), 3), o_maps=o_maps, out_f32=out_f32,
                    r_maps=rmaps, ktable=kt, tiles=tiles, box=box, ext=in_ext, block_n=bn, n_total=cout_pad,
                    bz_sel=(1, 0), sample_dim=sample_dim, bias=bias, stats=stats,
                    stats_ld=(stats.shape[1] if stats is not None else 0), stats_cpg=stats_cpg,
                    stages=stages, name=name, flops=2 * int(np.prod(in_ext)) * cout_pad * K * ncls)


def matrix_plan(a: TView, a_box_rows: Tuple[int, int, int, int], b: TView, o: TView, *, K: int,
                n_total: int, block_n: int, ext, tiles, bz_sel=(0, 0), sample_dim=0, out_f32=False,
                bias=None, bias_m=None, stats=None, stats_cpg=0, stages=0, name="",
                flops=0) -> GemmPlan:
    """Plain (batched) GEMM D = A B^T over explicit views: A rank-5 (K, x1..x4), B rank-4
    (K, rows, z1, z2), O rank-5 (N, x1..x4).  Used for the im2col'd first convolution and the
    attention products (unet_attention.py:49-51)."""
    n_kb = -(-K // BLOCK_K)
    kt = np.zeros((1, n_kb, 8), dtype=np.int32)
    kt[0, :, 1] = np.arange(n_kb) * BLOCK_K
    kt[0, :, 6] = np.arange(n_kb) * BLOCK_K
    chunk = _out_chunk(block_n, out_f32)
    es = 4 if out_f32 else 2
    return GemmPlan(a_maps=[MapSpec(a, (BLOCK_K,) + tuple(a_box_rows), 3)],
                    b_map=MapSpec(b, (BLOCK_K, block_n, 1, 1), 3),
                    o_maps=[MapSpec(o, (chunk,) + tuple(a_box_rows), _out_swizzle(chunk * es))],
                    ktable=kt, tiles=tuple(tiles), box=tuple(a_box_rows), ext=tuple(ext),
                    block_n=block_n, n_total=n_total, bz_sel=bz_sel, sample_dim=sample_dim,
                    out_f32=out_f32, bias=bias, bias_m=bias_m, stats=stats,
                    stats_ld=(stats.shape[1] if stats is not None else 0), stats_cpg=stats_cpg,
                    stages=stages, name=name, flops=flops)


def pick_wgrad_splits(base: int, total_mt: int, sms: int = 148, max_waves: int = 4) -> int:
    """Split count for the wgrad grid (grid = base * splits CTAs, ONE resident CTA per SM, all of
    about equal duration): the grid must FILL whole waves -- 300 CTAs on 148 SMs run as three
    waves, the last one with 4 CTAs.  Take the wave count (<= max_waves) whose grid uses the
    largest fraction of its SM-waves; ties go to fewer waves (fewer partial-sum reductions).
    Every CTA's share of the M tiles is rounded up, which is part of the cost."""
    best, best_cost = 1, None
    for waves in range(1, max_waves + 1):
        splits = min(total_mt, (waves * sms) // base)
        if splits < 1:
            continue
        w_real = -(-(base * splits) // sms)
        cost = w_real * -(-total_mt // splits)      # waves x tiles per CTA
        if best_cost is None or cost < best_cost * 0.97:
            best, best_cost = splits, cost
    if best_cost is None:                            # base > max_waves * sms: no split at all
        return 1
    return best


@dataclass
class WgradPlan:
    """dW[class][co][b_k + c] += sum_m dY[m, co] * A_kb[m, c] over the forward plan's k-table
    (MriWgradArgs).  `dy` is the gradient of the forward output tensor(s) `fwd.o_maps` view."""
    fwd: GemmPlan
    dy: torch.Tensor            # same shape/layout as the forward output base tensor
    dw: torch.Tensor            # fp32 [n_class, rows, K]
    n_total: int
    group: int = 4
    stages: int = 2
    splits: int = 0
    name: str = ""
    dy_views: Optional[List[TView]] = None   # explicit per-class dY views (attention products)
    _args: Optional[_lib.MriWgradArgs] = field(default=None, repr=False)
    _launches: list = field(default_factory=list, repr=False)
    _keep: list = field(default_factory=list, repr=False)

    def dy_specs(self) -> List[MapSpec]:
        if self.dy_views is not None:
            return [MapSpec(v, (BLOCK_K,) + tuple(self.fwd.box), 3) for v in self.dy_views]
        out = []
        for om in self.fwd.o_maps:
            v = om.view
            assert v.base.shape == self.dy.shape and v.base.dtype == torch.bfloat16
            out.append(MapSpec(TView(self.dy, v.dims, v.strides, v.offset),
                               (BLOCK_K,) + tuple(self.fwd.box), 3))
        return out

    def pick_splits(self, n_kb: Optional[int] = None, group: Optional[int] = None) -> int:
        if self.splits:
            return self.splits
        f = self.fwd
        co_blocks = -(-self.n_total // 128)
        n_kb = f.n_kb if n_kb is None else n_kb
        group = self.group if group is None else group
        base = f.n_class * co_blocks * (-(-n_kb // group))
        total_mt = int(np.prod(f.tiles))
        return pick_wgrad_splits(base, total_mt)

    def runs(self) -> List[Tuple[int, int, bool]]:
        """Cut the k-table into launches: (first entry, entries, xgroup).  Forward plans that share
        activation tiles (xreuse 1, one class, boxes of 8 x 16 positions) list their taps as kw
        triples; runs of whole triples go to the kernel's xgroup mode (one 10-wide tile per
        triple), everything else (folded 1x1 slabs, other plans) to the plain mode."""
        f = self.fwd
        use = (f.xreuse == 1 and f.n_class == 1 and self.dy_views is None and f.a_maps_std is not None
               and int(np.prod(f.box)) == BLOCK_M and f.box[0] == 8 and f.box[1] == 16
               and os.environ.get("MRI_WGRAD_XGROUP", "1") != "0")
        if not use:
            return [(0, f.n_kb, False)]
        kt = f.ktable[0]

        def is_triple(i):
            if i + 2 > f.n_kb - 1:
                return False
            r0, r1, r2 = kt[i], kt[i + 1], kt[i + 2]
            same = all(r0[c] == r1[c] == r2[c] for c in (0, 1, 3, 4, 5))
            return bool(same and r0[2] == -1 and r1[2] == 0 and r2[2] == 1
                        and r0[7] == 1 and r1[7] == 0 and r2[7] == 0)

        out, i = [], 0
        while i < f.n_kb:
            j = i
            if is_triple(i):
                while j < f.n_kb and is_triple(j):
                    j += 3
                out.append((i, j - i, True))
            else:
                while j < f.n_kb and not is_triple(j):
                    j += 1
                out.append((i, j - i, False))
            i = j
        return out

    def materialize(self, device) -> None:
        f = self.fwd
        std = list(f.a_maps_std or f.a_maps)
        maps = std + list(f.a_maps) + self.dy_specs()
        blob = encode_maps(maps, device)
        kt = torch.from_numpy(np.ascontiguousarray(f.ktable, dtype=np.int32)).to(device)
        self._keep = [blob, kt]
        assert self.dw.dtype == torch.float32 and self.dw.is_contiguous() and self.dw.dim() == 3
        assert self.dw.shape[0] == f.n_class and self.dw.shape[1] >= self.n_total
        self._launches = []
        for (k0, nk, xg) in self.runs():
            a = _lib.MriWgradArgs()
            a.a_maps = blob.data_ptr() + (128 * len(std) if xg else 0)
            a.dy_maps = blob.data_ptr() + 128 * (len(std) + len(f.a_maps))
            # sub-ranges of the table are only formed for single-class plans (runs())
            a.ktable = kt.data_ptr() + k0 * 8 * 4
            a.n_kb, a.n_class = nk, f.n_class
            for i in range(4):
                a.tiles[i], a.box[i] = f.tiles[i], f.box[i]
            a.n_total = self.n_total
            a.co_blocks = -(-self.n_total // 128)
            a.group = 6 if xg else self.group
            a.splits = self.pick_splits(nk, a.group)
            a.dw = self.dw.data_ptr()
            a.dw_rows, a.dw_ld = self.dw.shape[1], self.dw.shape[2]
            a.stages = 3 if xg else self.stages
            a.xgroup = 1 if xg else 0
            self._launches.append(a)
        self._args = self._launches[0]

    def launch(self, stream: Optional[int] = None) -> None:
        st = stream if stream is not None else _lib.current_stream_ptr()
        for a in self._launches:
            rc = _lib.load().mri_wgrad_launch(C.byref(a), st)
            _lib.check(rc, f"mri_wgrad_launch[{self.name}]")

    def simulate(self) -> None:
        """CPU emulation (tests): same tables, TMA zero fill."""
        f = self.fwd
        rows = int(np.prod(f.box))
        dys = self.dy_specs()
        for cls in range(f.n_class):
            for tix in itertools.product(*[range(t) for t in reversed(f.tiles)]):
                tix = tuple(reversed(tix))
                org = [tix[i] * f.box[i] for i in range(4)]
                dyt = _load_box(dys[cls].view, [0] + org,
                                (dys[cls].view.dims[0],) + tuple(f.box)).reshape(rows, -1)
                for kb in range(f.n_kb):
                    e = f.ktable[cls, kb]
                    am = f.a_maps[int(e[0])]
                    coords = [int(e[1])] + [org[i] + int(e[2 + i]) for i in range(4)]
                    a = _load_box(am.view, coords, (BLOCK_K,) + tuple(f.box)).reshape(rows, BLOCK_K)
                    bk = int(e[6])
                    self.dw[cls, :self.n_total, bk:bk + BLOCK_K] += (dyt.t() @ a)[:self.n_total]


# ======================================================================================
# Weight packing (fp32 reference layouts -> K-major bf16 matrices matching the k-tables)
# ======================================================================================

def pack_conv_weight(w: torch.Tensor, splits: Optional[Sequence[int]] = None, cout_pad: int = 0,
                     extra: Sequence[torch.Tensor] = ()) -> torch.Tensor:
    """nn.ConvNd weight [Cout, Cin, *k] -> [Cout_pad, K] with K ordered
    [source][tap (kd,kh,kw)][channel]; `splits` = channels per concatenated source.
    `extra`: 1x1 weights [Cout, C_j(,1,1,1)] appended as centre-tap-only sources."""
    cout, cin = w.shape[0], w.shape[1]
    taps = int(np.prod(w.shape[2:]))
    splits = list(splits) if splits else [cin]
    assert sum(splits) == cin
    wt = w.reshape(cout, cin, taps)
    cols, c0 = [], 0
    for ci in splits:
        cols.append(wt[:, c0:c0 + ci, :].permute(0, 2, 1).reshape(cout, taps * ci))
        c0 += ci
    for e in extra:
        cols.append(e.reshape(cout, -1))
    m = torch.cat(cols, dim=1)
    cp = cout_pad or cout
    out = torch.zeros(cp, m.shape[1], dtype=torch.bfloat16, device=w.device)
    out[:cout] = m.to(torch.bfloat16)
    return out.contiguous()


def pack_tap_weight(w: torch.Tensor, rows_pad: int) -> torch.Tensor:
    """nn.ConvNd weight [Cout, Cin, *k] -> [rows_pad, Cin] bf16 with row (tap * Cout + co), tap in
    (kd, kh, kw) order: the B operand of the "GEMM over taps" used for thin-Cout convolutions
    (mri_tap_gather finishes the convolution)."""
    cout, cin = w.shape[0], w.shape[1]
    taps = int(np.prod(w.shape[2:]))
    m = w.reshape(cout, cin, taps).permute(2, 0, 1).reshape(taps * cout, cin)
    out = torch.zeros(rows_pad, cin, dtype=torch.bfloat16, device=w.device)
    out[:taps * cout] = m.to(torch.bfloat16)
    return out.contiguous()


def pack_convT_weight(w: torch.Tensor, cout_pad: int = 0) -> torch.Tensor:
    """nn.ConvTransposeNd weight [Cin, Cout, 4,4(,4)] -> [2^d classes, Cout_pad, 2^d * Cin],
    class / tap order matching up_conv_plan (w fastest)."""
    cin, cout = w.shape[0], w.shape[1]
    ndim = w.dim() - 2
    cp = cout_pad or cout
    classes = list(itertools.product((0, 1), repeat=ndim))
    out = torch.zeros(len(classes), cp, (2 ** ndim) * cin, dtype=torch.bfloat16, device=w.device)
    for ci, rho in enumerate(classes):
        col = 0
        for combo in itertools.product((0, 1), repeat=ndim):
            ks = [_CT_TAPS[rho[i]][combo[i]][0] for i in range(ndim)]  # w, h, (d)
            idx = tuple(reversed(ks))  # weight dims are (kd, kh, kw)
            sl = w[(slice(None), slice(None)) + idx]  # [Cin, Cout]
            out[ci, :cout, col:col + cin] = sl.t().to(torch.bfloat16)
            col += cin
    return out.contiguous()


def unpack_conv_wgrad(dw: torch.Tensor, w_shape, splits: Optional[Sequence[int]] = None,
                      extra_shapes: Sequence = ()) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """Inverse of pack_conv_weight for an fp32 gradient matrix [Cout_pad, K]: returns the gradient
    in the nn.ConvNd layout [Cout, Cin, *k] and the gradients of the folded 1x1 weights."""
    cout, cin = w_shape[0], w_shape[1]
    taps = int(np.prod(w_shape[2:]))
    splits = list(splits) if splits else [cin]
    parts, col = [], 0
    for ci in splits:
        blk = dw[:cout, col:col + taps * ci].reshape(cout, taps, ci).permute(0, 2, 1)
        parts.append(blk)
        col += taps * ci
    g = torch.cat(parts, dim=1).reshape(w_shape)
    extras = []
    for shp in extra_shapes:
        n = int(np.prod(shp[1:]))
        extras.append(dw[:cout, col:col + n].reshape(shp))
        col += n
    return g, extras


def unpack_convT_wgrad(dw: torch.Tensor, w_shape) -> torch.Tensor:
    """Inverse of pack_convT_weight: dw [2^d, Cout_pad, 2^d * Cin] -> [Cin, Cout, 4, 4(, 4)]."""
    cin, cout = w_shape[0], w_shape[1]
    ndim = len(w_shape) - 2
    g = torch.zeros(w_shape, dtype=dw.dtype, device=dw.device)
    classes = list(itertools.product((0, 1), repeat=ndim))
    for ci, rho in enumerate(classes):
        col = 0
        for combo in itertools.product((0, 1), repeat=ndim):
            ks = [_CT_TAPS[rho[i]][combo[i]][0] for i in range(ndim)]
            idx = tuple(reversed(ks))
            g[(slice(None), slice(None)) + idx] = dw[ci, :cout, col:col + cin].t()
            col += cin
    return g
