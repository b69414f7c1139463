"""Split precision mode of the UNets (inference): parity with the reference's fp32 / TF32 sampling
path.

`ddpm_3d_ldm/show_model.py:254` samples WITHOUT autocast, i.e. with fp32 tensors and TF32 cuDNN
convolutions (1.1e-3 rel-L2 from fp32 on a B200, profiles/r04g_reference_precision_modes.json); the
default path of this package computes with bf16 operands like the reference's training path
(1.2e-2).  This mode closes the gap on the SAME tensor-core kernel, which only takes bf16
operands: an fp32 value v travels as two bf16 numbers hi = bf16(v), lo = bf16(v - hi) and a
product a * w is evaluated as a_hi * w_hi + a_lo * w_hi + a_hi * w_lo by widening the contraction
-- an activation with C channels is stored channels-last as [hi | lo | hi] (3C channels), the
weights as [w_hi | w_hi | w_lo] along their input-channel axis, which is an ordinary convolution
with 3C input channels for the implicit-GEMM plans; accumulation is fp32 and the output is written
as fp32 (`out_f32`, plain tiles).  Everything between two convolutions runs on fp32 tensors
(csrc/split_precision.cu): GroupNorm (+ SiLU, exact sigmoid) reads the fp32 convolution output and
writes the widened operand; the residual / skip connections enter conv2's K loop as centre-tap
sources with identity weights; the attention logits, probabilities and values are widened the
same way.  Cost: 3x the contraction length on plain (non-`swap_ab`) tiles, about 6x the bf16
path -- a parity mode, selected per model with `model.precision = "split"`.

Reference graph: ddpm_3d_ldm/unet_attention.py:157-200 (the same walk as engine.UNet3DProgram).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import _lib, ops
from . import plan as P
from .engine import UNetProgram, _pad_k, _pad_vec, _rup


# ---- host-side helpers (weights are widened on the host when they are packed) ------------------
def _hi_lo(w: torch.Tensor):
    hi = w.to(torch.bfloat16).to(torch.float32)
    lo = (w - hi).to(torch.bfloat16).to(torch.float32)
    return hi, lo


def widen_weight(w: torch.Tensor, splits: Optional[Sequence[int]] = None, transposed: bool = False) -> torch.Tensor:
    """nn.ConvNd weight [Cout, Cin, *k] (transposed: nn.ConvTransposeNd [Cin, Cout, *k]) -> the same
    layout with 3 * Cin input channels: per concatenated source (`splits`) [w_hi | w_hi | w_lo].
    The values are bf16-exact fp32, so the bf16 packers store them unchanged."""
    dim = 0 if transposed else 1
    cin = w.shape[dim]
    splits = list(splits) if splits else [cin]
    assert sum(splits) == cin
    hi, lo = _hi_lo(w.detach().to(torch.float32))
    parts, c0 = [], 0
    for c in splits:
        sl = [slice(None)] * w.dim()
        sl[dim] = slice(c0, c0 + c)
        parts += [hi[tuple(sl)], hi[tuple(sl)], lo[tuple(sl)]]
        c0 += c
    return torch.cat(parts, dim=dim).contiguous()


def identity_weight(C: int, device=None) -> torch.Tensor:
    """1x1 weight [C, 3C] that adds hi + lo of a widened activation: [I | I | 0]."""
    eye = torch.eye(C, device=device)
    return torch.cat([eye, eye, torch.zeros(C, C, device=device)], dim=1).contiguous()


def widen_activation(x_cl: torch.Tensor) -> torch.Tensor:
    """[..., C] fp32 -> [..., 3C] bf16 = [hi | lo | hi] (host reference of mri_gn_split / mri_split3)."""
    hi = x_cl.to(torch.bfloat16)
    lo = (x_cl - hi.to(torch.float32)).to(torch.bfloat16)
    return torch.cat([hi, lo, hi], dim=-1).contiguous()


class F32Act:
    """fp32 channels-last activation [N, *spatial, C] (+ GroupNorm partial sums of its producer)
    and, once somebody needs it as a convolution operand, its plain widened copy."""

    def __init__(self, t: torch.Tensor, stats: Optional[torch.Tensor] = None, cpg: int = 0):
        self.t, self.stats, self.cpg = t, stats, cpg
        self.raw3: Optional[torch.Tensor] = None
        self.keep = False          # a pending skip connection: not released by its first consumer

    @property
    def C(self) -> int:
        return self.t.shape[-1]

    @property
    def spatial(self) -> int:
        n = 1
        for s in self.t.shape[1:-1]:
            n *= s
        return n


class _SplitProgram(UNetProgram):
    """Buffers and op emitters shared by the split precision programs."""

    # ------------------------------------------------------------------ buffers
    def new_f32(self, sp: Sequence[int], C: int, with_stats: bool = True) -> F32Act:
        t = self.pool.get((self.B, *sp, C), torch.float32)
        if with_stats and C % (8 * self.groups) == 0:
            return F32Act(t, self.new_stats(self.groups), C // self.groups)
        return F32Act(t)

    def wide(self, like: torch.Tensor) -> torch.Tensor:
        return self.pool.get((*like.shape[:-1], 3 * like.shape[-1]), torch.bfloat16)

    def free(self, *items) -> None:
        """Buffers whose last consumer has been emitted go back to the pool (the program is
        straight-line: a released buffer is only handed to ops emitted later).  An F32Act that is
        still a pending skip connection (`keep`) stays."""
        for it in items:
            if it is None:
                continue
            if isinstance(it, F32Act):
                if it.keep:
                    continue
                self.pool.release(it.t)
                if it.raw3 is not None:
                    self.pool.release(it.raw3)
                it.t = it.raw3 = None
            else:
                self.pool.release(it)

    # ------------------------------------------------------------------ op emitters
    def gn_split(self, x: F32Act, gamma, beta, groups: int, eps: float, silu: bool, name: str,
                 c_off: int = 0, rowbias=None, rowbias_ld: int = 0) -> torch.Tensor:
        """[hi | lo | hi] of act(GroupNorm(x)) (+ rowbias); gamma / beta channels [c_off, c_off + C) apply."""
        assert x.stats is not None, f"{name}: input has no statistics"
        y = self.wide(x.t)
        B, S, C, xs, st, cpg = self.B, x.spatial, x.C, x.t, x.stats, x.cpg
        gm, bt = gamma[c_off:c_off + C], beta[c_off:c_off + C]
        self._add(name, lambda: ops.gn_split(xs, y, st, gm, bt, B, S, C, groups, cpg, eps, silu,
                                             rowbias=rowbias, rowbias_ld=rowbias_ld), [y])
        return y

    def gn_f32(self, x: F32Act, gamma, beta, groups: int, eps: float, silu: bool, residual: torch.Tensor,
               name: str) -> F32Act:
        """fp32 act(GroupNorm(x)) + residual: the post-norm block output (slice_cond_2d_ddpm/unet.py:52-56)."""
        assert x.stats is not None, f"{name}: input has no statistics"
        y = F32Act(self.pool.get(tuple(x.t.shape), torch.float32))
        B, S, C, xs, st, cpg, yt = self.B, x.spatial, x.C, x.t, x.stats, x.cpg, y.t
        self._add(name, lambda: ops.gn_split(xs, None, st, gamma, beta, B, S, C, groups, cpg, eps, silu,
                                             residual=residual, y32=yt), [yt])
        return y

    def raw(self, x: F32Act, name: str) -> torch.Tensor:
        """[hi | lo | hi] of x itself (a convolution operand without a norm in front of it)."""
        if x.raw3 is None:
            y = self.wide(x.t)
            B, S, C, xs = self.B, x.spatial, x.C, x.t
            self._add(f"{name}.split", lambda: ops.gn_split(xs, y, None, None, None, B, S, C, 1, 1, 0.0, False), [y])
            x.raw3 = y
        return x.raw3

    def conv(self, sources, wmat, cout: int, ksize: int, bias, *, rowbias=None, rowbias_ld=0,
             with_stats=True, name="conv") -> F32Act:
        sp = sources[0].x.shape[1:-1]
        y = self.new_f32(sp, cout, with_stats)
        self.gemm(P.conv_plan(sources, wmat, y.t, ksize, bias=bias, rowbias=rowbias, rowbias_ld=rowbias_ld,
                              stats=y.stats, stats_cpg=y.cpg, out_f32=True, name=name))
        return y

    def matrix_in_conv(self, col: torch.Tensor, wmat: torch.Tensor, S: int, kpad: int, cout: int, bias,
                       name: str, with_stats: bool) -> F32Act:
        """The first convolution as a GEMM over the patch matrix `col` [B * S, kpad]."""
        B = self.B
        h = self.new_f32(self.sp, cout, with_stats)
        a = P.TView(col, (kpad, S, B, 1, 1), (1, kpad, S * kpad, B * S * kpad, B * S * kpad))
        bv = P.TView(wmat, (kpad, cout, 1, 1), (1, kpad, kpad * cout, kpad * cout))
        o = P.TView(h.t, (cout, S, B, 1, 1), (1, cout, S * cout, B * S * cout, B * S * cout))
        self.gemm(P.matrix_plan(a, (128, 1, 1, 1), bv, o, K=kpad, n_total=cout, block_n=P.pick_block_n(cout),
                                ext=(S, B, 1, 1), tiles=(-(-S // 128), B, 1, 1), sample_dim=2, bias=bias,
                                stats=h.stats, stats_cpg=h.cpg, out_f32=True, name=name,
                                flops=2 * B * S * cout * kpad))
        return h

    def head(self, a3: torch.Tensor, oc, S: int, cin_pad: int = 0) -> None:
        """out_conv with fp32 output, channels-last [.., cout_pad] -> NC[D]HW.  What the samplers read
        (diffusion_base._reverse_loop_on) is eps in the state's own layout (ldc = 0)."""
        B, dev = self.B, self.device
        cout = oc.weight.shape[0]
        cp = _rup(cout, 16)
        sp = tuple(a3.shape[1:-1])

        def make():
            w = oc.weight.detach()
            if cin_pad and w.shape[1] != cin_pad:
                wp = torch.zeros(w.shape[0], cin_pad, *w.shape[2:], dtype=w.dtype, device=w.device)
                wp[:, :w.shape[1]] = w
                w = wp
            return P.pack_conv_weight(widen_weight(w), cout_pad=cp)

        w_out = self.packed(make)
        b_out = self.packed(lambda: _pad_vec(oc.bias.detach(), cp))
        y = torch.zeros(B, *sp, cp, device=dev)
        self.gemm(P.conv_plan([P.ConvSource(a3)], w_out, y, oc.weight.shape[2], bias=b_out, out_f32=True,
                              name="out_conv"))
        self.out = torch.zeros(B, cout, *sp, device=dev)
        src = y.view(B, S, cp)[:, :, :cout].permute(0, 2, 1)
        dst = self.out.view(B, cout, S)
        self._add("out.nchw", lambda: ops.copy_cast(src, dst), [self.out])
        self.eps_nhwc, self.cout, self.cout_pad = self.out, cout, 0
        self.fused_head = None


class UNet3DSplitProgram(_SplitProgram):
    """ddpm_3d_ldm/unet_attention.py:88-200 and ddpm_3d_ldm/unet.py:57-158 in split precision."""

    def __init__(self, model, batch: int, spatial: Sequence[int]):
        dev = next(model.parameters()).device
        super().__init__(dev, batch, spatial, groups=model.out_norm.num_groups, training=False)
        self.model = model
        B, (D, H, W) = batch, self.sp
        cin = model.in_channels
        chs = list(model.chs)
        L = len(chs)
        for s in self.sp:
            if s % (2 ** (L - 1)) != 0:
                raise _lib.MriError(f"spatial size {self.sp} must be divisible by {2 ** (L - 1)}")
        eps = model.out_norm.eps
        self.x_in = torch.zeros(B, cin, D, H, W, device=dev)
        self.t_in = torch.zeros(B, dtype=torch.int64, device=dev)

        tdim = model.time_mlp[1].in_features
        temb = self.time_embedding(self.t_in, model.time_mlp, tdim)
        blocks = []
        for blk in model.downs:
            blocks += [blk["res1"], blk["res2"]]
        blocks += [model.mid1, model.mid2]
        for blk in model.ups:
            blocks += [blk["res1"], blk["res2"]]
        tproj, toffs, tld = self.block_projections(temb, blocks, act=0)
        self._tproj = {id(b): (tproj[:, o:], tld, o) for b, o in zip(blocks, toffs)}
        self.tape.clear()        # inference only: nothing is differentiated in this mode

        # ---- in_conv: patch matrix over the channels [x | x - bf16(x) | x] ----------------------
        S = D * H * W
        ic = model.in_conv
        self.track(ic.weight, ic.bias)
        x6 = torch.zeros(B, 2 * cin, D, H, W, device=dev)
        kpad = _rup(27 * 3 * cin, 64)
        col = torch.zeros(B * S, kpad, dtype=torch.bfloat16, device=dev)
        x_in = self.x_in
        self._add("in_conv.lo", lambda: ops.bf16_residual_nchw(x_in, x6, B, cin * S), [x6])
        self._add("in_conv.im2col", lambda: ops.im2col(x_in, col, B, cin, D, H, W, 3, 3, kpad,
                                                       src2=x6, cin2=2 * cin), [col])
        w_in = self.packed(lambda: _pad_k(P.pack_conv_weight(widen_weight(ic.weight)), kpad))
        h = self.matrix_in_conv(col, w_in, S, kpad, chs[0], ic.bias, "in_conv", True)

        # ---- down path ---------------------------------------------------------------------------
        skips: List[F32Act] = []
        for i, blk in enumerate(model.downs):
            h = self.resblock(h, None, blk["res1"], eps, f"downs.{i}.res1")
            h = self.resblock(h, None, blk["res2"], eps, f"downs.{i}.res2")
            h.keep = True
            skips.append(h)
            if i != L - 1:
                dn = blk["down"]
                self.track(dn.weight, dn.bias)
                wd = self.packed(lambda dn=dn: P.pack_conv_weight(widen_weight(dn.weight)))
                y = self.new_f32([s // 2 for s in h.t.shape[1:-1]], chs[i + 1])
                self.gemm(P.down_conv_plan(self.raw(h, f"downs.{i}.down.in"), wd, y.t, bias=dn.bias,
                                           stats=y.stats, stats_cpg=y.cpg, out_f32=True, name=f"downs.{i}.down"))
                h = y

        # ---- bottleneck ----------------------------------------------------------------------------
        h = self.resblock(h, None, model.mid1, eps, "mid1")
        if hasattr(model, "mid_attn"):
            h = self.attention(h, model.mid_attn, "mid_attn")
        h = self.resblock(h, None, model.mid2, eps, "mid2")

        # ---- up path -------------------------------------------------------------------------------
        for j, blk in enumerate(model.ups):
            i = L - 1 - j
            if i != L - 1:
                up = blk["up"]
                self.track(up.weight, up.bias)
                wu = self.packed(lambda up=up: P.pack_convT_weight(widen_weight(up.weight, transposed=True)))
                u = self.new_f32([s * 2 for s in h.t.shape[1:-1]], chs[i])
                self.gemm(P.up_conv_plan(self.raw(h, f"ups.{j}.up.in"), wu, u.t, bias=up.bias, stats=u.stats,
                                         stats_cpg=u.cpg, out_f32=True, name=f"ups.{j}.up"))
                self.free(h)
                h = u
            skip = skips.pop()
            skip.keep = False       # its last consumer is the block below
            if tuple(skip.t.shape[1:-1]) != tuple(h.t.shape[1:-1]):
                raise _lib.MriError("skip/upsample shape mismatch (centre-crop path not supported)")
            h = self.resblock(h, skip, blk["res1"], eps, f"ups.{j}.res1")
            h = self.resblock(h, None, blk["res2"], eps, f"ups.{j}.res2")

        # ---- head --------------------------------------------------------------------------------
        on, oc = model.out_norm, model.out_conv
        self.track(on.weight, on.bias, oc.weight, oc.bias)
        a3 = self.gn_split(h, on.weight, on.bias, self.groups, eps, True, name="out_norm")
        self.free(h)
        self.head(a3, oc, S)
        self.params_changed()

    def resblock(self, x: F32Act, skip: Optional[F32Act], blk, eps: float, name: str) -> F32Act:
        """ResidualBlock3D (unet_attention.py:59-85); with `skip` the input is cat([x, skip], 1)."""
        n1, n2, c1, c2 = blk.norm1, blk.norm2, blk.conv1, blk.conv2
        self.track(n1.weight, n1.bias, n2.weight, n2.bias, c1.weight, c1.bias, c2.weight, c2.bias)
        cout = c1.weight.shape[0]
        rowbias, rb_ld, _ = self._tproj[id(blk)]
        srcs = [x] if skip is None else [x, skip]
        cins = [s.C for s in srcs]
        g_each = self.groups // len(srcs)   # GN(8, 2C) over a concat == GN(4) + GN(4) over the halves
        normed, c0 = [], 0
        for k, s in enumerate(srcs):
            normed.append(self.gn_split(s, n1.weight, n1.bias, g_each, eps, True, f"{name}.norm1.{k}", c_off=c0))
            c0 += s.C
        w1 = self.packed(lambda: P.pack_conv_weight(widen_weight(c1.weight, splits=cins),
                                                    splits=[3 * c for c in cins]))
        h = self.conv([P.ConvSource(a) for a in normed], w1, cout, 3, c1.bias, rowbias=rowbias,
                      rowbias_ld=rb_ld, name=f"{name}.conv1")
        self.free(*normed)
        a2 = self.gn_split(h, n2.weight, n2.bias, self.groups, eps, True, f"{name}.norm2")
        self.free(h)
        raws = [self.raw(s, f"{name}.in{k}") for k, s in enumerate(srcs)]
        if not isinstance(blk.skip, torch.nn.Identity):
            sk = blk.skip
            self.track(sk.weight, sk.bias)
            w2 = self.packed(lambda: P.pack_conv_weight(
                widen_weight(c2.weight),
                extra=[widen_weight(sk.weight.detach().reshape(cout, -1), splits=cins)]))
            sources = [P.ConvSource(a2)] + [P.ConvSource(r, taps=False) for r in raws]
            out = self.conv(sources, w2, cout, 3, c2.bias, rowbias=sk.bias, rowbias_ld=0,
                            name=f"{name}.conv2+skip")
        else:
            assert skip is None and x.C == cout
            w2 = self.packed(lambda: P.pack_conv_weight(widen_weight(c2.weight),
                                                        extra=[identity_weight(cout, self.device)]))
            out = self.conv([P.ConvSource(a2), P.ConvSource(raws[0], taps=False)], w2, cout, 3, c2.bias,
                            name=f"{name}.conv2+x")
        self.free(a2, *srcs)
        return out

    def attention(self, x: F32Act, blk, name: str) -> F32Act:
        """AttentionBlock3D (unet_attention.py:28-56) with every product in split precision."""
        B, C, dev = self.B, x.C, self.device
        heads = blk.num_heads
        d = C // heads
        n = x.spatial
        npad = _rup(n, 64)       # segment length of the widened key axis: whole K slabs
        C3 = 3 * C
        sp = tuple(x.t.shape[1:-1])
        self.track(blk.norm.weight, blk.norm.bias, blk.qkv.weight, blk.qkv.bias, blk.proj.weight,
                   blk.proj.bias)
        hn3 = self.gn_split(x, blk.norm.weight, blk.norm.bias, self.groups, blk.norm.eps, False, f"{name}.norm")
        wqkv = self.packed(lambda: P.pack_conv_weight(widen_weight(blk.qkv.weight)))
        qkv = self.conv([P.ConvSource(hn3)], wqkv, C3, 1, blk.qkv.bias, with_stats=False, name=f"{name}.qkv")
        # q -> [hi | lo | hi], k -> [hi | hi | lo] per head along d: [B, n, heads, 3d]
        q3 = torch.zeros(B, n, heads, 3 * d, dtype=torch.bfloat16, device=dev)
        k3 = torch.zeros(B, n, heads, 3 * d, dtype=torch.bfloat16, device=dev)
        qt = qkv.t
        self._add(f"{name}.q.split", lambda: ops.split3(qt, q3, B * n, heads, d, C3, d, heads * 3 * d, 3 * d, d, 0), [q3])
        self._add(f"{name}.k.split", lambda: ops.split3(qt, k3, B * n, heads, d, C3, d, heads * 3 * d, 3 * d, d, 1,
                                                        src_off=C), [k3])
        # S = q k^T (fp32 logits)
        Sm = torch.zeros(B, heads, n, npad, dtype=torch.float32, device=dev)
        ld = heads * 3 * d
        qa = P.TView(q3, (3 * d, n, heads, B, 1), (1, ld, 3 * d, n * ld, B * n * ld))
        kb = P.TView(k3, (3 * d, n, heads, B), (1, ld, 3 * d, n * ld))
        so = P.TView(Sm, (npad, n, heads, B, 1), (1, npad, n * npad, heads * n * npad, B * heads * n * npad))
        tiles = (-(-n // 128), heads, B, 1)
        self.gemm(P.matrix_plan(qa, (128, 1, 1, 1), kb, so, K=3 * d, n_total=npad, block_n=128,
                                ext=(n, heads, B, 1), tiles=tiles, bz_sel=(3, 4), out_f32=True,
                                name=f"{name}.qk^T", flops=2 * B * heads * n * n * 3 * d))
        P3 = torch.zeros(B, heads, n, 3 * npad, dtype=torch.bfloat16, device=dev)
        scale = float(d) ** -0.5
        self._add(f"{name}.softmax", lambda: ops.softmax_rows_split(Sm, P3, B * heads * n, n, npad, 3 * npad,
                                                                    npad, scale), [P3])
        # v^T [B, C, npad] fp32 = W_v . hn^T + b_v (keys contiguous), then widened [hi | hi | lo]
        wv = self.packed(lambda: widen_weight(blk.qkv.weight.detach()[2 * C:].reshape(C, C)).to(torch.bfloat16))
        bvv = self.packed(lambda: blk.qkv.bias.detach()[2 * C:].contiguous())
        vT = torch.zeros(B, C, npad, dtype=torch.float32, device=dev)
        a = P.TView(wv, (C3, C, 1, 1, 1), (1, C3, C * C3, C * C3, C * C3))
        b = P.TView(hn3, (C3, n, B, 1), (1, C3, n * C3, B * n * C3))
        sz = C * npad
        o_views = [P.TView(vT, (npad, C, 1, 1, 1), (1, npad, sz, sz, sz), offset=bi * sz) for bi in range(B)]
        pl = P.matrix_plan(a, (128, 1, 1, 1), b, o_views[0], K=C3, n_total=npad, block_n=128,
                           ext=(C, 1, 1, 1), tiles=(C // 128, 1, 1, 1), bz_sel=(1, 0), bias_m=bvv,
                           out_f32=True, name=f"{name}.vT", flops=2 * B * n * C * C3)
        chunk_box, swz = pl.o_maps[0].box, pl.o_maps[0].swizzle
        pl.o_maps = [P.MapSpec(v, chunk_box, swz) for v in o_views]
        pl.ktable = pl.ktable.repeat(B, axis=0)
        self.gemm(pl)
        vT3 = torch.zeros(B, C, 3 * npad, dtype=torch.bfloat16, device=dev)
        self._add(f"{name}.v.split", lambda: ops.split3(vT, vT3, B * C, 1, n, npad, 0, 3 * npad, 0, npad, 1), [vT3])
        # O = P v (fp32, token-major [B, n, C])
        O = F32Act(torch.zeros(B, *sp, C, dtype=torch.float32, device=dev))
        pa = P.TView(P3, (3 * npad, n, heads, B, 1),
                     (1, 3 * npad, n * 3 * npad, heads * n * 3 * npad, B * heads * n * 3 * npad))
        vb = P.TView(vT3, (3 * npad, d, heads, B), (1, 3 * npad, d * 3 * npad, C * 3 * npad))
        oo = P.TView(O.t, (d, n, heads, B, 1), (1, C, d, n * C, B * n * C))
        self.gemm(P.matrix_plan(pa, (128, 1, 1, 1), vb, oo, K=3 * npad, n_total=d, block_n=min(d, 128),
                                ext=(n, heads, B, 1), tiles=tiles, bz_sel=(3, 4), out_f32=True,
                                name=f"{name}.pv", flops=2 * B * heads * n * 3 * npad * d))
        wp = self.packed(lambda: P.pack_conv_weight(widen_weight(blk.proj.weight),
                                                    extra=[identity_weight(C, self.device)]))
        out = self.conv([P.ConvSource(self.raw(O, f"{name}.O")), P.ConvSource(self.raw(x, f"{name}.x"), taps=False)],
                        wp, C, 1, blk.proj.bias, name=f"{name}.proj+x")
        self.free(hn3, qkv, x)
        return out

    # ------------------------------------------------------------------ entry
    def forward(self, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        if self.params_changed():
            self.do_refresh()
        self.x_in.copy_(x)
        self.t_in.copy_(t)
        self.run()
        return self.out


class UNet2DSplitProgram(_SplitProgram):
    """slice_cond_2d_ddpm/unet.py:108-199 and ddpm_25d_all_modalities/unet.py:109-218 in split
    precision (their show_model / metrics scripts sample without autocast as well).  Post-norm
    blocks: every block input / output is an fp32 tensor; conv1 and the 1x1 res_conv read its plain
    widened copy, the second GroupNorm adds the fp32 residual."""

    def __init__(self, model, batch: int, spatial: Sequence[int], x_channels: int, ctx_channels: int):
        dev = next(model.parameters()).device
        super().__init__(dev, batch, spatial, groups=model.out_norm.num_groups, training=False)
        self.model = model
        B, (H, W) = batch, self.sp
        chs = list(model.chs)
        n_down = len(model.downs)
        if H % (2 ** n_down) or W % (2 ** n_down):
            raise _lib.MriError(f"split precision: image size {self.sp} must be divisible by {2 ** n_down}")
        eps = model.out_norm.eps
        cin = model.init_conv.weight.shape[1]
        if x_channels + ctx_channels != cin:
            raise _lib.MriError(f"init_conv expects {cin} input channels, got {x_channels} + {ctx_channels} context")
        self.x_in = torch.zeros(B, x_channels, H, W, device=dev)
        self.ctx_in = torch.zeros(B, ctx_channels, H, W, device=dev) if ctx_channels else None
        self.t_in = torch.zeros(B, dtype=torch.int64, device=dev)
        self.z_in = torch.zeros(B, 1, device=dev)

        # ---- conditioning: cond = time_mlp(t) + slice_mlp(z)  (unet.py:173-183), fp32 ---------------
        tdim = model.time_mlp[1].in_features
        temb = self.time_embedding(self.t_in, model.time_mlp, tdim)
        s0, s2 = model.slice_mlp[0], model.slice_mlp[2]
        self.track(s0.weight, s0.bias, s2.weight, s2.bias)
        zh = torch.zeros(B, s0.weight.shape[0], device=dev)
        cond = torch.zeros(B, tdim, device=dev)
        z_in = self.z_in
        self._add("slice_mlp.0", lambda: ops.linear(z_in, s0.weight, s0.bias, zh, act=1), [zh])
        self._add("slice_mlp.2", lambda: ops.linear(zh, s2.weight, s2.bias, cond, addend=temb), [cond])
        blocks = []
        for d in model.downs:
            blocks += [d.res1, d.res2]
        blocks += [model.mid_block1, model.mid_block2]
        for u in model.ups:
            blocks += [u.res1, u.res2]
        tproj, toffs, tld = self.block_projections(cond, blocks, act=1)   # SiLU on the projection (unet.py:48-50)
        self._tproj = {id(b): (tproj[:, o:], tld, o) for b, o in zip(blocks, toffs)}
        self.tape.clear()

        # ---- init_conv over the channels [x | lo(x) | x | ctx | lo(ctx) | ctx] ---------------------
        S = H * W
        ic = model.init_conv
        self.track(ic.weight, ic.bias)
        n2 = 2 * x_channels + 3 * ctx_channels
        buf = torch.zeros(B, n2, H, W, device=dev)
        x_in, ctx_in = self.x_in, self.ctx_in
        self._add("init_conv.lo", lambda: ops.bf16_residual_nchw(x_in, buf, B, x_channels * S, n2 * S), [buf])
        if ctx_channels:
            c0 = 2 * x_channels
            self._add("init_conv.ctx", lambda: ops.copy_cast(ctx_in.view(B, ctx_channels * S),
                                                             buf.view(B, n2 * S)[:, c0 * S:(c0 + ctx_channels) * S]), [buf])
            self._add("init_conv.ctx.lo", lambda: ops.bf16_residual_nchw(
                ctx_in, buf, B, ctx_channels * S, n2 * S, out_off=(c0 + ctx_channels) * S), [buf])
        kpad = _rup(9 * 3 * cin, 64)
        col = torch.zeros(B * S, kpad, dtype=torch.bfloat16, device=dev)
        self._add("init_conv.im2col", lambda: ops.im2col(x_in, col, B, x_channels, 1, H, W, 3, 2, kpad,
                                                         src2=buf, cin2=n2), [col])
        splits = [x_channels] + ([ctx_channels] if ctx_channels else [])
        w_in = self.packed(lambda: _pad_k(P.pack_conv_weight(widen_weight(ic.weight, splits=splits)), kpad))
        h = self.matrix_in_conv(col, w_in, S, kpad, chs[0], ic.bias, "init_conv", False)

        skips: List[F32Act] = []
        for i, d in enumerate(model.downs):
            h = self.resblock2d(h, None, d.res1, eps, f"downs.{i}.res1")
            h = self.resblock2d(h, None, d.res2, eps, f"downs.{i}.res2")
            h.keep = True
            skips.append(h)
            dn = d.down
            self.track(dn.weight, dn.bias)
            wd = self.packed(lambda dn=dn: P.pack_conv_weight(widen_weight(dn.weight)))
            y = self.new_f32([s // 2 for s in h.t.shape[1:-1]], dn.weight.shape[0], with_stats=False)
            self.gemm(P.down_conv_plan(self.raw(h, f"downs.{i}.down.in"), wd, y.t, bias=dn.bias, out_f32=True,
                                       name=f"downs.{i}.down"))
            h = y
        h = self.resblock2d(h, None, model.mid_block1, eps, "mid_block1")
        h = self.resblock2d(h, None, model.mid_block2, eps, "mid_block2")
        for j, u in enumerate(model.ups):
            skip = skips.pop()
            skip.keep = False
            up = u.up
            self.track(up.weight, up.bias)
            wu = self.packed(lambda up=up: P.pack_convT_weight(widen_weight(up.weight, transposed=True)))
            y = self.new_f32([s * 2 for s in h.t.shape[1:-1]], up.weight.shape[1], with_stats=False)
            self.gemm(P.up_conv_plan(self.raw(h, f"ups.{j}.up.in"), wu, y.t, bias=up.bias, out_f32=True,
                                     name=f"ups.{j}.up"))
            self.free(h)
            h = self.resblock2d(y, skip, u.res1, eps, f"ups.{j}.res1")
            h = self.resblock2d(h, None, u.res2, eps, f"ups.{j}.res2")

        on, oc = model.out_norm, model.out_conv
        self.track(on.weight, on.bias, oc.weight, oc.bias)
        h.stats = self.new_stats(self.groups)
        h.cpg = h.C // self.groups
        ht, hst, C0, G = h.t, h.stats, h.C, self.groups
        self._add("out_norm.stats", lambda: ops.stats_f32(ht, hst, B, S, C0, G), [hst])
        a3 = self.gn_split(h, on.weight, on.bias, self.groups, eps, True, name="out_norm")
        self.free(h)
        self.head(a3, oc, S)
        self.params_changed()

    def resblock2d(self, x: F32Act, skip: Optional[F32Act], blk, eps: float, name: str) -> F32Act:
        """ResidualBlock (slice_cond_2d_ddpm/unet.py:42-56), post-norm:
        h = silu(gn1(conv1(x))) + silu(lin(cond)); h = silu(gn2(conv2(h))); return h + res_conv(x).
        With `skip`, x is cat([x, skip], 1) (unet.py:101)."""
        n1, n2, c1, c2 = blk.norm1, blk.norm2, blk.conv1, blk.conv2
        self.track(n1.weight, n1.bias, n2.weight, n2.bias, c1.weight, c1.bias, c2.weight, c2.bias)
        cout = c1.weight.shape[0]
        rowbias, rb_ld, _ = self._tproj[id(blk)]
        srcs = [x] if skip is None else [x, skip]
        cins = [s.C for s in srcs]
        raws = [self.raw(s, f"{name}.in{k}") for k, s in enumerate(srcs)]
        w1 = self.packed(lambda: P.pack_conv_weight(widen_weight(c1.weight, splits=cins),
                                                    splits=[3 * c for c in cins]))
        h1 = self.conv([P.ConvSource(r) for r in raws], w1, cout, 3, c1.bias, name=f"{name}.conv1")
        a1 = self.gn_split(h1, n1.weight, n1.bias, self.groups, eps, True, f"{name}.norm1+temb",
                           rowbias=rowbias, rowbias_ld=rb_ld)
        self.free(h1)
        w2 = self.packed(lambda: P.pack_conv_weight(widen_weight(c2.weight)))
        h2 = self.conv([P.ConvSource(a1)], w2, cout, 3, c2.bias, name=f"{name}.conv2")
        self.free(a1)
        res_act = None
        if isinstance(blk.res_conv, torch.nn.Identity):
            assert skip is None and x.C == cout
            res = x.t
        else:
            rc = blk.res_conv
            self.track(rc.weight, rc.bias)
            wr = self.packed(lambda: P.pack_conv_weight(widen_weight(rc.weight, splits=cins),
                                                        splits=[3 * c for c in cins]))
            res_act = self.conv([P.ConvSource(r) for r in raws], wr, cout, 1, rc.bias, with_stats=False,
                                name=f"{name}.res_conv")
            res = res_act.t
        out = self.gn_f32(h2, n2.weight, n2.bias, self.groups, eps, True, res, f"{name}.norm2+res")
        self.free(h2, res_act, *srcs)
        return out

    # ------------------------------------------------------------------ entry
    def _load_inputs(self, x, t, z_pos, context):
        if self.params_changed():
            self.do_refresh()
        self.x_in.copy_(x)
        self.t_in.copy_(t)
        self.z_in.copy_(z_pos.reshape(-1, 1))
        if self.ctx_in is not None:
            self.ctx_in.copy_(context)

    def forward(self, x, t, z_pos, context=None) -> torch.Tensor:
        self._load_inputs(x, t, z_pos, context)
        self.run()
        return self.out


class VAE3DSplitProgram(_SplitProgram):
    """ddpm_3d_ldm/vae.py encode (`mode="encode"`, vae.py:49-55) / decode (`"decode"`, vae.py:82-87) in
    split precision: the reference's show_model.py decodes the sampled latents without autocast
    (show_model.py:255).  Same walk and the same channel padding as vae_engine.VAE3DProgram (32-channel
    levels run zero-padded to 64; GroupNorm(8, 32) over the padded tensor = 16 groups of 4)."""

    def __init__(self, vae, mode: str, batch: int, spatial: Sequence[int]):
        dev = next(vae.parameters()).device
        enc, dec = vae.encoder, vae.decoder
        first_block = enc.downs[0] if mode == "encode" else dec.ups[0]
        self.gn_groups = first_block.norm1.num_groups
        super().__init__(dev, batch, spatial, groups=self.gn_groups, training=False)
        self.mode = mode
        if len(self.sp) != 3:
            raise _lib.MriError("VAE3D expects 3 spatial dims")
        if mode == "encode":
            for s in self.sp:
                if s % (2 ** (enc.num_down - 1)) != 0:
                    raise _lib.MriError(f"volume size {self.sp} must be divisible by {2 ** (enc.num_down - 1)}")
            first, layers, last, prefix = enc.in_conv, enc.downs, enc.to_mu_logvar, "encoder"
        elif mode == "decode":
            first, layers, last, prefix = dec.from_latent, dec.ups, dec.out_conv, "decoder"
        else:
            raise ValueError(mode)
        B, (D, H, W) = batch, self.sp
        S = D * H * W
        cin, c0 = first.weight.shape[1], first.weight.shape[0]
        self.x_in = torch.zeros(B, cin, D, H, W, device=dev)
        self.track(first.weight, first.bias)
        x6 = torch.zeros(B, 2 * cin, D, H, W, device=dev)
        kpad = _rup(27 * 3 * cin, 64)
        col = torch.zeros(B * S, kpad, dtype=torch.bfloat16, device=dev)
        x_in, cp0 = self.x_in, self.cpad(c0)
        self._add(f"{prefix}.in.lo", lambda: ops.bf16_residual_nchw(x_in, x6, B, cin * S), [x6])
        self._add(f"{prefix}.in.im2col", lambda: ops.im2col(x_in, col, B, cin, D, H, W, 3, 3, kpad,
                                                            src2=x6, cin2=2 * cin), [col])
        w_in = self.packed(lambda: _pad_k(P.pack_conv_weight(widen_weight(first.weight), cout_pad=cp0), kpad))
        b_in = self.packed(lambda: _pad_vec(first.bias.detach(), cp0))
        h = self.matrix_in_conv(col, w_in, S, kpad, cp0, b_in, f"{prefix}.in", False)
        c_real = c0
        for i, layer in enumerate(layers):
            name = f"{prefix}.{i}"
            if isinstance(layer, torch.nn.ConvTranspose3d):
                c_real = layer.weight.shape[1]
                cip, cop = h.C, self.cpad(c_real)
                self.track(layer.weight, layer.bias)
                w = self.packed(lambda l=layer, a=cip, b=cop: P.pack_convT_weight(
                    widen_weight(self._pad2(l.weight.detach(), a, b), transposed=True), cout_pad=b))
                bb = self.packed(lambda l=layer, b=cop: _pad_vec(l.bias.detach(), b))
                y = self.new_f32([s * 2 for s in h.t.shape[1:-1]], cop, with_stats=False)
                self.gemm(P.up_conv_plan(self.raw(h, f"{name}.in"), w, y.t, bias=bb, out_f32=True, name=name))
                self.free(h)
                h = y
            elif isinstance(layer, torch.nn.Conv3d):
                c_real = layer.weight.shape[0]
                cip, cop = h.C, self.cpad(c_real)
                self.track(layer.weight, layer.bias)
                w = self.packed(lambda l=layer, a=cip, b=cop: P.pack_conv_weight(
                    widen_weight(self._pad2(l.weight.detach(), b, a)), cout_pad=b))
                bb = self.packed(lambda l=layer, b=cop: _pad_vec(l.bias.detach(), b))
                y = self.new_f32([s // 2 for s in h.t.shape[1:-1]], cop, with_stats=False)
                self.gemm(P.down_conv_plan(self.raw(h, f"{name}.in"), w, y.t, bias=bb, out_f32=True, name=name))
                self.free(h)
                h = y
            else:
                h, c_real = self.vae_resblock(h, layer, c_real, name)
        self.track(last.weight, last.bias)
        sp_out = tuple(h.t.shape[1:-1])
        self.head(self.raw(h, f"{prefix}.out.in"), last, sp_out[0] * sp_out[1] * sp_out[2], cin_pad=h.C)
        self.params_changed()

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def cpad(c: int) -> int:
        return _rup(c, 64)

    @staticmethod
    def _pad2(w: torch.Tensor, n0: int, n1: int) -> torch.Tensor:
        """Zero-pad dims 0 and 1 of a convolution weight to (n0, n1)."""
        if w.shape[0] == n0 and w.shape[1] == n1:
            return w
        out = torch.zeros(n0, n1, *w.shape[2:], dtype=w.dtype, device=w.device)
        out[:w.shape[0], :w.shape[1]] = w
        return out

    def norm_split(self, x: F32Act, norm, c_real: int, name: str) -> torch.Tensor:
        """[hi | lo | hi] of silu(GroupNorm(x)) over the padded tensor: groups of c_real / G channels
        (the all-zero padding groups give zeros: gamma = beta = 0 there)."""
        cpg = c_real // self.gn_groups
        if c_real % self.gn_groups:
            raise _lib.MriError(f"GroupNorm({self.gn_groups}, {c_real}) does not divide")
        cp = x.C
        G = cp // cpg
        x.stats, x.cpg = self.new_stats(G), cpg
        B, S, xs, st = self.B, x.spatial, x.t, x.stats
        self._add(f"{name}.stats", lambda: ops.stats_f32(xs, st, B, S, cp, G), [st])
        self.track(norm.weight, norm.bias)
        gm = self.packed(lambda: _pad_vec(norm.weight.detach(), cp))
        bt = self.packed(lambda: _pad_vec(norm.bias.detach(), cp))
        return self.gn_split(x, gm, bt, G, norm.eps, True, name)

    def vae_resblock(self, x: F32Act, blk, cin: int, name: str):
        """ResidualBlock3DNoTime.forward (vae.py:19-22)."""
        c1, c2 = blk.conv1, blk.conv2
        cout = c1.weight.shape[0]
        cip, cop = x.C, self.cpad(cout)
        self.track(c1.weight, c1.bias, c2.weight, c2.bias)
        a1 = self.norm_split(x, blk.norm1, cin, f"{name}.norm1")
        w1 = self.packed(lambda: P.pack_conv_weight(widen_weight(self._pad2(c1.weight.detach(), cop, cip))))
        b1 = self.packed(lambda: _pad_vec(c1.bias.detach(), cop))
        h = self.conv([P.ConvSource(a1)], w1, cop, 3, b1, with_stats=False, name=f"{name}.conv1")
        self.free(a1)
        a2 = self.norm_split(h, blk.norm2, cout, f"{name}.norm2")
        self.free(h)
        b2 = self.packed(lambda: _pad_vec(c2.bias.detach(), cop))
        xr = self.raw(x, f"{name}.in")
        if isinstance(blk.skip, torch.nn.Identity):
            w2 = self.packed(lambda: P.pack_conv_weight(widen_weight(self._pad2(c2.weight.detach(), cop, cop)),
                                                        extra=[identity_weight(cop, self.device)]))
            out = self.conv([P.ConvSource(a2), P.ConvSource(xr, taps=False)], w2, cop, 3, b2,
                            with_stats=False, name=f"{name}.conv2+x")
        else:
            sk = blk.skip
            self.track(sk.weight, sk.bias)
            w2 = self.packed(lambda: P.pack_conv_weight(
                widen_weight(self._pad2(c2.weight.detach(), cop, cop)),
                extra=[widen_weight(self._pad2(sk.weight.detach().reshape(cout, cin), cop, cip))]))
            bs = self.packed(lambda: _pad_vec(sk.bias.detach(), cop))
            out = self.conv([P.ConvSource(a2), P.ConvSource(xr, taps=False)], w2, cop, 3, b2, rowbias=bs,
                            rowbias_ld=0, with_stats=False, name=f"{name}.conv2+skip")
        self.free(a2, x)
        return out, cout

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.params_changed():
            self.do_refresh()
        self.x_in.copy_(x)
        self.run()
        return self.out
