"""Static-shape execution engine for the denoising UNets.

A :class:`UNetProgram` is built once per (model, batch, spatial size): it packs the weights into
the K-major bf16 matrices the tensor-core kernel wants, allocates every activation buffer
(channels-last bf16), builds one GemmPlan (TMA descriptors + k-table) per convolution and
records the launch sequence.  Running the program is a fixed list of C-ABI kernel launches on
the current CUDA stream, so it can be captured in a CUDA graph (see diffusion wrappers).

Fusions: conv epilogues add bias, the per-sample time-embedding projection and
the residual, and emit GroupNorm partial sums for the next norm; torch.cat is virtual (two TMA
sources in one K loop); the 1x1 skip convolutions of the 3D ResBlocks ride on conv2's K loop;
GroupNorm-apply + SiLU (+ time-embedding add + residual add for the 2D post-norm blocks) is one
bf16 pass.

Reference graphs: ddpm_3d_ldm/unet_attention.py:157-200 (UNet3DModelWithAttention.forward),
ddpm_3d_ldm/unet.py:115-158, slice_cond_2d_ddpm/unet.py:169-199,
ddpm_25d_all_modalities/unet.py:174-218.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, ops
from . import plan as P
from .backward import AttnRec, BackwardMixin, ConvRec, GnRec, ResizeRec, TimeRec, WindowRec


def _rup(n: int, m: int) -> int:
    return (n + m - 1) // m * m


@dataclass
class Act:
    """Channels-last bf16 activation [N, *spatial, C] (+ GroupNorm partial sums of its producer)."""
    t: torch.Tensor
    stats: Optional[torch.Tensor] = None  # fp64 [N, G, 2]
    cpg: int = 0                          # channels per statistics group

    @property
    def C(self) -> int:
        return self.t.shape[-1]

    @property
    def spatial(self) -> int:
        n = 1
        for s in self.t.shape[1:-1]:
            n *= s
        return n


class _Pool:
    """Recycles activation buffers by shape (the program is straight-line, so a buffer released
    by the builder is free for every later op in program order)."""

    def __init__(self, device):
        self.device = device
        self.free: Dict[Tuple, List[torch.Tensor]] = {}
        self.bytes = 0
        self.frozen = False

    def get(self, shape, dtype=torch.bfloat16) -> torch.Tensor:
        key = (tuple(shape), dtype)
        lst = self.free.get(key)
        if lst:
            return lst.pop()
        t = torch.zeros(*shape, dtype=dtype, device=self.device)
        self.bytes += t.numel() * t.element_size()
        return t

    def release(self, t: torch.Tensor) -> None:
        if self.frozen:  # training: every forward tensor is needed again by the backward pass
            return
        self.free.setdefault((tuple(t.shape), t.dtype), []).append(t)


class UNetProgram(BackwardMixin):
    """Common builder machinery; subclasses lay out a concrete UNet."""

    STATS_ARENA = 1 << 17  # doubles

    def __init__(self, device, batch: int, spatial: Sequence[int], groups: int = 8,
                 training: bool = False):
        _lib.require_device()
        self._binit()
        self.training = training
        self.device = device
        self.B = batch
        self.sp = tuple(int(s) for s in spatial)
        self.ndim = len(self.sp)
        self.groups = groups
        self.pool = _Pool(device)
        self.pool.frozen = training
        self.ops: List[Callable[[], None]] = []
        self.op_names: List[str] = []
        self.op_outs: List[List[torch.Tensor]] = []
        self.plans: List[P.GemmPlan] = []
        self.refresh: List[Callable[[], None]] = []  # re-pack weights after a parameter update
        self._packed: List[Tuple[torch.Tensor, Callable[[], torch.Tensor]]] = []
        self._gather = None
        self._rt_graphs: Dict[str, list] = {}
        self._arena = torch.zeros(self.STATS_ARENA, dtype=torch.float64, device=device)
        self._arena_used = 0
        self.gemm_flops = 0
        self.hbm_bytes_elementwise = 0
        self.gn_ops: List[Tuple[int, int]] = []
        self._param_versions: Optional[Tuple] = None
        self._params: List[torch.Tensor] = []
        self.fused_head: Optional[dict] = None   # set by thin_out_conv (inference programs)

    # ------------------------------------------------------------------ bookkeeping
    def _add(self, name: str, fn: Callable[[], None], outs: Sequence[torch.Tensor] = ()) -> None:
        self.op_names.append(name)
        self.ops.append(fn)
        self.op_outs.append([o for o in outs if o is not None])

    def run_traced(self) -> List[Tuple[str, List[torch.Tensor]]]:
        """Debug: run op by op (synchronising) and return a copy of every op's outputs."""
        trace = []
        ops.memset_zero(self._arena, max(self._arena_used, 4) * 8)
        for name, fn, outs in zip(self.op_names, self.ops, self.op_outs):
            fn()
            torch.cuda.synchronize(self.device)
            trace.append((name, [o.detach().clone() for o in outs]))
        return trace

    def new_stats(self, n_groups: int) -> torch.Tensor:
        n = self.B * n_groups * 2
        if self._arena_used + n > self.STATS_ARENA:
            raise _lib.MriError("statistics arena exhausted")
        s = self._arena[self._arena_used:self._arena_used + n].view(self.B, n_groups, 2)
        self._arena_used += _rup(n, 4)
        return s

    def new_act(self, sp: Sequence[int], C: int, with_stats: bool = True) -> Act:
        t = self.pool.get((self.B, *sp, C))
        if with_stats and C % (8 * self.groups) == 0:
            return Act(t, self.new_stats(self.groups), C // self.groups)
        return Act(t)

    def track(self, *params: torch.Tensor) -> None:
        self._params.extend(params)

    def packed(self, make: Callable[[], torch.Tensor]) -> torch.Tensor:
        """A derived weight buffer that is re-computed in place when parameters change."""
        buf = make()
        self.refresh.append(lambda: buf.copy_(make()))
        self._packed.append((buf, make))
        return buf

    # ------------------------------------------------------------------ refresh after a param update
    def do_refresh(self) -> None:
        """Bring every packed buffer up to date.  Almost all of them are pure gathers of parameter
        elements: those are refreshed by ONE mri_gather_pack launch driven by index maps that are
        derived once (see _build_gather); the few arithmetic ones (summed biases) re-run their
        torch expression."""
        if self._gather is None:
            self._gather = self._build_gather()
        table, n_segs, blocks, fallback, _keep = self._gather
        if n_segs:
            _lib.check(_lib.load().mri_gather_pack(table.data_ptr(), n_segs, blocks,
                                                   _lib.current_stream_ptr()), "mri_gather_pack")
        for i in fallback:
            self.refresh[i]()

    def _build_gather(self):
        """Derive, for each packed buffer, the index of the parameter element every output element
        copies.  The `make` closures are opaque torch expressions over the module parameters, so
        the map is found by probing: the parameters' data is temporarily replaced by base-128
        digits of each element's global address (+1; values 1..128 are exact in bf16 and fp32,
        0 stays reserved for padding), `make` is evaluated once per digit, and the digits read off
        the results.  A buffer whose map does not reproduce `make` on the real parameters
        (an arithmetic expression, not a gather) keeps its torch refresh."""
        import ctypes as C
        params, seen = [], set()
        for p_ in self._params:
            if id(p_) not in seen and p_.dtype == torch.float32:
                seen.add(id(p_))
                params.append(p_)
        offs, tot = [], 0
        for p_ in params:
            offs.append(tot)
            tot += p_.numel()
        n_dig = 1
        while (1 << (7 * n_dig)) <= tot + 1:
            n_dig += 1
        saved = [p_.data for p_ in params]
        addr = [None] * len(self._packed)
        try:
            with torch.no_grad():
                for k in range(n_dig):
                    for p_, off in zip(params, offs):
                        a = torch.arange(off + 1, off + 1 + p_.numel(), device=p_.device, dtype=torch.int64)
                        p_.data = ((((a >> (7 * k)) & 127) + 1).to(torch.float32)).view(p_.shape)
                    for i, (buf, make) in enumerate(self._packed):
                        if addr[i] is False:
                            continue
                        try:
                            d = make().reshape(-1).to(torch.float32)
                        except Exception:  # noqa: BLE001  (expression not defined on probe values)
                            addr[i] = False
                            continue
                        dl = d.round().to(torch.int64)
                        ok = bool(((d == dl.to(torch.float32)) & (dl >= 0) & (dl <= 128)).all().item())
                        if not ok or dl.numel() != buf.numel():
                            addr[i] = False
                            continue
                        if k == 0:
                            addr[i] = {"zero": dl == 0, "a": (dl - 1).clamp_min(0)}
                        else:
                            if not bool(((dl == 0) == addr[i]["zero"]).all().item()):
                                addr[i] = False
                                continue
                            addr[i]["a"] += (dl - 1).clamp_min(0) << (7 * k)
        finally:
            for p_, d in zip(params, saved):
                p_.data = d
        offs_t = torch.tensor(offs, device=self.device, dtype=torch.int64)
        segs, keep, fallback, blocks = [], [], [], 0
        flat = None
        for i, (buf, make) in enumerate(self._packed):
            info = addr[i]
            if not info or buf.dtype not in (torch.bfloat16, torch.float32) or not buf.is_contiguous():
                fallback.append(i)
                continue
            a = info["a"] - 1          # global element address, -1 for padding (a stored +1)
            a = torch.where(info["zero"], torch.full_like(a, -1), a)
            valid = a >= 0
            which = torch.searchsorted(offs_t, a.clamp_min(0), right=True) - 1
            used = torch.unique(which[valid]).tolist() if bool(valid.any().item()) else []
            if bool((a >= tot).any().item()):
                fallback.append(i)
                continue
            # a segment addresses at most 4 source tensors: a buffer that gathers from more (the
            # concatenated block projections) is cut into contiguous pieces of <= 4 sources each
            pieces = [(0, buf.numel(), used)]
            if len(used) > 4:
                wc = torch.where(valid, which, torch.full_like(which, -1)).cpu().numpy()
                cuts = np.flatnonzero(np.diff(wc)) + 1
                starts = np.concatenate([[0], cuts])
                ends = np.concatenate([cuts, [wc.size]])
                pieces, lo, cur = [], 0, []
                for st_, en_ in zip(starts.tolist(), ends.tolist()):
                    src_ = int(wc[st_])
                    if src_ >= 0 and src_ not in cur:
                        if len(cur) == 4:
                            pieces.append((lo, st_, cur))
                            lo, cur = st_, []
                        cur.append(src_)
                pieces.append((lo, wc.size, cur))
            local = torch.zeros_like(a)
            for pi in used:
                m = valid & (which == pi)
                local[m] = a[m] - offs[pi]
            if bool((local >= (1 << 28)).any().item()):
                fallback.append(i)
                continue
            slot = torch.zeros_like(a)
            for lo, hi, srcs_ in pieces:
                for s_, pi in enumerate(srcs_):
                    m = torch.zeros_like(valid)
                    m[lo:hi] = valid[lo:hi] & (which[lo:hi] == pi)
                    slot[m] = s_
            idx = torch.where(valid, (slot << 28) | local, torch.full_like(a, -1)).to(torch.int32)
            # verify against the real expression before trusting the map
            got = torch.zeros(buf.numel(), dtype=torch.float32, device=self.device)
            for pi in used:
                m = valid & (which == pi)
                got[m] = params[pi].detach().reshape(-1)[local[m]]
            if not torch.equal(got.to(buf.dtype), make().reshape(-1)):
                fallback.append(i)
                continue
            esz = buf.element_size()
            for lo, hi, srcs_ in pieces:
                if hi <= lo:
                    continue
                sg = _lib.MriGatherSeg()
                sg.dst, sg.idx, sg.n = buf.data_ptr() + lo * esz, idx.data_ptr() + lo * 4, hi - lo
                for s_, pi in enumerate(srcs_):
                    sg.src[s_] = params[pi].data_ptr()
                sg.block0 = blocks
                sg.dst_bf16 = 1 if buf.dtype == torch.bfloat16 else 0
                blocks += -(-(hi - lo) // 2048)
                segs.append(sg)
            keep.append(idx)
        table = None
        if segs:
            arr = (_lib.MriGatherSeg * len(segs))(*segs)
            raw = bytes(memoryview(arr))
            table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.device)
        self.gather_info = {"segments": len(segs), "fallback": len(fallback), "elements": sum(s_.n for s_ in segs)}
        return table, len(segs), blocks, fallback, keep

    def params_changed(self) -> bool:
        v = tuple(p._version for p in self._params)
        if self._param_versions is None:
            self._param_versions = v
            return False
        if v != self._param_versions:
            self._param_versions = v
            return True
        return False

    def run(self) -> None:
        """Enqueue the whole forward on the current stream (training programs: as one replayed
        CUDA graph after two eager runs -- the launch list and every buffer are static)."""
        if self.training:
            self._replay("fwd", self._run_eager)
        else:
            self._run_eager()

    def _run_eager(self) -> None:
        ops.memset_zero(self._arena, max(self._arena_used, 4) * 8)
        for fn in self.ops:
            fn()

    def run_fused_step(self, mode: int, **step_kw) -> None:
        """Inference programs with a thin output convolution: the whole forward EXCEPT the final
        tap gather, then mri_tap_gather_step, which finishes out_conv and applies the DDPM (mode 0)
        or DDIM (mode 1) update to self.x_in in place."""
        fh = self.fused_head
        assert fh["op_index"] == len(self.ops) - 1 and not self.training
        ops.memset_zero(self._arena, max(self._arena_used, 4) * 8)
        for fn in self.ops[:-1]:
            fn()
        ops.tap_gather_step(fh["y"], fh["bias"], self.B, fh["D"], fh["H"], fh["W"], fh["ndim"],
                            fh["cout"], fh["ldy"], self.x_in, mode, **step_kw)

    def _replay(self, key: str, body: Callable[[], None]) -> None:
        """`body` is a fixed launch sequence over static buffers: run it eagerly twice, then
        capture it once and replay the graph (removes ~10^2..10^3 launch overheads per step)."""
        if os.environ.get("MRI_NO_GRAPH") == "1" or torch.cuda.is_current_stream_capturing():
            body()
            return
        st = self._rt_graphs.setdefault(key, [0, None])
        st[0] += 1
        if st[1] is None:
            if st[0] < 3:
                body()
                return
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            # thread_local: the NCCL watchdog thread of an overlapped DDP step may query events
            # while this thread captures
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                body()
            st[1] = g
        st[1].replay()

    # ------------------------------------------------------------------ op emitters
    def gemm(self, pl: P.GemmPlan) -> None:
        pl.materialize(self.device)
        self.plans.append(pl)
        self.gemm_flops += pl.flops
        self._add(f"gemm:{pl.name}", pl.launch, [m.view.base for m in pl.o_maps[:1]] + [pl.stats])

    def gn(self, x: Act, gamma: torch.Tensor, beta: torch.Tensor, groups: int, eps: float,
           silu: bool, rowbias: Optional[torch.Tensor] = None, rowbias_ld: int = 0,
           residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
           name: str = "gn", c_off: int = 0, tproj_off: Optional[int] = None,
           gparam: Optional[torch.Tensor] = None, bparam: Optional[torch.Tensor] = None) -> torch.Tensor:
        """y = act(GroupNorm(x)) (+ rowbias[n, c]) (+ residual).  Uses x.stats (fine groups).
        gamma / beta are the module parameters; channels [c_off, c_off + C) of them apply."""
        assert x.stats is not None, f"{name}: input has no statistics"
        y = out if out is not None else self.pool.get(tuple(x.t.shape))
        B, S, C = self.B, x.spatial, x.C
        xs, st, cpg = x.t, x.stats, x.cpg
        gm, bt = gamma[c_off:c_off + C], beta[c_off:c_off + C]
        nbytes = 2 * xs.numel() * 2 + (residual.numel() * 2 if residual is not None else 0)
        self.hbm_bytes_elementwise += nbytes
        self.gn_ops.append((len(self.ops), nbytes))   # (op index, algorithmic HBM bytes) for bench.py

        def fn():
            ops.gn_apply(xs, y, st, gm, bt, B, S, C, groups, cpg, eps, silu, rowbias=rowbias,
                         rowbias_ld=rowbias_ld, residual=residual)

        self._add(name, fn, [y])
        self.tape.append(GnRec(x=x, y=y, gamma=gamma, beta=beta, c_off=c_off, groups=groups, eps=eps,
                               silu=silu, tproj_off=tproj_off, residual=residual, name=name,
                               gparam=gparam, bparam=bparam))
        return y

    def stats_of(self, x: Act, name: str = "gn_stats") -> None:
        """Standalone statistics pass for a tensor whose producer did not emit them."""
        x.stats = self.new_stats(self.groups)
        x.cpg = x.C // self.groups
        B, S, C, xs, st, cpg = self.B, x.spatial, x.C, x.t, x.stats, x.cpg
        self.hbm_bytes_elementwise += xs.numel() * 2
        self._add(name, lambda: ops.gn_stats(xs, st, B, S, C, cpg), [st])

    def conv(self, sources: Sequence[P.ConvSource], wmat: torch.Tensor, cout: int, ksize: int,
             bias: Optional[torch.Tensor], *, rowbias=None, rowbias_ld=0, residual=None,
             with_stats=True, name="conv", rec: Optional[dict] = None) -> Act:
        """Stride-1 convolution over (virtually concatenated) sources.  `rec` carries what the
        backward pass needs: weight (parameter), splits, extra_weight, bias_params, tproj_off,
        valid cout, need_dgrad."""
        sp = sources[0].x.shape[1:-1]
        y = self.new_act(sp, cout, with_stats)
        pl = P.conv_plan(sources, wmat, y.t, ksize, bias=bias, rowbias=rowbias, rowbias_ld=rowbias_ld,
                         residual=residual, stats=y.stats, stats_cpg=y.cpg, name=name)
        self.gemm(pl)
        if rec is not None:
            self.tape.append(ConvRec(kind="conv", plan=pl, y=y.t, ksize=ksize,
                                     sources=[(s_.x, s_.taps) for s_ in sources],
                                     weight=rec.get("weight"), splits=rec.get("splits", []),
                                     extra_weight=rec.get("extra_weight"),
                                     bias_params=rec.get("bias_params", []), residual=residual,
                                     tproj_off=rec.get("tproj_off"), cout=rec.get("cout", cout),
                                     need_dgrad=rec.get("need_dgrad", True),
                                     dgrad_dy=rec.get("dgrad_dy"), unpack=rec.get("unpack"),
                                     unpack_extra=rec.get("unpack_extra"), wfull=rec.get("wfull"),
                                     efull=rec.get("efull"), name=name))
        return y

    def thin_patch_matrix(self, x_in: torch.Tensor, conv, sp: Sequence[int], name: str):
        """Inference: patch matrix of a thin-channel input convolution through a channels-last
        bf16 copy padded to a multiple of 4 channels (mri_nchw_to_nhwc + mri_im2col4: 8-byte
        gathers instead of per-element ones).  Returns (col [B*S, kpad], packed weights
        [Cout, kpad] with column order (tap, padded channel), kpad)."""
        B = self.B
        nd = len(sp)
        cin, ksize = conv.weight.shape[1], conv.weight.shape[2]
        cp = _rup(cin, 4)
        S = 1
        for e in sp:
            S *= e
        kpad = _rup((ksize ** nd) * cp, 64)
        x4 = torch.zeros(B, S, cp, dtype=torch.bfloat16, device=self.device)
        col = torch.zeros(B * S, kpad, dtype=torch.bfloat16, device=self.device)
        sp3 = (1,) * (3 - nd) + tuple(sp)
        self._add(f"{name}.nhwc", lambda: ops.nchw_to_nhwc(x_in, x4, B, S, cin, cp), [x4])
        self._add(f"{name}.im2col",
                  lambda: ops.im2col4(x4, col, B, cp, sp3[0], sp3[1], sp3[2], ksize, nd, kpad), [col])

        def make():
            w = conv.weight.detach()
            if cp != cin:
                wp = torch.zeros(w.shape[0], cp, *w.shape[2:], dtype=w.dtype, device=w.device)
                wp[:, :cin] = w
                w = wp
            return _pad_k(P.pack_conv_weight(w), kpad)

        return col, self.packed(make), kpad

    def thin_in_conv(self, x_in: torch.Tensor, conv, sp: Sequence[int], name: str, with_stats: bool) -> Optional[Act]:
        """Inference: the first convolution (<= 4 input channels, 64 / 128 output channels, 3^d taps)
        as ONE launch of mri_thin_in_conv -- no layout pass, no im2col buffer, no generic GEMM whose
        K = 128 tiles are epilogue-bound.  Returns None when the shape is not covered."""
        cout, cin, ksize = conv.weight.shape[0], conv.weight.shape[1], conv.weight.shape[2]
        nd = len(sp)
        if (self.training or cin > 4 or cout not in (64, 128) or ksize != 3
                or os.environ.get("MRI_THIN_IN_CONV", "1") == "0"):
            return None
        B = self.B
        self.track(conv.weight, conv.bias)

        def make():
            w = conv.weight.detach()
            wp = torch.zeros(cout, 4, *w.shape[2:], dtype=w.dtype, device=w.device)
            wp[:, :cin] = w
            return _pad_k(P.pack_conv_weight(wp), 128)

        w128 = self.packed(make)
        y = self.new_act(sp, cout, with_stats)
        sp3 = (1,) * (3 - nd) + tuple(sp)
        bias, st = conv.bias, y.stats
        self._add(f"tc:{name}", lambda: ops.thin_in_conv(x_in, w128, bias, y.t, st, B, cin, sp3[0], sp3[1],
                                                         sp3[2], nd, cout), [y.t, st])
        S = sp3[0] * sp3[1] * sp3[2]
        self.gemm_flops += 2 * B * S * cout * (ksize ** nd) * cin
        return y

    def thin_out_conv(self, a: torch.Tensor, oc, name: str = "out_conv", cin_pad: int = 0) -> None:
        """Inference head for a k^d convolution with 1..8 output channels (out_conv): one GEMM
        computes every tap's product Y[q][tap*cout + co] = W[tap][co] . a[q] (K = Cin, the
        activation is read once instead of once per tap), mri_tap_gather sums the shifted taps.
        `a`: [B, *sp, Cin(_pad)] bf16.  Sets self.eps_nhwc / self.cout / self.cout_pad."""
        cout = oc.weight.shape[0]
        ksize = oc.weight.shape[2]
        sp = tuple(a.shape[1:-1])
        nd = len(sp)
        taps = ksize ** nd
        n_y = _rup(taps * cout, 16) if taps * cout <= 48 else _rup(taps * cout, 64)
        cin_pad = cin_pad or oc.weight.shape[1]
        self.track(oc.weight, oc.bias)

        def make():
            w = oc.weight.detach()
            if w.shape[1] != cin_pad:
                wp = torch.zeros(w.shape[0], cin_pad, *w.shape[2:], dtype=w.dtype, device=w.device)
                wp[:, :w.shape[1]] = w
                w = wp
            return P.pack_tap_weight(w, n_y)

        w_exp = self.packed(make)
        y = self.conv([P.ConvSource(a)], w_exp, n_y, 1, None, with_stats=False, name=f"{name}.taps")
        ldo = _rup(cout, 4)
        eps = self.pool.get((self.B, *sp, ldo))
        sp3 = (1,) * (3 - nd) + sp
        B, bias = self.B, oc.bias
        self._add(f"{name}.gather",
                  lambda: ops.tap_gather(y.t, eps, bias, B, sp3[0], sp3[1], sp3[2], ksize, nd, cout,
                                         n_y, ldo), [eps])
        # the samplers replace this last launch by mri_tap_gather_step (gather + reverse-step update
        # in one kernel: the predicted noise never goes to HBM); see DiffusionBase._reverse_loop_on
        if ksize == 3 and 1 <= cout <= 4:
            self.fused_head = dict(y=y.t, bias=bias, D=sp3[0], H=sp3[1], W=sp3[2], ndim=nd, cout=cout,
                                   ldy=n_y, op_index=len(self.ops) - 1)
        self.cout, self.cout_pad, self.eps_nhwc = cout, ldo, eps
        self.hbm_bytes_elementwise += y.t.numel() * 2 + eps.numel() * 2

    # ------------------------------------------------------------------ time embedding
    def time_embedding(self, t_in: torch.Tensor, time_mlp, dim: int) -> torch.Tensor:
        """SinusoidalPosEmb -> Linear -> SiLU -> Linear  (unet.py:124-129 / unet_attention.py:103-108)."""
        B, dev = self.B, self.device
        sin = torch.zeros(B, dim, device=dev)
        z1 = torch.zeros(B, dim * 4, device=dev)
        h1 = torch.zeros(B, dim * 4, device=dev)
        temb = torch.zeros(B, dim, device=dev)
        l1, l2 = time_mlp[1], time_mlp[3]
        self.track(l1.weight, l1.bias, l2.weight, l2.bias)
        freqs = ops.sinusoidal_freqs(dim, dev)
        self._add("sinusoidal", lambda: ops.sinusoidal(t_in, freqs, sin), [sin])
        if self.training:  # keep the pre-activation for the backward pass
            self._add("time_mlp.1", lambda: ops.linear(sin, l1.weight, l1.bias, z1), [z1])
            self._add("time_mlp.silu", lambda: ops.silu(z1, h1), [h1])
        else:
            self._add("time_mlp.1", lambda: ops.linear(sin, l1.weight, l1.bias, h1, act=1), [h1])
        self._add("time_mlp.3", lambda: ops.linear(h1, l2.weight, l2.bias, temb), [temb])
        self._time = dict(sin=sin, z1=z1, h1=h1, temb=temb, l1=l1, l2=l2)
        return temb

    def block_projections(self, cond: torch.Tensor, blocks: Sequence, act: int,
                          slice_rec: Optional[dict] = None) -> Tuple[torch.Tensor, List[int], int]:
        """All per-ResBlock Linear(t_dim, Cout) projections as ONE concatenated linear."""
        Ws = [b.time_mlp.weight for b in blocks]
        bs = [b.time_mlp.bias for b in blocks]
        self.track(*Ws, *bs)
        W_all = self.packed(lambda: torch.cat([w.detach() for w in Ws], 0).contiguous())
        b_all = self.packed(lambda: torch.cat([b.detach() for b in bs], 0).contiguous())
        total = W_all.shape[0]
        out = torch.zeros(self.B, total, device=self.device)
        zproj = torch.zeros(self.B, total, device=self.device)
        if self.training and act == 1:
            self._add("time_proj_all", lambda: ops.linear(cond, W_all, b_all, zproj), [zproj])
            self._add("time_proj_all.silu", lambda: ops.silu(zproj, out), [out])
        else:
            self._add("time_proj_all", lambda: ops.linear(cond, W_all, b_all, out, act=act), [out])
        offs, o = [], 0
        for w in Ws:
            offs.append(o)
            o += w.shape[0]
        self.dtproj = torch.zeros(self.B, total, device=self.device)
        t = self._time
        # recorded FIRST on the tape -> processed LAST in the backward pass
        self.tape.insert(0, TimeRec(sin=t["sin"], z1=t["z1"], h1=t["h1"], temb=t["temb"], l1=t["l1"],
                                    l2=t["l2"], cond=cond, W_all=W_all, zproj=zproj, act=act,
                                    blocks=list(blocks), offs=offs, slice=slice_rec))
        return out, offs, total


# ==========================================================================================
# 3D latent UNet (pre-norm ResBlocks, optional bottleneck attention)
# ==========================================================================================
class UNet3DProgram(UNetProgram):
    """ddpm_3d_ldm/unet_attention.py:88-200 and ddpm_3d_ldm/unet.py:57-158."""

    def __init__(self, model, batch: int, spatial: Sequence[int], training: bool = False):
        dev = next(model.parameters()).device
        super().__init__(dev, batch, spatial, groups=model.out_norm.num_groups, training=training)
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

        # ---- time embedding + all block projections ---------------------------------------
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

        # ---- in_conv: thin Cin -> explicit patch matrix + GEMM ------------------------------
        S = D * H * W
        ic = model.in_conv
        self.track(ic.weight, ic.bias)
        if training:
            kpad = _rup(27 * cin, 64)
            col = torch.zeros(B * S, kpad, dtype=torch.bfloat16, device=dev)
            x_in = self.x_in
            self._add("im2col", lambda: ops.im2col(x_in, col, B, cin, D, H, W, 3, 3, kpad), [col])
            w_in = self.packed(lambda: _pad_k(P.pack_conv_weight(ic.weight.detach()), kpad))
        h = self.thin_in_conv(self.x_in, ic, self.sp, "in_conv", with_stats=True)
        if h is None:
            if not training:
                col, w_in, kpad = self.thin_patch_matrix(self.x_in, ic, self.sp, "in_conv")
            h = self.new_act(self.sp, chs[0])
            pl = self._matrix_conv(col, w_in, h, S, kpad, ic.bias, "in_conv")
            self.gemm(pl)
            self.tape.append(ConvRec(kind="matrix", plan=pl, y=h.t, ksize=3, sources=[(col, True)],
                                     weight=ic.weight, splits=[cin], bias_params=[ic.bias], cout=chs[0],
                                     need_dgrad=False, kpad=kpad, name="in_conv"))

        # ---- down path -----------------------------------------------------------------------
        skips: List[Act] = []
        for i, blk in enumerate(model.downs):
            h = self.resblock(h, None, blk["res1"], eps, f"downs.{i}.res1")
            h = self.resblock(h, None, blk["res2"], eps, f"downs.{i}.res2")
            skips.append(h)
            if i != L - 1:
                dn = blk["down"]
                self.track(dn.weight, dn.bias)
                wd = self.packed(lambda dn=dn: P.pack_conv_weight(dn.weight.detach()))
                y = self.new_act([s // 2 for s in h.t.shape[1:-1]], chs[i + 1])
                pl = P.down_conv_plan(h.t, wd, y.t, bias=dn.bias, stats=y.stats, stats_cpg=y.cpg,
                                      name=f"downs.{i}.down")
                self.gemm(pl)
                self.tape.append(ConvRec(kind="down", plan=pl, y=y.t, ksize=4, sources=[(h.t, True)],
                                         weight=dn.weight, splits=[h.C], bias_params=[dn.bias],
                                         cout=chs[i + 1], name=f"downs.{i}.down"))
                h = y

        # ---- bottleneck ------------------------------------------------------------------------
        h = self.resblock(h, None, model.mid1, eps, "mid1", release_in=False)
        if hasattr(model, "mid_attn"):
            h = self.attention(h, model.mid_attn, "mid_attn")
        h = self.resblock(h, None, model.mid2, eps, "mid2")

        # ---- up path -----------------------------------------------------------------------------
        for j, blk in enumerate(model.ups):
            i = L - 1 - j
            if i != L - 1:
                up = blk["up"]
                self.track(up.weight, up.bias)
                wu = self.packed(lambda up=up: P.pack_convT_weight(up.weight.detach()))
                u = self.new_act([s * 2 for s in h.t.shape[1:-1]], chs[i])
                pl = P.up_conv_plan(h.t, wu, u.t, bias=up.bias, stats=u.stats, stats_cpg=u.cpg,
                                    name=f"ups.{j}.up")
                self.gemm(pl)
                self.tape.append(ConvRec(kind="up", plan=pl, y=u.t, ksize=4, sources=[(h.t, True)],
                                         weight=up.weight, splits=[h.C], bias_params=[up.bias],
                                         cout=chs[i], name=f"ups.{j}.up"))
                self.pool.release(h.t)
                h = u
            skip = skips.pop()
            if tuple(skip.t.shape[1:-1]) != tuple(h.t.shape[1:-1]):
                raise _lib.MriError("skip/upsample shape mismatch (centre-crop path not supported "
                                    "for sizes not divisible by the down-sampling factor)")
            h = self.resblock(h, skip, blk["res1"], eps, f"ups.{j}.res1")
            h = self.resblock(h, None, blk["res2"], eps, f"ups.{j}.res2")

        # ---- head --------------------------------------------------------------------------------
        on, oc = model.out_norm, model.out_conv
        self.track(on.weight, on.bias, oc.weight, oc.bias)
        a = self.gn(h, on.weight, on.bias, self.groups, eps, True, name="out_norm")
        self.cout = oc.weight.shape[0]
        self.deps64 = None
        if not training and self.cout <= 4:
            self.thin_out_conv(a, oc)
        else:
            self.cout_pad = _rup(self.cout, 16)
            w_out = self.packed(lambda: P.pack_conv_weight(oc.weight.detach(), cout_pad=self.cout_pad))
            b_out = self.packed(lambda: _pad_vec(oc.bias.detach(), self.cout_pad))
            if training:
                self.deps64 = torch.zeros(B, D, H, W, 64, dtype=torch.bfloat16, device=dev)
            y = self.conv([P.ConvSource(a)], w_out, self.cout_pad, 3, b_out, with_stats=False,
                          name="out_conv",
                          rec=dict(weight=oc.weight, splits=[a.shape[-1]], bias_params=[oc.bias],
                                   cout=self.cout, dgrad_dy=self.deps64))
            self.eps_nhwc = y.t  # [B, D, H, W, cout_pad] bf16
        self.out = torch.zeros(B, self.cout, D, H, W, device=dev)
        self.params_changed()
        if training:
            self.dout_in = torch.zeros(B, self.cout, D, H, W, device=dev)
            self.deps16 = torch.zeros_like(self.eps_nhwc)
            self.build_backward({id(self.eps_nhwc): self.deps16})

    # -------------------------------------------------------------------------------------------
    def _matrix_conv(self, col, wmat, y: Act, S: int, kpad: int, bias, name) -> P.GemmPlan:
        B, C = self.B, y.C
        a = P.TView(col, (kpad, S, B, 1, 1), (1, kpad, S * kpad, B * S * kpad, B * S * kpad))
        b = P.TView(wmat, (kpad, C, 1, 1), (1, kpad, kpad * C, kpad * C))
        o = P.TView(y.t, (C, S, B, 1, 1), (1, C, S * C, B * S * C, B * S * C))
        bn = P.pick_block_n(C)
        return P.matrix_plan(a, (128, 1, 1, 1), b, o, K=kpad, n_total=C, block_n=bn,
                             ext=(S, B, 1, 1), tiles=(-(-S // 128), B, 1, 1), sample_dim=2,
                             bias=bias, stats=y.stats, stats_cpg=y.cpg, name=name,
                             flops=2 * B * S * C * kpad)

    def resblock(self, x: Act, skip: Optional[Act], blk, eps: float, name: str,
                 release_in: bool = True) -> Act:
        """ResidualBlock3D (unet_attention.py:59-85).  With `skip`, the block input is the
        virtual concatenation [x, skip] (unet_attention.py:195)."""
        n1, n2, c1, c2 = blk.norm1, blk.norm2, blk.conv1, blk.conv2
        self.track(n1.weight, n1.bias, n2.weight, n2.bias, c1.weight, c1.bias, c2.weight, c2.bias)
        cout = c1.weight.shape[0]
        rowbias, rb_ld, t_off = self._tproj[id(blk)]
        srcs = [x] if skip is None else [x, skip]
        cins = [s.C for s in srcs]
        g_each = self.groups // len(srcs)  # GN(8, 2C) over a concat == GN(4)+GN(4) over halves
        normed, c0 = [], 0
        for k, s in enumerate(srcs):
            normed.append(self.gn(s, n1.weight, n1.bias, g_each, eps, True, c_off=c0,
                                  name=f"{name}.norm1.{k}"))
            c0 += s.C
        w1 = self.packed(lambda: P.pack_conv_weight(c1.weight.detach(), splits=cins))
        h = self.conv([P.ConvSource(a) for a in normed], w1, cout, 3, c1.bias, rowbias=rowbias,
                      rowbias_ld=rb_ld, name=f"{name}.conv1",
                      rec=dict(weight=c1.weight, splits=cins, bias_params=[c1.bias], tproj_off=t_off))
        for a in normed:
            self.pool.release(a)
        a2 = self.gn(h, n2.weight, n2.bias, self.groups, eps, True, name=f"{name}.norm2")
        self.pool.release(h.t)
        has_skip_conv = not isinstance(blk.skip, torch.nn.Identity)
        if has_skip_conv:
            sk = blk.skip
            self.track(sk.weight, sk.bias)
            extras = []
            c0 = 0
            for s in srcs:
                extras.append((c0, s.C))
                c0 += s.C
            w2 = self.packed(lambda: P.pack_conv_weight(
                c2.weight.detach(),
                extra=[sk.weight.detach().reshape(cout, -1)[:, a:a + n] for a, n in extras]))
            # the two bias vectors are added in the epilogue (skip bias as a row bias whose row
            # pitch is 0: the same row for every sample) -- no derived "sum of parameters" buffer
            # that would need a torch expression to refresh after every optimizer step
            sources = [P.ConvSource(a2)] + [P.ConvSource(s.t, taps=False) for s in srcs]
            out = self.conv(sources, w2, cout, 3, c2.bias, rowbias=sk.bias, rowbias_ld=0,
                            name=f"{name}.conv2+skip",
                            rec=dict(weight=c2.weight, splits=[cout], extra_weight=sk.weight,
                                     bias_params=[c2.bias, sk.bias]))
        else:
            assert skip is None and x.C == cout
            w2 = self.packed(lambda: P.pack_conv_weight(c2.weight.detach()))
            out = self.conv([P.ConvSource(a2)], w2, cout, 3, c2.bias, residual=x.t,
                            name=f"{name}.conv2",
                            rec=dict(weight=c2.weight, splits=[cout], bias_params=[c2.bias]))
        self.pool.release(a2)
        if release_in:
            for s in srcs:
                self.pool.release(s.t)
        return out

    def attention(self, x: Act, blk, name: str) -> Act:
        """AttentionBlock3D (unet_attention.py:28-56): GN -> qkv 1x1 -> softmax(q^T k / sqrt(d)) v
        -> proj 1x1 -> + x.  Layouts: qkv token-major [B, n, 3C]; k and v additionally transposed
        ([B, 2C, n], computed by a second GEMM W . hn^T) so that every product is K-major."""
        B, C, dev = self.B, x.C, self.device
        heads = blk.num_heads
        d = C // heads
        assert d % 64 == 0 and d <= 256
        n = x.spatial
        npad = _rup(n, 8)
        C3 = 3 * C
        self.track(blk.norm.weight, blk.norm.bias, blk.qkv.weight, blk.qkv.bias, blk.proj.weight,
                   blk.proj.bias)
        hn = self.gn(x, blk.norm.weight, blk.norm.bias, self.groups, blk.norm.eps, False,
                     name=f"{name}.norm")
        wqkv = self.packed(lambda: P.pack_conv_weight(blk.qkv.weight.detach()))
        qkv = self.conv([P.ConvSource(hn)], wqkv, C3, 1, blk.qkv.bias, with_stats=False,
                        name=f"{name}.qkv")
        qkv_plan = self.plans[-1]
        if (not self.training and d in (64, 128) and _lib.load().mri_attn_flash_supported(d)
                and os.environ.get("MRI_ATTN_FLASH", "1") != "0"):
            return self._attention_flash(x, blk, name, hn, qkv, heads, d, n, npad)
        # [k^T; v^T][b] = W[C:3C] . hn[b]^T  (+ bias along rows): classes = samples
        wkv = self.packed(lambda: blk.qkv.weight.detach()[C:].reshape(2 * C, C).to(torch.bfloat16).contiguous())
        bkv = self.packed(lambda: blk.qkv.bias.detach()[C:].contiguous())
        kvT = torch.zeros(B, 2 * C, npad, dtype=torch.bfloat16, device=dev)
        a = P.TView(wkv, (C, 2 * C, 1, 1, 1), (1, C, 2 * C * C, 2 * C * C, 2 * C * C))
        b = P.TView(hn, (C, n, B, 1), (1, C, n * C, B * n * C))
        sz = 2 * C * npad
        o_views = [P.TView(kvT, (npad, 2 * C, 1, 1, 1), (1, npad, sz, sz, sz), offset=bi * sz)
                   for bi in range(B)]
        pl = P.matrix_plan(a, (128, 1, 1, 1), b, o_views[0], K=C, n_total=npad, block_n=128,
                           ext=(2 * C, 1, 1, 1), tiles=(2 * C // 128, 1, 1, 1), bz_sel=(1, 0),
                           bias_m=bkv, name=f"{name}.kvT", flops=2 * B * n * 2 * C * C)
        chunk_box, swz = pl.o_maps[0].box, pl.o_maps[0].swizzle
        pl.o_maps = [P.MapSpec(v, chunk_box, swz) for v in o_views]
        pl.ktable = pl.ktable.repeat(B, axis=0)
        self.gemm(pl)
        # S = q k^T  (fp32 logits)
        S = torch.zeros(B, heads, n, npad, dtype=torch.float32, device=dev)
        qa = P.TView(qkv.t, (d, n, heads, B, 1), (1, C3, d, n * C3, B * n * C3))
        kb = P.TView(qkv.t, (d, n, heads, B), (1, C3, d, n * C3), offset=C)
        so = P.TView(S, (npad, n, heads, B, 1),
                     (1, npad, n * npad, heads * n * npad, B * heads * n * npad))
        self.gemm(P.matrix_plan(qa, (128, 1, 1, 1), kb, so, K=d, n_total=npad, block_n=128,
                                ext=(n, heads, B, 1), tiles=(-(-n // 128), heads, B, 1),
                                bz_sel=(3, 4), out_f32=True, name=f"{name}.qk^T",
                                flops=2 * B * heads * n * n * d))
        Pm = torch.zeros(B, heads, n, npad, dtype=torch.bfloat16, device=dev)
        scale = float(d) ** -0.5
        self._add(f"{name}.softmax",
                  lambda: ops.softmax_rows(S, Pm, B * heads * n, n, npad, npad, scale), [Pm])
        # O = P v  (token-major [B, n, C]); B operand = v^T rows of kvT
        O = self.pool.get(tuple(x.t.shape))
        pa = P.TView(Pm, (npad, n, heads, B, 1),
                     (1, npad, n * npad, heads * n * npad, B * heads * n * npad))
        vb = P.TView(kvT, (npad, d, heads, B), (1, npad, d * npad, sz), offset=C * npad)
        oo = P.TView(O, (d, n, heads, B, 1), (1, C, d, n * C, B * n * C))
        self.gemm(P.matrix_plan(pa, (128, 1, 1, 1), vb, oo, K=npad, n_total=d, block_n=min(d, 128),
                                ext=(n, heads, B, 1), tiles=(-(-n // 128), heads, B, 1),
                                bz_sel=(3, 4), name=f"{name}.pv", flops=2 * B * heads * n * n * d))
        wp = self.packed(lambda: P.pack_conv_weight(blk.proj.weight.detach()))
        out = self.conv([P.ConvSource(O)], wp, C, 1, blk.proj.bias, residual=x.t, name=f"{name}.proj")
        proj_plan = self.plans[-1]
        self.tape.append(AttnRec(x=x, hn=hn, qkv=qkv.t, kvT=kvT, Pm=Pm, O=O, out=out.t, blk=blk,
                                 qkv_plan=qkv_plan, proj_plan=proj_plan, heads=heads, n=n, npad=npad,
                                 name=name))
        # the norm's GnRec was recorded before the attention record; backward order must be
        # attention first, then the norm: move the norm record after... (tape is reversed, so
        # the later-recorded AttnRec is already processed first)
        self.pool.release(hn)
        self.pool.release(qkv.t)
        self.pool.release(O)
        self.pool.release(x.t)
        return out

    def _attention_flash(self, x: Act, blk, name: str, hn: torch.Tensor, qkv: Act, heads: int,
                         d: int, n: int, npad: int) -> Act:
        """Inference: softmax(q^T k / sqrt(d)) v in ONE kernel (mri_attn_flash_launch: S and O in
        TMEM, online softmax between the two MMAs; no logits / probabilities in HBM).  q and k are
        read straight from the token-major qkv rows; v^T [B, C, npad] comes from one extra GEMM
        W_v . hn^T (keys contiguous = the K-major B operand of P V)."""
        import ctypes as C_
        B, C, dev = self.B, x.C, self.device
        C3 = 3 * C
        wv = self.packed(lambda: blk.qkv.weight.detach()[2 * C:].reshape(C, C).to(torch.bfloat16).contiguous())
        bv = self.packed(lambda: blk.qkv.bias.detach()[2 * C:].contiguous())
        vT = torch.zeros(B, C, npad, dtype=torch.bfloat16, device=dev)
        a = P.TView(wv, (C, C, 1, 1, 1), (1, C, C * C, C * C, C * C))
        b = P.TView(hn, (C, n, B, 1), (1, C, n * C, B * n * C))
        sz = C * npad
        o_views = [P.TView(vT, (npad, C, 1, 1, 1), (1, npad, sz, sz, sz), offset=bi * sz)
                   for bi in range(B)]
        pl = P.matrix_plan(a, (128, 1, 1, 1), b, o_views[0], K=C, n_total=npad, block_n=128,
                           ext=(C, 1, 1, 1), tiles=(C // 128, 1, 1, 1), bz_sel=(1, 0),
                           bias_m=bv, name=f"{name}.vT", flops=2 * B * n * C * C)
        chunk_box, swz = pl.o_maps[0].box, pl.o_maps[0].swizzle
        pl.o_maps = [P.MapSpec(v, chunk_box, swz) for v in o_views]
        pl.ktable = pl.ktable.repeat(B, axis=0)
        self.gemm(pl)
        O = self.pool.get(tuple(x.t.shape))
        maps = P.encode_maps([
            P.MapSpec(P.TView(qkv.t, (C3, n, B), (1, C3, n * C3)), (64, 128, 1), 3),
            P.MapSpec(P.TView(vT, (npad, C, B), (1, npad, C * npad)), (64, d, 1), 3)], dev)
        args = _lib.MriAttnArgs()
        args.qk_map, args.vt_map, args.out = maps.data_ptr(), maps.data_ptr() + 128, O.data_ptr()
        args.batch, args.heads, args.n, args.d, args.C = B, heads, n, d, C
        args.k_col0, args.v_row0, args.ld_out = C, 0, C
        args.scale = float(d) ** -0.5
        self._keep_alive = getattr(self, "_keep_alive", []) + [maps, args, vT]
        lib = _lib.load()

        def flash():
            _lib.check(lib.mri_attn_flash_launch(C_.byref(args), _lib.current_stream_ptr()),
                       f"mri_attn_flash_launch[{name}]")

        self._add(f"attn:{name}.flash", flash, [O])
        self.gemm_flops += 2 * 2 * B * heads * n * n * d
        wp = self.packed(lambda: P.pack_conv_weight(blk.proj.weight.detach()))
        out = self.conv([P.ConvSource(O)], wp, C, 1, blk.proj.bias, residual=x.t, name=f"{name}.proj")
        self.pool.release(hn)
        self.pool.release(qkv.t)
        self.pool.release(O)
        self.pool.release(x.t)
        return out

    # -------------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """Eager entry: copy inputs into the static buffers, run, return fp32 NCDHW eps."""
        if self.params_changed():
            self.do_refresh()
        self.x_in.copy_(x)
        self.t_in.copy_(t)
        self.run()
        S = self.sp[0] * self.sp[1] * self.sp[2]
        ops.nhwc_to_nchw(self.eps_nhwc, self.out, self.B, S, self.cout, self.cout_pad)
        return self.out

    def backward(self, dout: torch.Tensor, sync=None) -> None:
        """Training: gradient of the loss w.r.t. the fp32 NCDHW output -> parameter gradients in
        self.pgrad (fp32, reference layouts).  `sync` (parallel.GradSync): all-reduce finished
        gradient buckets while the rest of the backward launch list runs."""
        self._backward(dout, self.sp[0] * self.sp[1] * self.sp[2], sync)


def _pad_k(w: torch.Tensor, kpad: int) -> torch.Tensor:
    out = torch.zeros(w.shape[0], kpad, dtype=w.dtype, device=w.device)
    out[:, :w.shape[1]] = w
    return out


def _pad_vec(v: torch.Tensor, n: int) -> torch.Tensor:
    out = torch.zeros(n, dtype=v.dtype, device=v.device)
    out[:v.numel()] = v
    return out


# ==========================================================================================
# 2D / 2.5D UNet (post-norm ResBlocks, slice-position conditioning, optional context channels)
# ==========================================================================================
class UNet2DProgram(UNetProgram):
    """slice_cond_2d_ddpm/unet.py:108-199 and ddpm_25d_all_modalities/unet.py:109-218."""

    def __init__(self, model, batch: int, spatial: Sequence[int], x_channels: int, ctx_channels: int,
                 training: bool = False):
        dev = next(model.parameters()).device
        super().__init__(dev, batch, spatial, groups=model.out_norm.num_groups, training=training)
        self.model = model
        B, (H, W) = batch, self.sp
        chs = list(model.chs)
        n_down = len(model.downs)
        if min(H, W) >> n_down < 1:
            raise _lib.MriError(f"image size {self.sp} is too small for {n_down} stride-2 levels")
        eps = model.out_norm.eps
        cin = model.init_conv.weight.shape[1]
        if x_channels + ctx_channels != cin:
            raise _lib.MriError(f"init_conv expects {cin} input channels, got {x_channels} + "
                                f"{ctx_channels} context")
        self.x_in = torch.zeros(B, x_channels, H, W, device=dev)
        self.ctx_in = torch.zeros(B, ctx_channels, H, W, device=dev) if ctx_channels else None
        self.t_in = torch.zeros(B, dtype=torch.int64, device=dev)
        self.z_in = torch.zeros(B, 1, device=dev)

        # ---- conditioning: cond = time_mlp(t) + slice_mlp(z)  (unet.py:173-183) -----------------
        tdim = model.time_mlp[1].in_features
        temb = self.time_embedding(self.t_in, model.time_mlp, tdim)
        s0, s2 = model.slice_mlp[0], model.slice_mlp[2]
        self.track(s0.weight, s0.bias, s2.weight, s2.bias)
        zz = torch.zeros(B, s0.weight.shape[0], device=dev)
        zh = torch.zeros(B, s0.weight.shape[0], device=dev)
        cond = torch.zeros(B, tdim, device=dev)
        z_in = self.z_in
        if training:
            self._add("slice_mlp.0", lambda: ops.linear(z_in, s0.weight, s0.bias, zz), [zz])
            self._add("slice_mlp.silu", lambda: ops.silu(zz, zh), [zh])
        else:
            self._add("slice_mlp.0", lambda: ops.linear(z_in, s0.weight, s0.bias, zh, act=1), [zh])
        self._add("slice_mlp.2", lambda: ops.linear(zh, s2.weight, s2.bias, cond, addend=temb), [cond])
        blocks = []
        for d in model.downs:
            blocks += [d.res1, d.res2]
        blocks += [model.mid_block1, model.mid_block2]
        for u in model.ups:
            blocks += [u.res1, u.res2]
        # the 2D block applies SiLU to the projected embedding (unet.py:48-50)
        tproj, toffs, tld = self.block_projections(
            cond, blocks, act=1, slice_rec=dict(z_in=z_in, zz=zz, zh=zh, s0=s0, s2=s2))
        self._tproj = {id(b): (tproj[:, o:], tld, o) for b, o in zip(blocks, toffs)}

        # ---- init_conv: thin Cin -> patch matrix + GEMM --------------------------------------------
        S = H * W
        ic = model.init_conv
        self.track(ic.weight, ic.bias)
        if training or ctx_channels:
            kpad = _rup(9 * cin, 64)
            col = torch.zeros(B * S, kpad, dtype=torch.bfloat16, device=dev)
            x_in, ctx_in = self.x_in, self.ctx_in
            self._add("im2col", lambda: ops.im2col(x_in, col, B, x_channels, 1, H, W, 3, 2, kpad,
                                                   src2=ctx_in, cin2=ctx_channels), [col])
            w_in = self.packed(lambda: _pad_k(P.pack_conv_weight(ic.weight.detach()), kpad))
        h = None if (training or ctx_channels) else self.thin_in_conv(self.x_in, ic, self.sp, "init_conv",
                                                                      with_stats=False)
        if h is None:
            if not (training or ctx_channels):
                col, w_in, kpad = self.thin_patch_matrix(self.x_in, ic, self.sp, "init_conv")
            h = self.new_act(self.sp, chs[0], with_stats=False)
            pl = self._matrix_conv2d(col, w_in, h, S, kpad, ic.bias, "init_conv")
            self.gemm(pl)
            self.tape.append(ConvRec(kind="matrix", plan=pl, y=h.t, ksize=3, sources=[(col, True)],
                                     weight=ic.weight, splits=[cin], bias_params=[ic.bias], cout=chs[0],
                                     need_dgrad=False, kpad=kpad, name="init_conv"))

        skips: List[Act] = []
        for i, d in enumerate(model.downs):
            h = self.resblock2d(h, None, d.res1, eps, f"downs.{i}.res1")
            h = self.resblock2d(h, None, d.res2, eps, f"downs.{i}.res2")
            skips.append(h)
            dn = d.down
            self.track(dn.weight, dn.bias)
            wd = self.packed(lambda dn=dn: P.pack_conv_weight(dn.weight.detach()))
            hh, ww = h.t.shape[1:-1]
            src = h.t
            if hh % 2 or ww % 2:
                # odd level: Conv2d(k=4, s=2, p=1) gives floor(n / 2) outputs, and output o reads
                # inputs 2o-1 .. 2o+2 <= n-1 -- the same values as on the level zero-padded to an
                # even size, whose extra last output is dropped again
                src = torch.zeros(B, hh + hh % 2, ww + ww % 2, h.C, dtype=torch.bfloat16, device=dev)
                self._add(f"downs.{i}.pad", lambda s_=h.t, d_=src, n=(hh, ww): ops.copy_window(
                    s_, d_, (0, 0), (0, 0), n), [src])
                self.tape.append(WindowRec(h.t, src, (0, 0), (0, 0), (hh, ww), name=f"downs.{i}.pad"))
            y = self.new_act([s // 2 for s in src.shape[1:-1]], dn.weight.shape[0], with_stats=False)
            pl = P.down_conv_plan(src, wd, y.t, bias=dn.bias, name=f"downs.{i}.down")
            self.gemm(pl)
            self.tape.append(ConvRec(kind="down", plan=pl, y=y.t, ksize=4, sources=[(src, True)],
                                     weight=dn.weight, splits=[h.C], bias_params=[dn.bias],
                                     cout=dn.weight.shape[0], name=f"downs.{i}.down"))
            if src is not h.t:
                yc = self.new_act([hh // 2, ww // 2], dn.weight.shape[0], with_stats=False)
                self._add(f"downs.{i}.crop", lambda s_=y.t, d_=yc.t, n=(hh // 2, ww // 2): ops.copy_window(
                    s_, d_, (0, 0), (0, 0), n), [yc.t])
                self.tape.append(WindowRec(y.t, yc.t, (0, 0), (0, 0), (hh // 2, ww // 2),
                                           name=f"downs.{i}.crop"))
                self.pool.release(y.t)
                y = yc
            h = y
        h = self.resblock2d(h, None, model.mid_block1, eps, "mid_block1")
        h = self.resblock2d(h, None, model.mid_block2, eps, "mid_block2")
        for j, u in enumerate(model.ups):
            skip = skips.pop()
            up = u.up
            self.track(up.weight, up.bias)
            wu = self.packed(lambda up=up: P.pack_convT_weight(up.weight.detach()))
            y = self.new_act([s * 2 for s in h.t.shape[1:-1]], up.weight.shape[1], with_stats=False)
            pl = P.up_conv_plan(h.t, wu, y.t, bias=up.bias, name=f"ups.{j}.up")
            self.gemm(pl)
            self.tape.append(ConvRec(kind="up", plan=pl, y=y.t, ksize=4, sources=[(h.t, True)],
                                     weight=up.weight, splits=[h.C], bias_params=[up.bias],
                                     cout=up.weight.shape[1], name=f"ups.{j}.up"))
            self.pool.release(h.t)
            if tuple(y.t.shape[1:-1]) != tuple(skip.t.shape[1:-1]):
                # unet.py:98-99: F.interpolate(x, size=skip.shape[-2:], mode="bilinear")
                y2 = self.pool.get((B, *skip.t.shape[1:-1], y.C))
                self._add(f"ups.{j}.resize", lambda s_=y.t, d_=y2: ops.resize_bilinear(s_, d_), [y2])
                self.tape.append(ResizeRec(y.t, y2, name=f"ups.{j}.resize"))
                self.pool.release(y.t)
                y = Act(y2)
            h = self.resblock2d(y, skip, u.res1, eps, f"ups.{j}.res1")
            h = self.resblock2d(h, None, u.res2, eps, f"ups.{j}.res2")

        on, oc = model.out_norm, model.out_conv
        self.track(on.weight, on.bias, oc.weight, oc.bias)
        self.stats_of(h, "out_norm.stats")
        a = self.gn(h, on.weight, on.bias, self.groups, eps, True, name="out_norm")
        self.cout = oc.weight.shape[0]
        self.deps64 = None
        if not training and self.cout <= 4:
            self.thin_out_conv(a, oc)
        else:
            self.cout_pad = _rup(self.cout, 16)
            w_out = self.packed(lambda: P.pack_conv_weight(oc.weight.detach(), cout_pad=self.cout_pad))
            b_out = self.packed(lambda: _pad_vec(oc.bias.detach(), self.cout_pad))
            if training:
                self.deps64 = torch.zeros(B, H, W, 64, dtype=torch.bfloat16, device=dev)
            y = self.conv([P.ConvSource(a)], w_out, self.cout_pad, 3, b_out, with_stats=False,
                          name="out_conv",
                          rec=dict(weight=oc.weight, splits=[a.shape[-1]], bias_params=[oc.bias],
                                   cout=self.cout, dgrad_dy=self.deps64))
            self.eps_nhwc = y.t
        self.out = torch.zeros(B, self.cout, H, W, device=dev)
        self.params_changed()
        if training:
            self.dout_in = torch.zeros(B, self.cout, H, W, device=dev)
            self.deps16 = torch.zeros_like(self.eps_nhwc)
            self.build_backward({id(self.eps_nhwc): self.deps16})

    def _matrix_conv2d(self, col, wmat, y: Act, S: int, kpad: int, bias, name) -> P.GemmPlan:
        B, C = self.B, y.C
        a = P.TView(col, (kpad, S, B, 1, 1), (1, kpad, S * kpad, B * S * kpad, B * S * kpad))
        b = P.TView(wmat, (kpad, C, 1, 1), (1, kpad, kpad * C, kpad * C))
        o = P.TView(y.t, (C, S, B, 1, 1), (1, C, S * C, B * S * C, B * S * C))
        return P.matrix_plan(a, (128, 1, 1, 1), b, o, K=kpad, n_total=C, block_n=P.pick_block_n(C),
                             ext=(S, B, 1, 1), tiles=(-(-S // 128), B, 1, 1), sample_dim=2,
                             bias=bias, stats=y.stats, stats_cpg=y.cpg, name=name,
                             flops=2 * B * S * C * kpad)

    def resblock2d(self, x: Act, skip: Optional[Act], blk, eps: float, name: str) -> Act:
        """ResidualBlock (slice_cond_2d_ddpm/unet.py:42-56), post-norm:
        h = silu(gn1(conv1(x))) + silu(lin(cond)); h = silu(gn2(conv2(h))); return h + res_conv(x).
        With `skip`, x is the virtual concatenation [x, skip] (unet.py:101)."""
        n1, n2, c1, c2 = blk.norm1, blk.norm2, blk.conv1, blk.conv2
        self.track(n1.weight, n1.bias, n2.weight, n2.bias, c1.weight, c1.bias, c2.weight, c2.bias)
        cout = c1.weight.shape[0]
        rowbias, rb_ld, t_off = self._tproj[id(blk)]
        srcs = [x] if skip is None else [x, skip]
        cins = [s.C for s in srcs]
        w1 = self.packed(lambda: P.pack_conv_weight(c1.weight.detach(), splits=cins))
        h1 = self.conv([P.ConvSource(s.t) for s in srcs], w1, cout, 3, c1.bias, name=f"{name}.conv1",
                       rec=dict(weight=c1.weight, splits=cins, bias_params=[c1.bias]))
        a1 = self.gn(h1, n1.weight, n1.bias, self.groups, eps, True, rowbias=rowbias, rowbias_ld=rb_ld,
                     name=f"{name}.norm1+temb", tproj_off=t_off)
        self.pool.release(h1.t)
        w2 = self.packed(lambda: P.pack_conv_weight(c2.weight.detach()))
        h2 = self.conv([P.ConvSource(a1)], w2, cout, 3, c2.bias, name=f"{name}.conv2",
                       rec=dict(weight=c2.weight, splits=[cout], bias_params=[c2.bias]))
        self.pool.release(a1)
        if isinstance(blk.res_conv, torch.nn.Identity):
            assert skip is None and x.C == cout
            res = x.t
            res_tmp = None
        else:
            rc = blk.res_conv
            self.track(rc.weight, rc.bias)
            wr = self.packed(lambda: P.pack_conv_weight(rc.weight.detach(), splits=cins))
            res_tmp = self.conv([P.ConvSource(s.t) for s in srcs], wr, cout, 1, rc.bias,
                                with_stats=False, name=f"{name}.res_conv",
                                rec=dict(weight=rc.weight, splits=cins, bias_params=[rc.bias]))
            res = res_tmp.t
        out = self.gn(h2, n2.weight, n2.bias, self.groups, eps, True, residual=res,
                      name=f"{name}.norm2+res")
        self.pool.release(h2.t)
        if res_tmp is not None:
            self.pool.release(res_tmp.t)
        for s in srcs:
            self.pool.release(s.t)
        return Act(out)

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
        ops.nhwc_to_nchw(self.eps_nhwc, self.out, self.B, self.sp[0] * self.sp[1], self.cout,
                         self.cout_pad)
        return self.out

    def backward(self, dout: torch.Tensor, sync=None) -> None:
        self._backward(dout, self.sp[0] * self.sp[1], sync)
