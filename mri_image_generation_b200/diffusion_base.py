"""Shared implementation of the diffusion wrappers (q_sample / p_sample / loops) on the C ABI.

The three public classes (model_scripts/*/diffusion.py) differ only in schedule, the extra
model arguments (z_pos, context) and the loss; everything that touches the GPU is here:
  - q_sample / p_sample / DDIM arithmetic: one fused fp32 kernel each (mri_q_sample,
    mri_ddpm_step, mri_ddim_step), bit-exact w.r.t. the reference's eager association order;
  - the reverse loop: when the denoiser is one of this package's UNets, one CUDA graph holds a
    whole reverse step (time embedding, UNet, fused update with the noise drawn INSIDE the
    kernel, t -= 1) and is replayed T times.  The kernels draw Philox4x32-10 + Box-Muller
    normals with ATen's (seed, subsequence, offset) mapping (csrc/diffusion_ops.cu), seeded from
    torch's CUDA generator, and the host advances that generator by what the draws consumed: for
    a given torch.manual_seed the noise is bit-identical to what `torch.randn_like` would have
    produced at the reference's call sites, and no ATen kernel runs inside a replayed step.
"""
from __future__ import annotations

import os
import weakref
from typing import Callable, Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib, ops
from .modules import EngineModule


def _require_cuda(x: torch.Tensor, what: str) -> None:
    if not x.is_cuda:
        raise _lib.MriError(f"{what}: tensors must live on a B200 GPU (got {x.device}); "
                            "this framework has no CPU fallback")


class _LossFn(torch.autograd.Function):
    """min-SNR / plain MSE loss with its gradient kernel (ddpm_3d_ldm/diffusion.py:91-99)."""

    @staticmethod
    def forward(ctx, pred, noise, t, snr, gamma):
        B = pred.shape[0]
        per = torch.empty(B, dtype=torch.float32, device=pred.device)
        loss = torch.empty(1, dtype=torch.float32, device=pred.device)
        ops.minsnr_loss(pred, noise, t, snr, gamma, per, loss)
        ctx.save_for_backward(pred, noise, t, snr)
        ctx.gamma = gamma
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        pred, noise, t, snr = ctx.saved_tensors
        dpred = torch.empty_like(pred)
        ops.minsnr_loss_bwd(pred, noise, t, snr, ctx.gamma, g.reshape(1).float().contiguous(), dpred)
        return dpred, None, None, None, None


class DiffusionBase(nn.Module):
    """Host logic common to GaussianDiffusion (2D, 2.5D) and GaussianDiffusionLatent3D."""

    model: nn.Module
    timesteps: int

    # ------------------------------------------------------------------ device-side generator state
    def _rng(self, dev) -> ops.DeviceRng:
        rngs = self.__dict__.setdefault("_mri_rng", {})
        key = torch.device(dev)
        if key not in rngs:
            rngs[key] = ops.DeviceRng(key)
        return rngs[key]

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_mri_rng", None)
        return state

    def _apply(self, fn, *args, **kwargs):
        self.__dict__.pop("_mri_rng", None)
        return super()._apply(fn, *args, **kwargs)

    def _randn(self, shape, device) -> torch.Tensor:
        """torch.randn(shape, device=device) drawn by mri_randn (bit-identical, generator advanced)."""
        out = torch.empty(shape, dtype=torch.float32, device=device)
        _require_cuda(out, "sample")
        if out.numel() == 0:
            return out
        with torch.cuda.device(out.device):
            rng = self._rng(out.device)
            off = rng.load()
            ops.randn(out, rng)
            rng.commit(off + ops.randn_offset_increment(out.numel()))
        return out

    # ------------------------------------------------------------------ fused arithmetic
    def _q_sample(self, x_start, t, noise):
        _require_cuda(x_start, "q_sample")
        x0 = x_start.float().contiguous()
        out = torch.empty_like(x0)
        with torch.cuda.device(x0.device):
            ops.q_sample(x0, noise.float().contiguous(), t.to(x0.device).long().contiguous(),
                         self.sqrt_alphas_cumprod, self.sqrt_one_minus_alphas_cumprod, out)
        return out

    def _q_sample_draw(self, x_start, t) -> Tuple[torch.Tensor, torch.Tensor]:
        """q_sample with `noise = torch.randn_like(x_start)` drawn inside the kernel
        (ddpm_3d_ldm/diffusion.py:75-82): returns (x_t, noise)."""
        _require_cuda(x_start, "q_sample")
        x0 = x_start.float().contiguous()
        out, noise = torch.empty_like(x0), torch.empty_like(x0)
        with torch.cuda.device(x0.device):
            rng = self._rng(x0.device)
            off = rng.load()
            ops.q_sample_rng(x0, rng, t.to(x0.device).long().contiguous(), self.sqrt_alphas_cumprod,
                             self.sqrt_one_minus_alphas_cumprod, out, noise)
            rng.commit(off + ops.randn_offset_increment(x0.numel()))
        return out, noise

    def _p_update(self, x, t, eps, noise=None):
        """x_{t-1} from (x_t, eps, z): ddpm_3d_ldm/diffusion.py:118-126.  noise=None: z =
        randn_like(x) is drawn inside the kernel."""
        _require_cuda(x, "p_sample")
        xc = x.float().contiguous()
        out = torch.empty_like(xc)
        tl = t.to(xc.device).long().contiguous()
        with torch.cuda.device(xc.device):
            if noise is None:
                rng = self._rng(xc.device)
                off = rng.load()
                ops.ddpm_step_rng(xc, eps.float().contiguous(), rng, tl, self.betas,
                                  self.sqrt_one_minus_alphas_cumprod, self.sqrt_recip_alphas,
                                  self.posterior_variance, out)
                rng.commit(off + ops.randn_offset_increment(xc.numel()))
            else:
                ops.ddpm_step(xc, eps.float().contiguous(), noise.float().contiguous(), tl, self.betas,
                              self.sqrt_one_minus_alphas_cumprod, self.sqrt_recip_alphas,
                              self.posterior_variance, out)
        return out

    def _ddim_update(self, x, t, t_prev, eps):
        _require_cuda(x, "p_sample_ddim")
        xc = x.float().contiguous()
        out = torch.empty_like(xc)
        with torch.cuda.device(xc.device):
            ops.ddim_step(xc, eps.float().contiguous(), t.to(xc.device).long().contiguous(),
                          t_prev.to(xc.device).long().contiguous(), self.alphas_cumprod, out)
        return out

    def _loss(self, pred, noise, t, gamma: float):
        """min-SNR weighted (gamma > 0) or plain (gamma <= 0) MSE, fused reduce."""
        _require_cuda(pred, "p_losses")
        snr = getattr(self, "snr", None)
        if gamma > 0 and snr is None:
            raise _lib.MriError("min-SNR loss needs the `snr` buffer")
        with torch.cuda.device(pred.device):
            return _LossFn.apply(pred.float().contiguous(), noise.float().contiguous(),
                                 t.to(pred.device).long().contiguous(),
                                 snr if snr is not None else self.betas, float(gamma))

    # ------------------------------------------------------------------ graph-replayed loops
    def _engine_model(self) -> Optional[EngineModule]:
        m = self.model
        inner = getattr(m, "module", None)  # DataParallel / DDP wrappers
        if isinstance(inner, EngineModule):
            m = inner
        return m if isinstance(m, EngineModule) else None

    def _schedule_ptrs(self) -> Tuple[int, ...]:
        return tuple(getattr(self, n).data_ptr() for n in
                     ("betas", "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas",
                      "posterior_variance", "alphas_cumprod"))

    def _step_graph(self, prog, mode: str, one_step, tprev):
        """The CUDA graph of one reverse step on `prog`.  The graph lives ON the program (it
        replays launches into the program's buffers, so it must die with it: evicted or dropped
        programs take their graphs along) and remembers which diffusion object and which schedule
        buffers it was captured for; anything else re-captures."""
        dev = prog.x_in.device
        graphs = prog.__dict__.setdefault("_step_graphs", {})
        hit = graphs.get(mode)
        ptrs = self._schedule_ptrs()
        if hit is not None and hit[0]() is self and hit[1] == ptrs:
            return hit[2]
        prog.x_in.zero_()
        prog.t_in.fill_(1)
        tprev.fill_(0)
        self._rng(dev).load()           # a valid (seed, offset) for the warm-up / capture launches
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            one_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        prog.t_in.fill_(1)
        with torch.cuda.graph(graph):
            one_step()
        torch.cuda.synchronize(dev)
        graphs[mode] = (weakref.ref(self), ptrs, graph)
        return graph

    def _p_sample_on(self, prog, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """One ancestral step x_t -> x_{t-1} on `prog` (whose conditioning inputs the caller has
        set): eager the first time a program is used, a replay of the captured step afterwards
        (the same launches, so the same values and the same draws)."""
        calls = prog.__dict__.get("_p_sample_calls", 0)
        prog.__dict__["_p_sample_calls"] = calls + 1
        out = self._reverse_loop(prog, x.float(), 0, 1, "ddpm", use_graph=calls >= 1, t_vec=t)
        return out

    def _reverse_loop(self, prog, img: torch.Tensor, start_t: int, n_steps: int, mode: str,
                      use_graph: bool = True, t_vec: Optional[torch.Tensor] = None,
                      stride: int = 1) -> torch.Tensor:
        """Run n_steps reverse steps (i = start_t, start_t - stride, ...) on `prog` (a UNetProgram
        whose x_in is the sampler state).  mode: 'ddpm' | 'ddim' (t_prev = max(t - stride, 0))."""
        dev = img.device
        with torch.cuda.device(dev):
            return self._reverse_loop_on(prog, img, start_t, n_steps, mode, use_graph, t_vec, stride)

    def _reverse_loop_on(self, prog, img, start_t, n_steps, mode, use_graph, t_vec, stride):
        dev = img.device
        if prog.params_changed():
            prog.do_refresh()
        B = prog.B
        per_sample = img[0].numel()
        tprev = prog.__dict__.setdefault("_tprev_buf", torch.zeros(B, dtype=torch.int64, device=dev))
        C, ldc = prog.cout, prog.cout_pad
        rng = self._rng(dev)
        inc = ops.randn_offset_increment(prog.x_in.numel())
        if stride != 1 and mode != "ddim":
            raise _lib.MriError("strided timesteps need the DDIM update")
        # the graph bakes the stride in: one graph per (mode, stride)
        gkey = mode if stride == 1 else f"{mode}/{stride}"

        fused = (getattr(prog, "fused_head", None) is not None and not prog.training
                 and os.environ.get("MRI_FUSED_STEP", "1") != "0")

        def one_step():
            if fused:  # out_conv's tap sum + the update in one kernel: eps never goes to HBM
                if mode == "ddpm":
                    prog.run_fused_step(0, rng=rng, t=prog.t_in, betas=self.betas,
                                        sqrt_1mac=self.sqrt_one_minus_alphas_cumprod,
                                        sqrt_recip_alphas=self.sqrt_recip_alphas,
                                        post_var=self.posterior_variance)
                    ops.step_advance(prog.t_in, -1, rng=rng, rng_increment=inc)
                else:
                    prog.run_fused_step(1, t=prog.t_in, t_prev=tprev, alphas_cumprod=self.alphas_cumprod)
                    ops.step_advance(prog.t_in, -stride, t_prev=tprev)
                return
            prog.run()
            if mode == "ddpm":
                ops.ddpm_step_rng(prog.x_in, prog.eps_nhwc, rng, prog.t_in, self.betas,
                                  self.sqrt_one_minus_alphas_cumprod, self.sqrt_recip_alphas,
                                  self.posterior_variance, prog.x_in, eps_nhwc_ldc=ldc, channels=C)
                ops.step_advance(prog.t_in, -1, rng=rng, rng_increment=inc)
            else:
                ops.ddim_step(prog.x_in, prog.eps_nhwc, prog.t_in, tprev, self.alphas_cumprod,
                              prog.x_in, eps_nhwc_ldc=ldc, channels=C)
                ops.step_advance(prog.t_in, -stride, t_prev=tprev)

        graph = None
        if use_graph and (n_steps >= 3 or t_vec is not None):
            graph = self._step_graph(prog, gkey, one_step, tprev)
        prog.x_in.copy_(img)
        if t_vec is not None:  # per-sample timesteps (p_sample called directly)
            prog.t_in.copy_(t_vec.to(dev).long())
        else:
            prog.t_in.fill_(start_t)
        tprev.fill_(max(start_t - stride, 0))
        off = rng.load() if mode == "ddpm" else 0
        for _ in range(n_steps):
            if graph is not None:
                graph.replay()
            else:
                one_step()
        if mode == "ddpm":  # z = randn_like(x) is drawn at every step, t = 0 included
            rng.commit(off + n_steps * inc)
        assert per_sample == prog.x_in[0].numel()
        return prog.x_in.clone()
