"""Shared implementation of the diffusion wrappers (q_sample / p_sample / loops) on the C ABI.

The three public classes (model_scripts/*/diffusion.py) differ only in schedule, the extra
model arguments (z_pos, context) and the loss; everything that touches the GPU is here:
  - q_sample / p_sample / DDIM arithmetic: one fused fp32 kernel each (mri_q_sample,
    mri_ddpm_step, mri_ddim_step), bit-exact w.r.t. the reference's eager association order;
  - the reverse loop: when the denoiser is one of this package's UNets, one CUDA graph holds a
    whole reverse step (time embedding, UNet, noise draw, fused update, t -= 1) and is replayed
    T times; the noise comes from torch's own Philox stream (normal_ on a static buffer), so
    for a given torch.manual_seed the draws are the ones the reference would make on this GPU.
"""
from __future__ import annotations

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

    # ------------------------------------------------------------------ pickling / graphs
    def _graphs(self) -> Dict:
        g = self.__dict__.get("_mri_graphs")
        if g is None:
            g = {}
            self.__dict__["_mri_graphs"] = g
        return g

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_mri_graphs", None)
        return state

    def _apply(self, fn, *args, **kwargs):
        self.__dict__.pop("_mri_graphs", None)
        return super()._apply(fn, *args, **kwargs)

    # ------------------------------------------------------------------ fused arithmetic
    def _q_sample(self, x_start, t, noise):
        _require_cuda(x_start, "q_sample")
        x0 = x_start.float().contiguous()
        out = torch.empty_like(x0)
        ops.q_sample(x0, noise.float().contiguous(), t.to(x0.device).long().contiguous(),
                     self.sqrt_alphas_cumprod, self.sqrt_one_minus_alphas_cumprod, out)
        return out

    def _p_update(self, x, t, eps, noise):
        """x_{t-1} from (x_t, eps, z): ddpm_3d_ldm/diffusion.py:118-126."""
        _require_cuda(x, "p_sample")
        xc = x.float().contiguous()
        out = torch.empty_like(xc)
        ops.ddpm_step(xc, eps.float().contiguous(), noise.float().contiguous(),
                      t.to(xc.device).long().contiguous(), self.betas,
                      self.sqrt_one_minus_alphas_cumprod, self.sqrt_recip_alphas,
                      self.posterior_variance, out)
        return out

    def _ddim_update(self, x, t, t_prev, eps):
        _require_cuda(x, "p_sample_ddim")
        xc = x.float().contiguous()
        out = torch.empty_like(xc)
        ops.ddim_step(xc, eps.float().contiguous(), t.to(xc.device).long().contiguous(),
                      t_prev.to(xc.device).long().contiguous(), self.alphas_cumprod, out)
        return out

    def _loss(self, pred, noise, t, gamma: float):
        """min-SNR weighted (gamma > 0) or plain (gamma <= 0) MSE, fused reduce."""
        _require_cuda(pred, "p_losses")
        snr = getattr(self, "snr", None)
        if gamma > 0 and snr is None:
            raise _lib.MriError("min-SNR loss needs the `snr` buffer")
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

    def _step_graph(self, prog, mode: str, one_step, tprev):
        """The CUDA graph of one reverse step on `prog` (captured once per program and mode)."""
        dev = prog.x_in.device
        key = (id(prog), mode)
        graphs = self._graphs()
        graph = graphs.get(key)
        if graph is None:
            # warm-up + capture must not disturb the caller's RNG stream
            rng = torch.cuda.get_rng_state(dev)
            prog.x_in.zero_()
            prog.t_in.fill_(1)
            tprev.fill_(0)
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
            torch.cuda.set_rng_state(rng, dev)
            graphs[key] = graph
        return graph

    def _p_sample_on(self, prog, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """One ancestral step x_t -> x_{t-1} on `prog` (whose conditioning inputs the caller has
        set): eager the first time a program is used, a replay of the captured step afterwards
        (the same launches, so the same values and the same draws from torch's Philox stream)."""
        calls = prog.__dict__.get("_p_sample_calls", 0)
        prog.__dict__["_p_sample_calls"] = calls + 1
        out = self._reverse_loop(prog, x.float(), 0, 1, "ddpm", use_graph=calls >= 1, t_vec=t)
        return out

    def _reverse_loop(self, prog, img: torch.Tensor, start_t: int, n_steps: int, mode: str,
                      use_graph: bool = True, t_vec: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Run n_steps reverse steps (i = start_t, start_t-1, ...) on `prog` (a UNetProgram whose
        x_in is the sampler state).  mode: 'ddpm' | 'ddim'."""
        dev = img.device
        if prog.params_changed():
            prog.do_refresh()
        B = prog.B
        per_sample = img[0].numel()
        noise = prog.__dict__.setdefault("_noise_buf", torch.empty_like(prog.x_in))
        tprev = prog.__dict__.setdefault("_tprev_buf", torch.zeros(B, dtype=torch.int64, device=dev))
        C, ldc = prog.cout, prog.cout_pad

        def one_step():
            prog.run()
            if mode == "ddpm":
                noise.normal_()
                ops.ddpm_step(prog.x_in, prog.eps_nhwc, noise, prog.t_in, self.betas,
                              self.sqrt_one_minus_alphas_cumprod, self.sqrt_recip_alphas,
                              self.posterior_variance, prog.x_in, eps_nhwc_ldc=ldc, channels=C)
            else:
                ops.ddim_step(prog.x_in, prog.eps_nhwc, prog.t_in, tprev, self.alphas_cumprod,
                              prog.x_in, eps_nhwc_ldc=ldc, channels=C)
                ops.add_i64(tprev, -1)
            ops.add_i64(prog.t_in, -1)

        graph = None
        if use_graph and (n_steps >= 3 or t_vec is not None):
            graph = self._step_graph(prog, mode, one_step, tprev)
        prog.x_in.copy_(img)
        if t_vec is not None:  # per-sample timesteps (p_sample called directly)
            prog.t_in.copy_(t_vec.to(dev).long())
        else:
            prog.t_in.fill_(start_t)
        tprev.fill_(start_t - 1)
        for _ in range(n_steps):
            if graph is not None:
                graph.replay()
            else:
                one_step()
        assert per_sample == prog.x_in[0].numel()
        return prog.x_in.clone()
