"""Shared host-side pieces of the drop-in model classes.

The classes in mri_image_generation_b200.model_scripts.* keep the reference's constructor
signatures, attribute names and state_dict keys, so they hold their parameters in the same
torch.nn containers (nn.Conv3d, nn.GroupNorm, nn.Linear ...) the reference uses -- as parameter
holders only.  Their forward never calls those containers: it runs the UNetProgram (engine.py),
i.e. the sm_100a kernels behind the C ABI.
"""
from __future__ import annotations

import functools
from typing import Dict, Tuple

import torch
import torch.nn as nn

from . import _lib


def on_input_device(fn):
    """Run a module entry point with the input tensor's GPU as the current device: kernels are
    launched on `torch.cuda.current_stream()` and TMA descriptors are encoded for the current
    context, so a model living on cuda:1 while cuda:0 is current (model.to('cuda:1') without
    set_device, nn.DataParallel replicas) would otherwise launch on the wrong device."""
    @functools.wraps(fn)
    def wrapped(self, x, *args, **kwargs):
        if torch.is_tensor(x) and x.is_cuda and x.device.index != torch.cuda.current_device():
            with torch.cuda.device(x.device):
                return fn(self, x, *args, **kwargs)
        return fn(self, x, *args, **kwargs)
    return wrapped


class SinusoidalHolder(nn.Module):
    """Parameter-free placeholder for index 0 of `time_mlp` (SinusoidalPosEmb,
    slice_cond_2d_ddpm/unet.py:7-25).  The embedding itself is computed by mri_sinusoidal."""

    def __init__(self, dim: int):
        super().__init__()
        self.dim = dim

    def forward(self, t):  # pragma: no cover - never part of the product path
        raise _lib.MriError("SinusoidalHolder is a placeholder; the embedding runs in mri_sinusoidal")


class EngineModule(nn.Module):
    """Base for the drop-in UNets: caches one UNetProgram per (batch, spatial...) key."""

    def _programs(self) -> Dict[Tuple, object]:
        progs = self.__dict__.get("_mri_programs")
        if progs is None:
            progs = {}
            self.__dict__["_mri_programs"] = progs
        return progs

    def __getstate__(self):
        # programs hold ctypes handles / device tables: never pickled (mlflow.pytorch.log_model,
        # slice_cond_2d_ddpm/model.py:320, pickles the module)
        state = self.__dict__.copy()
        state.pop("_mri_programs", None)
        state.pop("_mri_grad_sync", None)   # process-group handle of parallel.DistributedDataParallel
        return state

    def _apply(self, fn, *args, **kwargs):
        # .to()/.cuda()/.half() may re-allocate parameter storage: drop cached programs
        self.__dict__.pop("_mri_programs", None)
        return super()._apply(fn, *args, **kwargs)

    def _check_input(self, x: torch.Tensor) -> None:
        if getattr(self, "_is_replica", False):
            # nn.DataParallel (slice_cond_2d_ddpm/model.py:113-115) re-creates shallow replicas of
            # the module on every forward; a UNetProgram is a static launch list bound to ONE set of
            # parameter tensors, so replicas would rebuild it every step -- and share the cached
            # programs of the original through the copied __dict__.  Refuse instead of being slow
            # and wrong: the multi-GPU path is one process per GPU.
            raise _lib.MriError(
                "nn.DataParallel replication is not supported by the B200 drop-in: run one process "
                "per GPU (torchrun) and wrap the UNet in DistributedDataParallel "
                "(mri_image_generation_b200.parallel.DistributedDataParallel or torch's), as "
                "ddpm_3d_ldm/train.py:232-233 does.  With a single visible GPU DataParallel calls "
                "the module directly and works.")
        if not x.is_cuda:
            raise _lib.MriError(
                f"{type(self).__name__} runs on B200 GPUs only (input is on {x.device}); this "
                "framework has no CPU fallback -- the reference implementation covers CPU.")

    def _needs_grad(self) -> bool:
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    MAX_PROGRAMS = 4  # static-shape programs own their activation memory: keep few

    def get_program(self, key: Tuple, build):
        progs = self._programs()
        prog = progs.get(key)
        if prog is None:
            if len(progs) >= self.MAX_PROGRAMS:
                progs.pop(next(iter(progs)))
            prog = build()
            progs[key] = prog
        return prog


class UNetFunction(torch.autograd.Function):
    """Autograd bridge: forward = the training UNetProgram's forward launch list, backward = its
    backward launch list.  All parameters are passed as inputs so that autograd (and DDP's
    reducer hooks, ddpm_3d_ldm/train.py:232-233) see their gradients.

    The program owns static activation buffers: exactly one backward may follow each forward
    (the reference training loops do precisely that, train.py:395-400)."""

    @staticmethod
    def forward(ctx, prog, run_forward, n_params, *tensors):
        ctx.prog = prog
        ctx.n_params = n_params
        out = run_forward()
        return out.clone()

    @staticmethod
    def backward(ctx, dout):
        prog = ctx.prog
        sync = getattr(prog, "grad_sync", None)
        with torch.cuda.device(dout.device):
            prog.backward(dout.contiguous().float(), sync=sync)
        if sync is not None and sync.active():
            sync.reduce_pending(prog.param_list)   # gradients left in .grad by no_sync() steps
        # one copy of the whole gradient arena (the program overwrites it next step); every
        # parameter's gradient is a view of that copy
        flat = prog.garena[:prog._garena_used].clone()
        grads = []
        for p in prog.param_list:
            loc = prog._g_off.get(id(p))
            if loc is None or not p.requires_grad:
                grads.append(None)
            else:
                grads.append(flat[loc[0]:loc[0] + loc[1]].view(p.shape))
        extra = len(ctx.needs_input_grad) - 3 - len(grads)
        return (None, None, None, *grads, *([None] * extra))


class ProgramFunction(torch.autograd.Function):
    """UNetFunction for programs that may also return the gradient of their input (the VAE
    decoder's latent, ddpm_3d_ldm/vae.py:106-118): apply(prog, run_forward, x_or_None, *params).
    The program leaves the input gradient in `prog.dz_out`."""

    @staticmethod
    def forward(ctx, prog, run_forward, x, *params):
        ctx.prog = prog
        return run_forward().clone()

    @staticmethod
    def backward(ctx, dout):
        prog = ctx.prog
        with torch.cuda.device(dout.device):
            prog.backward(dout.contiguous().float())
        flat = prog.garena[:prog._garena_used].clone()
        grads = []
        for p in prog.param_list:
            loc = prog._g_off.get(id(p))
            if loc is None or not p.requires_grad:
                grads.append(None)
            else:
                grads.append(flat[loc[0]:loc[0] + loc[1]].view(p.shape))
        dx = prog.dz_out.clone() if (ctx.needs_input_grad[2] and prog.dz_out is not None) else None
        return (None, None, dx, *grads)
