"""Fused optimizer step for the training path.

`Adam` is a drop-in for `torch.optim.Adam(params, lr, betas, eps, weight_decay)` as the reference
constructs it (ddpm_3d_ldm/train.py:243, slice_cond_2d_ddpm/model.py:126): same constructor,
`state_dict()` layout (`step`, `exp_avg`, `exp_avg_sq` per parameter), and
`torch.amp.GradScaler` support through the fused-optimizer protocol (`grad_scale` /
`found_inf`, no host synchronisation).  The whole update is ONE launch of mri_adam_step over all
parameters instead of torch's ~8 multi-tensor kernels per chunk of tensors.

fp32 parameters on a B200 only; anything else raises (no fallback).
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable

import torch

from . import _lib


def _bump_versions(tensors) -> None:
    """The kernel updated the parameters in place; advance their autograd version counters so
    that UNetProgram.params_changed() (engine.py) notices and re-packs the bf16 weight buffers."""
    inc = getattr(torch._C, "_increment_version", None)
    if inc is not None:
        try:
            inc(tensors)
            return
        except TypeError:
            for t in tensors:
                inc(t)
            return
    torch._foreach_add_(tensors, 0.0)


class Adam(torch.optim.Optimizer):
    _step_supports_amp_scaling = True  # GradScaler hands over grad_scale / found_inf

    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        # grad_scale / found_inf are set (and deleted again) by torch.amp.GradScaler around step()
        self._tables = {}
        self._group_steps = {}

    def load_state_dict(self, state_dict):
        """Loaded moment tensors replace the ones the cached segment tables point at."""
        out = super().load_state_dict(state_dict)
        self._tables.clear()
        self._group_steps.clear()
        return out

    def add_param_group(self, param_group):
        out = super().add_param_group(param_group)
        if hasattr(self, "_tables"):
            self._tables.clear()
        return out

    def _group_step(self, gi: int, plist) -> torch.Tensor:
        """ONE device-resident step counter per param group (the kernel applies one bias
        correction per launch).  It starts from the LARGEST step any parameter of the group
        already has -- after load_state_dict, or when parameters that had no gradient before
        join -- never from whichever parameter happens to come first."""
        st = self._group_steps.get(gi)
        dev = plist[0].device
        if st is None:
            have = [self.state[p]["step"] for p in plist if "step" in self.state[p]]
            if have:
                st = torch.stack([h.detach().to(dev, torch.float32).reshape(()) for h in have]).max()
            else:
                st = torch.zeros((), dtype=torch.float32, device=dev)
            st = st.clone()
            self._group_steps[gi] = st
        return st

    def _table(self, gi: int, plist, grads):
        """Device segment table of one param group; rebuilt when a pointer changed (new .grad
        tensors appear after zero_grad(set_to_none=True))."""
        # the table bakes in all four pointers: parameter, gradient AND both moments
        key = tuple((p.data_ptr(), g.data_ptr(), self.state[p]["exp_avg"].data_ptr(),
                     self.state[p]["exp_avg_sq"].data_ptr()) for p, g in zip(plist, grads))
        hit = self._tables.get(gi)
        if hit is not None and hit[0] == key:
            return hit[1], hit[2], hit[3]
        segs, blocks = [], 0
        for p, g in zip(plist, grads):
            st = self.state[p]
            sg = _lib.MriAdamSeg()
            sg.p, sg.g = p.data_ptr(), g.data_ptr()
            sg.m, sg.v = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            sg.n, sg.block0 = p.numel(), blocks
            blocks += -(-p.numel() // 1024)
            segs.append(sg)
        arr = (_lib.MriAdamSeg * len(segs))(*segs)
        table = torch.frombuffer(bytearray(bytes(memoryview(arr))), dtype=torch.uint8).to(plist[0].device)
        self._tables[gi] = (key, table, len(segs), blocks)
        return table, len(segs), blocks

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            plist, grads = [], []
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                    raise _lib.MriError("mri Adam: fp32 CUDA parameters and gradients only (no fallback)")
                if p.grad.is_sparse or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise _lib.MriError("mri Adam: dense contiguous parameters / gradients only")
                st = self.state[p]
                if len(st) == 0:
                    # device-resident step counter (as torch's fused Adam keeps it): it must not
                    # advance when GradScaler skips the update, and the host never syncs on that
                    st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                plist.append(p)
                grads.append(p.grad)
            if not plist:
                continue
            # one counter per group: every parameter's state["step"] aliases it
            step_t = self._group_step(gi, plist)
            for p in plist:
                self.state[p]["step"] = step_t
            table, n, blocks = self._table(gi, plist, grads)
            dev = plist[0].device
            gs, fi = getattr(self, "grad_scale", None), getattr(self, "found_inf", None)
            gs = gs.to(dev, torch.float32).reshape(-1) if gs is not None else None
            fi = fi.to(dev, torch.float32).reshape(-1) if fi is not None else None
            with torch.cuda.device(dev):
                rc = lib.mri_adam_step(table.data_ptr(), n, blocks, float(group["lr"]),
                                       float(group["betas"][0]), float(group["betas"][1]),
                                       float(group["eps"]), float(group["weight_decay"]),
                                       step_t.data_ptr(),
                                       gs.data_ptr() if gs is not None else None,
                                       fi.data_ptr() if fi is not None else None,
                                       torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(rc, "mri_adam_step")
            _bump_versions(plist)  # parameters changed behind autograd's back: packed buffers re-gather
        return loss
