"""Drop-in for model_scripts/ddpm_3d_ldm/unet_attention.py (reference file:line in docstrings).

Same classes, constructor signatures, attributes and state_dict keys; forward runs the B200
engine (tcgen05 implicit-GEMM convolutions, fused GroupNorm/SiLU passes, tensor-core attention).
"""
import torch
import torch.nn as nn

from ... import _lib
from ...engine import UNet3DProgram
from ...split_engine import UNet3DSplitProgram
from ...modules import EngineModule, SinusoidalHolder, UNetFunction, on_input_device

# name kept for importers of the reference module (unet_attention.py:7)
SinusoidalPositionEmbeddings = SinusoidalHolder


class AttentionBlock3D(nn.Module):
    """Parameter holder for AttentionBlock3D (unet_attention.py:28-35)."""

    def __init__(self, channels, num_heads=4, groups=8):
        super().__init__()
        self.channels = channels
        self.num_heads = num_heads
        self.norm = nn.GroupNorm(groups, channels)
        self.qkv = nn.Conv3d(channels, channels * 3, 1)
        self.proj = nn.Conv3d(channels, channels, 1)


class ResidualBlock3D(nn.Module):
    """Parameter holder for ResidualBlock3D (unet_attention.py:59-77)."""

    def __init__(self, in_channels, out_channels, time_emb_dim=None, groups=8):
        super().__init__()
        self.time_emb_dim = time_emb_dim
        self.norm1 = nn.GroupNorm(groups, in_channels)
        self.act1 = nn.SiLU()
        self.conv1 = nn.Conv3d(in_channels, out_channels, 3, padding=1)
        if time_emb_dim is not None:
            self.time_mlp = nn.Linear(time_emb_dim, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels)
        self.act2 = nn.SiLU()
        self.conv2 = nn.Conv3d(out_channels, out_channels, 3, padding=1)
        if in_channels != out_channels:
            self.skip = nn.Conv3d(in_channels, out_channels, 1)
        else:
            self.skip = nn.Identity()


class _UNet3DBase(EngineModule):
    SUPPORTS_GRAD_SYNC = True   # backward can hand gradient buckets to parallel.GradSync

    def _build(self, in_channels, base_channels, channel_mults, time_emb_dim, groups, num_heads):
        self.in_channels = in_channels
        self.time_mlp = nn.Sequential(
            SinusoidalHolder(time_emb_dim),
            nn.Linear(time_emb_dim, time_emb_dim * 4),
            nn.SiLU(),
            nn.Linear(time_emb_dim * 4, time_emb_dim),
        )
        self.num_levels = len(channel_mults)
        chs = [base_channels * m for m in channel_mults]
        self.chs = chs
        self.in_conv = nn.Conv3d(in_channels, chs[0], 3, padding=1)
        downs = []
        for i in range(self.num_levels):
            ch = chs[i]
            res1 = ResidualBlock3D(ch, ch, time_emb_dim, groups)
            res2 = ResidualBlock3D(ch, ch, time_emb_dim, groups)
            if i != self.num_levels - 1:
                down = nn.Conv3d(ch, chs[i + 1], 4, stride=2, padding=1)
            else:
                down = nn.Identity()
            downs.append(nn.ModuleDict({"res1": res1, "res2": res2, "down": down}))
        self.downs = nn.ModuleList(downs)
        self.mid1 = ResidualBlock3D(chs[-1], chs[-1], time_emb_dim, groups)
        if num_heads is not None:
            self.mid_attn = AttentionBlock3D(chs[-1], num_heads=num_heads, groups=groups)
        self.mid2 = ResidualBlock3D(chs[-1], chs[-1], time_emb_dim, groups)
        ups = []
        cur_ch = chs[-1]
        for i in reversed(range(self.num_levels)):
            ch = chs[i]
            if i != self.num_levels - 1:
                up = nn.ConvTranspose3d(cur_ch, ch, 4, stride=2, padding=1)
            else:
                up = nn.Identity()
            res1 = ResidualBlock3D(ch * 2, ch, time_emb_dim, groups)
            res2 = ResidualBlock3D(ch, ch, time_emb_dim, groups)
            ups.append(nn.ModuleDict({"up": up, "res1": res1, "res2": res2}))
            cur_ch = ch
        self.ups = nn.ModuleList(ups)
        self.out_norm = nn.GroupNorm(groups, chs[0])
        self.out_act = nn.SiLU()
        self.out_conv = nn.Conv3d(chs[0], in_channels, 3, padding=1)

    # "bf16": bf16 operands / fp32 accumulation (the reference's training precision, autocast(bf16));
    # "split": parity with its fp32 / TF32 sampling path (show_model.py:254) -- every operand travels
    # as two bf16 numbers through the same kernels (split_engine.py); inference only, ~6x the time
    precision = "bf16"

    def program(self, batch: int, spatial, training: bool = False):
        if self.precision == "split":
            if training:
                raise _lib.MriError("precision = 'split' is an inference mode (run under torch.no_grad(), "
                                    "or set model.precision = 'bf16' for training)")
            key = (int(batch), tuple(int(s) for s in spatial), "split")
            return self.get_program(key, lambda: UNet3DSplitProgram(self, key[0], key[1]))
        if self.precision != "bf16":
            raise _lib.MriError(f"unknown precision {self.precision!r}: 'bf16' or 'split'")
        key = (int(batch), tuple(int(s) for s in spatial), bool(training))
        return self.get_program(key, lambda: UNet3DProgram(self, key[0], key[1], training=key[2]))

    @on_input_device
    def forward(self, x, t):
        """x: (B, C, D, H, W) fp32, t: (B,) int64 -> predicted noise (B, C, D, H, W) fp32
        (unet_attention.py:157-200)."""
        self._check_input(x)
        if x.dim() != 5 or x.shape[1] != self.in_channels:
            raise _lib.MriError(f"expected input (B, {self.in_channels}, D, H, W), got {tuple(x.shape)}")
        xf, tl = x.float().contiguous(), t.to(x.device).long()
        if self._needs_grad():
            # training: forward + backward launch lists behind one autograd node
            prog = self.program(x.shape[0], x.shape[2:], training=True)
            prog.param_list = list(self.parameters())
            sync = self.__dict__.get("_mri_grad_sync")  # set by parallel.DistributedDataParallel
            prog.grad_sync = sync.take() if sync is not None else None
            return UNetFunction.apply(prog, lambda: prog.forward(xf, tl), len(prog.param_list),
                                      *prog.param_list)
        prog = self.program(x.shape[0], x.shape[2:])
        return prog.forward(xf, tl).clone()


class UNet3DModelWithAttention(_UNet3DBase):
    """unet_attention.py:88-155."""

    def __init__(self, in_channels, base_channels=64, channel_mults=(1, 2, 4), time_emb_dim=256,
                 groups=8, num_heads=4):
        super().__init__()
        self._build(in_channels, base_channels, channel_mults, time_emb_dim, groups, num_heads)
