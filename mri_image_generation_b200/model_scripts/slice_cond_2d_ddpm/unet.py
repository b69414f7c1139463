"""Drop-in for model_scripts/slice_cond_2d_ddpm/unet.py (reference file:line in docstrings).

Same classes, constructor signatures, attributes and state_dict keys; forward runs the B200
engine (engine.UNet2DProgram)."""
import torch
import torch.nn as nn

from ... import _lib
from ...engine import UNet2DProgram
from ...split_engine import UNet2DSplitProgram
from ...modules import EngineModule, SinusoidalHolder, UNetFunction, on_input_device

SinusoidalPosEmb = SinusoidalHolder  # unet.py:7


class ResidualBlock(nn.Module):
    """Parameter holder for ResidualBlock (unet.py:28-40)."""

    def __init__(self, in_ch: int, out_ch: int, t_dim: int, groups: int = 8):
        super().__init__()
        self.conv1 = nn.Conv2d(in_ch, out_ch, 3, padding=1)
        self.conv2 = nn.Conv2d(out_ch, out_ch, 3, padding=1)
        self.time_mlp = nn.Linear(t_dim, out_ch)
        self.norm1 = nn.GroupNorm(groups, out_ch)
        self.norm2 = nn.GroupNorm(groups, out_ch)
        self.act = nn.SiLU()
        self.res_conv = nn.Conv2d(in_ch, out_ch, 1) if in_ch != out_ch else nn.Identity()


class DownBlock(nn.Module):
    """unet.py:59-70."""

    def __init__(self, in_ch: int, out_ch: int, t_dim: int):
        super().__init__()
        self.res1 = ResidualBlock(in_ch, out_ch, t_dim)
        self.res2 = ResidualBlock(out_ch, out_ch, t_dim)
        self.down = nn.Conv2d(out_ch, out_ch, 4, stride=2, padding=1)


class UpBlock(nn.Module):
    """unet.py:80-93 (the constructor prints, as the reference's does)."""

    def __init__(self, in_ch: int, skip_ch: int, out_ch: int, t_dim: int):
        super().__init__()
        self.up = nn.ConvTranspose2d(in_ch, out_ch, 4, stride=2, padding=1)
        self.res1 = ResidualBlock(out_ch + skip_ch, out_ch, t_dim)
        self.res2 = ResidualBlock(out_ch, out_ch, t_dim)
        print(f"UpBlock: in_ch={in_ch}, skip_ch={skip_ch}, out_ch={out_ch}")


class _UNet2DBase(EngineModule):
    SUPPORTS_GRAD_SYNC = True   # backward can hand gradient buckets to parallel.GradSync

    def _build(self, in_channels, out_channels, base_channels, channel_mults, time_emb_dim):
        self.time_mlp = nn.Sequential(
            SinusoidalHolder(time_emb_dim),
            nn.Linear(time_emb_dim, time_emb_dim * 4),
            nn.SiLU(),
            nn.Linear(time_emb_dim * 4, time_emb_dim),
        )
        self.slice_mlp = nn.Sequential(
            nn.Linear(1, time_emb_dim * 4),
            nn.SiLU(),
            nn.Linear(time_emb_dim * 4, time_emb_dim),
        )
        self.chs = [base_channels * m for m in channel_mults]
        self.init_conv = nn.Conv2d(in_channels, self.chs[0], 3, padding=1)
        downs = []
        for in_ch, out_ch in zip(self.chs[:-1], self.chs[1:]):
            downs.append(DownBlock(in_ch, out_ch, time_emb_dim))
        self.downs = nn.ModuleList(downs)
        self.mid_block1 = ResidualBlock(self.chs[-1], self.chs[-1], time_emb_dim)
        self.mid_block2 = ResidualBlock(self.chs[-1], self.chs[-1], time_emb_dim)
        ups = []
        skip_chs = self.chs[1:]
        in_ch = self.chs[-1]
        for skip_ch, out_ch in zip(reversed(skip_chs), reversed(self.chs[:-1])):
            ups.append(UpBlock(in_ch, skip_ch, out_ch, time_emb_dim))
            in_ch = out_ch
        self.ups = nn.ModuleList(ups)
        self.out_norm = nn.GroupNorm(8, self.chs[0])
        self.out_conv = nn.Conv2d(self.chs[0], out_channels, 3, padding=1)

    # "bf16" (default) | "split": parity with the fp32 / TF32 sampling of the reference's show_model /
    # metrics scripts (no autocast there); see split_engine.py.  Inference only.
    precision = "bf16"

    def program(self, batch: int, spatial, x_channels: int, ctx_channels: int = 0,
                training: bool = False):
        if self.precision == "split":
            if training:
                raise _lib.MriError("precision = 'split' is an inference mode (run under torch.no_grad(), "
                                    "or set model.precision = 'bf16' for training)")
            skey = (int(batch), tuple(int(s) for s in spatial), int(x_channels), int(ctx_channels), "split")
            return self.get_program(skey, lambda: UNet2DSplitProgram(self, skey[0], skey[1], skey[2], skey[3]))
        if self.precision != "bf16":
            raise _lib.MriError(f"unknown precision {self.precision!r}: 'bf16' or 'split'")
        key = (int(batch), tuple(int(s) for s in spatial), int(x_channels), int(ctx_channels),
               bool(training))
        return self.get_program(key, lambda: UNet2DProgram(self, key[0], key[1], key[2], key[3],
                                                           training=key[4]))

    def _forward(self, x, t, z_pos, context=None):
        self._check_input(x)
        if x.dim() != 4:
            raise _lib.MriError(f"expected input (B, C, H, W), got {tuple(x.shape)}")
        t = t.to(x.device).long()
        z_pos = z_pos.to(x.device).float()
        cc = 0 if context is None else context.shape[1]
        ctx = None if context is None else context.to(x.device).float().contiguous()
        xf = x.float().contiguous()
        if self._needs_grad():
            prog = self.program(x.shape[0], x.shape[2:], x.shape[1], cc, training=True)
            prog.param_list = list(self.parameters())
            sync = self.__dict__.get("_mri_grad_sync")  # set by parallel.DistributedDataParallel
            prog.grad_sync = sync.take() if sync is not None else None
            return UNetFunction.apply(prog, lambda: prog.forward(xf, t, z_pos, ctx),
                                      len(prog.param_list), *prog.param_list)
        prog = self.program(x.shape[0], x.shape[2:], x.shape[1], cc)
        return prog.forward(xf, t, z_pos, ctx).clone()


class UNet(_UNet2DBase):
    """unet.py:108-199: time + slice-position conditioned UNet for 1-channel 2D slices."""

    def __init__(self, img_channels: int = 1, base_channels: int = 64, channel_mults=(1, 2, 4, 8),
                 time_emb_dim: int = 256):
        super().__init__()
        self._build(img_channels, img_channels, base_channels, channel_mults, time_emb_dim)

    @on_input_device
    def forward(self, x: torch.Tensor, t: torch.Tensor, z_pos: torch.Tensor) -> torch.Tensor:
        """unet.py:169-199."""
        return self._forward(x, t, z_pos)
