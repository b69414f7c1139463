"""Drop-in for model_scripts/ddpm_25d_all_modalities/unet.py: the 2D UNet with separate
in/out channel counts and optional context slices concatenated on channels (unet.py:109-218)."""
from typing import Union

import torch

from ...modules import on_input_device

from ..slice_cond_2d_ddpm.unet import (DownBlock, ResidualBlock, SinusoidalPosEmb,  # noqa: F401
                                       UpBlock, _UNet2DBase)


class UNet(_UNet2DBase):
    """ddpm_25d_all_modalities/unet.py:109-172."""

    def __init__(self, in_channels: int = 1, out_channels: int = 1, base_channels: int = 64,
                 channel_mults=(1, 2, 4, 8), time_emb_dim: int = 256):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self._build(in_channels, out_channels, base_channels, channel_mults, time_emb_dim)

    @on_input_device
    def forward(self, x: torch.Tensor, t: torch.Tensor, z_pos: torch.Tensor,
                context: Union[torch.Tensor, None] = None) -> torch.Tensor:
        """x: (B, C_target, H, W) noisy centre-slice modalities; context: (B, C_context, H, W)
        clean neighbouring slices, concatenated on channels (unet.py:174-218)."""
        return self._forward(x, t, z_pos, context)
