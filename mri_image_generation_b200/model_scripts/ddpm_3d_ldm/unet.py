"""Drop-in for model_scripts/ddpm_3d_ldm/unet.py: the 3D latent UNet without attention
(unet.py:57-158).  Same engine as unet_attention.py minus the bottleneck attention block."""
from .unet_attention import (ResidualBlock3D, SinusoidalPositionEmbeddings,  # noqa: F401
                             _UNet3DBase)


class UNet3DModel(_UNet3DBase):
    """unet.py:57-113."""

    def __init__(self, in_channels, base_channels=64, channel_mults=(1, 2, 4), time_emb_dim=256,
                 groups=8):
        super().__init__()
        self._build(in_channels, base_channels, channel_mults, time_emb_dim, groups, None)
