"""CPU oracle for the diffusion hot path -- TEST INFRASTRUCTURE ONLY.

A plain-PyTorch fp32 restatement of the reference's algorithm (NickB42/mri-image-generation,
model_scripts/{slice_cond_2d_ddpm,ddpm_25d_all_modalities,ddpm_3d_ldm}/{unet,unet_attention,
diffusion}.py), written functionally over a state_dict so that it shares nothing with the
product code in mri_image_generation_b200/.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module; the product path never
does (it raises without the CUDA library).

Pinning: the reference ships no tests or golden vectors for this path, so the
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: oracle/make_golden.py imports the
unmodified reference modules from /root/reference, runs them on seeded inputs and commits the
results under tests/golden/; tests/test_oracle_golden.py checks this file against them (and,
where /root/reference is present, against the live reference).  The arithmetic lives in
PyTorch (requirements.txt:1, unpinned; this image: torch 2.11.0+cu128).

Every function cites the reference file:line it follows (paths relative to model_scripts/).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# ------------------------------------------------------------------------------------------
# embeddings / small pieces
# ------------------------------------------------------------------------------------------
def sinusoidal(t: torch.Tensor, dim: int) -> torch.Tensor:
    """SinusoidalPosEmb.forward -- slice_cond_2d_ddpm/unet.py:12-25 (same in the other copies)."""
    half = dim // 2
    f = math.log(10000) / (half - 1)
    emb = torch.exp(torch.arange(half, device=t.device) * -f)
    emb = t.float().unsqueeze(1) * emb.unsqueeze(0)
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if dim % 2 == 1:
        emb = F.pad(emb, (0, 1))
    return emb


def _lin(sd: SD, p: str, x):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


def _gn(sd: SD, p: str, x, groups=8):
    return F.group_norm(x, groups, sd[p + ".weight"], sd[p + ".bias"], eps=1e-5)


def _conv(sd: SD, p: str, x, nd: int, **kw):
    fn = F.conv3d if nd == 3 else F.conv2d
    return fn(x, sd[p + ".weight"], sd[p + ".bias"], **kw)


def _convT(sd: SD, p: str, x, nd: int, **kw):
    fn = F.conv_transpose3d if nd == 3 else F.conv_transpose2d
    return fn(x, sd[p + ".weight"], sd[p + ".bias"], **kw)


def _time_mlp(sd: SD, t, dim):
    """time_mlp Sequential -- unet.py:124-129 / unet_attention.py:103-108."""
    e = sinusoidal(t, dim)
    return _lin(sd, "time_mlp.3", F.silu(_lin(sd, "time_mlp.1", e)))


# ------------------------------------------------------------------------------------------
# 3D latent UNet (ddpm_3d_ldm)
# ------------------------------------------------------------------------------------------
def resblock3d(sd: SD, p: str, x, temb, groups=8):
    """ResidualBlock3D.forward -- ddpm_3d_ldm/unet_attention.py:79-85 (pre-norm; the time
    projection is added WITHOUT an activation)."""
    h = _conv(sd, p + ".conv1", F.silu(_gn(sd, p + ".norm1", x, groups)), 3, padding=1)
    h = h + _lin(sd, p + ".time_mlp", temb)[:, :, None, None, None]
    h = _conv(sd, p + ".conv2", F.silu(_gn(sd, p + ".norm2", h, groups)), 3, padding=1)
    if (p + ".skip.weight") in sd:
        return h + _conv(sd, p + ".skip", x, 3)
    return h + x


def attention3d(sd: SD, p: str, x, heads=4, groups=8):
    """AttentionBlock3D.forward -- ddpm_3d_ldm/unet_attention.py:37-56."""
    B, C, D, H, W = x.shape
    h = _gn(sd, p + ".norm", x, groups)
    q, k, v = _conv(sd, p + ".qkv", h, 3).chunk(3, dim=1)
    q = q.reshape(B, heads, C // heads, D * H * W)
    k = k.reshape(B, heads, C // heads, D * H * W)
    v = v.reshape(B, heads, C // heads, D * H * W)
    scale = (C // heads) ** -0.5
    attn = torch.softmax(torch.einsum("bhcn,bhcm->bhnm", q, k) * scale, dim=-1)
    h = torch.einsum("bhnm,bhcm->bhcn", attn, v).reshape(B, C, D, H, W)
    return x + _conv(sd, p + ".proj", h, 3)


def unet3d_forward(sd: SD, x, t, heads: int = 4, groups: int = 8):
    """UNet3DModelWithAttention.forward (unet_attention.py:157-200) / UNet3DModel.forward
    (ddpm_3d_ldm/unet.py:115-158); attention is present iff the state_dict has mid_attn.*"""
    tdim = sd["time_mlp.1.weight"].shape[1]
    temb = _time_mlp(sd, t, tdim)
    levels = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("downs."))
    h = _conv(sd, "in_conv", x, 3, padding=1)
    skips = []
    for i in range(levels):
        h = resblock3d(sd, f"downs.{i}.res1", h, temb, groups)
        h = resblock3d(sd, f"downs.{i}.res2", h, temb, groups)
        skips.append(h)
        if f"downs.{i}.down.weight" in sd:
            h = _conv(sd, f"downs.{i}.down", h, 3, stride=2, padding=1)
    h = resblock3d(sd, "mid1", h, temb, groups)
    if "mid_attn.qkv.weight" in sd:
        h = attention3d(sd, "mid_attn", h, heads, groups)
    h = resblock3d(sd, "mid2", h, temb, groups)
    for j in range(levels):
        if f"ups.{j}.up.weight" in sd:
            h = _convT(sd, f"ups.{j}.up", h, 3, stride=2, padding=1)
        skip = skips.pop()
        if h.shape[-3:] != skip.shape[-3:]:  # centre crop, unet_attention.py:184-193
            dz = (skip.shape[-3] - h.shape[-3]) // 2
            dy = (skip.shape[-2] - h.shape[-2]) // 2
            dx = (skip.shape[-1] - h.shape[-1]) // 2
            skip = skip[..., dz:dz + h.shape[-3], dy:dy + h.shape[-2], dx:dx + h.shape[-1]]
        h = torch.cat([h, skip], dim=1)
        h = resblock3d(sd, f"ups.{j}.res1", h, temb, groups)
        h = resblock3d(sd, f"ups.{j}.res2", h, temb, groups)
    return _conv(sd, "out_conv", F.silu(_gn(sd, "out_norm", h, groups)), 3, padding=1)


# ------------------------------------------------------------------------------------------
# 3D VAE (ddpm_3d_ldm/vae.py) -- the step either side of the diffusion path
# ------------------------------------------------------------------------------------------
def vae_resblock(sd: SD, p: str, x, groups=8):
    """ResidualBlock3DNoTime.forward -- ddpm_3d_ldm/vae.py:19-22."""
    h = _conv(sd, p + ".conv1", F.silu(_gn(sd, p + ".norm1", x, groups)), 3, padding=1)
    h = _conv(sd, p + ".conv2", F.silu(_gn(sd, p + ".norm2", h, groups)), 3, padding=1)
    if (p + ".skip.weight") in sd:
        return h + _conv(sd, p + ".skip", x, 3)
    return h + x


def _vae_layers(sd: SD, prefix: str, h, groups, transposed: bool):
    i = 0
    while any(k.startswith(f"{prefix}.{i}.") for k in sd):
        p = f"{prefix}.{i}"
        if (p + ".norm1.weight") in sd:
            h = vae_resblock(sd, p, h, groups)
        elif transposed:   # nn.ConvTranspose3d(c, c, 4, stride=2, padding=1), vae.py:75-79
            h = _convT(sd, p, h, 3, stride=2, padding=1)
        else:              # nn.Conv3d(c, c, 4, stride=2, padding=1), vae.py:41-43
            h = _conv(sd, p, h, 3, stride=2, padding=1)
        i += 1
    return h


def vae3d_encode(sd: SD, x, groups: int = 8):
    """VAE3D.encode -> Encoder3D.forward (vae.py:49-55, 102-104): returns (mu, logvar)."""
    h = _conv(sd, "encoder.in_conv", x, 3, padding=1)
    h = _vae_layers(sd, "encoder.downs", h, groups, transposed=False)
    stats = _conv(sd, "encoder.to_mu_logvar", h, 3, padding=1)
    mu, logvar = torch.chunk(stats, 2, dim=1)
    return mu, logvar


def vae3d_decode(sd: SD, z, groups: int = 8):
    """VAE3D.decode -> Decoder3D.forward (vae.py:82-87, 111-112)."""
    h = _conv(sd, "decoder.from_latent", z, 3, padding=1)
    h = _vae_layers(sd, "decoder.ups", h, groups, transposed=True)
    return _conv(sd, "decoder.out_conv", h, 3, padding=1)


# ------------------------------------------------------------------------------------------
# 2D / 2.5D UNet (slice_cond_2d_ddpm, ddpm_25d_all_modalities)
# ------------------------------------------------------------------------------------------
def resblock2d(sd: SD, p: str, x, cond):
    """ResidualBlock.forward -- slice_cond_2d_ddpm/unet.py:42-56 (post-norm; SiLU IS applied
    to the projected embedding)."""
    h = F.silu(_gn(sd, p + ".norm1", _conv(sd, p + ".conv1", x, 2, padding=1)))
    h = h + F.silu(_lin(sd, p + ".time_mlp", cond))[:, :, None, None]
    h = F.silu(_gn(sd, p + ".norm2", _conv(sd, p + ".conv2", h, 2, padding=1)))
    if (p + ".res_conv.weight") in sd:
        return h + _conv(sd, p + ".res_conv", x, 2)
    return h + x


def unet2d_forward(sd: SD, x, t, z_pos, context: Optional[torch.Tensor] = None):
    """UNet.forward -- slice_cond_2d_ddpm/unet.py:169-199; with `context`,
    ddpm_25d_all_modalities/unet.py:174-218 (context concatenated on channels, :198-199)."""
    t = t.to(x.device)
    z_pos = z_pos.to(x.device).float()
    tdim = sd["time_mlp.1.weight"].shape[1]
    temb = _time_mlp(sd, t, tdim)
    zemb = _lin(sd, "slice_mlp.2", F.silu(_lin(sd, "slice_mlp.0", z_pos.unsqueeze(-1))))
    cond = temb + zemb
    if context is not None:
        x = torch.cat([x, context], dim=1)
    x = _conv(sd, "init_conv", x, 2, padding=1)
    n_down = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("downs."))
    skips = []
    for i in range(n_down):
        x = resblock2d(sd, f"downs.{i}.res1", x, cond)
        x = resblock2d(sd, f"downs.{i}.res2", x, cond)
        skips.append(x)
        x = _conv(sd, f"downs.{i}.down", x, 2, stride=2, padding=1)
    x = resblock2d(sd, "mid_block1", x, cond)
    x = resblock2d(sd, "mid_block2", x, cond)
    for j in range(n_down):
        skip = skips.pop()
        x = _convT(sd, f"ups.{j}.up", x, 2, stride=2, padding=1)
        if x.shape[-2:] != skip.shape[-2:]:  # unet.py:98-99
            x = F.interpolate(x, size=skip.shape[-2:], mode="bilinear", align_corners=False)
        x = torch.cat([x, skip], dim=1)
        x = resblock2d(sd, f"ups.{j}.res1", x, cond)
        x = resblock2d(sd, f"ups.{j}.res2", x, cond)
    return _conv(sd, "out_conv", F.silu(_gn(sd, "out_norm", x)), 2, padding=1)


# ------------------------------------------------------------------------------------------
# diffusion process
# ------------------------------------------------------------------------------------------
def linear_betas(T: int, beta_start=1e-4, beta_end=0.02) -> torch.Tensor:
    """slice_cond_2d_ddpm/diffusion.py:23."""
    return torch.linspace(beta_start, beta_end, T, dtype=torch.float32)


def cosine_betas(T: int, s: float = 0.008) -> torch.Tensor:
    """cosine_beta_schedule -- ddpm_3d_ldm/diffusion.py:50-56."""
    steps = T + 1
    x = torch.linspace(0, T, steps, dtype=torch.float32)
    ac = torch.cos(((x / T) + s) / (1 + s) * math.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = 1 - (ac[1:] / ac[:-1])
    return torch.clamp(betas, 1e-8, 0.999)


def schedule_buffers(betas: torch.Tensor, with_snr: bool = True) -> SD:
    """The registered buffers -- slice_cond_2d_ddpm/diffusion.py:23-49,
    ddpm_3d_ldm/diffusion.py:23-48 (2.5D: same minus `snr`, ddpm_25d.../diffusion.py:22-47)."""
    alphas = 1.0 - betas
    ac = torch.cumprod(alphas, dim=0)
    ac_prev = torch.cat([torch.tensor([1.0], dtype=torch.float32), ac[:-1]], dim=0)
    out = {
        "betas": betas, "alphas": alphas, "alphas_cumprod": ac, "alphas_cumprod_prev": ac_prev,
        "sqrt_alphas_cumprod": torch.sqrt(ac),
        "sqrt_one_minus_alphas_cumprod": torch.sqrt(1.0 - ac),
        "sqrt_recip_alphas": torch.sqrt(1.0 / alphas),
    }
    if with_snr:
        out["snr"] = ac / (1.0 - ac)
    pv = betas * (1.0 - ac_prev) / (1.0 - ac)
    out["posterior_variance"] = pv
    out["posterior_log_variance_clipped"] = torch.log(torch.clamp(pv, min=1e-20))
    return out


def _extract(a, t, x):
    """_extract -- ddpm_3d_ldm/diffusion.py:58-66 (rank-generic form)."""
    return a.gather(-1, t).view(t.shape[0], *([1] * (x.dim() - 1)))


def q_sample(buf: SD, x0, t, noise):
    """q_sample -- slice_cond_2d_ddpm/diffusion.py:60-75, ddpm_3d_ldm/diffusion.py:68-82."""
    return _extract(buf["sqrt_alphas_cumprod"], t, x0) * x0 + \
        _extract(buf["sqrt_one_minus_alphas_cumprod"], t, x0) * noise


def p_sample_update(buf: SD, x, t, eps, noise):
    """The arithmetic of p_sample after the model call -- slice_cond_2d_ddpm/diffusion.py:115-132,
    ddpm_3d_ldm/diffusion.py:106-126.  No clamp exists in the reference."""
    betas_t = _extract(buf["betas"], t, x)
    s1m = _extract(buf["sqrt_one_minus_alphas_cumprod"], t, x)
    sra = _extract(buf["sqrt_recip_alphas"], t, x)
    pv = _extract(buf["posterior_variance"], t, x)
    mean = sra * (x - betas_t / s1m * eps)
    mask = (t != 0).float().view(t.shape[0], *([1] * (x.dim() - 1)))
    return mean + mask * torch.sqrt(pv) * noise


def ddim_update(buf: SD, x, t, t_prev, eps):
    """p_sample_ddim after the model call -- ddpm_3d_ldm/diffusion.py:173-186."""
    a_t = _extract(buf["alphas_cumprod"], t, x)
    a_prev = _extract(buf["alphas_cumprod"], t_prev, x)
    x0 = (x - torch.sqrt(1.0 - a_t) * eps) / torch.clamp(torch.sqrt(a_t), min=1e-8)
    return torch.sqrt(a_prev) * x0 + torch.sqrt(1.0 - a_prev) * eps


def minsnr_loss(buf: SD, pred, noise, t, gamma: float = 5.0):
    """p_losses tail -- ddpm_3d_ldm/diffusion.py:91-99 (rank-generic mean over non-batch dims;
    the 2D copy's hard-coded dim=(1,2,3,4) crashes on 4-D input)."""
    mse = ((pred - noise) ** 2).mean(dim=tuple(range(1, pred.dim())))
    snr_t = buf["snr"].gather(-1, t)
    w = torch.minimum(snr_t, torch.tensor(gamma)) / snr_t
    return (w * mse).mean()


def mse_loss(pred, noise):
    """ddpm_25d_all_modalities/diffusion.py:89 (F.mse_loss)."""
    return F.mse_loss(pred, noise)


def sample_loop(buf: SD, model_fn, x_T, noises: Sequence[torch.Tensor], T: int):
    """p_sample_loop -- ddpm_3d_ldm/diffusion.py:128-141 with the per-step noise injected
    (noises[k] is the draw of step i = T-1-k), so that trajectories are comparable across
    devices with different RNG implementations."""
    img = x_T
    B = x_T.shape[0]
    for k, i in enumerate(reversed(range(T))):
        t = torch.full((B,), i, dtype=torch.long)
        img = p_sample_update(buf, img, t, model_fn(img, t), noises[k])
    return img


# ------------------------------------------------------------------------------------------
# data path: what Dataset.__getitem__ does to a loaded volume (numpy + F.interpolate on the CPU)
# ------------------------------------------------------------------------------------------
def _zscore_nonzero(a, floor):
    """In place: a[a != 0] <- (a - mean) / std over those entries; `floor(std)` is the reference's
    guard against a vanishing deviation.  Returns False when there is no non-zero entry."""
    import numpy as np
    nz = a != 0
    if not np.any(nz):
        return False
    m, s = a[nz].mean(), floor(a[nz].std())
    a[nz] = (a[nz] - m) / s
    return True


def preprocess_slice(slice_2d, image_size: int) -> torch.Tensor:
    """slice_cond_2d_ddpm/dataset.py:71-98 == ddpm_25d_all_modalities/dataset.py:79-103:
    z-score of the non-zero pixels, clip +-5, [0, 1], bilinear resize, [-1, 1] -> (1, S, S).
    (Works on a copy: the 2-D reference mutates its cached volume through the view, see
    mri_image_generation_b200/model_scripts/slice_cond_2d_ddpm/dataset.py.)"""
    import numpy as np
    a = np.array(slice_2d, dtype=np.float32, copy=True)
    _zscore_nonzero(a, lambda s: s if s > 0 else 1.0)
    a = (np.clip(a, -5, 5) + 5) / 10.0
    t = F.interpolate(torch.from_numpy(a)[None, None], size=(image_size, image_size),
                      mode="bilinear", align_corners=False)[0]
    return t * 2.0 - 1.0


def normalize_volume(vol, eps: float = 1e-6, clip_val: float = 5.0):
    """ddpm_3d_ldm/dataset.py:11-41 (float32 numpy in, float32 numpy out, input untouched)."""
    import numpy as np
    a = np.array(vol, dtype=np.float32, copy=True)
    floor = lambda s: 1.0 if s < eps else s
    if not _zscore_nonzero(a, floor):
        a = (a - a.mean()) / floor(a.std())      # dataset.py:26-32: nothing but background
    a = np.clip(a, -clip_val, clip_val)
    a = (a + clip_val) / (2.0 * clip_val)
    return a * 2.0 - 1.0


def pad_to_min_shape(vol, target_shape):
    """ddpm_3d_ldm/dataset.py:44-77: symmetric zero padding of (C, D, H, W), the odd voxel after."""
    import numpy as np
    widths = [(0, 0)]
    for have, want in zip(vol.shape[1:], target_shape):
        missing = max(want - have, 0)
        widths.append((missing // 2, missing - missing // 2))
    return np.pad(vol, widths, mode="constant") if any(b or a for b, a in widths) else vol


def crop_patch(vol, patch_size, random_crop: bool = True, rng=None):
    """ddpm_3d_ldm/dataset.py:80-105: centre crop, or a random one drawing z, y, x starts from
    Python's `random` (only along axes with room)."""
    import random as _random
    rng = rng or _random
    starts = []
    for have, want in zip(vol.shape[1:], patch_size):
        if have < want:
            raise ValueError("Volume is smaller than patch even after padding.")
        room = have - want
        starts.append((rng.randint(0, room) if room > 0 else 0) if random_crop else room // 2)
    z, y, x = starts
    return vol[:, z:z + patch_size[0], y:y + patch_size[1], x:x + patch_size[2]]


def load_volume_patch(vols_hwd, patch_size, random_crop: bool = True, rng=None):
    """ddpm_3d_ldm/dataset.py:160-185 after the file read: (H, W, D) modality arrays ->
    (C, pd, ph, pw) float32."""
    import numpy as np
    stack = np.stack([normalize_volume(np.transpose(v, (2, 0, 1))) for v in vols_hwd], axis=0)
    return crop_patch(pad_to_min_shape(stack, patch_size), patch_size, random_crop, rng)
