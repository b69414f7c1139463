"""VAE3D drop-in (encode_to_latent / decode_from_latent) on the B200 engine vs the fp32 CPU oracle
and the committed reference outputs (tests/golden/vae3d.pt).

Tolerance: bf16 operands / fp32 accumulation through ~14 convolutions and GroupNorms -> rel-L2
<= 3e-2 against the fp32 result (torch autocast-bf16 of the reference shows the same order)."""
import contextlib
import io

import pytest
import torch

from helpers import load_gold, rel_l2, synthetic_state_dict
from oracle import reference_oracle as O

pytestmark = pytest.mark.gpu
TOL = 3e-2


def build(kwargs, shapes, seed):
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.vae import VAE3D
    sd = synthetic_state_dict(shapes, seed)
    m = VAE3D(**kwargs)
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [tuple(s) for s in shapes]
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval(), sd


@pytest.mark.parametrize("name", ["vae_b32", "vae_b64_l8"])
def test_vae_encode_decode_vs_reference_fixture(name):
    g = load_gold("vae3d.pt")[name]
    m, sd = build(g["kwargs"], g["shapes"], g["seed"])
    with torch.no_grad():
        mu, logvar = m.encode(g["x"].cuda())
        lat = m.encode_to_latent(g["x"].cuda())
        rec = m.decode_from_latent(g["mu"].cuda())
    assert mu.shape == g["mu"].shape and rec.shape == g["recon_of_mu"].shape
    e_mu, e_lv, e_rec = rel_l2(mu, g["mu"]), rel_l2(logvar, g["logvar"]), rel_l2(rec, g["recon_of_mu"])
    print(f"{name}: rel-L2 mu {e_mu:.3e} logvar {e_lv:.3e} recon {e_rec:.3e}")
    assert e_mu < TOL and e_lv < TOL and e_rec < TOL
    assert torch.equal(lat, mu)
    # forward(): recon, mu, logvar with the module's own reparameterisation noise
    with torch.no_grad():
        torch.manual_seed(3)
        r2, mu2, lv2 = m(g["x"].cuda())
    assert torch.equal(mu2, mu) and r2.shape == g["x"].shape and torch.isfinite(r2).all()


def test_vae_full_shape_roundtrip_matches_oracle():
    """BASELINE cfg4 geometry at reduced size: VAE3D(4, 32, 3, latent 3) on 4x32x48x32 volumes
    (latents 3x8x12x8), odd box counts, batch 3."""
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.vae import VAE3D
    torch.manual_seed(0)
    ref = VAE3D(4, 32, 3, 3)
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    m = ref.cuda().eval()
    x = torch.randn(3, 4, 32, 48, 32)
    with torch.no_grad():
        z = m.encode_to_latent(x.cuda())
        want_mu, _ = O.vae3d_encode(sd, x)
        rec = m.decode_from_latent(want_mu.cuda())
        want_rec = O.vae3d_decode(sd, want_mu)
    assert rel_l2(z, want_mu) < TOL, rel_l2(z, want_mu)
    assert rel_l2(rec, want_rec) < TOL, rel_l2(rec, want_rec)


def test_vae_requires_cuda():
    from mri_image_generation_b200 import _lib
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.vae import VAE3D
    m = VAE3D(4, 32, 3, 3).cuda()
    x = torch.randn(1, 4, 16, 16, 16, device="cuda")
    with pytest.raises(_lib.MriError), torch.no_grad():
        m.encode(x.cpu())
