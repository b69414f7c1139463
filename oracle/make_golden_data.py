"""Generate tests/golden/data_path.pt from the UNMODIFIED reference Dataset code (build container
only; needs /root/reference).

    python oracle/make_golden_data.py

nibabel is absent from this image and is only used by the reference for file IO, so an empty
stand-in module is registered before the import; the functions exercised here
(`BraTSSliceDataset._preprocess_slice`, the 2-D `__getitem__` arithmetic through a fake volume
cache, `_normalize_volume`, `_pad_to_min_shape`, `_random_or_center_crop`,
`BraTS3DVolumeDataset._load_volume` with a fake `nib.load`) never touch a file.  The script also
asserts that oracle/reference_oracle.py reproduces every fixture bit for bit before writing.
"""
from __future__ import annotations

import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import reference_oracle as O  # noqa: E402


def synthetic_volume(shape, seed, background=0.45, scale=400.0):
    """MRI-like: a zero background, positive intensities elsewhere (float32, (H, W, D))."""
    rng = np.random.default_rng(seed)
    v = rng.gamma(2.0, scale / 2.0, size=shape).astype(np.float32)
    v[rng.random(shape) < background] = 0.0
    return v


def main():
    nib = types.ModuleType("nibabel")
    sys.modules["nibabel"] = nib
    from model_scripts.ddpm_25d_all_modalities.dataset import BraTSSliceDataset as DS25
    from model_scripts.slice_cond_2d_ddpm.dataset import BraTSSliceDataset as DS2
    from model_scripts.ddpm_3d_ldm import dataset as d3

    out = {}
    # ---- 2-D / 2.5-D slices ------------------------------------------------------------------
    vol = synthetic_volume((24, 20, 7), 11)
    vol[:, :, 2] = 0.0                      # an empty slice
    vol[:, :, 3] = np.where(vol[:, :, 3] != 0, 123.0, 0.0)   # constant foreground: std == 0
    out["slice_vol"] = torch.from_numpy(vol.copy())
    fake = types.SimpleNamespace(image_size=16)
    res = torch.stack([DS25._preprocess_slice(fake, vol[:, :, z]) for z in range(vol.shape[2])])
    out["slices_16"] = res                                   # (7, 1, 16, 16), downsampling
    fake.image_size = 24
    fake32 = types.SimpleNamespace(image_size=32)
    out["slices_32"] = torch.stack([DS25._preprocess_slice(fake32, vol[:, :, z])
                                    for z in range(vol.shape[2])])   # upsampling, H != W
    for z in range(vol.shape[2]):
        assert torch.equal(res[z], O.preprocess_slice(vol[:, :, z], 16)), z
        assert torch.equal(out["slices_32"][z], O.preprocess_slice(vol[:, :, z], 32)), z
    # the 2-D dataset's own __getitem__ (first visit of each slice; it mutates its cache)
    ds = DS2.__new__(DS2)
    ds.image_size, ds.slice_tuples = 16, [("v", z) for z in range(vol.shape[2])]
    ds._load_volume = lambda path, _v=vol.copy(): _v
    items = [ds[z] for z in range(vol.shape[2])]
    out["items2d_z_pos"] = torch.tensor([float(zp) for _, zp in items])
    for z, (s, zp) in enumerate(items):
        assert torch.equal(s, res[z]), z
        assert zp == np.float32(z / (vol.shape[2] - 1))
    # ---- 3-D volumes ---------------------------------------------------------------------------
    vols = [synthetic_volume((10, 12, 6), 20 + m, scale=300.0 + 100 * m) for m in range(4)]
    vols[3][...] = 0.0                      # a modality that is background only
    out["vols_hwd"] = torch.from_numpy(np.stack(vols))
    norm = [d3._normalize_volume(np.transpose(v, (2, 0, 1)).copy()) for v in vols]
    out["normalized"] = torch.from_numpy(np.stack(norm))     # (4, 6, 10, 12)
    for v, n in zip(vols, norm):
        assert np.array_equal(O.normalize_volume(np.transpose(v, (2, 0, 1))), n)

    class FakeImg:
        def __init__(self, a):
            self.a = a

        def get_fdata(self):
            return self.a.astype(np.float64)

    nib.load = lambda p: FakeImg(vols[int(p)])
    ds3 = d3.BraTS3DVolumeDataset.__new__(d3.BraTS3DVolumeDataset)
    cases = {}
    for name, patch, rnd, seed in [("center_pad", (8, 8, 16), False, 0),   # pad D and W, crop H
                                   ("center_crop", (4, 6, 8), False, 0),
                                   ("random_crop", (3, 7, 5), True, 1234),
                                   ("random_pad", (7, 9, 13), True, 99),
                                   ("identity", (6, 10, 12), True, 5)]:
        ds3.patch_size, ds3.random_crop = patch, rnd
        random.seed(seed)
        got = ds3._load_volume(("0", "1", "2", "3"))
        after = random.random()
        random.seed(seed)
        mine = O.load_volume_patch(vols, patch, rnd)
        assert np.array_equal(got, mine), name
        assert after == random.random(), name
        cases[name] = {"patch": patch, "random_crop": rnd, "seed": seed,
                       "out": torch.from_numpy(np.ascontiguousarray(got)), "next_random": after}
    out["patches"] = cases
    torch.save(out, os.path.join(GOLD, "data_path.pt"))
    print(f"data_path.pt: {os.path.getsize(os.path.join(GOLD, 'data_path.pt')) / 1024:.1f} KiB "
          "(oracle == reference, bit-exact)")


if __name__ == "__main__":
    main()
