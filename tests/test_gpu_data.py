"""Data-path kernels (csrc/data_path.cu through mri_image_generation_b200/data.py and the Dataset
mirrors) against the oracle's restatement of the reference Dataset arithmetic and against the
committed outputs of the unmodified reference (tests/golden/data_path.pt).

Tolerance: the reference computes the slice / volume mean and deviation with numpy's float32
pairwise sums, the kernels with fp64 sums rounded once; both are within a few fp32 ulp of the
exact value, which moves the z-scores by <~ 2e-6 relative.  Outputs live in [-1, 1]:
|difference| <= 2e-6 is asserted (5e-6 where clipping at +-5 sigma is hit by design)."""
import random

import numpy as np
import pytest
import torch

from helpers import load_gold
from oracle import reference_oracle as O

pytestmark = pytest.mark.gpu
ATOL = 2e-6


def mri_like(shape, seed, background=0.4, scale=400.0):
    rng = np.random.default_rng(seed)
    v = rng.gamma(2.0, scale / 2.0, size=shape).astype(np.float32)
    v[rng.random(shape) < background] = 0.0
    return v


def test_masked_stats_both_layouts():
    from mri_image_generation_b200 import ops
    vol = mri_like((37, 29, 45), 1)
    vol[:, :, 5] = 0
    vol[:, :, 6] = np.where(vol[:, :, 6] != 0, 77.5, 0)
    x = torch.from_numpy(vol).cuda()
    want = np.zeros((45, 2), np.float32)
    for z in range(45):
        s = vol[:, :, z]
        nz = s[s != 0]
        want[z] = (nz.mean(), nz.std() if nz.std() > 0 else 1.0) if nz.size else (0.0, 1.0)
    got_items = ops.masked_stats(x, 2).cpu().numpy()                           # item axis contiguous
    got_cols = ops.masked_stats(x.permute(2, 0, 1).contiguous(), 0).cpu().numpy()   # columns contiguous
    got_one = np.stack([ops.masked_stats(x[:, :, z:z + 1], 2).cpu().numpy()[0] for z in (0, 5, 6, 44)])
    assert np.allclose(got_items, want, rtol=2e-6, atol=1e-6)
    assert np.array_equal(got_items, got_cols)
    assert np.array_equal(got_one, got_items[[0, 5, 6, 44]])
    assert tuple(got_items[5]) == (0.0, 1.0) and got_items[6][1] == 1.0
    # a whole volume as one item, 3-D rule (std < eps -> 1)
    nz = vol[vol != 0]
    whole = ops.masked_stats(x, None, eps=1e-6).cpu().numpy()[0]
    assert np.allclose(whole, (nz.mean(dtype=np.float64), nz.std(dtype=np.float64)), rtol=1e-6)
    tiny = torch.full((4, 4, 4), 3.0, device="cuda")
    tiny[0, 0, 0] = 3.0 + 2.4e-7
    assert ops.masked_stats(tiny, None, eps=1e-6).cpu()[0, 1] == 1.0
    assert 0 < ops.masked_stats(tiny, None, eps=0.0).cpu()[0, 1] < 1e-6


@pytest.mark.parametrize("shape,size", [((240, 240, 155), 128), ((240, 240, 12), 240),
                                        ((61, 83, 9), 40), ((16, 12, 3), 64), ((5, 7, 2), 1)])
def test_preprocess_slices_equal_the_oracle(shape, size):
    from mri_image_generation_b200 import data
    vol = mri_like(shape, shape[0] + size)
    vol[:, :, 0] = 0
    got = data.preprocess_slices(torch.from_numpy(vol).cuda(), size).cpu()
    assert got.shape == (shape[2], size, size)
    zs = range(shape[2]) if shape[2] <= 12 else (0, 1, 40, 77, 154)
    for z in zs:
        want = O.preprocess_slice(vol[:, :, z], size)[0]
        assert (got[z] - want).abs().max().item() <= ATOL, z
    assert torch.all(got[0] == 0)
    # a sub-range writes the same values
    if shape[2] > 2:
        part = data.preprocess_slices(torch.from_numpy(vol).cuda(), size, z0=1, z1=3).cpu()
        assert torch.equal(part, got[1:3])


def test_slices_and_volumes_match_the_reference_fixtures():
    from mri_image_generation_b200 import data
    g = load_gold("data_path.pt")
    vol = g["slice_vol"].cuda()
    for size, key in ((16, "slices_16"), (32, "slices_32")):
        got = data.preprocess_slices(vol, size).cpu()
        assert (got - g[key][:, 0]).abs().max().item() <= ATOL, key
    vols = [v.cuda() for v in g["vols_hwd"]]
    for m, v in enumerate(vols):
        n = data.normalize_volume(v.permute(2, 0, 1)).cpu()
        assert (n - g["normalized"][m]).abs().max().item() <= ATOL, m
        n2 = data.normalize_volume(v.permute(2, 0, 1).contiguous()).cpu()     # direct kernel
        assert torch.equal(n, n2)
    for name, c in g["patches"].items():
        random.seed(c["seed"])
        got = data.load_patch(vols, c["patch"], c["random_crop"]).cpu()
        assert random.random() == c["next_random"], name
        assert got.shape == c["out"].shape, name
        assert (got - c["out"]).abs().max().item() <= ATOL, name
    # the helper trio composed like the reference composes it
    stack = torch.stack([data.normalize_volume(v.permute(2, 0, 1)) for v in vols])
    c = g["patches"]["random_pad"]
    random.seed(c["seed"])
    out = data.random_or_center_crop(data.pad_to_min_shape(stack, c["patch"]), c["patch"], True)
    assert (out.cpu() - c["out"]).abs().max().item() <= ATOL


def test_full_size_volume_patch_equals_the_oracle():
    """BraTS geometry: (240, 240, 155) volumes, the training patch (128, 160, 160) and a patch
    that needs padding along D (160 > 155)."""
    from mri_image_generation_b200 import data
    vols = [mri_like((240, 240, 155), 50 + m, scale=300 + 150 * m) for m in range(2)]
    dev = [torch.from_numpy(v).cuda() for v in vols]
    for patch, rnd, seed in (((128, 160, 160), True, 3), ((160, 192, 250), False, 0)):
        random.seed(seed)
        got = data.load_patch(dev, patch, rnd).cpu().numpy()
        random.seed(seed)
        want = O.load_volume_patch(vols, patch, rnd)
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= ATOL
        assert got.min() >= -1 and got.max() <= 1


def _write_subjects(root, n, shape, seed=0):
    rng = np.random.default_rng(seed)
    vols = {}
    for i in range(n):
        d = root / f"sub{i:02d}"
        d.mkdir()
        for m in ("flair", "t1", "t1ce", "t2"):
            v = rng.gamma(2.0, 200.0, size=shape).astype(np.float32)
            v[rng.random(shape) < 0.4] = 0
            with open(d / f"sub{i:02d}_{m}.nii.gz", "wb") as f:
                np.save(f, v)
            vols[(i, m)] = v
    return vols


def test_dataset_mirrors_return_what_the_reference_returns(tmp_path):
    from mri_image_generation_b200 import data
    from mri_image_generation_b200.model_scripts.ddpm_25d_all_modalities.dataset import BraTSSliceDataset as DS25
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.dataset import BraTS3DVolumeDataset
    from mri_image_generation_b200.model_scripts.slice_cond_2d_ddpm.dataset import BraTSSliceDataset as DS2
    vols = _write_subjects(tmp_path, 2, (30, 26, 20))
    src = data.NpySource()
    ds2 = DS2(tmp_path, image_size=16, source=src)
    for idx in (0, 7, len(ds2) - 1):
        path, z = ds2.slice_tuples[idx]
        i = int(path.parent.name[3:])
        s, zp = ds2[idx]
        assert s.is_cuda and s.shape == (1, 16, 16) and zp == np.float32(z / 19)
        assert (s.cpu() - O.preprocess_slice(vols[(i, "flair")][:, :, z], 16)).abs().max() <= ATOL
    assert len(ds2._cache) == 2
    ds25 = DS25(tmp_path, image_size=16, slice_radius=1, source=src)
    xc, xx, zp = ds25[3]
    path, z = ds25.slice_tuples[3]
    order = ("t1", "t1ce", "t2", "flair")
    want_c = torch.cat([O.preprocess_slice(vols[(0, m)][:, :, z], 16) for m in order])
    want_x = torch.cat([O.preprocess_slice(vols[(0, m)][:, :, zz], 16) for zz in (z - 1, z + 1) for m in order])
    assert xc.shape == (4, 16, 16) and xx.shape == (8, 16, 16) and zp == np.float32(z / 19)
    assert (xc.cpu() - want_c).abs().max() <= ATOL and (xx.cpu() - want_x).abs().max() <= ATOL
    # through a DataLoader, as model.py:103-110 builds it (device tensors: no workers, no pinning)
    xb, cb, zb = next(iter(torch.utils.data.DataLoader(ds25, batch_size=4, shuffle=False)))
    assert xb.is_cuda and xb.shape == (4, 4, 16, 16) and cb.shape == (4, 8, 16, 16) and zb.shape == (4,)
    ds3 = BraTS3DVolumeDataset(tmp_path, patch_size=(16, 32, 24), random_crop=True, source=src)
    random.seed(11)
    got = ds3[0]
    random.seed(11)
    case = ds3.cases[0]
    i = int(case[0].parent.name[3:])
    want = O.load_volume_patch([vols[(i, m)] for m in ("flair", "t1", "t1ce", "t2")], (16, 32, 24), True)
    assert got.is_cuda and got.shape == (4, 16, 32, 24)
    assert np.abs(got.cpu().numpy() - want).max() <= ATOL
