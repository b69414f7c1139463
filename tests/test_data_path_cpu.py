"""Data path, CPU side: the oracle's restatement of the reference Dataset arithmetic against the
committed outputs of the unmodified reference (tests/golden/data_path.pt, made by
oracle/make_golden_data.py), and the host logic of mri_image_generation_b200/data.py (index
building, pad / crop geometry, `random` consumption, refusal to run without CUDA)."""
import random

import numpy as np
import pytest
import torch

from helpers import load_gold
from oracle import reference_oracle as O


def test_oracle_slices_match_reference_outputs():
    g = load_gold("data_path.pt")
    vol = g["slice_vol"].numpy()
    for z in range(vol.shape[2]):
        assert torch.allclose(O.preprocess_slice(vol[:, :, z], 16), g["slices_16"][z], atol=1e-6), z
        assert torch.allclose(O.preprocess_slice(vol[:, :, z], 32), g["slices_32"][z], atol=1e-6), z
    # the empty slice is the mid-grey 0.0 everywhere; the constant one is 0.0 on background,
    # (0 + 5) / 10 * 2 - 1 = 0 on foreground too (z-score of a constant with std -> 1)
    assert torch.all(g["slices_16"][2] == 0) and torch.all(g["slices_16"][3] == 0)
    assert vol[:, :, 0].any() and not torch.all(g["slices_16"][0] == 0)


def test_oracle_volume_pipeline_matches_reference_outputs():
    g = load_gold("data_path.pt")
    vols = [v.numpy() for v in g["vols_hwd"]]
    for m, v in enumerate(vols):
        n = O.normalize_volume(np.transpose(v, (2, 0, 1)))
        assert np.allclose(n, g["normalized"][m].numpy(), atol=1e-6), m
    assert torch.all(g["normalized"][3] == 0)            # background-only modality
    for name, c in g["patches"].items():
        random.seed(c["seed"])
        got = O.load_volume_patch(vols, c["patch"], c["random_crop"])
        assert got.shape == (4,) + tuple(c["patch"]), name
        assert np.allclose(got, c["out"].numpy(), atol=1e-6), name
        assert random.random() == c["next_random"], name   # same draws from `random`


def test_pad_and_crop_geometry_equals_oracle():
    from mri_image_generation_b200 import data
    rng = np.random.default_rng(0)
    for _ in range(50):
        shape = tuple(int(v) for v in rng.integers(1, 12, size=3))
        patch = tuple(int(v) for v in rng.integers(1, 12, size=3))
        pads = data.pad_amounts(shape, patch)
        vol = rng.standard_normal((2,) + shape).astype(np.float32)
        padded = O.pad_to_min_shape(vol, patch)
        assert padded.shape[1:] == tuple(s + b + a for s, (b, a) in zip(shape, pads))
        for rnd in (False, True):
            random.seed(7)
            start = data.crop_start(padded.shape[1:], patch, rnd)
            r1 = random.random()
            random.seed(7)
            want = O.crop_patch(padded, patch, rnd)
            assert r1 == random.random()
            z, y, x = start
            assert np.array_equal(padded[:, z:z + patch[0], y:y + patch[1], x:x + patch[2]], want)
    with pytest.raises(ValueError):
        data.crop_start((4, 4, 4), (4, 5, 4), False)


def _write_subjects(root, n, shape, seed=0, missing=None):
    rng = np.random.default_rng(seed)
    for i in range(n):
        d = root / f"sub{i:02d}"
        d.mkdir()
        for m in ("flair", "t1", "t1ce", "t2"):
            if missing == (i, m):
                continue
            v = rng.gamma(2.0, 200.0, size=shape).astype(np.float32)
            v[rng.random(shape) < 0.4] = 0
            with open(d / f"sub{i:02d}_{m}.nii.gz", "wb") as f:
                np.save(f, v)


def test_dataset_index_mirrors_the_reference(tmp_path, capsys):
    from mri_image_generation_b200 import data
    from mri_image_generation_b200.model_scripts.ddpm_25d_all_modalities.dataset import BraTSSliceDataset as DS25
    from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.dataset import BraTS3DVolumeDataset
    from mri_image_generation_b200.model_scripts.slice_cond_2d_ddpm.dataset import BraTSSliceDataset as DS2
    _write_subjects(tmp_path, 3, (6, 5, 20), missing=(1, "t2"))
    src = data.NpySource()
    ds2 = DS2(tmp_path, image_size=8, device="cuda:0", source=src)
    # slice_cond_2d_ddpm/dataset.py:34-38: z in [int(0.1 D), int(0.9 D))
    assert len(ds2) == 3 * 16 and [z for _, z in ds2.slice_tuples[:16]] == list(range(2, 18))
    ds25 = DS25(tmp_path, image_size=8, slice_radius=2, device="cuda:0", source=src)
    # ddpm_25d_all_modalities/dataset.py:47-50: shrunk by the radius on both sides
    assert len(ds25) == 3 * 12 and ds25.slice_tuples[0][1] == 4 and ds25.slice_tuples[11][1] == 15
    ds3 = BraTS3DVolumeDataset(tmp_path, patch_size=(4, 4, 4), device="cuda:0", source=src)
    assert len(ds3) == 2                                   # the subject without t2 is dropped
    assert [p.name.split("_")[1].split(".")[0] for p in ds3.cases[0]] == ["flair", "t1", "t1ce", "t2"]
    assert "Found 2 BraTS subjects." in capsys.readouterr().out
    with pytest.raises(RuntimeError):
        DS2(tmp_path / "sub00", modality_suffix="_nothing.nii.gz", device="cuda:0", source=src)


def test_no_cpu_path():
    from mri_image_generation_b200 import _lib, data, ops
    with pytest.raises(_lib.MriError):
        data._device("cpu")
    with pytest.raises(_lib.MriError):
        ops.masked_stats(torch.zeros(4, 4, 4), 2)
    with pytest.raises(_lib.MriError):
        data.preprocess_slices(torch.zeros(4, 4, 4), 8)


def test_device_dataloader_drops_worker_options_only_for_device_datasets():
    """overlay --device-datasets: the scripts' DataLoader(..., num_workers=8, pin_memory=True)
    (ddpm_3d_ldm/train.py:180-188) must not fork workers / pin when items are device tensors,
    also behind Subset and random_split (slice_cond_2d_ddpm/model.py:73-82)."""
    import torch
    import torch.utils.data as tud

    from mri_image_generation_b200 import data

    class Dev(tud.Dataset):
        MRI_DEVICE_DATASET = True

        def __len__(self):
            return 10

        def __getitem__(self, i):
            return torch.full((2,), float(i))

    class Host(Dev):
        MRI_DEVICE_DATASET = False

    DL = data.device_dataloader(tud.DataLoader)
    sub = tud.Subset(Dev(), list(range(8)))
    a, b = tud.random_split(sub, [6, 2])
    for ds in (Dev(), sub, a, tud.ConcatDataset([Host(), Dev()])):
        dl = DL(ds, batch_size=4, shuffle=True, num_workers=8, pin_memory=True, worker_init_fn=print,
                prefetch_factor=2)
        assert dl.num_workers == 0 and dl.pin_memory is False and dl.worker_init_fn is None
        assert next(iter(dl)).shape == (4, 2)
    dl = DL(Host(), batch_size=2, num_workers=2, pin_memory=False)
    assert dl.num_workers == 2
    assert not data.is_device_dataset(Host()) and data.is_device_dataset(a)
