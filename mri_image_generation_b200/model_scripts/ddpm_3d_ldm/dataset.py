"""Drop-in for model_scripts/ddpm_3d_ldm/dataset.py (BraTS3DVolumeDataset and its helpers) with
the volume arithmetic on the device (mri_image_generation_b200/data.py)."""
from collections import OrderedDict
from pathlib import Path

import torch

from ... import data

# the reference's module-level helpers, for CUDA tensors (dataset.py:11-105)
_normalize_volume = data.normalize_volume
_pad_to_min_shape = data.pad_to_min_shape
_random_or_center_crop = data.random_or_center_crop


class BraTS3DVolumeDataset(torch.utils.data.Dataset):
    """ddpm_3d_ldm/dataset.py:108-191: the four modalities of every subject that has all of them,
    as one (4, D, H, W) patch in [-1, 1].  The raw volumes and their statistics stay on the device
    (least recently used subjects beyond `cache_size` are dropped); a visit costs one patch kernel
    per modality.  Random crops consume Python's `random` exactly as the reference does.  Tensors
    live on `device`: DataLoader with num_workers=0, pin_memory=False."""

    MRI_DEVICE_DATASET = True

    def __init__(self, root_dir, patch_size=(128, 160, 160), random_crop=True,
                 modalities=("flair", "t1", "t1ce", "t2"), device=None, source=None, cache_size=32):
        super().__init__()
        self.root_dir = Path(root_dir)
        self.patch_size = tuple(patch_size)
        self.random_crop = random_crop
        self.modalities = modalities
        self.device = data._device(device)
        self.source = source if source is not None else data.NibabelSource(fdata=True)
        self.cache_size = cache_size
        self._cache = OrderedDict()
        self.cases = self._find_cases()
        if len(self.cases) == 0:
            raise ValueError(f"No BraTS cases found in {root_dir}")
        print(f"Found {len(self.cases)} BraTS subjects.")

    def _find_cases(self):
        # dataset.py:139-154
        cases = []
        for flair_path in list(self.root_dir.rglob("*_flair.nii.gz")):
            base = str(flair_path).replace("_flair.nii.gz", "")
            paths = {"flair": Path(flair_path), "t1": Path(base + "_t1.nii.gz"),
                     "t1ce": Path(base + "_t1ce.nii.gz"), "t2": Path(base + "_t2.nii.gz")}
            if all(p.exists() for p in paths.values()):
                cases.append(tuple(paths[m] for m in self.modalities))
        return cases

    def __len__(self):
        return len(self.cases)

    def _resident(self, paths):
        key = tuple(str(p) for p in paths)
        if key in self._cache:
            self._cache.move_to_end(key)
            return self._cache[key]
        vols, stats = [], []
        for p in paths:
            vol = self.source.load(p)
            if vol.ndim == 4:                      # dataset.py:171-173
                vol = vol[..., 0]
            v = data.to_device(vol, self.device)
            vols.append(v)
            stats.append(data._volume_stats(v, 1e-6))
        self._cache[key] = (vols, stats)
        if self.cache_size is not None and len(self._cache) > self.cache_size:
            self._cache.popitem(last=False)
        return vols, stats

    def _load_volume(self, paths):
        """dataset.py:160-185 -> (4, D, H, W) CUDA tensor."""
        with torch.cuda.device(self.device):
            vols, stats = self._resident(paths)
            return data.load_patch(vols, self.patch_size, self.random_crop, stats=stats)

    def __getitem__(self, idx):
        return self._load_volume(self.cases[idx])
