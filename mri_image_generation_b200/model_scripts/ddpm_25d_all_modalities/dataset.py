"""Drop-in for model_scripts/ddpm_25d_all_modalities/dataset.py (multi-modal 2.5-D
BraTSSliceDataset) with the slice arithmetic on the device (mri_image_generation_b200/data.py)."""
from pathlib import Path

import numpy as np
import torch

from ... import data


class BraTSSliceDataset(data.SliceDatasetBase):
    """ddpm_25d_all_modalities/dataset.py:10-155: item = (x_center (4, S, S): t1, t1ce, t2, flair
    at z; x_context (4 * 2 * slice_radius, S, S): the neighbours z-r..z+r without z, modalities
    inner; z / (D - 1)).  Tensors live on `device` (DataLoader: num_workers=0, pin_memory=False);
    each subject is preprocessed once into a (D, 4, S, S) device tensor, so x_center is a view."""

    def __init__(self, root_dir, image_size=128, slice_radius=1, device=None, source=None,
                 cache_size=None):
        super().__init__()
        self.root_dir = Path(root_dir)
        self.image_size = image_size
        self.slice_radius = slice_radius
        self.modalities = ["_t1.nii.gz", "_t1ce.nii.gz", "_t2.nii.gz", "_flair.nii.gz"]
        self.device = data._device(device)
        self.source = source if source is not None else data.NibabelSource()
        self._build_index(self.flair_suffix, slice_radius)
        self._cache = data._SliceCache(cache_size)

    def subject(self, flair_path) -> torch.Tensor:
        """All preprocessed slices of one subject, (D, 4, S, S) on the device."""
        return self._preprocessed(flair_path, self.modalities, self.flair_suffix)

    def __getitem__(self, idx):
        flair_path, z = self.slice_tuples[idx]
        vol = self.subject(flair_path)
        r = self.slice_radius
        x_context = torch.cat([vol[z + dz] for dz in range(-r, r + 1) if dz != 0], dim=0)
        z_pos = np.float32(z / (vol.shape[0] - 1))
        return vol[z], x_context, z_pos
