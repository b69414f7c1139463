"""Drop-in for model_scripts/slice_cond_2d_ddpm/dataset.py (BraTSSliceDataset) with the slice
arithmetic on the device (mri_image_generation_b200/data.py)."""
from pathlib import Path

import numpy as np
import torch

from ... import data


class BraTSSliceDataset(data.SliceDatasetBase):
    """slice_cond_2d_ddpm/dataset.py:10-104: every *flair.nii.gz under root_dir, the central 80 %
    of its slices; item = (slice in [-1, 1] of shape (1, S, S), z / (D - 1) as np.float32).

    Differences from the reference, both deliberate: the tensors live on `device` (use the
    DataLoader with num_workers=0, pin_memory=False); a volume is preprocessed ONCE, all slices in
    one launch, and kept on the device, so the reference's side effect of z-scoring the cached
    volume in place (dataset.py:68,76: `slice_2d` is a view of the cached array, a second visit
    normalises already-normalised values) does not occur -- every visit returns what the
    reference returns on the first one."""

    def __init__(self, root_dir, modality_suffix="_flair.nii.gz", image_size=128, device=None,
                 source=None, cache_size=None):
        super().__init__()
        self.root_dir = Path(root_dir)
        self.image_size = image_size
        self.modality_suffix = modality_suffix
        self.device = data._device(device)
        self.source = source if source is not None else data.NibabelSource()
        self._build_index(modality_suffix, 0)
        self._cache = data._SliceCache(cache_size)

    def volume(self, path) -> torch.Tensor:
        """All preprocessed slices of one volume, (D, 1, S, S) on the device."""
        return self._preprocessed(path, [self.modality_suffix], self.modality_suffix)

    def __getitem__(self, idx):
        path, z = self.slice_tuples[idx]
        vol = self.volume(path)
        z_pos = np.float32(z / (vol.shape[0] - 1))
        return vol[z], z_pos
