"""The data path in front of the UNets, on the device.

The reference's Dataset classes normalise / resize / crop ONE slice (or volume) per __getitem__
on the CPU with numpy (slice_cond_2d_ddpm/dataset.py:67-104, ddpm_25d_all_modalities/
dataset.py:79-155, ddpm_3d_ldm/dataset.py:11-105,160-185).  Here a loaded volume is uploaded once
and preprocessed by three kernels behind the C ABI (csrc/data_path.cu): statistics of the
non-zero voxels (`mri_masked_stats`), normalise + bilinear resize of every slice of the volume in
one launch (`mri_slice_normalize_resize`), normalise + zero pad + crop + axis transposition of a
3-D patch in one launch (`mri_volume_normalize_patch`).  With 180 GB of HBM the preprocessed
slices of a whole training set stay resident: __getitem__ becomes a view.

Reading NIfTI files is host IO and stays with nibabel (a `source` object with `load(path)` /
`shape(path)`; anything with those two methods works).  There is no CPU arithmetic path.
"""
from __future__ import annotations

import random
from collections import OrderedDict
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, ops


# ---- host IO -----------------------------------------------------------------------------------
class NibabelSource:
    """load(path) -> float32 ndarray as the reference reads it; shape(path) -> header shape."""

    def __init__(self, fdata: bool = False):
        self.fdata = fdata        # ddpm_3d_ldm/dataset.py:168 uses get_fdata(); the 2-D sets dataobj

    @staticmethod
    def _nib():
        try:
            import nibabel
        except ImportError as e:  # pragma: no cover - nibabel is absent in the build image
            raise _lib.MriError("reading NIfTI files needs nibabel; pass source= for other "
                                "formats") from e
        return nibabel

    def shape(self, path) -> Tuple[int, ...]:
        return tuple(self._nib().load(str(path)).shape)

    def load(self, path) -> np.ndarray:
        img = self._nib().load(str(path))
        if self.fdata:
            return img.get_fdata().astype(np.float32)
        return np.asanyarray(img.dataobj).astype(np.float32)


class NpySource:
    """Volumes stored with numpy.save under the same file names (tests, pre-converted data)."""

    def shape(self, path) -> Tuple[int, ...]:
        return tuple(np.load(str(path), mmap_mode="r").shape)

    def load(self, path) -> np.ndarray:
        return np.load(str(path)).astype(np.float32)


def to_device(vol: np.ndarray, device) -> torch.Tensor:
    t = torch.from_numpy(np.ascontiguousarray(vol, dtype=np.float32))
    return t.to(device, non_blocking=False)


def _device(device) -> torch.device:
    dev = torch.device(device if device is not None else "cuda")
    if dev.type != "cuda":
        raise _lib.MriError("the data path runs on a CUDA device (no CPU path); the reference's "
                            "Dataset classes are the CPU implementation")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


# ---- 2-D slices ----------------------------------------------------------------------------------
def preprocess_slices(vol_hwd: torch.Tensor, image_size, out: Optional[torch.Tensor] = None,
                      z0: int = 0, z1: Optional[int] = None) -> torch.Tensor:
    """Every slice vol[:, :, z], z0 <= z < z1, of an (H, W, D) volume through the reference's
    per-slice pipeline: z-score over the slice's non-zero pixels, clip to +-5, map to [0, 1],
    bilinear resize to image_size, map to [-1, 1] (slice_cond_2d_ddpm/dataset.py:71-98,
    ddpm_25d_all_modalities/dataset.py:79-103).  Returns `out` ([z1 - z0, S, S]; a view with
    contiguous slices is accepted, e.g. cache[:, m] of a [D, 4, S, S] tensor)."""
    if vol_hwd.dim() != 3:
        raise _lib.MriError(f"preprocess_slices: expected an (H, W, D) volume, got {tuple(vol_hwd.shape)}")
    z1 = vol_hwd.shape[2] if z1 is None else z1
    if not 0 <= z0 < z1 <= vol_hwd.shape[2]:
        raise _lib.MriError(f"preprocess_slices: slice range [{z0}, {z1}) outside depth {vol_hwd.shape[2]}")
    sh, sw = (image_size, image_size) if isinstance(image_size, int) else image_size
    x = vol_hwd[:, :, z0:z1]
    if out is None:
        out = torch.empty(z1 - z0, sh, sw, dtype=torch.float32, device=vol_hwd.device)
    ms = ops.masked_stats(x, item_dim=2, eps=0.0)
    ops.slice_normalize_resize(x, 2, ms, out)
    return out


class _SliceCache:
    """path -> preprocessed [D, M, S, S] tensor on the device (M modalities), least recently used
    evicted beyond `cache_size` entries (None: keep everything -- a BraTS training set of 1251
    subjects x 155 slices x 128^2 fp32 is 12.7 GB per modality)."""

    def __init__(self, cache_size: Optional[int]):
        self.cache_size = cache_size
        self._d: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def get(self, key: str, make):
        if key in self._d:
            self._d.move_to_end(key)
            return self._d[key]
        val = make()
        self._d[key] = val
        if self.cache_size is not None and len(self._d) > self.cache_size:
            self._d.popitem(last=False)
        return val

    def __len__(self):
        return len(self._d)


class SliceDatasetBase(torch.utils.data.Dataset):
    """Shared by the 2-D and 2.5-D BraTSSliceDataset mirrors: volume discovery, the (path, z)
    index and the device cache of preprocessed volumes."""

    MRI_DEVICE_DATASET = True

    flair_suffix = "_flair.nii.gz"

    def _build_index(self, anchor_suffix: str, radius: int):
        # slice_cond_2d_ddpm/dataset.py:24-38, ddpm_25d_all_modalities/dataset.py:36-52
        self.volume_paths = sorted(self.root_dir.rglob(f"*{anchor_suffix}"))
        if not self.volume_paths:
            raise RuntimeError(f"No FLAIR files (*{anchor_suffix}) found under {self.root_dir}")
        self.slice_tuples = []
        self._depth = {}
        for p in self.volume_paths:
            shape = self.source.shape(p)
            if len(shape) != 3:
                continue
            D = shape[2]
            self._depth[str(p)] = D
            for z in range(int(0.1 * D) + radius, int(0.9 * D) - radius):
                self.slice_tuples.append((p, z))
        print(f"Found {len(self.volume_paths)} volumes.")
        print(f"Built {len(self.slice_tuples)} (volume, slice) pairs.")

    def _preprocessed(self, anchor_path, suffixes: Sequence[str], anchor_suffix: str) -> torch.Tensor:
        def make():
            S = self.image_size
            D = self._depth[str(anchor_path)]
            cache = torch.empty(D, len(suffixes), S, S, dtype=torch.float32, device=self.device)
            for m, suf in enumerate(suffixes):
                path = str(anchor_path).replace(anchor_suffix, suf)
                vol = to_device(self.source.load(path), self.device)
                preprocess_slices(vol, S, out=cache[:, m])
            return cache
        with torch.cuda.device(self.device):
            return self._cache.get(str(anchor_path), make)

    def __len__(self):
        return len(self.slice_tuples)


# ---- 3-D volumes ---------------------------------------------------------------------------------
def normalize_volume(vol: torch.Tensor, eps: float = 1e-6, clip_val: float = 5.0) -> torch.Tensor:
    """_normalize_volume (ddpm_3d_ldm/dataset.py:11-41) of a (D, H, W)-indexed CUDA tensor of any
    strides: a new contiguous tensor in [-1, 1]; the input is not modified."""
    if vol.dim() != 3:
        raise _lib.MriError(f"normalize_volume: expected (D, H, W), got {tuple(vol.shape)}")
    ms = _volume_stats(vol, eps)
    out = torch.empty(tuple(vol.shape), dtype=torch.float32, device=vol.device)
    ops.volume_normalize_patch(vol, ms, (0, 0, 0), out, clip=clip_val)
    return out


def _volume_stats(vol: torch.Tensor, eps: float) -> torch.Tensor:
    if vol.is_contiguous():
        return ops.masked_stats(vol, None, eps)
    order = sorted(range(3), key=lambda d: -vol.stride(d))
    dense = vol.permute(*order)
    if not dense.is_contiguous():
        raise _lib.MriError("volume statistics need a permutation of a contiguous volume")
    return ops.masked_stats(dense, None, eps)


def pad_amounts(shape_dhw, target_shape):
    """(before, after) per axis of _pad_to_min_shape (ddpm_3d_ldm/dataset.py:51-63)."""
    res = []
    for s, t in zip(shape_dhw, target_shape):
        p = max(t - s, 0)
        res.append((p // 2, p - p // 2))
    return res


def pad_to_min_shape(vol: torch.Tensor, target_shape) -> torch.Tensor:
    """_pad_to_min_shape (ddpm_3d_ldm/dataset.py:44-77) of a (C, D, H, W) CUDA tensor."""
    pads = pad_amounts(vol.shape[1:], target_shape)
    if not any(b or a for b, a in pads):
        return vol
    shape = [vol.shape[0]] + [s + b + a for s, (b, a) in zip(vol.shape[1:], pads)]
    with torch.cuda.device(vol.device):
        out = torch.empty(shape, dtype=torch.float32, device=vol.device)
        ops.memset_zero(out)
        inner = out[:, pads[0][0]:pads[0][0] + vol.shape[1], pads[1][0]:pads[1][0] + vol.shape[2],
                    pads[2][0]:pads[2][0] + vol.shape[3]]
        ops.copy_cast(vol, inner)
    return out


def crop_start(shape_dhw, patch_size, random_crop: bool = True):
    """Start indices of _random_or_center_crop (ddpm_3d_ldm/dataset.py:86-100): the same calls to
    Python's `random` in the same order (z, y, x; a draw only where there is room)."""
    for s, p in zip(shape_dhw, patch_size):
        if s < p:
            raise ValueError("Volume is smaller than patch even after padding.")
    if random_crop:
        return tuple(random.randint(0, s - p) if s - p > 0 else 0 for s, p in zip(shape_dhw, patch_size))
    return tuple((s - p) // 2 for s, p in zip(shape_dhw, patch_size))


def random_or_center_crop(vol: torch.Tensor, patch_size, random_crop: bool = True) -> torch.Tensor:
    """_random_or_center_crop (ddpm_3d_ldm/dataset.py:80-105): a view, like the numpy slice."""
    sz, sy, sx = crop_start(vol.shape[1:], patch_size, random_crop)
    pd, ph, pw = patch_size
    return vol[:, sz:sz + pd, sy:sy + ph, sx:sx + pw]


def load_patch(vols_hwd: Sequence[torch.Tensor], patch_size, random_crop: bool = True,
               stats: Optional[Sequence[torch.Tensor]] = None, eps: float = 1e-6,
               clip_val: float = 5.0) -> torch.Tensor:
    """BraTS3DVolumeDataset._load_volume after the file read (ddpm_3d_ldm/dataset.py:166-185) for
    the raw (H, W, D) modality volumes of one subject: each is z-scored over its non-zero voxels,
    clipped, mapped to [-1, 1], transposed to (D, H, W), zero padded to at least patch_size and
    cropped -- one statistics pass and one patch kernel per modality, nothing else materialised.
    Returns (C, pd, ph, pw)."""
    H, W, D = vols_hwd[0].shape
    for v in vols_hwd:
        if tuple(v.shape) != (H, W, D):
            raise _lib.MriError("load_patch: the modalities of one subject must share a shape")
    pads = pad_amounts((D, H, W), patch_size)
    padded = [s + b + a for s, (b, a) in zip((D, H, W), pads)]
    start = crop_start(padded, patch_size, random_crop)
    origin = [s - b for s, (b, _) in zip(start, pads)]
    dev = vols_hwd[0].device
    with torch.cuda.device(dev):
        out = torch.empty((len(vols_hwd),) + tuple(patch_size), dtype=torch.float32, device=dev)
        for c, v in enumerate(vols_hwd):
            ms = stats[c] if stats is not None else _volume_stats(v, eps)
            ops.volume_normalize_patch(v.permute(2, 0, 1), ms, origin, out[c], clip=clip_val)
    return out


# ---- DataLoader over device-resident datasets ---------------------------------------------------
def is_device_dataset(ds) -> bool:
    """True for the Dataset mirrors of this package, also behind torch's Subset / random_split /
    ConcatDataset wrappers (slice_cond_2d_ddpm/model.py:71-82, ddpm_3d_ldm/train.py:160-167)."""
    seen = 0
    while ds is not None and seen < 8:
        if getattr(ds, "MRI_DEVICE_DATASET", False):
            return True
        if hasattr(ds, "datasets"):
            return any(is_device_dataset(d) for d in ds.datasets)
        ds = getattr(ds, "dataset", None)
        seen += 1
    return False


def device_dataloader(base):
    """A `torch.utils.data.DataLoader` whose worker / pinning options are overridden when the
    dataset hands out CUDA tensors: the reference's scripts ask for `num_workers=8,
    pin_memory=True` (ddpm_3d_ldm/train.py:180-188), which is how a CPU dataset is fed to a GPU;
    items that are already device tensors are batched in the main process (torch.stack on the
    device) and need neither.  Everything else (batch size, shuffling, samplers) is untouched."""

    class DataLoader(base):
        def __init__(self, dataset, *args, **kwargs):
            if is_device_dataset(dataset):
                if len(args) > 4:
                    raise _lib.MriError("DataLoader over a device dataset: pass num_workers / "
                                        "pin_memory by keyword")
                kwargs["num_workers"] = 0
                kwargs["pin_memory"] = False
                kwargs["persistent_workers"] = False
                kwargs.pop("prefetch_factor", None)
                kwargs.pop("worker_init_fn", None)
            super().__init__(dataset, *args, **kwargs)

    DataLoader.__qualname__ = "DataLoader"
    return DataLoader
