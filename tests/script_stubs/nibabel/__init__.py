"""Stand-in for nibabel: `load(path)` fabricates a deterministic BraTS-shaped (240, 240, 155)
volume from the file NAME (zero background, positive 'brain' ellipsoid, modality-dependent
texture); `Nifti1Image` / `save` write the array with numpy.  Enough for
model_scripts/*/dataset.py and ddpm_3d_ldm/show_model.py:157-169."""
import zlib

import numpy as np

SHAPE = (240, 240, 155)
_cache = {}


def _volume(path: str) -> np.ndarray:
    key = path.replace("\\", "/").split("/")[-1]
    vol = _cache.get(key)
    if vol is None:
        rng = np.random.default_rng(zlib.crc32(key.encode()))
        H, W, D = SHAPE
        # coarse random field, repeated up to full resolution (cheap), inside an ellipsoid
        coarse = rng.uniform(200.0, 1200.0, size=(H // 8, W // 8, (D + 4) // 5)).astype(np.float32)
        vol = np.repeat(np.repeat(np.repeat(coarse, 8, 0), 8, 1), 5, 2)[:H, :W, :D]
        vol = vol + rng.normal(0.0, 25.0, size=vol.shape).astype(np.float32)
        hh, ww, dd = np.ogrid[:H, :W, :D]
        inside = (((hh - H / 2) / (0.38 * H)) ** 2 + ((ww - W / 2) / (0.32 * W)) ** 2
                  + ((dd - D / 2) / (0.45 * D)) ** 2) <= 1.0
        vol = np.where(inside, np.maximum(vol, 1.0), 0.0).astype(np.float32)
        if len(_cache) >= 24:
            _cache.pop(next(iter(_cache)))
        _cache[key] = vol
    return vol


class _Image:
    def __init__(self, path=None, data=None, affine=None):
        self._path, self._data = path, data
        self.affine = np.eye(4) if affine is None else affine
        self.header = {}

    @property
    def shape(self):
        return SHAPE if self._data is None else tuple(self._data.shape)

    @property
    def dataobj(self):
        return _volume(self._path) if self._data is None else self._data

    def get_fdata(self, dtype=np.float64):
        return np.asarray(self.dataobj).astype(dtype)


def load(path):
    return _Image(path=str(path))


class Nifti1Image(_Image):
    def __init__(self, dataobj, affine, header=None):
        super().__init__(data=np.asarray(dataobj), affine=affine)


def save(img, path):
    with open(str(path), "wb") as f:
        np.save(f, np.asarray(img.dataobj))
