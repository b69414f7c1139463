import numpy as np


class _Axes:
    def __init__(self, fig):
        self._fig = fig

    def imshow(self, img, **_kw):
        a = np.asarray(img)
        assert a.ndim in (2, 3)
        self._fig.images += 1

    def axis(self, *_a, **_k):
        return None

    def set_ylabel(self, *_a, **_k):
        return None

    def set_title(self, *_a, **_k):
        return None


class _Figure:
    def __init__(self):
        self.images = 0
        self.title = ""

    def suptitle(self, t):
        self.title = t


_current = None


def subplots(nrows=1, ncols=1, figsize=None, **_kw):
    global _current
    fig = _Figure()
    _current = fig
    axes = np.empty((nrows, ncols), dtype=object)
    for i in range(nrows):
        for j in range(ncols):
            axes[i, j] = _Axes(fig)
    if nrows == 1 and ncols == 1:
        return fig, axes[0, 0]
    if nrows == 1 or ncols == 1:
        return fig, axes.reshape(-1)
    return fig, axes


def tight_layout(*_a, **_k):
    return None


def savefig(path, **_kw):
    with open(str(path), "w") as f:
        f.write(f"stub figure: {_current.images if _current else 0} images, title {_current.title!r}\n")


def close(*_a, **_k):
    return None
