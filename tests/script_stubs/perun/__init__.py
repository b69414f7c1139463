"""Stand-in for perun (energy tracking): the decorator just calls the function
(helpers/perun_utils.py:141-143, ddpm_25d_all_modalities/model.py:372-381)."""
import functools

config = None
_callbacks = []


def perun(data_out=None, format="json", **_kw):
    def deco(fn):
        @functools.wraps(fn)
        def wrapped(*a, **k):
            return fn(*a, **k)
        return wrapped
    return deco


def register_callback(fn):
    _callbacks.append(fn)


from . import processing  # noqa: E402,F401
from .data_model import data  # noqa: E402,F401
