import io

import torch


def log_model(model, artifact_path=None, **_kw):
    """mlflow pickles the whole module (slice_cond_2d_ddpm/model.py:320): do the same, and load it
    back, so that an unpicklable drop-in would fail here as it would under the real mlflow."""
    from . import RECORD
    buf = io.BytesIO()
    torch.save(model, buf)
    buf.seek(0)
    again = torch.load(buf, weights_only=False)
    assert type(again) is type(model)
    RECORD["models"].append({"artifact_path": artifact_path, "class": type(model).__name__,
                             "pickled_bytes": buf.getbuffer().nbytes})
