"""Minimal in-memory stand-in for the mlflow calls the reference scripts make
(model_scripts/*/model.py, ddpm_3d_ldm/train.py): experiments, one active run, params / metrics /
artifacts recorded in RECORD (printed by tools/run_reference_scripts.py)."""
import contextlib
import json
import os

RECORD = {"experiment": None, "params": {}, "metrics": {}, "artifacts": [], "models": []}
_active = None


class _Run:
    def __init__(self, name):
        self.info = type("Info", (), {"run_id": name or "run", "run_name": name})()


def set_experiment(name):
    RECORD["experiment"] = name


@contextlib.contextmanager
def start_run(run_name=None, **_kw):
    global _active
    _active = _Run(run_name)
    try:
        yield _active
    finally:
        _active = None
        out = os.environ.get("MRI_STUB_MLFLOW_OUT")
        if out:
            with open(out, "w") as f:
                json.dump(RECORD, f, indent=1, default=str)


def active_run():
    return _active


def log_param(key, value):
    RECORD["params"][key] = value


def log_params(d):
    RECORD["params"].update(d)


def log_metric(key, value, step=None):
    RECORD["metrics"].setdefault(key, []).append((step, float(value)))


def log_artifact(path, artifact_path=None):
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    RECORD["artifacts"].append(str(path))


from . import pytorch  # noqa: E402,F401
