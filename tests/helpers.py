"""Shared test helpers (tests only)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def synthetic_state_dict(shapes, seed: int):
    """Must match oracle/make_golden.py::synthetic_state_dict."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in shapes:
        w = torch.randn(*shp, generator=g) * 0.05
        if ("norm" in k) and k.endswith(".weight"):
            w = w + 1.0
        sd[k] = w
    return sd


def shapes_of(module):
    return [(k, tuple(v.shape)) for k, v in module.state_dict().items()]


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def load_gold(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)
