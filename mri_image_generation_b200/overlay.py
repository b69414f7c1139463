"""Run the reference's OWN scripts over the B200 path without editing them.

The reference's drivers (`model_scripts/<pkg>/model.py`, `train.py`, `show_model.py`, ...) pull the
hot-path classes in with relative imports (`from .unet import UNet`,
`slice_cond_2d_ddpm/model.py:15-17`; `from .unet_attention import UNet3DModelWithAttention`,
`ddpm_3d_ldm/train.py:23-26`).  A relative import resolves through `sys.modules` under the
absolute name `model_scripts.<pkg>.<module>` first, so registering the drop-in modules under those
names BEFORE the script is imported makes every such line bind to the B200 classes; everything
else in the package (dataset, helpers, the script itself) is still the reference's file.

    python -m mri_image_generation_b200.overlay -m model_scripts.ddpm_3d_ldm.train

from the root of a reference checkout is `python -m model_scripts.ddpm_3d_ldm.train` with the
UNet / diffusion (/ VAE) modules replaced.  `install()` is the same thing as a function, e.g. for
a `sitecustomize.py`.  Nothing here falls back: a hot-path module that fails to import raises.
"""
from __future__ import annotations

import argparse
import importlib
import runpy
import sys
from typing import Dict, Iterable, List, Optional

# reference module -> what it defines on the hot path
HOT_MODULES: Dict[str, List[str]] = {
    "slice_cond_2d_ddpm": ["unet", "diffusion"],
    "ddpm_25d_all_modalities": ["unet", "diffusion"],
    "ddpm_3d_ldm": ["unet", "unet_attention", "diffusion", "vae"],
}
# optional: the datasets with the slice / volume arithmetic on the device (need num_workers=0)
DATA_MODULES: Dict[str, List[str]] = {k: ["dataset"] for k in HOT_MODULES}


def install(packages: Optional[Iterable[str]] = None, *, vae: bool = True, datasets: bool = False,
            overlap_ddp: bool = False, fused_adam: bool = False, precision: Optional[str] = None,
            root: str = "model_scripts") -> List[str]:
    """Alias the drop-in modules as `<root>.<pkg>.<module>`.  Returns the aliased names.

    vae=False keeps the reference's `vae.py` (stage 1 of ddpm_3d_ldm/train.py on the reference
    implementation); datasets=True also replaces `dataset.py` by the device data path and
    `torch.utils.data.DataLoader` by one that drops the worker / pinning options for datasets
    whose items already are device tensors (data.device_dataloader);
    overlap_ddp=True makes `from torch.nn.parallel import DistributedDataParallel` in the scripts
    (ddpm_3d_ldm/train.py:16) resolve to the wrapper that overlaps the gradient all-reduce with
    the backward launch list (modules it does not know are handed to torch's wrapper);
    fused_adam=True makes `torch.optim.Adam(...)` (train.py:242-243, model.py:126) the one-launch
    Adam of this package (same constructor, update rule, state_dict and GradScaler protocol);
    precision="split" makes every UNet / VAE the script builds run in split precision (fp32-class
    parity with the un-autocast sampling of the show_model scripts; inference only)."""
    done = []
    for pkg in (packages or HOT_MODULES):
        if pkg not in HOT_MODULES:
            raise ValueError(f"unknown reference package {pkg!r}; have {sorted(HOT_MODULES)}")
        mods = list(HOT_MODULES[pkg]) + (DATA_MODULES[pkg] if datasets else [])
        for mod in mods:
            if mod == "vae" and not vae:
                continue
            name = f"{root}.{pkg}.{mod}"
            if name in sys.modules and not sys.modules[name].__name__.startswith(__package__ + "."):
                raise RuntimeError(f"{name} was imported before overlay.install(): the script "
                                   "already holds the reference classes")
            sys.modules[name] = importlib.import_module(f"{__package__}.model_scripts.{pkg}.{mod}")
            done.append(name)
    if datasets:
        import torch.utils.data as tud

        from . import data
        if not hasattr(tud, "_mri_torch_dataloader"):
            tud._mri_torch_dataloader = tud.DataLoader
        tud.DataLoader = data.device_dataloader(tud._mri_torch_dataloader)
        done.append("torch.utils.data.DataLoader")
    if precision is not None:
        if precision not in ("bf16", "split"):
            raise ValueError(f"unknown precision {precision!r}")
        from .model_scripts.ddpm_3d_ldm import unet_attention as _u3, vae as _v
        from .model_scripts.slice_cond_2d_ddpm import unet as _u2
        for cls in (_u3._UNet3DBase, _u2._UNet2DBase, _v.VAE3D):
            cls.precision = precision
        done.append(f"precision={precision}")
    if fused_adam:
        import torch.optim as topt

        from . import optim
        if not hasattr(topt, "_mri_torch_adam"):
            topt._mri_torch_adam = topt.Adam
        topt.Adam = optim.Adam
        done.append("torch.optim.Adam")
    if overlap_ddp:
        import torch.nn.parallel as tnp

        from . import parallel
        if not hasattr(tnp, "_mri_torch_ddp"):
            tnp._mri_torch_ddp = tnp.DistributedDataParallel
        tnp.DistributedDataParallel = parallel.ddp_for_scripts(tnp._mri_torch_ddp)
        done.append("torch.nn.parallel.DistributedDataParallel")
    return done


def uninstall(root: str = "model_scripts") -> None:
    for name in [n for n, m in sys.modules.items()
                 if n.startswith(root + ".") and m.__name__.startswith(__package__ + ".")]:
        del sys.modules[name]
    import torch.nn.parallel as tnp
    import torch.utils.data as tud
    if hasattr(tnp, "_mri_torch_ddp"):
        tnp.DistributedDataParallel = tnp._mri_torch_ddp
        del tnp._mri_torch_ddp
    if hasattr(tud, "_mri_torch_dataloader"):
        tud.DataLoader = tud._mri_torch_dataloader
        del tud._mri_torch_dataloader
    import torch.optim as topt
    if hasattr(topt, "_mri_torch_adam"):
        topt.Adam = topt._mri_torch_adam
        del topt._mri_torch_adam
    from .model_scripts.ddpm_3d_ldm import unet_attention as _u3, vae as _v
    from .model_scripts.slice_cond_2d_ddpm import unet as _u2
    for cls in (_u3._UNet3DBase, _u2._UNet2DBase, _v.VAE3D):
        cls.precision = "bf16"


def main(argv=None) -> None:
    ap = argparse.ArgumentParser(prog="python -m mri_image_generation_b200.overlay",
                                 description=__doc__.split("\n\n")[0])
    ap.add_argument("--keep-vae", action="store_true", help="leave ddpm_3d_ldm/vae.py to the reference")
    ap.add_argument("--device-datasets", action="store_true", help="also replace <pkg>/dataset.py")
    ap.add_argument("--overlap-ddp", action="store_true",
                    help="DistributedDataParallel -> the wrapper overlapping all-reduce and backward")
    ap.add_argument("--fused-adam", action="store_true", help="torch.optim.Adam -> the one-launch Adam")
    ap.add_argument("--precision", choices=["bf16", "split"], default=None,
                    help="split: fp32-class parity for the sampling scripts (inference only)")
    ap.add_argument("--path", action="append", default=[], help="prepend to sys.path (stub modules ...)")
    ap.add_argument("-m", dest="module", required=True, help="the reference script, as for python -m")
    ap.add_argument("args", nargs=argparse.REMAINDER)
    ns = ap.parse_args(argv)
    for p in reversed(ns.path):
        sys.path.insert(0, p)
    if "" not in sys.path and "." not in sys.path:
        sys.path.insert(0, "")      # what `python -m` itself does: the reference checkout is the cwd
    names = install(vae=not ns.keep_vae, datasets=ns.device_datasets, overlap_ddp=ns.overlap_ddp,
                    fused_adam=ns.fused_adam, precision=ns.precision, root=ns.module.split(".")[0])
    print(f"[mri_b200.overlay] {len(names)} modules bound to the B200 path: {', '.join(names)}", flush=True)
    sys.argv = [ns.module] + list(ns.args)
    runpy.run_module(ns.module, run_name="__main__", alter_sys=True)


if __name__ == "__main__":
    main()
