#!/usr/bin/env python
"""Execute the reference's OWN driver scripts, unmodified, over the B200 drop-in.

    python tools/run_reference_scripts.py --reference /path/to/mri-image-generation \
        [--scripts train3d,show3d,model2d,show2d,model25d] [--keep-vae] [--nproc 2 --overlap-ddp]
        [--device-datasets] [--fused-adam] [--split-sampling]

What it does, per script:
  * copies `<reference>/model_scripts` (the .py files) into a scratch work tree -- the scripts
    write checkpoints / samples beside themselves and the checkout may be read-only; not one byte
    of them is edited;
  * fabricates the dataset layout the scripts expect (`../datasets/{train,val,dataset}/<case>/
    <case>_{flair,t1,t1ce,t2}.nii.gz`, empty placeholder files; `tests/script_stubs/nibabel`
    turns each name into a deterministic 240 x 240 x 155 volume) and, for the sampling scripts, the
    checkpoints they load (from the training script's own output when it ran first, else
    random-init drop-in weights);
  * runs `python -m mri_image_generation_b200.overlay -m model_scripts.<pkg>.<script>` in that
    tree with stub mlflow / perun / nibabel / matplotlib on the path: the hot-path modules
    (unet, unet_attention, diffusion, vae) bind to the sm_100a kernels, everything else (dataset,
    helpers, training loop, optimizer, GradScaler, checkpointing) is the reference's code;
  * checks the exit code, that every loss the script printed is finite, and that the files the
    script promises exist; writes `<out>/summary.json` and one log per script.

Needs a B200 (the drop-in has no CPU path) and a reference checkout; `tests/test_gpu_scripts.py`
runs it as a `-m gpu` test whenever both are there.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
STUBS = REPO / "tests" / "script_stubs"
RUN_ID = "b200run"          # SLURM_JOB_ID -> RUN_IDENTIFIER of the training scripts

SCRIPTS = {
    # name: (module, needs, description)
    "train3d": ("model_scripts.ddpm_3d_ldm.train", "ddpm_3d_ldm/train.py: VAE stage + LDM stage (DEBUG_FAST as committed)"),
    "show3d": ("model_scripts.ddpm_3d_ldm.show_model", "ddpm_3d_ldm/show_model.py: VAE sanity, latent stats, DDIM round trips, eps-MSE, sampling + decode"),
    "model2d": ("model_scripts.slice_cond_2d_ddpm.model", "slice_cond_2d_ddpm/model.py: 2D training loop"),
    "show2d": ("model_scripts.slice_cond_2d_ddpm.show_model", "slice_cond_2d_ddpm/show_model.py: pseudo-3D brain, 155 slices x 800 steps"),
    "model25d": ("model_scripts.ddpm_25d_all_modalities.model", "ddpm_25d_all_modalities/model.py: 2.5D training loop"),
}


def find_reference(arg: str | None) -> Path | None:
    for cand in (arg, os.environ.get("MRI_REFERENCE_DIR"), REPO / "baseline" / "_ref", "/root/reference"):
        if cand and (Path(cand) / "model_scripts" / "ddpm_3d_ldm" / "train.py").is_file():
            return Path(cand)
    return None


def stage_tree(ref: Path, work: Path) -> None:
    dst = work / "model_scripts"
    if dst.exists():
        return
    shutil.copytree(ref / "model_scripts", dst,
                    ignore=shutil.ignore_patterns("__pycache__", "*.pt", "*.png", "*.ipynb", "models",
                                                  "samples*", "perun_results"))


def stage_datasets(work: Path, n_train: int, n_val: int, n_plain: int) -> None:
    for sub, n in (("train", n_train), ("val", n_val), ("dataset", n_plain)):
        for i in range(n):
            case = f"BraTS_synth_{sub}_{i:03d}"
            d = work / "datasets" / sub / case
            d.mkdir(parents=True, exist_ok=True)
            for m in ("flair", "t1", "t1ce", "t2"):
                (d / f"{case}_{m}.nii.gz").touch()


def stage_checkpoints(work: Path, name: str) -> str:
    """Checkpoints the sampling scripts load, at the paths and in the formats they expect."""
    sys.path.insert(0, str(REPO))
    import contextlib
    import io

    import torch
    note = ""
    if name == "show3d":
        from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.diffusion import GaussianDiffusionLatent3D
        from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.unet_attention import UNet3DModelWithAttention
        from mri_image_generation_b200.model_scripts.ddpm_3d_ldm.vae import VAE3D
        root = work / "model_scripts" / "ddpm_3d_ldm" / "models"
        dst = root / "1594474"                         # show_model.py:391 RUN_ID
        dst.mkdir(parents=True, exist_ok=True)
        trained = root / RUN_ID
        torch.manual_seed(0)
        vae = VAE3D(in_channels=4, base_channels=32, num_down=3, latent_channels=16)
        unet = UNet3DModelWithAttention(in_channels=16, base_channels=128, channel_mults=(1, 2, 4),
                                        time_emb_dim=256, groups=8, num_heads=4)
        if (trained / "vae3d_final.pt").is_file():
            vae.load_state_dict(torch.load(trained / "vae3d_final.pt", map_location="cpu"), strict=True)
            note += "vae from train.py; "
        if (trained / "3d_ldm_diffusion_best.pt").is_file():
            # train.py:607-608 saves the UNet's state_dict; show_model.py:225 loads the file into the
            # diffusion wrapper with strict=True (reference defect 2): stage what it expects
            unet.load_state_dict(torch.load(trained / "3d_ldm_diffusion_best.pt", map_location="cpu"), strict=True)
            note += "unet from train.py; "
        with contextlib.redirect_stdout(io.StringIO()):
            diff = GaussianDiffusionLatent3D(unet, 16, timesteps=400)
        torch.save(vae.state_dict(), dst / "vae3d_final.pt")
        torch.save(diff.state_dict(), dst / "3d_ldm_diffusion_best.pt")
    elif name == "show2d":
        from mri_image_generation_b200.model_scripts.slice_cond_2d_ddpm.diffusion import GaussianDiffusion
        from mri_image_generation_b200.model_scripts.slice_cond_2d_ddpm.unet import UNet
        root = work / "model_scripts" / "slice_cond_2d_ddpm" / "models"
        dst = root / "1591447"                         # show_model.py:218
        dst.mkdir(parents=True, exist_ok=True)
        torch.manual_seed(0)
        unet = UNet(img_channels=1, base_channels=64, channel_mults=(1, 2, 4, 8), time_emb_dim=256)
        trained = root / RUN_ID / "2d_central_ddpm_flair_best.pt"
        if trained.is_file():
            sd = torch.load(trained, map_location="cpu")
            unet.load_state_dict({k[len("model."):]: v for k, v in sd.items() if k.startswith("model.")},
                                 strict=True)
            note += "unet from model.py; "
        with contextlib.redirect_stdout(io.StringIO()):
            diff = GaussianDiffusion(unet, image_size=128, channels=1, timesteps=800)   # show_model.py:21
        torch.save(diff.state_dict(), dst / "2d_central_ddpm_flair_best.pt")
    return note.strip()


LOSS_RE = re.compile(r"(?:loss|Loss|eps-MSE)[^0-9\-naNif]*[:=]?\s*(-?(?:\d+\.\d+(?:e[-+]?\d+)?|nan|inf))")


def check_log(text: str) -> dict:
    vals = [float(v) for v in LOSS_RE.findall(text)]
    return {"losses_seen": len(vals), "all_finite": all(math.isfinite(v) for v in vals),
            "first": vals[:3], "last": vals[-3:]}


def expected_files(work: Path, name: str) -> list:
    ms = work / "model_scripts"
    if name == "train3d":
        return [ms / "ddpm_3d_ldm/models" / RUN_ID / "vae3d_final.pt",
                ms / "ddpm_3d_ldm/models" / RUN_ID / "3d_ldm_diffusion_best.pt"]
    if name == "show3d":
        out = ms / "ddpm_3d_ldm/samples_inference/1594474"
        return [out / "sample_000.pt", out / "sample_000_mod0.nii.gz", out / "vae_recon_sanity_recon.png",
                out / "roundtrip_t0399_recon.png"]
    if name == "model2d":
        return [ms / "slice_cond_2d_ddpm/models" / RUN_ID / "2d_central_ddpm_flair_best.pt"]
    if name == "show2d":
        return [ms / "slice_cond_2d_ddpm/samples_inference/brain7_all_slices.png"]
    if name == "model25d":
        return [ms / "ddpm_25d_all_modalities/models" / RUN_ID / "2d_central_ddpm_flair_best.pt"]
    return []


def run_one(name: str, work: Path, out: Path, args) -> dict:
    module, desc = SCRIPTS[name]
    note = stage_checkpoints(work, name) if name.startswith("show") else ""
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([str(REPO)] + ([env["PYTHONPATH"]] if env.get("PYTHONPATH") else []))
    env["SLURM_JOB_ID"] = RUN_ID
    env["MRI_STUB_MLFLOW_OUT"] = str(out / f"{name}_mlflow.json")
    env.setdefault("PYTHONUNBUFFERED", "1")
    overlay = ["-m", "mri_image_generation_b200.overlay", "--path", str(STUBS)]
    if args.keep_vae:
        overlay.append("--keep-vae")
    if args.overlap_ddp:
        overlay.append("--overlap-ddp")
    if args.device_datasets:
        overlay.append("--device-datasets")
    if args.fused_adam:
        overlay.append("--fused-adam")
    if args.split_sampling and name.startswith("show"):
        overlay += ["--precision", "split"]
    overlay += ["-m", module]
    if args.nproc > 1 and name == "train3d":
        cmd = [sys.executable, "-m", "torch.distributed.run", "--standalone", "--local-addr", "127.0.0.1",
               "--nproc-per-node", str(args.nproc)] + overlay
    else:
        cmd = [sys.executable] + overlay
    log = out / f"{name}.log"
    t0 = time.time()
    with open(log, "w") as f:
        f.write(f"$ (cd {work}) {' '.join(cmd)}\n")
        f.flush()
        try:
            rc = subprocess.run(cmd, cwd=work, env=env, stdout=f, stderr=subprocess.STDOUT,
                                timeout=args.timeout).returncode
        except subprocess.TimeoutExpired:
            rc = -9
    text = log.read_text(errors="replace")
    files = expected_files(work, name)
    res = {"script": desc, "module": module, "returncode": rc, "seconds": round(time.time() - t0, 1),
           "overlay_bound": "[mri_b200.overlay]" in text, "staged": note,
           "files": {str(p.relative_to(work)): p.is_file() for p in files}, **check_log(text)}
    res["ok"] = bool(rc == 0 and res["overlay_bound"] and res["all_finite"] and all(res["files"].values())
                     and (res["losses_seen"] > 0 or name == "show2d"))
    mf = out / f"{name}_mlflow.json"
    if mf.is_file():
        rec = json.loads(mf.read_text())
        res["mlflow"] = {"params": len(rec["params"]), "metrics": {k: len(v) for k, v in rec["metrics"].items()},
                         "artifacts": len(rec["artifacts"]), "models": rec["models"]}
    tail = "\n".join(text.splitlines()[-12:])
    print(f"--- {name}: rc={rc} ok={res['ok']} {res['seconds']} s\n{tail}\n", flush=True)
    return res


def main() -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--reference", default=None, help="reference checkout (default: $MRI_REFERENCE_DIR, baseline/_ref)")
    ap.add_argument("--scripts", default="train3d,show3d,model2d,show2d,model25d")
    ap.add_argument("--work", default=None, help="scratch tree (default: a fresh temp dir)")
    ap.add_argument("--out", default=str(REPO / "gpurun_out" / "scripts"))
    ap.add_argument("--keep-vae", action="store_true", help="leave vae.py to the reference")
    ap.add_argument("--overlap-ddp", action="store_true")
    ap.add_argument("--fused-adam", action="store_true", help="torch.optim.Adam -> the one-launch Adam")
    ap.add_argument("--split-sampling", action="store_true",
                    help="the show_model scripts sample / decode in split precision (fp32-class parity)")
    ap.add_argument("--device-datasets", action="store_true",
                    help="dataset.py -> the device data path (normalise / resize / pad / crop kernels)")
    ap.add_argument("--nproc", type=int, default=1, help="train3d under torchrun with this many ranks")
    ap.add_argument("--timeout", type=int, default=1500, help="seconds per script")
    ap.add_argument("--cases", default="6,3,4", help="synthetic subjects in datasets/train,val,dataset")
    args = ap.parse_args()
    ref = find_reference(args.reference)
    if ref is None:
        print("no reference checkout found (--reference / MRI_REFERENCE_DIR / baseline/_ref)")
        return 2
    work = Path(args.work) if args.work else Path(tempfile.mkdtemp(prefix="mri_scripts_"))
    out = Path(args.out).resolve()
    out.mkdir(parents=True, exist_ok=True)
    stage_tree(ref, work)
    stage_datasets(work, *[int(v) for v in args.cases.split(",")])
    summary = {"reference": str(ref), "work": str(work), "keep_vae": args.keep_vae, "nproc": args.nproc,
               "device_datasets": args.device_datasets, "overlap_ddp": args.overlap_ddp,
               "fused_adam": args.fused_adam,
               "results": {}}
    for name in [s for s in args.scripts.split(",") if s]:
        summary["results"][name] = run_one(name, work, out, args)
        (out / "summary.json").write_text(json.dumps(summary, indent=1))
    ok = all(r["ok"] for r in summary["results"].values())
    print(json.dumps({k: {"ok": v["ok"], "rc": v["returncode"], "s": v["seconds"], "losses": v["losses_seen"]}
                      for k, v in summary["results"].items()}))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
