"""overlay.install(): the reference's scripts, imported unmodified, bind to the drop-in classes
(build container only: needs a reference checkout; the GPU box runs the scripts for real through
tests/test_gpu_scripts.py when one was staged)."""
import os
import subprocess
import sys
import textwrap

import pytest

from helpers import ROOT

REF = next((d for d in (os.environ.get("MRI_REFERENCE_DIR"), "/root/reference", os.path.join(ROOT, "baseline", "_ref"))
            if d and os.path.isfile(os.path.join(d, "model_scripts", "ddpm_3d_ldm", "train.py"))), None)
pytestmark = pytest.mark.skipif(REF is None, reason="no reference checkout here")

PROBE = textwrap.dedent("""
    import sys
    sys.path[:0] = [{stubs!r}, {ref!r}, {root!r}]
    from mri_image_generation_b200 import overlay
    names = overlay.install({kw})
    import model_scripts.ddpm_3d_ldm.show_model as s3      # no side effects at import (main() is guarded)
    import model_scripts.slice_cond_2d_ddpm.show_model as s2
    import model_scripts.ddpm_25d_all_modalities.generate_pseudo3d_volume as g25
    ours = "mri_image_generation_b200.model_scripts."
    for cls in (s3.UNet3DModel, s3.UNet3DModelWithAttention, s3.GaussianDiffusionLatent3D, s2.UNet,
                s2.GaussianDiffusion, g25.UNet, g25.GaussianDiffusion):
        assert cls.__module__.startswith(ours), cls
    print("VAE", s3.VAE3D.__module__)
    print("DATASET", s3.BraTS3DVolumeDataset.__module__, g25.BraTSSliceDataset.__module__)
    assert s3.__file__.startswith({ref!r}), s3.__file__      # the script itself is the reference's file
    print("BOUND", len(names))
""")


def _probe(kw):
    code = PROBE.format(stubs=os.path.join(ROOT, "tests", "script_stubs"), ref=REF, root=ROOT, kw=kw)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_scripts_bind_to_the_drop_in_and_keep_their_own_dataset():
    out = _probe("")
    assert "VAE mri_image_generation_b200.model_scripts.ddpm_3d_ldm.vae" in out
    assert "DATASET model_scripts.ddpm_3d_ldm.dataset model_scripts.ddpm_25d_all_modalities.dataset" in out
    assert "BOUND 8" in out


def test_keep_vae_and_device_datasets_options():
    out = _probe("vae=False, datasets=True")
    assert "VAE model_scripts.ddpm_3d_ldm.vae" in out
    assert "DATASET mri_image_generation_b200.model_scripts.ddpm_3d_ldm.dataset" in out


def test_install_after_the_script_was_imported_is_refused():
    code = textwrap.dedent(f"""
        import sys
        sys.path[:0] = [{os.path.join(ROOT, 'tests', 'script_stubs')!r}, {REF!r}, {ROOT!r}]
        import model_scripts.slice_cond_2d_ddpm.unet
        from mri_image_generation_b200 import overlay
        try:
            overlay.install(["slice_cond_2d_ddpm"])
        except RuntimeError as e:
            print("REFUSED", e)
    """)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "REFUSED" in r.stdout, r.stdout + r.stderr
