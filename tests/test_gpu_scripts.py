"""The reference's own scripts, executed unmodified over the drop-in on a B200
(tools/run_reference_scripts.py).  Needs a reference checkout next to a GPU: point
MRI_REFERENCE_DIR at one (the build container has the sources but no GPU, the GPU boxes have no
/root/reference; the runs recorded under profiles/ were made with the checkout staged in the
git-ignored baseline/_ref).  MRI_SCRIPTS selects the scripts (default: the 2D training loop, the
fastest one; "all" = train3d,show3d,model2d,show2d,model25d, about ten minutes)."""
import json
import os
import subprocess
import sys

import pytest

from helpers import ROOT

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_reference_scripts_run_over_the_drop_in(tmp_path):
    import run_reference_scripts as H
    ref = H.find_reference(None)
    if ref is None or str(ref) == "/root/reference" and not os.access("/root/reference", os.R_OK):
        pytest.skip("no reference checkout on this box (MRI_REFERENCE_DIR / baseline/_ref)")
    scripts = os.environ.get("MRI_SCRIPTS", "model2d")
    if scripts == "all":
        scripts = ",".join(H.SCRIPTS)
    out = tmp_path / "out"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_reference_scripts.py"), "--reference",
                        str(ref), "--scripts", scripts, "--work", str(tmp_path / "work"), "--out", str(out)],
                       capture_output=True, text=True, timeout=3600)
    print(r.stdout[-4000:], r.stderr[-2000:])
    summary = json.loads((out / "summary.json").read_text())
    for name, res in summary["results"].items():
        assert res["ok"], (name, res)
    assert r.returncode == 0
