#!/usr/bin/env python
"""Install the UNMODIFIED reference into baseline/_ref (git-ignored, NOT gpurun-ignored: it travels to the GPU box).

The contract's recipe is
    python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference
It fails in this image: the reference's build backend is hatchling (pyproject.toml [build-system]) and neither the
interpreter nor /opt/wheelhouse has it (``ModuleNotFoundError: No module named 'hatchling'``, also with --no-deps from a
/tmp copy).  The package is pure Python and its wheel would contain exactly ``packages = ["edge_diffusion_tts"]``
(pyproject.toml [tool.hatch.build.targets.wheel]), so this script places the same files: a verbatim copy of the package
directory (plus the top-level long-form script, which ``bench.py`` never runs), nothing edited.  The run-time shims the
reference needs in this image live in the CALLER (bench.py ``_import_reference``): ``matplotlib`` stubbed in sys.modules
before import (utils/visualization.py:8), cwd moved to a temp dir while ``CFG()`` runs (config.py:165-166 creates
./data and ./run_edge_diffusion), ``SemanticEncoder`` never built (encoder.py:35 downloads HuBERT).  SURVEY.md F13.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("EDTTS_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def _tree_sha(path: str) -> str:
    h = hashlib.sha256()
    for d, _, fs in sorted(os.walk(path)):
        for f in sorted(fs):
            if f.endswith(".py"):
                p = os.path.join(d, f)
                h.update(os.path.relpath(p, path).encode())
                h.update(open(p, "rb").read())
    return h.hexdigest()[:16]


def install(verbose: bool = True) -> bool:
    """Returns True when baseline/_ref holds the reference package afterwards."""
    pkg_src = os.path.join(SRC, "edge_diffusion_tts")
    pkg_dst = os.path.join(DST, "edge_diffusion_tts")
    if not os.path.isdir(pkg_src):
        return os.path.isdir(pkg_dst)                       # GPU box: use what travelled
    if os.path.isdir(pkg_dst) and _tree_sha(pkg_dst) == _tree_sha(pkg_src):
        return True
    os.makedirs(DST, exist_ok=True)
    outcome = "pip: not attempted"
    if os.environ.get("EDTTS_TRY_PIP", "0") == "1":
        tmp = "/tmp/edtts_refcopy"
        shutil.rmtree(tmp, ignore_errors=True)
        shutil.copytree(SRC, tmp)
        r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--find-links",
                            "/opt/wheelhouse", "--no-deps", "--target", DST, tmp], capture_output=True, text=True)
        outcome = "pip: ok" if r.returncode == 0 else "pip: failed (" + r.stderr.strip().splitlines()[-1][:120] + ")"
    if not os.path.isdir(pkg_dst) or _tree_sha(pkg_dst) != _tree_sha(pkg_src):
        shutil.rmtree(pkg_dst, ignore_errors=True)
        shutil.copytree(pkg_src, pkg_dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        for extra in ("inference_pipeline.py", "pyproject.toml"):
            if os.path.exists(os.path.join(SRC, extra)):
                shutil.copy2(os.path.join(SRC, extra), os.path.join(DST, extra))
        outcome += "; verbatim copy of the package directory"
    json.dump({"source": SRC, "sha16_py_tree": _tree_sha(pkg_dst), "how": outcome},
              open(os.path.join(DST, "INSTALL.json"), "w"), indent=1)
    if verbose:
        print(f"[install_reference] {pkg_dst}: {outcome}", flush=True)
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
