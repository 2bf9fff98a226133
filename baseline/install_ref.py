#!/usr/bin/env python
"""Install the UNMODIFIED reference (krasserm/perceiver-io, /root/reference) into the git-ignored
``baseline/_ref/`` so that it travels to the GPU box with the gpurun snapshot.

Why not ``pip install --target baseline/_ref /root/reference``: the reference's build backend is poetry-core
(pyproject.toml), which is not installed in this image and cannot be fetched (no network) — recorded in
DESIGN.md.  The package is pure Python, so the install pip would perform is a file copy of the ``perceiver``
package; this script does exactly that, byte for byte (a SHA-256 manifest is written next to it), and adds the
three-line ``fairscale`` stand-in the hot-path file needs at import time (modules.py:5 imports
``fairscale.nn.checkpoint_wrapper``, only reached with ``activation_checkpointing=True``).

Used by: ``bench.py --impl reference`` (times the reference's own ``CrossAttention.forward`` on the host cores),
the ``-m gpu`` test that runs ``patch()`` on a real reference model, and nothing in the product package.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("PCV_REFERENCE_ROOT", "/root/reference")

FAIRSCALE_STUB = '''"""Stand-in for fairscale.nn (not installed here): the reference only calls checkpoint_wrapper when
activation_checkpointing=True."""


def checkpoint_wrapper(module, offload_to_cpu=False):
    return module
'''


def install() -> str:
    pkg = os.path.join(SRC, "perceiver")
    if not os.path.isdir(pkg):
        raise RuntimeError(f"{SRC} does not hold the reference (perceiver/ missing)")
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    shutil.copytree(pkg, os.path.join(DEST, "perceiver"), ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    os.makedirs(os.path.join(DEST, "fairscale", "nn"))
    with open(os.path.join(DEST, "fairscale", "__init__.py"), "w") as f:
        f.write('"""stand-in, see nn/__init__.py"""\n')
    with open(os.path.join(DEST, "fairscale", "nn", "__init__.py"), "w") as f:
        f.write(FAIRSCALE_STUB)
    manifest = {}
    for dirpath, _, files in os.walk(os.path.join(DEST, "perceiver")):
        for name in sorted(files):
            path = os.path.join(dirpath, name)
            with open(path, "rb") as fh:
                manifest[os.path.relpath(path, DEST)] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=1, sort_keys=True)
    return DEST


def import_reference_core():
    """The reference's ``perceiver.model.core`` package from baseline/_ref (raises if it was never installed)."""
    if not os.path.isdir(os.path.join(DEST, "perceiver", "model", "core")):
        raise RuntimeError("baseline/_ref is empty: run `python baseline/install_ref.py` where /root/reference exists")
    if DEST not in sys.path:
        sys.path.insert(0, DEST)
    import perceiver.model.core as core

    return core


def available() -> bool:
    return os.path.isdir(os.path.join(DEST, "perceiver", "model", "core"))


if __name__ == "__main__":
    print("installed the reference into", install())
