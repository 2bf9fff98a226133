"""Make the real reference importable in the AUTHORING container — TEST INFRASTRUCTURE ONLY.

/root/reference is pure Python but its hot-path file imports ``fairscale.nn.checkpoint_wrapper``
(modules.py:5), which is not installed here; a three-line stand-in module is registered instead (the
wrapper is only reached with ``activation_checkpointing=True``, which no golden case uses).
/root/reference does not exist on the GPU box: nothing under tests marked ``gpu``, ``smoke()`` or
``bench.py`` may call this; it is used by ``gen_golden.py`` and by CPU-side oracle pinning tests,
which skip when the directory is absent.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "perceiver", "model", "core"))


def import_reference_core():
    """Returns the reference's ``perceiver.model.core`` package."""
    if not reference_available():
        raise RuntimeError(f"{REFERENCE_ROOT} is not present in this environment")
    if "fairscale" not in sys.modules:
        fairscale = types.ModuleType("fairscale")
        fairscale_nn = types.ModuleType("fairscale.nn")
        fairscale_nn.checkpoint_wrapper = lambda module, offload_to_cpu=False: module
        fairscale.nn = fairscale_nn
        sys.modules["fairscale"] = fairscale
        sys.modules["fairscale.nn"] = fairscale_nn
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import perceiver.model.core as core  # noqa: E402

    return core
