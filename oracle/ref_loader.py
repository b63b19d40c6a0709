"""Import the UNMODIFIED reference staged under ``oracle/_ref/reference`` -- TEST INFRASTRUCTURE ONLY.

``oracle/make_ref.sh`` (run by ``make`` / ``__graft_entry__.build()`` whenever ``/root/reference`` exists) stages the
reference's own Python files there so that they travel to the GPU box.  Only ``tests/``, ``bench.py``'s
``--impl reference`` / ``cpu_baseline`` legs and ``__graft_entry__.smoke()`` may call this; the product package never
does (``tests/test_abi.py::test_product_never_imports_oracle``).
"""
from __future__ import annotations

import importlib
import os
import sys

REF_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "EigenTrajectory", "descriptor.py"))


def load(*names):
    """Import reference top-level packages (``"EigenTrajectory"``, ``"utils"``, ``"baseline"``) and return them.

    The staged tree goes to the front of ``sys.path`` (the reference uses absolute imports of its own top-level
    packages); bytecode is not written into it."""
    if not available():
        raise ImportError(f"{REF_ROOT} is not staged: run oracle/make_ref.sh where /root/reference exists")
    sys.dont_write_bytecode = True
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    mods = tuple(importlib.import_module(n) for n in names)
    return mods[0] if len(mods) == 1 else mods
