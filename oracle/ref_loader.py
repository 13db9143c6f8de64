"""Import the UNMODIFIED reference (/root/reference) on top of the shims.

TEST INFRASTRUCTURE ONLY.  Used by oracle/gen_golden.py in the build container
to produce tests/golden/*; /root/reference does not exist on the GPU box: there only
bench.py's CPU-baseline / `--impl reference` legs import this module, on the copy that
oracle/make_ref.py staged under oracle/_ref/ (kind "reference-on-shims").
"""
from __future__ import annotations

import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIMS = os.path.join(_HERE, "shims")


def _default_root():
    """/root/reference in the build container; on the GPU box the byte-for-byte copy staged by oracle/make_ref.py."""
    for cand in (os.environ.get("CARLABEV_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "CarlaBEV")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _default_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "CarlaBEV"))


def load_reference():
    """Return the imported `CarlaBEV` package of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    for p in (_SHIMS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    # CarlaBEV.src.actors.actor imports CarlaBEV.src.gui.settings; the gui package
    # __init__ would pull the whole pygame GUI.  Register the package without
    # executing its __init__ so only `settings` (pure dataclasses) is loaded.
    if "CarlaBEV.src.gui" not in sys.modules:
        gui = types.ModuleType("CarlaBEV.src.gui")
        gui.__path__ = [os.path.join(REFERENCE_ROOT, "CarlaBEV", "src", "gui")]
        sys.modules["CarlaBEV.src.gui"] = gui
    import CarlaBEV  # noqa: WPS433

    return CarlaBEV
