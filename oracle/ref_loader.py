"""TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference implementation.

Imports /root/reference/code/loss.py (read-only mount, exists only in the build
container, never on the GPU box) with stub modules for the three mesh libraries
its `utils.py` drags in at import time (utils.py:4,154-156: openmesh, trimesh, igl),
none of which the hot path uses.  Used by oracle/make_golden.py to mint the golden
vectors under tests/golden/ and by tests that validate the restatements when the
reference is present.  Nothing in the product path imports this file.
"""
import importlib
import os
import sys
import types
import warnings

REFERENCE_CODE = os.environ.get("RRL_REFERENCE_CODE", "/root/reference/code")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_CODE, "loss.py"))


def load():
    """Return the reference `loss` module (cached)."""
    if "_rrl_reference_loss" in sys.modules:
        return sys.modules["_rrl_reference_loss"]
    if not available():
        raise RuntimeError("reference sources not present at " + REFERENCE_CODE)
    for name in ("igl", "openmesh", "trimesh"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    saved_path = list(sys.path)
    saved_loss = sys.modules.pop("loss", None)
    sys.path.insert(0, REFERENCE_CODE)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mod = importlib.import_module("loss")
    finally:
        sys.path[:] = saved_path
        sys.modules.pop("loss", None)
        if saved_loss is not None:
            sys.modules["loss"] = saved_loss
    sys.modules["_rrl_reference_loss"] = mod
    return mod
