"""Copies the files of the UNMODIFIED reference that its CPU path needs into baseline/_ref/ (git-ignored, NOT
gpurun-ignored: it travels to the GPU box), so that `bench.py --impl reference` can time the reference itself on the
box's host cores (SURVEY 8(c) "GPU box"):

    code/loss.py, code/utils.py, code/LieAlgebra/*.py, code/sample_data/challenge_data/*

Runs in the build container only (needs /root/reference); __graft_entry__.build() calls it.  The files are never edited
and never enter the history; nothing in the product imports them.  bench.py loads loss.py with stub modules for
igl / openmesh / trimesh (utils.py:4,154-156 import them at module level; the hot path uses none of them)."""
import os
import shutil
import sys

SRC = os.environ.get("RRL_REFERENCE_CODE", "/root/reference/code")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def fetch(verbose=False):
    if not os.path.isfile(os.path.join(SRC, "loss.py")):
        return False
    os.makedirs(DST, exist_ok=True)
    for name in ("loss.py", "utils.py"):
        shutil.copyfile(os.path.join(SRC, name), os.path.join(DST, name))
    for sub in ("LieAlgebra", os.path.join("sample_data", "challenge_data")):
        s, d = os.path.join(SRC, sub), os.path.join(DST, sub)
        if os.path.isdir(s):
            os.makedirs(d, exist_ok=True)
            for f in os.listdir(s):
                if os.path.isfile(os.path.join(s, f)) and (f.endswith(".py") or f.endswith(".obj")):
                    shutil.copyfile(os.path.join(s, f), os.path.join(d, f))
    if verbose:
        print("reference files copied to", DST)
    return True


def load():
    """the reference's `loss` module from baseline/_ref (None when it was never fetched)"""
    import importlib
    import types
    import warnings
    if "_rrl_baseline_loss" in sys.modules:
        return sys.modules["_rrl_baseline_loss"]
    if not os.path.isfile(os.path.join(DST, "loss.py")):
        return None
    for name in ("igl", "openmesh", "trimesh"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    saved_path, saved_loss, saved_utils = list(sys.path), sys.modules.pop("loss", None), sys.modules.pop("utils", None)
    sys.path.insert(0, DST)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mod = importlib.import_module("loss")
    finally:
        sys.path[:] = saved_path
        sys.modules.pop("loss", None)
        sys.modules.pop("utils", None)
        if saved_loss is not None:
            sys.modules["loss"] = saved_loss
        if saved_utils is not None:
            sys.modules["utils"] = saved_utils
    sys.modules["_rrl_baseline_loss"] = mod
    return mod


if __name__ == "__main__":
    sys.exit(0 if fetch(verbose=True) else 1)
