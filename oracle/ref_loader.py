"""TEST INFRASTRUCTURE — imports the staged, unmodified reference (oracle/_ref/, see make_ref.py).

Only tests/, __graft_entry__ and bench.py's CPU legs (--impl reference, cpu_baseline) may use this.  One tree per
process: the current tree ("current": dual-source semantics) and the snapshot ("snapshot": vanilla semantics, the
only one in which guidance / uncond gnet / SR run) define colliding module names.
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "_ref")
_loaded = {}


def available():
    return os.path.isfile(os.path.join(ROOT, "MANIFEST.json"))


def load(tree="snapshot"):
    """Returns a namespace with the reference modules: models, generate_images, encoders, utils."""
    if _loaded:
        if tree not in _loaded:
            raise RuntimeError(f"reference tree {next(iter(_loaded))!r} already imported in this process")
        return _loaded[tree]
    if not available():
        raise FileNotFoundError("oracle/_ref is not staged: run `python oracle/make_ref.py` in the build container")
    for m in ("kornia", "litdata"):          # absent in this image and never touched on the denoising path
        sys.modules.setdefault(m, types.ModuleType(m))
    cur = os.path.join(ROOT, "current")
    sys.path[:0] = [os.path.join(ROOT, "snapshot"), cur] if tree == "snapshot" else [cur]
    ns = types.SimpleNamespace(tree=tree)
    ns.models = importlib.import_module("training.models")
    ns.encoders = importlib.import_module("training.encoders")
    ns.utils = importlib.import_module("training.utils")
    ns.generate_images = importlib.import_module("generate_images")
    ns.persistence = importlib.import_module("torch_utils.persistence")
    _loaded[tree] = ns
    return ns
