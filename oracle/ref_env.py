"""Loader of the UNMODIFIED reference (test infrastructure only, like everything under oracle/).

Finds the reference tree -- /root/reference in the authoring container, else the copy staged under oracle/_ref/ by
oracle/fetch_reference.py (which travels to the GPU box) -- and imports its modules without leaving them on
sys.path / in sys.modules:

  reference_root()                      the tree, or None
  load_model_module(ours=True)          the reference's models/pointnet2_sem_seg.py on top of THIS repo's
                                        models/pointnet2_utils.py (the drop-in configuration), or (ours=False) on
                                        top of its own models/pointnet2_utils.py (the pure reference)
  load_localfunctions()                 /root/reference/localfunctions.py (modelTraining, modelTesting, add_vote)
  load_script(name)                     sem_seg_training / sem_seg_testing (dataset classes)

The reference's scripts import IO / plotting packages this image does not have (laspy, open3d, h5py, matplotlib,
pytz, torchviz); none of them is used by the functions exercised here, so empty stand-in modules satisfy the import
statements.
"""
import datetime
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
_CANDIDATES = (os.environ.get("PN2_REFERENCE_ROOT", "/root/reference"), os.path.join(HERE, "_ref"))


def reference_root():
    for c in _CANDIDATES:
        if os.path.isfile(os.path.join(c, "models", "pointnet2_sem_seg.py")) and os.path.isfile(
                os.path.join(c, "models", "pointnet2_utils.py")):
            return c
    return None


def reference_kind():
    r = reference_root()
    return None if r is None else ("mounted" if r == _CANDIDATES[0] else "staged copy (oracle/_ref)")


def install_stand_ins():
    for name in ("laspy", "open3d", "h5py", "matplotlib", "matplotlib.pyplot", "pytz", "torchviz"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    mpl = sys.modules.get("matplotlib")
    if mpl is not None and not hasattr(mpl, "pyplot"):
        mpl.pyplot = sys.modules["matplotlib.pyplot"]
    if not hasattr(sys.modules["pytz"], "timezone"):          # localfunctions.py:102 builds a module-level tz with it
        sys.modules["pytz"].timezone = lambda name: datetime.timezone(datetime.timedelta(hours=8))


_OWNED = ("models", "pointnet2_sem_seg", "pointnet2_utils", "localfunctions", "provider", "sem_seg_training",
          "sem_seg_testing", "geofunction")


class _Scoped:
    """sys.path / sys.modules as they were, minus anything the import inside the block added under the reference's
    module names (the imported module objects stay alive through the references the caller keeps)."""

    def __init__(self, paths):
        self.paths = paths

    def __enter__(self):
        self.saved_path = list(sys.path)
        self.saved_mods = {k: v for k, v in sys.modules.items() if k.split(".")[0] in _OWNED}
        for k in self.saved_mods:
            del sys.modules[k]
        repo_paths = {os.path.abspath(ROOT), os.path.abspath(os.path.join(ROOT, "tests"))}
        rest = [p for p in sys.path if os.path.abspath(p or ".") not in repo_paths or ROOT in self.paths]
        sys.path[:] = list(self.paths) + [p for p in rest if p not in self.paths]
        return self

    def __exit__(self, *exc):
        for k in [k for k in sys.modules if k.split(".")[0] in _OWNED]:
            del sys.modules[k]
        sys.modules.update(self.saved_mods)
        sys.path[:] = self.saved_path
        return False


def load_model_module(ours=True):
    """(pointnet2_sem_seg module, models.pointnet2_utils module).  ours=True: the reference's model file, unchanged,
    importing `models.pointnet2_utils` from THIS repo (sem_seg_training.py:542 / pointnet2_sem_seg.py:3)."""
    ref = reference_root()
    if ref is None:
        raise FileNotFoundError("no reference tree (/root/reference or oracle/_ref); run oracle/fetch_reference.py where it is mounted")
    paths = [ROOT, os.path.join(ref, "models")] if ours else [ref, os.path.join(ref, "models")]
    with _Scoped(paths):
        mod = importlib.import_module("pointnet2_sem_seg")
        utils = importlib.import_module("models.pointnet2_utils")
    want = ROOT if ours else ref
    assert mod.__file__.startswith(ref), mod.__file__
    assert os.path.abspath(utils.__file__).startswith(os.path.abspath(want)), (utils.__file__, want)
    return mod, utils


def load_localfunctions():
    ref = reference_root()
    if ref is None:
        raise FileNotFoundError("no reference tree")
    install_stand_ins()
    with _Scoped([ref]):
        lf = importlib.import_module("localfunctions")
    assert lf.__file__.startswith(ref), lf.__file__
    return lf


def load_script(name):
    """sem_seg_training / sem_seg_testing as modules (their argparse only runs under __main__ / main())."""
    ref = reference_root()
    if ref is None:
        raise FileNotFoundError("no reference tree")
    install_stand_ins()
    argv, sys.argv = sys.argv, sys.argv[:1]
    try:
        with _Scoped([ref, os.path.join(ref, "models")]):
            try:
                importlib.import_module("geofunction")
            except Exception:                                   # needs a real open3d: nothing exercised here calls it
                g = sys.modules["geofunction"] = types.ModuleType("geofunction")
                g.cal_geofeature = None
            mod = importlib.import_module(name)
    finally:
        sys.argv = argv
    assert mod.__file__.startswith(ref), mod.__file__
    return mod
