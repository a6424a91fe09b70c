"""Import the UNMODIFIED reference (hjxwhy/mipnerf_pl) from the git-ignored `baseline/_ref/`.

`tools/install_ref.py` puts the reference's own files there (verbatim copies, sha256 in MANIFEST.json).  This loader
is used only by `bench.py`'s CPU arm (`--impl reference`, `cpu_baseline`) and by `tests/test_reference_arm.py`;
nothing under `mipnerf_pl_b200/` imports it.
"""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "models", "mip_nerf.py"))


def load():
    """-> (MipNerf class, Rays namedtuple, models.mip module) of the reference; raises ImportError if not installed."""
    if not available():
        raise ImportError(f"{REF_DIR} is empty: run `python tools/install_ref.py` where /root/reference exists")
    for name in ("models", "datasets"):  # the reference imports itself through these top-level names
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__file__", "").startswith(REF_DIR):
            raise ImportError(f"another top-level package named {name!r} is already imported ({mod.__file__})")
    sys.path.insert(0, REF_DIR)
    try:
        mip_nerf = importlib.import_module("models.mip_nerf")
        mip = importlib.import_module("models.mip")
        ds = importlib.import_module("datasets.datasets")
    finally:
        sys.path.remove(REF_DIR)
    return mip_nerf.MipNerf, ds.Rays, mip
