#!/usr/bin/env python
"""Place the UNMODIFIED reference files of the hot path in the git-ignored `baseline/_ref/`.

    python tools/install_ref.py            # copies from $MIPNERF_REFERENCE or /root/reference

The reference ships no setup.py / pyproject.toml, so `pip install /root/reference` has nothing to build; what
`bench.py --impl reference` and `cpu_baseline` need is the four files `MipNerf.forward` imports
(models/mip_nerf.py:1-4, models/mip.py:1-4): models/{__init__,mip,mip_nerf}.py and datasets/{__init__,datasets}.py.
They are copied byte for byte (sha256 recorded in baseline/_ref/MANIFEST.json) and are never tracked by git
(`.gitignore: baseline/_ref/`); gpurun ships the directory to the GPU box, where /root/reference does not exist.
(models/nerf_system.py needs pytorch-lightning, which this image lacks; it is not on the per-ray path.)
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
DEST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["models/__init__.py", "models/mip.py", "models/mip_nerf.py", "datasets/__init__.py", "datasets/datasets.py"]
EMPTY = []


def install(src: str = None, quiet: bool = False) -> bool:
    src = src or os.environ.get("MIPNERF_REFERENCE", "/root/reference")
    if not os.path.isdir(src):
        if not quiet:
            print(f"install_ref: {src} not present; keeping whatever is in {DEST}", file=sys.stderr)
        return os.path.exists(os.path.join(DEST, "MANIFEST.json"))
    manifest = {"source": src, "files": {}}
    for rel in FILES:
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), dst)
        with open(dst, "rb") as f:
            manifest["files"][rel] = hashlib.sha256(f.read()).hexdigest()
    for rel in EMPTY:
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        open(dst, "w").close()
        manifest["files"][rel] = "empty package marker (the reference's imports pytorch-lightning)"
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    if not quiet:
        print(f"install_ref: {len(FILES)} files -> {DEST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
