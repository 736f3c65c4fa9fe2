"""Put the UNMODIFIED reference (rmazzier/OpenSetGaitRecognition_PCAA) where the GPU box can import it.

The reference is a flat directory of scripts (no setup.py / pyproject.toml: ``pip install --target baseline/_ref
/root/reference`` has nothing to build), so the install is a file copy of its ``*.py`` into ``baseline/_ref/``.
``baseline/_ref/`` is git-ignored (reference sources never enter this repository's history) but not gpurun-ignored:
it travels to the GPU box with the snapshot exactly like the built ``.so``.  ``__graft_entry__.build()`` runs this
whenever ``/root/reference`` is present (i.e. in the build container); on the GPU box the prebuilt copy is used.

    python baseline/install_ref.py            # -> baseline/_ref/{models,utils,constants,PCAA_ablation,...}.py + MANIFEST.json
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("PCAA_REFERENCE", "/root/reference")


def install(src: str = SRC, dest: str = DEST) -> dict:
    """Copy every top-level ``*.py`` of the reference; write a manifest of sha256 digests so a test can tell that the
    files on the box are the ones that were in /root/reference (nothing is patched)."""
    if not os.path.isdir(src):
        raise RuntimeError(f"reference tree not found at {src}")
    os.makedirs(dest, exist_ok=True)
    manifest = {}
    for name in sorted(os.listdir(src)):
        if not name.endswith(".py"):
            continue
        with open(os.path.join(src, name), "rb") as f:
            data = f.read()
        manifest[name] = hashlib.sha256(data).hexdigest()
        shutil.copyfile(os.path.join(src, name), os.path.join(dest, name))
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=1, sort_keys=True)
    return manifest


def verify(dest: str = DEST) -> bool:
    """True when every file listed in the manifest is present with the recorded digest."""
    mpath = os.path.join(dest, "MANIFEST.json")
    if not os.path.exists(mpath):
        return False
    with open(mpath) as f:
        files = json.load(f)["files"]
    for name, digest in files.items():
        p = os.path.join(dest, name)
        if not os.path.exists(p):
            return False
        with open(p, "rb") as f:
            if hashlib.sha256(f.read()).hexdigest() != digest:
                return False
    return bool(files)


if __name__ == "__main__":
    m = install()
    print(f"installed {len(m)} reference files into {DEST}")
    sys.exit(0 if verify() else 1)
