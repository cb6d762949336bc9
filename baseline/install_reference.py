#!/usr/bin/env python
"""Install the UNMODIFIED reference (Datou/Datou-gomoku-muzero) into baseline/_ref/ for bench.py's
`--impl reference` arm and `cpu_baseline` leg.

    python baseline/install_reference.py [/root/reference]

The reference is a flat directory of Python modules with no setup.py / pyproject.toml, so
`pip install --target baseline/_ref /root/reference` has nothing to build: "installing" it is
placing its modules on an import path, which is all this script does (byte-for-byte copies of the
*.py files, plus MANIFEST.json with their SHA-256 so the GPU-box run can show they are unmodified).
baseline/_ref/ is git-ignored (no reference source enters the history) but NOT gpurun-ignored, so it
travels to the GPU box like the built .so files.  Run by __graft_entry__.build() whenever the
reference tree is present; on the GPU box (no /root/reference) the copy that travelled is used.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")


def install(src="/root/reference"):
    if not os.path.isdir(src):
        return None
    os.makedirs(DEST, exist_ok=True)
    manifest = {}
    for name in sorted(os.listdir(src)):
        path = os.path.join(src, name)
        if name.endswith(".py") and os.path.isfile(path):
            shutil.copyfile(path, os.path.join(DEST, name))
            manifest[name] = hashlib.sha256(open(path, "rb").read()).hexdigest()
    json.dump({"source": src, "files": manifest}, open(os.path.join(DEST, "MANIFEST.json"), "w"), indent=1)
    return DEST


def verify():
    """True if baseline/_ref holds exactly the files its manifest lists, unmodified."""
    try:
        m = json.load(open(os.path.join(DEST, "MANIFEST.json")))["files"]
        return bool(m) and all(hashlib.sha256(open(os.path.join(DEST, n), "rb").read()).hexdigest() == h for n, h in m.items())
    except Exception:
        return False


if __name__ == "__main__":
    d = install(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print("installed" if d else "reference tree not found", d or "", "verified" if verify() else "")
