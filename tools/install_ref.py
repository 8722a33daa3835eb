#!/usr/bin/env python
"""Stage the UNMODIFIED reference package for the bench's `cpu_baseline_python` leg.

    python tools/install_ref.py            # /root/reference/gym_space -> baseline/_ref/gym_space

`baseline/_ref/` is git-ignored (reference sources never enter the history) but NOT gpurun-ignored, so the copy
travels with the tree to the GPU box, where /root/reference does not exist.  bench.py imports it behind the `gym`
shim of oracle/gym (the image has neither gym nor gymnasium) and steps it one env per worker process
(last bullet).  Nothing in the product package or the tests' GPU path reads it.
A byte-for-byte copy: the sha256 of every file is written to baseline/_ref/MANIFEST.json.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/gym_space"
DST = os.path.join(ROOT, "baseline", "_ref", "gym_space")


def install(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print(f"install_ref: {SRC} not present (GPU box): using the staged copy" if os.path.isdir(DST)
                  else f"install_ref: neither {SRC} nor {DST} exists")
        return os.path.isdir(DST)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "assets"))
    manifest = {}
    for base, _, files in os.walk(DST):
        for f in sorted(files):
            p = os.path.join(base, f)
            manifest[os.path.relpath(p, DST)] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    with open(os.path.join(os.path.dirname(DST), "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"install_ref: {len(manifest)} files -> {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
