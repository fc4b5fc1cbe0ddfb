#!/usr/bin/env python
"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference package, made importable on the GPU box.

TEST / BASELINE INFRASTRUCTURE -- never imported by the product (mogptk_b200/).

GAMES-UChile/mogptk is pure Python (31 files, no native build), so "building" it is copying the
package directory byte for byte from the read-only checkout at /root/reference into
``oracle/_ref/mogptk``.  ``oracle/_ref/`` is listed in .gitignore (the reference's sources never
enter this repository's history) but NOT in .gpurunignore, so the copy travels to the GPU box with
the snapshot, where /root/reference does not exist.  A MANIFEST with the sha256 of every file is
written beside it so that tests can prove the copy is unmodified.

The reference imports matplotlib / IPython at module import and neither is installed in this
image; ``oracle/ref_loader.py`` pre-seeds ``sys.modules`` with inert stand-ins for them (plotting is
never exercised) -- the reference's own files are not touched.

Usage:  python oracle/build_ref.py            (also run by __graft_entry__.build())
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/mogptk"
DST_ROOT = os.path.join(HERE, "_ref")
DST = os.path.join(DST_ROOT, "mogptk")


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def manifest(root):
    out = {}
    for d, _, files in os.walk(root):
        if "__pycache__" in d:
            continue
        for f in sorted(files):
            if f.endswith((".pyc", ".pyo")):
                continue
            p = os.path.join(d, f)
            out[os.path.relpath(p, root)] = _sha(p)
    return out


def build(verbose=False):
    """Copy the reference package; returns True if oracle/_ref is usable afterwards."""
    if not os.path.isdir(SRC):
        ok = os.path.isdir(DST)
        if verbose:
            print("build_ref: %s not present (GPU box?) -- %s" % (SRC, "using the shipped copy" if ok else "no copy"))
        return ok
    want = manifest(SRC)
    mpath = os.path.join(DST_ROOT, "MANIFEST.json")
    if os.path.isdir(DST) and os.path.exists(mpath):
        try:
            with open(mpath) as f:
                have = json.load(f)
            if have.get("files") == want and manifest(DST) == want:
                if verbose:
                    print("build_ref: oracle/_ref is up to date (%d files)" % len(want))
                return True
        except Exception:
            pass
    if os.path.isdir(DST_ROOT):
        shutil.rmtree(DST_ROOT)
    os.makedirs(DST_ROOT)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for extra in ("LICENSE", "setup.py"):
        p = os.path.join(os.path.dirname(SRC), extra)
        if os.path.exists(p):
            shutil.copy(p, os.path.join(DST_ROOT, extra + ".reference"))
    with open(mpath, "w") as f:
        json.dump({"source": SRC, "files": want}, f, indent=1, sort_keys=True)
    if verbose:
        print("build_ref: copied %d files into %s" % (len(want), DST))
    return True


if __name__ == "__main__":
    sys.exit(0 if build(verbose=True) else 1)
