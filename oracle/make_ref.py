"""Materialise ``oracle/_ref/``: the UNMODIFIED reference modules, so they can travel to the GPU box.

TEST / BENCH INFRASTRUCTURE.  ``/root/reference`` exists only in the build container; ``oracle/_ref/`` is
git-ignored (never part of the history) but NOT gpurun-ignored, so it ships with the snapshot like a built ``.so``.
``__graft_entry__.build()`` calls :func:`materialise` whenever ``/root/reference`` is present.  The files are copied
byte for byte (``model/*.py``; the reference is pure Python, there is nothing to compile) and a manifest with their
SHA-256 digests is written next to them; ``oracle/refshim.py`` imports them from ``/root/reference`` when that tree
exists and from ``oracle/_ref`` otherwise, and verifies the digests before importing the copy.

    python -m oracle.make_ref
"""
import hashlib
import json
import os
import shutil

SRC_ROOT = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST_ROOT = os.path.join(HERE, "_ref")
PACKAGES = ("model", "data")   # model/: the hot path, E2E / Decoder, ModelBase.load_model; data/: imported by model/fsrnn.py


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def materialise(src_root=SRC_ROOT, dst_root=DST_ROOT):
    """Copy the reference's Python packages into ``dst_root``; returns the manifest dict (or None if no source)."""
    if not os.path.isdir(os.path.join(src_root, "model")):
        return None
    manifest = {"source": src_root, "files": {}}
    for pkg in PACKAGES:
        sdir, ddir = os.path.join(src_root, pkg), os.path.join(dst_root, pkg)
        os.makedirs(ddir, exist_ok=True)
        for name in sorted(os.listdir(sdir)):
            if not name.endswith(".py"):
                continue
            s, d = os.path.join(sdir, name), os.path.join(ddir, name)
            if not (os.path.exists(d) and _sha(d) == _sha(s)):
                shutil.copyfile(s, d)
            manifest["files"][pkg + "/" + name] = _sha(d)
    with open(os.path.join(dst_root, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    return manifest


def verify(dst_root=DST_ROOT):
    """True when ``dst_root`` holds exactly the files of its manifest, unmodified."""
    mpath = os.path.join(dst_root, "MANIFEST.json")
    if not os.path.exists(mpath):
        return False
    with open(mpath) as f:
        manifest = json.load(f)
    return all(os.path.exists(os.path.join(dst_root, rel)) and _sha(os.path.join(dst_root, rel)) == digest
               for rel, digest in manifest["files"].items()) and len(manifest["files"]) > 0


if __name__ == "__main__":
    m = materialise()
    print("oracle/_ref: %s" % ("%d files" % len(m["files"]) if m else "no reference tree at %s" % SRC_ROOT))
