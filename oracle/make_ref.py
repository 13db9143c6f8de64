"""Stage the UNMODIFIED reference package next to the shims so that it can be timed on the GPU box.

TEST / BASELINE INFRASTRUCTURE ONLY.  `python oracle/make_ref.py` copies /root/reference/CarlaBEV (3.3 MB of
Python sources and assets, byte for byte, nothing edited) into oracle/_ref/CarlaBEV.  oracle/_ref/ is git-ignored
(reference sources never enter this repository's history) but not gpurun-ignored, so the copy travels with the
snapshot to the box, where /root/reference does not exist.  `bench.py --impl reference` and `cpu_baseline` import it
from there on top of oracle/shims (pygame-lite / gymnasium-lite: neither library is installable in this image) and
label the result `kind: "reference-on-shims"`; without the staged copy they time the oracle port (`kind: "port"`).
__graft_entry__.build() runs this recipe whenever /root/reference is present."""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("CARLABEV_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")


def tree_digest(root):
    h = hashlib.sha256()
    n = 0
    for d, dirs, files in os.walk(root):
        dirs.sort()
        dirs[:] = [x for x in dirs if x != "__pycache__"]
        for f in sorted(files):
            if f.endswith(".pyc"):
                continue
            p = os.path.join(d, f)
            h.update(os.path.relpath(p, root).encode())
            with open(p, "rb") as fh:
                h.update(fh.read())
            n += 1
    return h.hexdigest(), n


def stage(force=False):
    src = os.path.join(SRC, "CarlaBEV")
    if not os.path.isdir(src):
        return None
    dst = os.path.join(DST, "CarlaBEV")
    want, n = tree_digest(src)
    manifest = os.path.join(DST, "MANIFEST.json")
    if not force and os.path.exists(manifest):
        try:
            with open(manifest) as f:
                if json.load(f).get("sha256") == want and os.path.isdir(dst):
                    return dst
        except Exception:  # noqa: BLE001
            pass
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(DST, exist_ok=True)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    got, _ = tree_digest(dst)
    assert got == want, "staged copy differs from the reference"
    with open(manifest, "w") as f:
        json.dump({"source": src, "files": n, "sha256": want,
                   "note": "byte-for-byte copy of the reference package; git-ignored, travels to the GPU box"}, f, indent=1)
    return dst


if __name__ == "__main__":
    out = stage(force="-f" in sys.argv)
    print(out if out else f"{SRC}/CarlaBEV not found: nothing staged")
