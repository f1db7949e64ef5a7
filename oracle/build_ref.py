"""oracle/_ref: the UNMODIFIED reference, staged for the GPU box.  TEST INFRASTRUCTURE ONLY.

The reference (cherise215/MaxStyle) is pure Python: "building" it means placing its own files where the
parity tests and `bench.py --impl reference` can import them on a box that has no /root/reference.
`build()` copies -- byte for byte, nothing edited -- the files the hot path's callers need:

    src/**/*.py                       the reference package (MaxStyle layer, MyDecoder / UnetDecoder, the solver)
    notebooks/data/{image,label}.npy  the 20 x 192 x 192 ACDC batch the reference notebook uses
    notebooks/model/*.pth             the trained FCN_16 weights that go with it

into oracle/_ref/ (git-ignored: the sources never enter this repository's history; NOT gpurun-ignored,
so the directory travels with the snapshot).  A MANIFEST with sha256 sums is written beside them so a
test can prove the staged files are the reference's.  Run by `__graft_entry__.build()` whenever
/root/reference is present; on the GPU box the prebuilt directory is used as it came.

Only tests/, __graft_entry__.smoke() and bench.py's reference legs may import from oracle/ (tests/test_no_fallback.py).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DST = os.path.join(HERE, "_ref")
FIXTURES = ["notebooks/data/image.npy", "notebooks/data/label.npy",
            "notebooks/model/image_decoder.pth", "notebooks/model/image_encoder.pth", "notebooks/model/segmentation_decoder.pth"]


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def wanted_files(src: str = REF_SRC):
    out = []
    for root, _dirs, files in os.walk(os.path.join(src, "src")):
        for fn in files:
            if fn.endswith(".py"):
                out.append(os.path.relpath(os.path.join(root, fn), src))
    out += [f for f in FIXTURES if os.path.exists(os.path.join(src, f))]
    return sorted(out)


def build(src: str = REF_SRC, dst: str = REF_DST) -> str | None:
    """Stage the reference under oracle/_ref.  Returns the directory, or None when the reference is not mounted."""
    if not os.path.isdir(os.path.join(src, "src", "advanced")):
        return dst if os.path.exists(os.path.join(dst, "MANIFEST.json")) else None
    files = wanted_files(src)
    manifest = {}
    for rel in files:
        s, d = os.path.join(src, rel), os.path.join(dst, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        digest = _sha(s)
        if not (os.path.exists(d) and _sha(d) == digest):
            shutil.copyfile(s, d)
        manifest[rel] = digest
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=1, sort_keys=True)
    return dst


def available(dst: str = REF_DST) -> bool:
    return os.path.exists(os.path.join(dst, "MANIFEST.json")) and os.path.exists(os.path.join(dst, "src", "advanced", "maxstyle.py"))


if __name__ == "__main__":
    print(build())
