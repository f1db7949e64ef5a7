"""oracle/_ref (the unmodified reference staged for the GPU box) and the import shims around it -- CPU checks."""
import hashlib
import json
import os

import pytest
import torch

from oracle import build_ref, ref_shims

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def _staged():
    if not build_ref.available():
        if os.path.isdir("/root/reference/src"):
            build_ref.build()
        else:
            pytest.skip("oracle/_ref not staged and /root/reference not mounted")
    return build_ref.REF_DST


def test_staged_files_are_the_reference_byte_for_byte():
    dst = _staged()
    with open(os.path.join(dst, "MANIFEST.json")) as f:
        manifest = json.load(f)["files"]
    assert "src/advanced/maxstyle.py" in manifest and "notebooks/data/image.npy" in manifest
    for rel, digest in manifest.items():
        assert _sha(os.path.join(dst, rel)) == digest, f"{rel} was modified after staging"
        src = os.path.join("/root/reference", rel)
        if os.path.exists(src):
            assert _sha(src) == digest, f"{rel} differs from /root/reference"


def test_staging_is_git_ignored_and_travels():
    with open(os.path.join(ROOT, ".gitignore")) as f:
        assert "oracle/_ref/" in f.read()
    ignore = os.path.join(ROOT, ".gpurunignore")
    if os.path.exists(ignore):
        with open(ignore) as f:
            assert "oracle/_ref" not in f.read()


def test_reference_imports_through_the_shims_and_its_loop_runs():
    """The reference solver + notebook fixtures + the reference's own layer, on CPU (n_iter = 1 keeps it to seconds)."""
    from oracle import ref_loop
    _staged()
    ref = ref_shims.load()
    assert ref.MaxStyle.__module__ == "src.advanced.maxstyle"
    solver = ref_loop.build_solver(ref, use_gpu=False)
    image, label = ref_loop.load_fixture(ref, torch.device("cpu"))
    assert tuple(image.shape) == (20, 1, 192, 192) and tuple(label.shape) == (20, 192, 192)
    made = []
    out = ref_loop.run_loop(ref, solver, image, label, ref.MaxStyle, seed=7, p=1.0, n_iter=1, capture=made)
    assert tuple(out.shape) == (20, 1, 192, 192) and torch.isfinite(out).all()
    assert [m.num_feature for m in made] == [16, 16, 1]                       # FCN_16 splice points 3, 4, 5
    assert all(m.gamma_std is not None for m in made)


def test_reference_layer_timer_honours_steps_and_warmup():
    from oracle.ref_layer_bench import time_reference_layer
    _staged()
    res = time_reference_layer(4, 8, 32, 32, steps=3, warmup=2, threads=2)
    assert res["iters"] == 3 and res["kind"] == "reference" and res["samples_per_s"] > 0
