"""The product path has no CPU / eager fallback: without the compiled library it raises."""
import pytest


def test_missing_library_raises(monkeypatch):
    from maxstyle_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libmaxstyle_b200.so")
    with pytest.raises(_lib.MaxStyleLibraryError, match="no fallback"):
        _lib.get_lib()


def test_functional_wrappers_reject_cpu_tensors():
    import torch
    from maxstyle_b200 import functional as F
    with pytest.raises(RuntimeError, match="no CPU path"):
        F.instance_stats(torch.randn(2, 2, 4, 4), 1e-6, torch.zeros(1024, dtype=torch.uint8))


def test_product_code_never_touches_the_oracle():
    """oracle/ is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs
    may import it.  The package and the development tools must not."""
    import glob
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    offenders = []
    for path in glob.glob(os.path.join(root, "maxstyle_b200", "**", "*"), recursive=True) + glob.glob(os.path.join(root, "tools", "*")):
        if os.path.isfile(path) and path.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
            src = open(path, errors="ignore").read()
            if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "/root/reference" in src:
                offenders.append(os.path.relpath(path, root))
    assert not offenders, f"these product / tool files reference oracle/ or the reference checkout: {offenders}"
    bench = open(os.path.join(root, "bench.py")).read()
    # bench.py: the oracle / the staged reference only inside the sanctioned functions -- the CPU timing shared by the
    # `--impl reference` arm and the cpu_baseline leg, and the untimed parity check that runs after the timed region
    uses = [m.start() for m in re.finditer(r"^\s*(from|import)\s+oracle\b", bench, flags=re.M)]
    assert uses, "bench.py should time the reference and check parity"
    spans = []
    for name in ("cpu_reference_timing", "parity_check"):
        start = bench.index(f"def {name}(")
        nxt = re.search(r"^def ", bench[start + 4:], flags=re.M)
        spans.append((start, start + 4 + (nxt.start() if nxt else len(bench))))
    assert all(any(a <= u < b for a, b in spans) for u in uses), "oracle imported outside cpu_reference_timing / parity_check"
    timed = bench[bench.index("# ---- timed region"):bench.index("launches = F.launches.kernels - launches0")]
    assert "oracle" not in timed and "parity" not in timed
    assert "/root/reference" not in bench


def test_host_side_entry_points_need_cuda():
    """Executor, graphed step, host pipeline and MixStyle refuse to run without a CUDA device instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from maxstyle_b200 import StyleLoopExecutor, GraphedLayerStep, HostStepPipeline, MaxStyle, MixStyle
    with pytest.raises(RuntimeError, match="no CPU path"):
        StyleLoopExecutor(lambda c, l: c, lambda r: r.sum(), 4, {3: 2})
    layer = MaxStyle(4, 2, p=1.0, use_gpu=False)
    x = torch.randn(4, 2, 8, 8)
    with pytest.raises(RuntimeError, match="no CPU path"):
        GraphedLayerStep(layer, x, x.clone())
    with pytest.raises(RuntimeError, match="no CPU path"):
        HostStepPipeline(layer, (4, 2, 8, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        layer(x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        MixStyle(p=1.0)(x)
