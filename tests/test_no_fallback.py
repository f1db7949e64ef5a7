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
