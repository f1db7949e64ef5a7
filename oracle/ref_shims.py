"""Import shims that let the UNMODIFIED reference package import in this image.  TEST INFRASTRUCTURE ONLY.

The reference's solver (src/models/advanced_triplet_recon_segmentation_model.py) imports plotting, medical-imaging I/O and
GUI modules at module level that the hot path never calls (`from tkinter import E` model:8,
`from numpy.lib.function_base import copy` basic_operations.py:5, SimpleITK basic_operations.py:17, medpy + IPython
metrics.py:5,7, monai unetr.py:18-21, `collections.MutableMapping` data_structure.py:1, matplotlib / seaborn / skimage
save.py:4-7, vis.py:3-4).  `install()` registers a meta-path finder that serves EMPTY stand-in modules for exactly those
top-level packages when (and only when) the real one is not installed, and restores two names that newer numpy / Python
dropped.  No reference file is edited; none of the stubbed modules is touched by MaxStyle, MyDecoder, UnetDecoder or
generate_max_style_image.

`load(root)` puts `root` (the reference checkout: /root/reference here, oracle/_ref on the GPU box) on sys.path and returns
the imported modules the tests use.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

STUB_ROOTS = ("tkinter", "matplotlib", "mpl_toolkits", "seaborn", "skimage", "SimpleITK", "medpy", "monai", "IPython", "torchio",
              "tensorboardX", "PIL", "cv2", "nibabel", "tqdm_stub")


class _Anything:
    """Stands in for any class / function / constant of a stubbed module; usable as a base class and callable."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        full = self.__name__ + "." + name
        if full in sys.modules:
            return sys.modules[full]
        return _Anything if name[:1].isupper() else _Anything()


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, roots):
        self.roots = set(roots)

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in self.roots:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_installed = False


def install():
    """Idempotent.  Stubs only what is missing; real packages always win."""
    global _installed
    if _installed:
        return
    missing = []
    for root in STUB_ROOTS:
        try:
            if importlib.util.find_spec(root) is None:
                missing.append(root)
        except (ImportError, ValueError):
            missing.append(root)
    sys.meta_path.append(_StubFinder(missing))
    import collections
    import collections.abc
    if not hasattr(collections, "MutableMapping"):          # removed in Python 3.10 (data_structure.py:1)
        collections.MutableMapping = collections.abc.MutableMapping
    import numpy as np
    if importlib.util.find_spec("numpy.lib.function_base") is None:      # gone in numpy 2 (basic_operations.py:5)
        m = types.ModuleType("numpy.lib.function_base")
        m.copy = np.copy
        sys.modules["numpy.lib.function_base"] = m
    try:
        import scipy.misc                                    # noqa: F401  (save.py:13)
    except Exception:                                        # noqa: BLE001
        sys.modules["scipy.misc"] = types.ModuleType("scipy.misc")
    _installed = True


def reference_root() -> str | None:
    """oracle/_ref when it has been staged (it travels to the GPU box), else the mounted checkout, else None."""
    here = os.path.dirname(os.path.abspath(__file__))
    staged = os.path.join(here, "_ref")
    if os.path.exists(os.path.join(staged, "src", "advanced", "maxstyle.py")):
        return staged
    if os.path.exists("/root/reference/src/advanced/maxstyle.py"):
        return "/root/reference"
    return None


def load(root: str | None = None):
    """Import the reference package from `root`.  Returns a namespace with
    MaxStyle, MixStyle, MyDecoder, UnetDecoder, UnetEncoder, Solver (AdvancedTripletReconSegmentationModel), solver_module, root."""
    root = root or reference_root()
    if root is None:
        raise RuntimeError("the reference is not available: neither oracle/_ref (run __graft_entry__.build() where /root/reference "
                           "is mounted) nor /root/reference exists")
    install()
    if root not in sys.path:
        sys.path.insert(0, root)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        solver_module = importlib.import_module("src.models.advanced_triplet_recon_segmentation_model")
        unet = importlib.import_module("src.models.segmentation_models.unet")
        encdec = importlib.import_module("src.models.ebm.encoder_decoder")
        maxstyle = importlib.import_module("src.advanced.maxstyle")
        mixstyle = importlib.import_module("src.advanced.mixstyle")
        basic_ops = importlib.import_module("src.common_utils.basic_operations")
        custom_loss = importlib.import_module("src.models.custom_loss")
    return types.SimpleNamespace(root=root, solver_module=solver_module, Solver=solver_module.AdvancedTripletReconSegmentationModel,
                                 MaxStyle=maxstyle.MaxStyle, MixStyle=mixstyle.MixStyle, MyDecoder=encdec.MyDecoder,
                                 UnetDecoder=unet.UnetDecoder, UnetEncoder=unet.UnetEncoder, basic_operations=basic_ops,
                                 custom_loss=custom_loss, maxstyle_module=maxstyle)
