"""Harness around the UNMODIFIED reference callers of the layer.  TEST INFRASTRUCTURE ONLY.

Builds the reference's own solver (`AdvancedTripletReconSegmentationModel`, model:41) exactly as its notebook does
(notebooks/vis_hard_example.ipynb cells 5-9: FCN_16_standard_no_STN, 4 classes, shipped weights, the 20 x 192 x 192 batch) and
runs its `generate_max_style_image` (model:458-571) with either the reference's `MaxStyle` or a replacement class swapped in
by name -- the drop-in point INTEGRATION.md describes (the solver looks the class up in its module globals at model:523).
Nothing of the reference is edited: the swap is `setattr(solver_module, "MaxStyle", cls)`.

On a CPU-only box two test-side patches are needed: the weights were saved from CUDA (`torch.load` needs
map_location) and the solver constructs `MaxStyle` with the default use_gpu=True.  Neither applies on the GPU box.
"""
from __future__ import annotations

import contextlib
import os

import numpy as np
import torch

from . import ref_shims


def load_fixture(ref, device):
    """The notebook's batch: image [20,1,192,192] float in [0,1] (rescaled like cell 7), label [20,192,192] int64."""
    image = np.load(os.path.join(ref.root, "notebooks", "data", "image.npy"))
    label = np.load(os.path.join(ref.root, "notebooks", "data", "label.npy"))
    image_v = torch.from_numpy(image[:, np.newaxis, :, :]).float().to(device)
    label_v = torch.from_numpy(label).long().to(device)
    image_v = ref.basic_operations.rescale_intensity(image_v)
    return image_v, label_v


@contextlib.contextmanager
def _cpu_patches(ref, use_gpu: bool):
    if use_gpu:
        yield
        return
    real_load = torch.load

    def load_cpu(path, *a, **k):
        k.setdefault("map_location", "cpu")
        return real_load(path, *a, **k)

    torch.load = load_cpu
    try:
        yield
    finally:
        torch.load = real_load


def build_solver(ref, network_type="FCN_16_standard_no_STN", use_gpu=True, pretrained=True, image_size=192):
    """The reference solver, constructed like notebook cell 5 (pretrained=False: kaiming-initialised networks, used for FCN_64
    for which the reference ships no weights)."""
    ckpt = os.path.join(ref.root, "notebooks", "model") if pretrained else None
    import io
    with _cpu_patches(ref, use_gpu), contextlib.redirect_stdout(io.StringIO()):
        solver = ref.Solver(network_type=network_type, checkpoint_dir=ckpt, num_classes=4, use_gpu=use_gpu, debug=False,
                            image_size=image_size)
        solver.eval()
    return solver


def _widened(cls):
    """Factory: construct `cls` as usual (float32 draws from the generators), then widen the module to float64."""
    def make(*a, **k):
        return cls(*a, **k).double()
    return make


class _CpuLayer:
    """On a CPU box the reference layer must be built with use_gpu=False; the solver does not pass the argument."""

    def __init__(self, cls):
        self.cls = cls

    def __call__(self, *a, **k):
        k.setdefault("use_gpu", False)
        return self.cls(*a, **k)


def run_loop(ref, solver, image_v, label_v, layer_cls, *, seed=7, p=1.0, n_iter=5, layers=(3, 4, 5), channel_num=(128, 64, 32, 16, 16, 1),
             always_use_beta=True, capture=None, double=False):
    """solver.generate_max_style_image with `layer_cls` as the MaxStyle class.  `capture`: optional list that receives the
    layer modules the solver constructed (their parameters / gradients can be inspected afterwards).  `double=True`: networks,
    layers (the same float32 draws, widened) and Adam state in float64 -- the yardstick float32 runs are measured against."""
    if double:
        for m in solver.model.values():
            m.double()
        try:
            return run_loop(ref, solver, image_v.double(), label_v, _widened(layer_cls), seed=seed, p=p, n_iter=n_iter, layers=layers,
                            channel_num=channel_num, always_use_beta=always_use_beta, capture=capture)
        finally:
            for m in solver.model.values():
                m.float()
    use_gpu = image_v.is_cuda
    with torch.no_grad():
        (z_i, z_s), _ = solver.fast_predict(image_v)
    made = []

    def factory(*a, **k):
        if not use_gpu:
            k.setdefault("use_gpu", False)
        m = layer_cls(*a, **k)
        made.append(m)
        return m

    old = ref.solver_module.MaxStyle
    ref.solver_module.MaxStyle = factory
    try:
        out = solver.generate_max_style_image(image_code=z_i, decoder_layers_indexes=list(layers), channel_num=list(channel_num), p=p,
                                              n_iter=n_iter, mix_style=True, lr=0.1, no_noise=False, reference_image=image_v,
                                              reference_segmentation=label_v, noise_learnable=True, mix_learnable=True,
                                              loss_types=["seg"], loss_weights=[1], always_use_beta=always_use_beta, debug=False,
                                              fix_seed=seed)
    finally:
        ref.solver_module.MaxStyle = old
    if capture is not None:
        capture.extend(made)
    return out


def first_iteration_grads(ref, solver, image_v, label_v, layer_cls, *, seed=7, p=1.0, layers=(3, 4, 5),
                          channel_num=(128, 64, 32, 16, 16, 1), always_use_beta=True, double=False):
    """Gradients of the reference's loop at its first iteration, without the optimiser step: the loop body of model:539-566
    (decode with the layers, re-encode, segment, -CE, backward) driven through the reference's own methods.
    `double=True` runs networks and layers in float64 (the same weights and the same float32 parameter draws, widened):
    the yardstick against which float32 runs of the reference layer and of a replacement are both measured."""
    use_gpu = image_v.is_cuda
    if double:
        for m in solver.model.values():
            m.double()
        try:
            return _first_iteration(ref, solver, image_v.double(), label_v, layer_cls, seed, p, layers, channel_num, always_use_beta, True)
        finally:
            for m in solver.model.values():
                m.float()                  # float(double(w)) == w: the float32 weights come back bit for bit
    return _first_iteration(ref, solver, image_v, label_v, layer_cls, seed, p, layers, channel_num, always_use_beta, False)


def _first_iteration(ref, solver, image_v, label_v, layer_cls, seed, p, layers, channel_num, always_use_beta, double):
    use_gpu = image_v.is_cuda
    torch.manual_seed(seed)
    with torch.no_grad():
        (z_i, z_s), _ = solver.fast_predict(image_v)
    n = z_i.size(0)
    mods = {}
    for i in layers:
        kw = dict(p=p, mix_style=True, no_noise=False, mix_learnable=True, noise_learnable=True, always_use_beta=always_use_beta, debug=False)
        if not use_gpu:
            kw["use_gpu"] = False
        mods[str(i)] = layer_cls(n, channel_num[i], **kw)
    md = torch.nn.ModuleDict(mods)
    if double:
        md.double()
    for m in solver.model.values():
        ref.basic_operations.set_grad(m, requires_grad=False)
    captured = {}

    def grab(key):
        def hook(_mod, inputs, output):
            captured[key] = [inputs[0].detach().clone(), None]
            if output.requires_grad:
                output.register_hook(lambda g: captured[key].__setitem__(1, g.detach().clone()))
        return hook

    handles = [m.register_forward_hook(grab(k)) for k, m in mods.items()]
    recon = solver.model["image_decoder"].apply_max_style(z_i, decoder_layers_indexes=list(layers), nn_style_augmentor_dict=md)
    for h_ in handles:
        h_.remove()
    zi2, zs2 = solver.encode_image(recon, disable_track_bn_stats=True)
    pred = solver.decoder_inference(decoder=solver.model["segmentation_decoder"], latent_code=zs2, eval=False, disable_track_bn_stats=True)
    loss = -ref.custom_loss.basic_loss_fn(pred=pred, target=label_v, loss_type="cross entropy", class_weights=None, use_gpu=use_gpu)
    loss.backward()
    grads = {k: {name: (prm.grad.detach().clone() if prm.grad is not None else None) for name, prm in m.named_parameters()}
             for k, m in mods.items()}
    for k, m in mods.items():
        m.captured_io = tuple(captured.get(k, (None, None)))       # (input feature map, upstream gradient) of this layer in the loop
    return recon.detach(), float(loss.detach()), grads, mods
