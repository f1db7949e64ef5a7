"""The layer fused with its neighbours in the reference's decoder.

In `MyDecoder.apply_max_style` (src/models/ebm/encoder_decoder.py:598-631) every spliced layer follows an activation that is a
kernel of its own in the reference: the LeakyReLU(0.2) that ends `res_up_family.forward` (:337-357: `get_features` + `non_linear`)
in front of layers 0-4, and the sigmoid `last_act` (:624-627) in front of layer 5; the loop's output then goes through
`rescale_intensity` (src/common_utils/basic_operations.py:257-281, called at model:868-869), two more reductions over the image.
`apply_max_style_fused` is the same splice with those three neighbours folded into the layer's kernels:

    x = up_k.get_features(x)                    # stock PyTorch / cuDNN: ConvTranspose + residual conv branch (unchanged)
    x = layer_k.forward_fused(x, "leaky_relu")  # LeakyReLU applied as the layer loads x: one full-tensor pass less
    ...
    x = final_conv(x)
    y, mm = layer_5.forward_fused(x, "sigmoid", collect_minmax=True)     # sigmoid on load, per-plane min / max on store
    image = rescale_intensity(y, mm)            # one pass instead of two reductions + four elementwise kernels

Blocks whose layer is not spliced (or whose `dropout` is set: the reference applies it between activation and layer) run
exactly as in the reference.
"""
from __future__ import annotations

import contextlib

import torch

from . import _lib as L
from . import functional as F


@contextlib.contextmanager
def _batch_stats_only(module):
    """What the reference's `_disable_tracking_bn_stats` (src/models/model_util.py:468-509) does around every decoder block of
    `apply_max_style`, side effects included: BatchNorm2d/3d layers get `track_running_stats = False` and their weight / bias
    `requires_grad_(False)` for the duration of the block, and on the way out `track_running_stats` is restored and weight /
    bias `requires_grad_` is set to that restored value (the reference does exactly this, model_util.py:489-495)."""
    saved = []
    for m in module.modules():
        if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
            saved.append((m, m.track_running_stats))
            m.track_running_stats = False
            if getattr(m, "weight", None) is not None:
                m.weight.requires_grad_(False)
            if getattr(m, "bias", None) is not None:
                m.bias.requires_grad_(False)
    try:
        yield
    finally:
        for m, v in saved:
            m.track_running_stats = v
            if getattr(m, "weight", None) is not None:
                m.weight.requires_grad_(bool(v))
            if getattr(m, "bias", None) is not None:
                m.bias.requires_grad_(bool(v))


def rescale_intensity(y: torch.Tensor, minmax, new_min: float = 0.0, new_max: float = 1.0, eps: float = 1e-20) -> torch.Tensor:
    """`rescale_intensity(y, new_min, new_max)` of the reference (basic_operations.py:257-281) from the per-plane extremes
    `forward_fused(..., collect_minmax=True)` returned: out = (y - min) / (max - min + eps) * (new_max - new_min) + new_min."""
    if minmax is None:
        raise RuntimeError("maxstyle_b200: no min/max were collected (the layer was inactive for this draw); use the reference's "
                           "rescale_intensity on this tensor")
    if not y.is_cuda:
        raise RuntimeError("maxstyle_b200: rescale_intensity got a CPU tensor (there is no CPU path)")
    y = y.contiguous()
    n, c, h, w = y.shape
    out = torch.empty_like(y)
    with F.device_guard(y.device):
        rc = L.get_lib().maxstyle_rescale(y.data_ptr(), minmax[0].data_ptr(), minmax[1].data_ptr(), out.data_ptr(), float(new_min),
                                          float(new_max), float(eps), n, c, h, w, F.dtype_code(y), F._stream())
    L.check(rc, "maxstyle_rescale")
    F.launches.kernels += 1
    return out


def apply_max_style_fused(decoder, image_code, nn_style_augmentor_dict, decoder_layers_indexes=(3, 4, 5), collect_minmax: bool = False):
    """`decoder.apply_max_style(image_code, nn_style_augmentor_dict, decoder_layers_indexes)` for the reference's `MyDecoder`
    (same blocks, same order, same BatchNorm handling) with each spliced layer fused with the activation in front of it.
    Returns the decoded image; with `collect_minmax=True` (and layer 5 spliced and active) also the (min, max) pair for
    `rescale_intensity`."""
    idx = set(int(i) for i in decoder_layers_indexes)
    layers = nn_style_augmentor_dict

    def fusable(block):
        return hasattr(block, "get_features") and getattr(block, "dropout", None) is None and isinstance(
            getattr(block, "last_act", None), torch.nn.LeakyReLU)

    x = image_code.detach().clone()
    if 0 in idx:
        x = layers["0"](x)
    for k, name in enumerate(("up1", "up2", "up3", "up4"), start=1):
        block = getattr(decoder, name)
        with _batch_stats_only(block):
            if k in idx and fusable(block):
                x = layers[str(k)].forward_fused(block.get_features(x), "leaky_relu", block.last_act.negative_slope)
            else:
                x = block(x)
                if k in idx:
                    x = layers[str(k)](x)
    with _batch_stats_only(decoder.final_conv):
        x = decoder.final_conv(x)
    mm = None
    last_act = getattr(decoder, "last_act", None)
    if 5 in idx and isinstance(last_act, torch.nn.Sigmoid):
        res = layers["5"].forward_fused(x, "sigmoid", collect_minmax=collect_minmax)
        x, mm = res if collect_minmax else (res, None)
    else:
        if last_act is not None:
            x = last_act(x)
        if 5 in idx:
            x = layers["5"](x)
    return (x, mm) if collect_minmax else x
