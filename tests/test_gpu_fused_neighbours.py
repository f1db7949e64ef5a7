"""the layer fused with the activation in front of it (LeakyReLU(0.2) of res_up_family, the sigmoid in front of
layer 5) and with the min / max that rescale_intensity needs behind it -- against the reference's own ops in sequence
(torch activation -> reference MaxStyle -> reference rescale_intensity), the unmodified reference imported from oracle/_ref."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref():
    from oracle import ref_shims
    if ref_shims.reference_root() is None:
        pytest.skip("oracle/_ref not staged")
    return ref_shims.load()


def rel(a, b):
    a = a.detach().double().cpu().numpy(); b = b.detach().double().cpu().numpy()
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _pair(ref, n, c, seed):
    """Reference layer (float64 yardstick and float32) and the replacement, same draws."""
    from maxstyle_b200 import MaxStyle
    out = []
    for cls, dt in ((ref.MaxStyle, torch.float64), (ref.MaxStyle, torch.float32), (MaxStyle, torch.float32)):
        torch.manual_seed(seed)
        layer = cls(n, c, p=1.0)
        out.append(layer.to(dt) if dt == torch.float64 else layer)
    return out


@pytest.mark.parametrize("shape,pre,slope", [((20, 16, 96, 96), "leaky_relu", 0.2), ((20, 16, 192, 192), "leaky_relu", 0.2),
                                             ((6, 3, 50, 46), "leaky_relu", 0.01), ((20, 1, 192, 192), "sigmoid", 0.0),
                                             ((4, 2, 33, 21), "sigmoid", 0.0), ((20, 64, 224, 224), "leaky_relu", 0.2)])
def test_activation_fused_into_the_layer(shape, pre, slope):
    """y, dz and the parameter gradients of layer(act(z)) -- first forward (batch std computed) and a cached-std forward."""
    ref = _ref()
    n, c, h, w = shape
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(n * 13 + h)
    truth, ref32, ours = _pair(ref, n, c, 5 + n)
    act = (lambda t: torch.nn.functional.leaky_relu(t, slope)) if pre == "leaky_relu" else torch.sigmoid
    for rep in range(2):                                         # rep 1: gamma_std / beta_std come from the cache
        z = torch.randn(n, c, h, w, device=dev, generator=g) * (2.0 if pre == "sigmoid" else 1.5) + 0.3
        dy = torch.randn(n, c, h, w, device=dev, generator=g)
        res = {}
        for name, layer, dt in (("truth", truth, torch.float64), ("ref32", ref32, torch.float32), ("ours", ours, torch.float32)):
            zi = z.to(dt).clone().requires_grad_(True)
            layer.zero_grad()
            y = ours.forward_fused(zi, pre, slope) if name == "ours" else layer(act(zi))
            y.backward(dy.to(dt))
            res[name] = dict(y=y.detach(), dz=zi.grad, **{k: p.grad.clone() for k, p in layer.named_parameters()})
        for key, tol in (("y", 1e-5), ("dz", 1e-4), ("gamma_noise", 1e-4), ("beta_noise", 1e-4), ("lmda", 1e-4)):
            e_o, e_r = rel(res["ours"][key], res["truth"][key]), rel(res["ref32"][key], res["truth"][key])
            assert e_o <= max(tol, 3 * e_r), f"rep {rep} {key}: fused {e_o:.2e} (reference ops in fp32: {e_r:.2e})"


@pytest.mark.parametrize("shape", [(20, 1, 192, 192), (5, 3, 40, 36), (8, 2, 224, 224)])
def test_minmax_collected_for_rescale_intensity(shape):
    """sigmoid -> layer -> rescale_intensity (the tail of the reference's loop, model:868-869) in two kernels."""
    from maxstyle_b200 import rescale_intensity
    ref = _ref()
    n, c, h, w = shape
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(7)
    _, ref32, ours = _pair(ref, n, c, 21)
    z = torch.randn(n, c, h, w, device=dev, generator=g) * 2.0
    want_y = ref32(torch.sigmoid(z))
    want = ref.basic_operations.rescale_intensity(want_y, 0, 1)
    y, mm = ours.forward_fused(z, "sigmoid", collect_minmax=True)
    got = rescale_intensity(y, mm, 0.0, 1.0)
    assert rel(y, want_y) < 1e-5
    assert float((got - want).abs().max()) < 1e-5                               # images in [0, 1]
    assert float(got.min()) >= 0.0 and float(got.amax(dim=(2, 3)).min()) > 0.999


def test_decoder_splice_fused_matches_reference_splice():
    """`apply_max_style_fused` on the reference's MyDecoder (notebook weights) == the reference's `apply_max_style` with the
    reference layers, image and first-pass layer gradients."""
    from maxstyle_b200 import MaxStyle, apply_max_style_fused
    from oracle import ref_loop
    ref = _ref()
    dev = torch.device("cuda:0")
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.allow_tf32 = False            # TF32 convolutions would turn last-bit differences of y into 1e-4 ones
    torch.backends.cuda.matmul.allow_tf32 = False
    solver = ref_loop.build_solver(ref, "FCN_16_standard_no_STN", use_gpu=True, pretrained=True)
    image, label = ref_loop.load_fixture(ref, dev)
    with torch.no_grad():
        (z_i, z_s), _ = solver.fast_predict(image)
    dec = solver.model["image_decoder"]
    for m in solver.model.values():
        ref.basic_operations.set_grad(m, requires_grad=False)
    target = torch.rand_like(image)
    outs = {}
    for name, cls in (("ref", ref.MaxStyle), ("ours", MaxStyle)):
        torch.manual_seed(3)
        mods = torch.nn.ModuleDict({str(k): cls(20, ch, p=1.0) for k, ch in ((3, 16), (4, 16), (5, 1))})
        if name == "ref":
            out = dec.apply_max_style(z_i, decoder_layers_indexes=[3, 4, 5], nn_style_augmentor_dict=mods)
        else:
            out = apply_max_style_fused(dec, z_i, mods, [3, 4, 5])
        ((out - target) ** 2).mean().backward()
        outs[name] = (out.detach(), {k: {n: p.grad.clone() for n, p in m.named_parameters()} for k, m in mods.items()})
    assert rel(outs["ours"][0], outs["ref"][0]) < 1e-5
    for k in outs["ref"][1]:
        for n, gr in outs["ref"][1][k].items():
            assert rel(outs["ours"][1][k][n], gr) < 2e-3, f"layer {k} d{n}"      # through 3 conv blocks: see test_gpu_reference_callers


def test_fused_identity_cases_and_errors():
    from maxstyle_b200 import MaxStyle
    torch.manual_seed(0)
    layer = MaxStyle(4, 3, p=0.0)                                  # never active
    z = torch.randn(4, 3, 16, 16, device="cuda")
    assert torch.equal(layer.forward_fused(z, "leaky_relu", 0.2), torch.nn.functional.leaky_relu(z, 0.2))
    y, mm = layer.forward_fused(z, "sigmoid", collect_minmax=True)
    assert mm is None and torch.equal(y, torch.sigmoid(z))
    with pytest.raises(ValueError):
        layer.forward_fused(z, "tanh")
