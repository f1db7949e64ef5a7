"""The replacement layer inside the reference's OWN callers, unmodified and imported from
oracle/_ref (staged by `__graft_entry__.build()`; the GPU box has no /root/reference):

  * `AdvancedTripletReconSegmentationModel.generate_max_style_image` (model:458-571) on the reference notebook's fixtures
    (notebooks/data/image.npy, label.npy, notebooks/model/*.pth; vis_hard_example.ipynb cells 5-9) -- run once with the
    reference's MaxStyle on CUDA and once with maxstyle_b200.MaxStyle swapped in by name, same `fix_seed`;
  * `MyDecoder.apply_max_style` (encoder_decoder.py:598-631) -- inside the solver above, and for the FCN_64 widths;
  * `UnetDecoder.apply_max_style` (unet.py:104-137).

Tolerances.  The decoded image of the first pass carries BASELINE.json's 1e-5 (max-norm relative).  The layer's own gradients
are pinned at 1e-4 against the reference in tests/test_gpu_parity.py (same x, same dy).  HERE the gradients have travelled
through ~30 frozen conv / BatchNorm(batch statistics) / LeakyReLU layers after leaving the layer and before coming back to
it, which amplify last-bit differences of y; so the yardstick is a float64 run of the reference (same weights, same draws,
widened): the replacement must be as close to it as the reference's own float32 run is (within 3x, or 1e-4).  After n_iter
Adam(lr=0.1) steps on a non-convex loss, rounding differences are amplified step by step (a +-1-ulp difference in a gradient
near zero flips the sign of the first Adam update, which is +-lr whatever the gradient's size); the returned image is
therefore compared with a looser, stated bound and the reference-vs-reference run-to-run spread is reported beside it.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref():
    from oracle import ref_shims
    if ref_shims.reference_root() is None:
        pytest.skip("oracle/_ref not staged (run __graft_entry__.build() where /root/reference is mounted)")
    return ref_shims.load()


def rel(a, b):
    a = a.detach().double().cpu().numpy(); b = b.detach().double().cpu().numpy()
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope="module")
def setup():
    from oracle import ref_loop
    ref = _ref()
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda:0")
    solver = ref_loop.build_solver(ref, "FCN_16_standard_no_STN", use_gpu=True, pretrained=True)
    image, label = ref_loop.load_fixture(ref, dev)
    return ref, solver, image, label


def test_first_decode_and_first_iteration_gradients_match_reference(setup):
    """model:539-566, first pass: decode through the three layers (MyDecoder.apply_max_style), re-encode, segment, -CE,
    backward.  Same seed => same layer state (RNG contract on the CUDA generator); y within 1e-5; gradients as close to the
    float64 run of the reference as the reference's own float32 run."""
    from maxstyle_b200 import MaxStyle
    from oracle import ref_loop
    ref, solver, image, label = setup
    for seed, beta in ((7, True), (47, False)):
        recon_r, loss_r, grads_r, mods_r = ref_loop.first_iteration_grads(ref, solver, image, label, ref.MaxStyle, seed=seed, always_use_beta=beta)
        recon_o, loss_o, grads_o, mods_o = ref_loop.first_iteration_grads(ref, solver, image, label, MaxStyle, seed=seed, always_use_beta=beta)
        for k in mods_r:
            assert torch.equal(mods_r[k].perm, mods_o[k].perm), f"layer {k}: perm differs"
            for name in ("gamma_noise", "beta_noise", "lmda"):
                assert torch.equal(getattr(mods_r[k], name).detach(), getattr(mods_o[k], name).detach()), f"layer {k}: {name} drawn differently"
        assert rel(recon_o, recon_r) < 1e-5, f"decoded image: {rel(recon_o, recon_r):.2e}"
        assert abs(loss_o - loss_r) <= 1e-5 * max(abs(loss_r), 1e-3)
        recon_t, loss_t, grads_t, _ = ref_loop.first_iteration_grads(ref, solver, image, label, ref.MaxStyle, seed=seed, always_use_beta=beta, double=True)
        assert rel(recon_o, recon_t) < 1e-5
        # (1) through the network: the float32 noise floor of this loop is what the reference's own float32 run shows against
        #     its float64 run (1e-4 .. 6e-4 here); the replacement has to stay within 5x of the worst of those
        floor = max(rel(grads_r[k][name], grads_t[k][name]) for k in grads_r for name in grads_r[k])
        for k in grads_r:
            for name, g_r in grads_r[k].items():
                g_o, g_t = grads_o[k][name], grads_t[k][name]
                assert g_o is not None and g_r is not None and g_t is not None
                e_ref, e_ours = rel(g_r, g_t), rel(g_o, g_t)
                print(f"seed {seed} layer {k} d{name}: reference fp32 vs fp64 {e_ref:.2e}, replacement vs fp64 {e_ours:.2e}, vs each other {rel(g_o, g_r):.2e}")
                assert e_ours <= max(5 * floor, 1e-4), f"layer {k} d{name}: ours {e_ours:.2e}, reference floor {floor:.2e} (seed {seed})"
        # (2) the layers themselves, on the activations and upstream gradients they met inside the reference's loop: the
        #     replacement against the float64 reference layer on the SAME (x, dy) -- BASELINE.json's 1e-5 / 1e-4
        for k, m_r in mods_r.items():
            x_in, g_up = m_r.captured_io
            assert x_in is not None and g_up is not None
            n, c = x_in.shape[0], x_in.shape[1]
            kw = dict(p=1.0, always_use_beta=beta)
            out = {}
            for tag, cls, dt in (("truth", ref.MaxStyle, torch.float64), ("ref32", ref.MaxStyle, torch.float32), ("ours", MaxStyle, torch.float32)):
                layer = cls(n, c, **kw)
                layer.perm = m_r.perm.clone()
                if hasattr(layer, "_perm_dev"):
                    layer._perm_dev = None
                with torch.no_grad():
                    for name in ("gamma_noise", "beta_noise", "lmda"):
                        getattr(layer, name).copy_(getattr(m_r, name).detach())
                layer.rand_p = torch.zeros(1)
                layer = layer.to(dt) if dt == torch.float64 else layer
                xi = x_in.to(dt).clone().requires_grad_(True)
                yi = layer(xi)
                yi.backward(g_up.to(dt))
                out[tag] = dict(y=yi.detach(), dx=xi.grad, **{name: prm.grad for name, prm in layer.named_parameters()})
            for key, tol in (("y", 1e-5), ("dx", 1e-4), ("gamma_noise", 1e-4), ("beta_noise", 1e-4), ("lmda", 1e-4)):
                e_o, e_r = rel(out["ours"][key], out["truth"][key]), rel(out["ref32"][key], out["truth"][key])
                print(f"seed {seed} layer {k} standalone {key}: replacement vs fp64 reference {e_o:.2e} (reference fp32: {e_r:.2e})")
                assert e_o <= tol, f"layer {k} {key}: {e_o:.2e} > {tol:g} on the loop's own activations (seed {seed})"


@pytest.mark.parametrize("n_iter", [0, 1, 5])
def test_generate_max_style_image_with_replacement(setup, n_iter):
    """The whole loop (model:458-571), p=1 so all three layers are active, notebook fixtures, the reference's own Adam."""
    from maxstyle_b200 import MaxStyle
    from oracle import ref_loop
    ref, solver, image, label = setup
    seed = 7
    made_r, made_o = [], []
    out_r = ref_loop.run_loop(ref, solver, image, label, ref.MaxStyle, seed=seed, p=1.0, n_iter=n_iter, capture=made_r)
    out_r2 = ref_loop.run_loop(ref, solver, image, label, ref.MaxStyle, seed=seed, p=1.0, n_iter=n_iter)
    out_o = ref_loop.run_loop(ref, solver, image, label, MaxStyle, seed=seed, p=1.0, n_iter=n_iter, capture=made_o)
    assert out_o.shape == out_r.shape == (20, 1, 192, 192)
    noise = float((out_r - out_r2).abs().max())
    err = float((out_o - out_r).abs().max())
    mean_err = float((out_o - out_r).abs().mean())
    # float64 yardstick: the reference loop with everything widened (same draws).  After n_iter Adam(lr=0.1) steps the float32
    # trajectories of reference AND replacement have drifted from it; the replacement must not drift more than the reference.
    out_t = ref_loop.run_loop(ref, solver, image, label, ref.MaxStyle, seed=seed, p=1.0, n_iter=n_iter)  if n_iter == 0 else \
        ref_loop.run_loop(ref, solver, image, label, ref.MaxStyle, seed=seed, p=1.0, n_iter=n_iter, double=True)
    dr_max, dr_mean = float((out_r.double() - out_t.double()).abs().max()), float((out_r.double() - out_t.double()).abs().mean())
    do_max, do_mean = float((out_o.double() - out_t.double()).abs().max()), float((out_o.double() - out_t.double()).abs().mean())
    print(f"n_iter={n_iter}: |ours - ref| max {err:.3e} mean {mean_err:.3e}; ref run-to-run max {noise:.3e}; vs float64 loop: "
          f"reference max {dr_max:.3e} mean {dr_mean:.3e}, replacement max {do_max:.3e} mean {do_mean:.3e}")
    # images live in [0, 1] (sigmoid output)
    bound = {0: 1e-5, 1: 2e-4, 5: 5e-3}[n_iter]
    if n_iter <= 1:
        assert err <= max(bound, 10 * noise, 3 * dr_max), f"n_iter={n_iter}: max abs err {err:.3e} (reference's own float32 drift {dr_max:.3e})"
    else:
        # two float32 trajectories that have each drifted from the float64 one are up to dr + do apart from EACH OTHER, so the
        # replacement is held against the yardstick instead: its mean drift within 1.5x of the reference's own (measured: 2.8e-4
        # vs 3.3e-4), its worst pixel (a heavy-tailed maximum over 737k pixels of an amplifying loop) within 5x.
        assert do_max <= max(bound, 5 * dr_max), f"n_iter={n_iter}: worst-pixel drift {do_max:.3e} vs the reference's {dr_max:.3e}"
        assert do_mean <= max(bound / 50, 1.5 * dr_mean), f"n_iter={n_iter}: mean drift {do_mean:.3e} vs the reference's {dr_mean:.3e}"
    assert do_mean <= max(bound / 10, 3 * dr_mean), f"n_iter={n_iter}: mean drift {do_mean:.3e} vs the reference's {dr_mean:.3e}"
    for m_r, m_o in zip(made_r, made_o):
        for name in ("gamma_noise", "beta_noise", "lmda"):
            a, b = getattr(m_o, name).detach(), getattr(m_r, name).detach()
            d = (a - b).abs()
            if n_iter == 0:
                assert float(d.max()) == 0.0, f"{name}: drawn differently"
            else:
                # an Adam step is +-lr whatever the gradient's size: a parameter whose gradient is rounding noise may step the
                # other way (2 * lr apart).  Such parameters must be rare; everything else agrees closely.
                assert float(d.max()) <= 2 * 0.1 * n_iter + 1e-6
                moved = float((d > 5e-3 * n_iter).float().mean())
                print(f"n_iter={n_iter} {name}: max |diff| {float(d.max()):.3e}, fraction further apart than {5e-3 * n_iter:g}: {moved:.3f}")
                if n_iter == 1:            # after more steps the trajectories of the sign-flipped entries have spread; the image bound above is the check
                    assert moved <= 0.02, f"{name} after one step: {moved:.3f} of the entries moved apart"


def test_default_probability_and_inactive_layers(setup):
    """p=0.5 as shipped (train...py:263): some layers draw inactive and are the identity; same draws on both sides."""
    from maxstyle_b200 import MaxStyle
    from oracle import ref_loop
    ref, solver, image, label = setup
    for seed in (7, 47, 87, 127):
        a, b = [], []
        out_r = ref_loop.run_loop(ref, solver, image, label, ref.MaxStyle, seed=seed, p=0.5, n_iter=1, capture=a)
        out_o = ref_loop.run_loop(ref, solver, image, label, MaxStyle, seed=seed, p=0.5, n_iter=1, capture=b)
        assert [float(m.rand_p) for m in a] == [float(m.rand_p) for m in b]
        assert float((out_o - out_r).abs().max()) <= 2e-4


def test_unet_decoder_apply_max_style():
    """UnetDecoder.apply_max_style (unet.py:104-137): kaiming-initialised reference UNet (reduce_factor 4), layers after up3 /
    up4 / the output conv; outputs within 1e-5 and layer gradients within 1e-4 of the reference layer in the same network."""
    from maxstyle_b200 import MaxStyle
    ref = _ref()
    dev = torch.device("cuda:0")
    torch.backends.cudnn.deterministic = True
    torch.manual_seed(3)
    enc = ref.UnetEncoder(input_channel=1, reduce_factor=4, encoder_dropout=None, norm=torch.nn.BatchNorm2d).to(dev)
    dec = ref.UnetDecoder(n_classes=4, reduce_factor=4, decoder_dropout=None, norm=torch.nn.BatchNorm2d, last_act=None).to(dev)
    x = torch.rand(12, 1, 128, 128, device=dev)
    with torch.no_grad():
        feats = enc(x)
    target = torch.randn(12, 4, 128, 128, device=dev)
    results = {}
    for name, cls in (("ref", ref.MaxStyle), ("ours", MaxStyle)):
        torch.manual_seed(11)
        mods = torch.nn.ModuleDict({"3": cls(12, 16, p=1.0), "4": cls(12, 16, p=1.0), "5": cls(12, 4, p=1.0)})
        out = dec.apply_max_style(feats, decoder_layers_indexes=[3, 4, 5], nn_style_augmentor_dict=mods)
        ((out - target) ** 2).mean().backward()
        results[name] = (out.detach(), {k: {n: p.grad.clone() for n, p in m.named_parameters()} for k, m in mods.items()})
        dec.zero_grad()
    assert rel(results["ours"][0], results["ref"][0]) < 1e-5
    # float64 yardstick: the same networks and draws, widened
    enc.double(); dec.double()
    torch.manual_seed(11)
    mods = torch.nn.ModuleDict({"3": ref.MaxStyle(12, 16, p=1.0), "4": ref.MaxStyle(12, 16, p=1.0), "5": ref.MaxStyle(12, 4, p=1.0)}).double()
    out = dec.apply_max_style([f.double() for f in feats], decoder_layers_indexes=[3, 4, 5], nn_style_augmentor_dict=mods)
    ((out - target.double()) ** 2).mean().backward()
    truth = {k: {n: p.grad.clone() for n, p in m.named_parameters()} for k, m in mods.items()}
    for k in results["ref"][1]:
        for n, g in results["ref"][1][k].items():
            e_ref, e_ours = rel(g, truth[k][n]), rel(results["ours"][1][k][n], truth[k][n])
            assert e_ours <= max(3 * e_ref, 1e-4), f"layer {k} d{n}: ours {e_ours:.2e}, reference {e_ref:.2e}"


def test_mydecoder_fcn64_config1_shapes():
    """MyDecoder.apply_max_style at the FCN_64 widths (channel_num [512,256,128,64,64,1], train...py:257) on a 224 x 224 batch of
    20: layer 4 sees BASELINE config 1's 20 x 64 x 224 x 224.  Kaiming-initialised reference decoder (no FCN_64 weights ship)."""
    from maxstyle_b200 import MaxStyle
    from oracle import ref_loop
    ref = _ref()
    dev = torch.device("cuda:0")
    torch.backends.cudnn.deterministic = True
    torch.manual_seed(5)
    solver = ref_loop.build_solver(ref, "FCN_64_standard_no_STN", use_gpu=True, pretrained=False, image_size=224)
    image = torch.rand(20, 1, 224, 224, device=dev)
    label = torch.randint(0, 4, (20, 224, 224), device=dev)
    chans = (512, 256, 128, 64, 64, 1)
    recon_r, loss_r, grads_r, _ = ref_loop.first_iteration_grads(ref, solver, image, label, ref.MaxStyle, seed=9, channel_num=chans, always_use_beta=False)
    recon_o, loss_o, grads_o, mods_o = ref_loop.first_iteration_grads(ref, solver, image, label, MaxStyle, seed=9, channel_num=chans, always_use_beta=False)
    assert tuple(mods_o["4"].data.shape) == (20, 64, 224, 224)
    assert rel(recon_o, recon_r) < 1e-5
    _, _, grads_t, _ = ref_loop.first_iteration_grads(ref, solver, image, label, ref.MaxStyle, seed=9, channel_num=chans, always_use_beta=False, double=True)
    for k in grads_r:
        for n, g in grads_r[k].items():
            e_ref, e_ours = rel(g, grads_t[k][n]), rel(grads_o[k][n], grads_t[k][n])
            assert e_ours <= max(3 * e_ref, 1e-4), f"layer {k} d{n}: ours {e_ours:.2e}, reference {e_ref:.2e}"
