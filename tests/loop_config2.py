#!/usr/bin/env python
"""BASELINE config 2: decoder with MaxStyle after its last 3 blocks, the full inner style-optimisation loop on one GPU.

TEST / MEASUREMENT INFRASTRUCTURE (it imports oracle/).  The reference model (AdvancedTripletReconSegmentationModel,
`generate_max_style_image`, src/models/advanced_triplet_recon_segmentation_model.py:458-571, and
`MyDecoder.apply_max_style`, src/models/ebm/encoder_decoder.py:598-631) cannot travel to the GPU box, so this file
carries a stock-PyTorch stand-in with the same block structure and activation shapes (bilinear up, 1x1 skip + two 3x3
convs with InstanceNorm, LeakyReLU(0.2); channel widths of FCN_16 / FCN_64) -- the convolutions are out of scope and stay
cuDNN -- and runs the reference's loop around it twice: once with the reference layer's eager op chain on the GPU
(oracle.torch_port.StylePort, what `MaxStyle(use_gpu=True)` executes) and once with `maxstyle_b200.MaxStyle` (+ the
fused style step), from identical parameters.  Reports loop time, the layers' share, and output parity.

    python tests/loop_config2.py [--width 64] [--batch 20] [--size 224] [--n-iter 5] [--reps 5]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch
import torch.nn as nn
import torch.nn.functional as TF


class ResUp(nn.Module):
    """res_up_family stand-in (encoder_decoder.py:300-357)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
        self.conv = nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1, bias=False), nn.InstanceNorm2d(cout), nn.LeakyReLU(0.2),
                                  nn.Conv2d(cout, cout, 3, padding=1, bias=False), nn.InstanceNorm2d(cout))
        self.conv_input = nn.Conv2d(cin, cout, 1, bias=False)
        self.act = nn.LeakyReLU(0.2)

    def forward(self, x):
        x = self.up(x)
        return self.act(self.conv_input(x) + self.conv(x))


class Decoder(nn.Module):
    """MyDecoder stand-in: four up blocks + 1x1 head; `apply_max_style` splices the style layers exactly where
    encoder_decoder.py:598-631 does (indices 0..5)."""

    def __init__(self, cin, width, cout, last_act=None):
        super().__init__()
        self.widths = [4 * width, 2 * width, width, width]
        chans = [cin] + self.widths
        self.ups = nn.ModuleList(ResUp(chans[i], chans[i + 1]) for i in range(4))
        self.final_conv = nn.Conv2d(width, cout, 1)
        self.last_act = last_act

    def channel_num(self, cin):
        return [cin] + self.widths + [self.final_conv.out_channels]        # style layer i sees channel_num[i] channels

    def forward(self, x):
        return self._run(x, {}, [])

    def apply_max_style(self, code, layers, idx):
        return self._run(code.detach().clone(), layers, idx)      # the reference detaches + clones the code (:600-602)

    def _run(self, x, layers, idx):
        if 0 in idx:
            x = layers["0"](x)
        for i, up in enumerate(self.ups):
            x = up(x)
            if i + 1 in idx:
                x = layers[str(i + 1)](x)
        x = self.final_conv(x)
        if self.last_act is not None:
            x = self.last_act(x)
        if 5 in idx:
            x = layers["5"](x)
        return x


class Encoder(nn.Module):
    def __init__(self, cin, width, cz):
        super().__init__()
        chans = [cin, width, 2 * width, 4 * width, cz]
        self.net = nn.Sequential(*[nn.Sequential(nn.Conv2d(chans[i], chans[i + 1], 3, stride=2, padding=1, bias=False),
                                                 nn.InstanceNorm2d(chans[i + 1]), nn.LeakyReLU(0.2)) for i in range(4)])

    def forward(self, x):
        return self.net(x)


class PortLayer(nn.Module):
    """The reference layer's eager op chain (oracle.torch_port.StylePort) as a module on the GPU."""

    def __init__(self, perm, gamma, beta, lmda):
        super().__init__()
        from oracle.torch_port import StylePort
        self.port = StylePort(perm, gamma.detach().cpu().reshape(gamma.shape[0], -1), beta.detach().cpu().reshape(beta.shape[0], -1),
                              lmda.detach().cpu().reshape(-1))
        dev = gamma.device
        self.gamma_noise = nn.Parameter(self.port.gamma_noise.detach().to(dev))
        self.beta_noise = nn.Parameter(self.port.beta_noise.detach().to(dev))
        self.lmda = nn.Parameter(self.port.lmda.detach().to(dev))
        self.port.gamma_noise, self.port.beta_noise, self.port.lmda = self.gamma_noise, self.beta_noise, self.lmda
        self.port.perm = self.port.perm        # CPU index tensor, like the reference (H2D copy per forward)

    def forward(self, x):
        return self.port.forward(x)


def build(width, size, batch, seed=0, dev="cuda"):
    torch.manual_seed(seed)
    cz = 8 * width
    enc = Encoder(1, width, cz).to(dev)
    dec = Decoder(cz, width, 1, last_act=torch.sigmoid).to(dev)
    seg = Decoder(cz, width, 4).to(dev)
    for m in (enc, dec, seg):
        for p in m.parameters():
            p.requires_grad_(False)
    g = torch.Generator().manual_seed(seed + 1)
    image = torch.rand(batch, 1, size, size, generator=g).to(dev)
    label = torch.randint(0, 4, (batch, size, size), generator=g).to(dev)
    return enc, dec, seg, image, label


def make_layers(kind, dec, cz, batch, idx, seed):
    """Three style layers from identical random state: ours first (draws the state), the port copies it."""
    from maxstyle_b200 import MaxStyle
    torch.manual_seed(seed)
    chans = dec.channel_num(cz)
    ours = {str(i): MaxStyle(batch, chans[i], p=1.0) for i in idx}
    if kind == "ours":
        return nn.ModuleDict(ours)
    return nn.ModuleDict({k: PortLayer(m.perm, m.gamma_noise, m.beta_noise, m.lmda) for k, m in ours.items()})


def inner_loop(kind, enc, dec, seg, image, label, idx, n_iter, lr, seed, fused=True, timers=None, ce=TF.cross_entropy):
    """generate_max_style_image (model:458-571): n_iter+1 decoder passes, n_iter encoder+segmentation passes with
    loss = -CE, backward into the style parameters only, Adam(lr) step."""
    code = enc(image).detach()
    layers = make_layers(kind, dec, code.shape[1], image.shape[0], idx, seed)
    if kind == "ours" and fused:
        from maxstyle_b200 import FusedStyleOptimizer
        opt = FusedStyleOptimizer(layers.values(), lr=lr)
    else:
        opt = torch.optim.Adam(layers.parameters(), lr=lr)
    recon = None
    for i in range(n_iter + 1):
        layers.zero_grad()
        if i > 0:
            opt.zero_grad()
            p = seg(enc(recon))
            loss = -ce(p, label)
            loss.backward()
            opt.step()
            layers.zero_grad()
        recon = dec.apply_max_style(code, layers, idx)
    return recon.detach().clone(), layers


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def layer_only_ms(kind, dec, cz, batch, size, idx, n_iter, reps, seed):
    """Time of the style layers alone on the loop's activation shapes: (n_iter+1) forwards and n_iter backward+step."""
    layers = make_layers(kind, dec, cz, batch, idx, seed)
    chans = dec.channel_num(cz)
    sizes = {i: size // 2 ** max(0, 4 - i) for i in range(6)}
    xs = {k: torch.rand(batch, chans[int(k)], sizes[int(k)], sizes[int(k)], device="cuda").requires_grad_(True) for k in layers}
    dys = {k: torch.randn_like(v) for k, v in xs.items()}
    opt = torch.optim.Adam(layers.parameters(), lr=0.1)

    def run():
        for i in range(n_iter + 1):
            for k, m in layers.items():
                y = m(xs[k])
                if i > 0:
                    y.backward(dys[k])
            if i > 0:
                opt.step(); opt.zero_grad()
    return timed(run, reps)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=64, help="decoder base width: 64 = FCN_64 (style layers see 64, 64, 1 channels), 16 = FCN_16")
    ap.add_argument("--batch", type=int, default=20)
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--n-iter", type=int, default=5)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--our-ce", action="store_true", help="replacement arms also use maxstyle_b200.cross_entropy_2D (SURVEY 8f-4) "
                    "instead of torch's fused F.cross_entropy; the reference arm then uses the reference's op chain for the loss")
    args = ap.parse_args()
    idx = [3, 4, 5]
    enc, dec, seg, image, label = build(args.width, args.size, args.batch)
    out = {"config": f"BASELINE config 2 stand-in: width {args.width}, batch {args.batch}, {args.size}x{args.size}, layers {idx}, n_iter {args.n_iter}",
           "style_layer_shapes": [[args.batch, dec.channel_num(8 * args.width)[i], args.size // 2 ** max(0, 4 - i), args.size // 2 ** max(0, 4 - i)] for i in idx]}
    ce_ours, ce_ref = TF.cross_entropy, TF.cross_entropy
    if args.our_ce:
        from maxstyle_b200 import cross_entropy_2D as ce_ours

        def ce_ref(x, t):                      # custom_loss.py:1058-1078 as eager ops
            n, c, h, w = x.shape
            lp = TF.log_softmax(x, dim=1).transpose(1, 2).transpose(2, 3).contiguous().view(-1, c)
            mask = torch.ones(n, 1, h, w, device=x.device).reshape(n * h * w, 1)
            return torch.sum(TF.nll_loss(lp, t.view(-1), reduction="none") * mask.flatten()) / float(mask.numel())
        out["loss"] = "reference op chain (port arm) vs maxstyle_b200.cross_entropy_2D (replacement arms)"
    rec = {}
    for kind in ("port", "ours"):
        ce = ce_ref if kind == "port" else ce_ours
        r, layers = inner_loop(kind, enc, dec, seg, image, label, idx, args.n_iter, 0.1, seed=7, ce=ce)
        rec[kind] = (r, {k: [p.detach().clone() for p in (m.gamma_noise, m.beta_noise, m.lmda)] for k, m in layers.items()})
        out[f"loop_ms_{kind}"] = round(timed(lambda: inner_loop(kind, enc, dec, seg, image, label, idx, args.n_iter, 0.1, seed=7, ce=ce), args.reps), 2)
        out[f"layers_only_ms_{kind}"] = round(layer_only_ms(kind, dec, 8 * args.width, args.batch, args.size, idx, args.n_iter, args.reps, seed=7), 2)
    # the same loop through the CUDA-graphed executor (maxstyle_b200.StyleLoopExecutor)
    from maxstyle_b200 import StyleLoopExecutor
    code = enc(image).detach()
    chans = dec.channel_num(code.shape[1])
    ex = StyleLoopExecutor(lambda cd, layers: dec.apply_max_style(cd, layers, idx),
                           lambda img, lab: -ce_ours(seg(enc(img)), lab),
                           args.batch, {i: chans[i] for i in idx}, n_iter=args.n_iter, lr=0.1, p=1.0)

    def graphed():
        torch.manual_seed(7)
        return ex.run(enc(image).detach(), label)
    r_ex = graphed()
    out["loop_ms_executor"] = round(timed(graphed, args.reps), 2)
    out["executor_vs_eager_ours_rel_diff"] = float((r_ex - rec["ours"][0]).abs().max() / rec["ours"][0].abs().max())
    out["loop_speedup_executor"] = round(out["loop_ms_port"] / out["loop_ms_executor"], 3)
    a, b = rec["ours"][0], rec["port"][0]
    out["recon_max_abs_diff"] = float((a - b).abs().max())
    out["recon_rel_diff"] = float((a - b).abs().max() / b.abs().max())
    out["note"] = ("recon after n_iter Adam steps differs where a gradient component is rounding noise (Adam turns it into a "
                   "+-lr step either way); first-pass output and first-iteration gradients are compared in tests/test_gpu_loop.py")
    r0a, _ = inner_loop("ours", enc, dec, seg, image, label, idx, 0, 0.1, seed=7, ce=ce_ours)
    r0b, _ = inner_loop("port", enc, dec, seg, image, label, idx, 0, 0.1, seed=7, ce=ce_ref)
    out["first_pass_rel_diff"] = float((r0a - r0b).abs().max() / r0b.abs().max())
    out["loop_speedup"] = round(out["loop_ms_port"] / out["loop_ms_ours"], 3)
    out["layers_speedup"] = round(out["layers_only_ms_port"] / out["layers_only_ms_ours"], 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
