"""Training path of the generator: one torch.autograd.Function around the fused forward and the backward
kernels of libhwg_b200 (reference: the autograd graph PyTorch records for model/pure_gen.py:42-50,202-216).

The style vector path (PixelNorm, six Linear+LeakyReLU, the ten AdaIN projections — [B,128] matrices) stays in
torch autograd (plain library GEMMs on tiny operands); its outputs `s` and `gb` are inputs of the Function,
which returns their gradients.  Everything that touches an image-sized tensor runs on the library:
  forward   tcgen05 convs with fused bias/noise/LeakyReLU/statistics, blur, AdaIN apply, output head
  backward  output head, AdaIN+LeakyReLU(+noise-weight) backward in two passes, blur (self-adjoint),
            dgrad = hwg_conv_fprop on the gradient (strided taps for the up-sampling convs),
            wgrad = hwg_conv_wgrad (per output phase for the up-sampling convs).
The weight re-parameterisations (EqualLR scales, FusedUpsample's 4x4 averaged kernel, the row-parity sums of
the nearest-upsample convs) are tiny and linear; their adjoints are applied on the host side.
"""
from math import sqrt

import torch
import torch.nn.functional as F

from . import conv, ops
from ._lib import ACT_LRELU, ACT_NONE
from .pure_gen import TAPS3x3, conv1_forward

_SEL = {0: [(0, 1), (-1, 3)], 1: [(1, 0), (0, 2)]}  # FusedUpsample: output parity -> [(input offset, kernel index)]


def _vert_src(par, kh):
    return (par + kh - 1) // 2


# ----------------------------------------------------------------------------------------------------------
def _param_list(m):
    """Conv-side parameters in a fixed order (inputs of the Function after content, s, gb)."""
    ps = []
    for blk in m.conv:
        c1 = blk.conv1 if blk.kind in ("initial", "plain") else (blk.conv1[1] if blk.kind == "vert_up" else blk.conv1[0])
        ps += [c1.weight, c1.bias, blk.noise1.weight_orig, blk.conv2.weight, blk.conv2.bias, blk.noise2.weight_orig]
    ps += [m.out[0].conv.weight_orig, m.out[0].conv.bias]
    return ps


def _style_path(m, style):
    """torch-autograd part: style -> s [B,S] and gb [B, sum 2C] (pure_gen.py:31-39,46,57,63)."""
    s = style.float()
    s = s / torch.sqrt((s * s).mean(1, keepdim=True) + 1e-8)
    for mod in m.style_emb:
        if isinstance(mod, torch.nn.Linear):
            s = F.leaky_relu(F.linear(s, mod.weight, mod.bias), 0.2)
        elif isinstance(mod, torch.nn.Dropout) and m.training:
            s = F.dropout(s, mod.p, True)
    ws, bs = [], []
    for blk in m.conv:
        for ad in (blk.adain1, blk.adain2):
            ws.append(ad.style.weight)
            bs.append(ad.style.bias)
    gb = F.linear(s, torch.cat(ws, 0), torch.cat(bs, 0))
    return s, gb


def forward_train(m, content, s, gb, noise):
    """CUDA forward that keeps what the backward needs.  Returns (image, ctx)."""
    c = m._packed()
    T, B, ncls = content.shape
    dev = content.device
    x = ops.gen_pack_input(content.float(), s if m.append_style else None, c["cin_pad"])
    seed, seed_dev = None, None
    if noise is None:
        seed, seed_dev = m._noise_seed()
    else:
        noise = [z.permute(0, 2, 3, 1).contiguous().float() for z in noise]
    gbs = gb.stride(0)
    recs = []
    off, k = 0, 0
    H, W = 1, T
    out = None
    nblk = len(c["blocks"])
    for bi, e in enumerate(c["blocks"]):
        C = e["C"]
        st = torch.zeros((B, C, 2), device=dev, dtype=torch.float32)
        nz = None if noise is None else noise[k]
        x_in, Hin, Win = x, H, W
        a, Ho, Wo = conv1_forward(x, e, B, H, W, st, nz, k, seed, seed_dev)
        H, W = Ho, Wo
        coef, save = ops.adain_coeffs(st, gb[:, off:], gb[:, off + C:], gbs, B, C, H * W, save=True)
        recs.append(dict(x_in=x_in, Hin=Hin, Win=Win, a=a, coef=coef, save=save, nz=nz, subseq=16 * k, off=off))
        x = ops.scale_shift_act(a, coef, True, out=torch.empty_like(a))
        off += 2 * C
        k += 1
        st = torch.zeros((B, C, 2), device=dev, dtype=torch.float32)
        nz = None if noise is None else noise[k]
        a = conv.conv_fprop(x, e["w2"], TAPS3x3, H, W, bias=e["b2"], act=ACT_LRELU, slope=0.2, noise=nz,
                            noise_w=e["nw2"], noise_seed=seed, noise_subseq=16 * k, noise_seed_dev=seed_dev, stats=st)
        coef, save = ops.adain_coeffs(st, gb[:, off:], gb[:, off + C:], gbs, B, C, H * W, save=True)
        recs.append(dict(x_in=x, Hin=H, Win=W, a=a, coef=coef, save=save, nz=nz, subseq=16 * k, off=off))
        if bi == nblk - 1:
            out = ops.gen_output(a, coef, c["w_out"], c["b_out"])
        else:
            x = ops.scale_shift_act(a, coef, True, out=torch.empty_like(a))
        off += 2 * C
        k += 1
    ctx = dict(recs=recs, seed=seed or 0, seed_dev=seed_dev, out=out, T=T, B=B, ncls=ncls, gb_width=gb.size(1))
    return out, ctx


# ----------------------------------------------------------------------------------------------------------
def _taps_f32(w4d):
    co, ci, kh, kw = w4d.shape
    return w4d.detach().float().permute(2, 3, 0, 1).reshape(kh * kw, co, ci).contiguous()


def _w4(dw, kh, kw):
    t, co, ci = dw.shape
    return dw.view(kh, kw, co, ci).permute(2, 3, 0, 1).contiguous()


def _fused_w4(mod, w):
    wp = F.pad(w * mod.multiplier, [1, 1, 1, 1])
    return (wp[:, :, 1:, 1:] + wp[:, :, :-1, 1:] + wp[:, :, 1:, :-1] + wp[:, :, :-1, :-1]) / 4


def backward_train(m, ctx, g_out):
    """Returns (g_content [T,B,ncls], g_s [B,S], g_gb [B,sum 2C], [conv-side parameter grads in _param_list order])."""
    recs, seed, seed_dev = ctx["recs"], ctx["seed"], ctx["seed_dev"]
    c = m._packed()
    B, T, ncls = ctx["B"], ctx["T"], ctx["ncls"]
    dev = g_out.device
    g_gb = torch.zeros((B, ctx["gb_width"]), device=dev, dtype=torch.float32)
    pgrads = []
    # ---- output head
    last = recs[-1]
    g, dw_out, db0 = ops.gen_output_bwd(g_out.contiguous().float(), ctx["out"], last["a"], last["coef"], c["w_out"])
    w_orig = m.out[0].conv.weight_orig
    g_wout = (dw_out * sqrt(2.0 / (w_orig.size(1) * w_orig[0][0].numel()))).view_as(w_orig)
    g_bout = db0.view(1)
    blocks = list(m.conv)
    for bi in range(len(blocks) - 1, -1, -1):
        blk, e = blocks[bi], c["blocks"][bi]
        C = e["C"]
        r1, r2 = recs[2 * bi], recs[2 * bi + 1]
        # ---------------- second half: conv2 + noise2 + lrelu + adain2
        gy, dgam, dbet, dbias2, dnw2 = ops.adain_lrelu_bwd(g, r2["a"], r2["save"], r2["coef"], 0.2, r2["nz"], seed,
                                                           r2["subseq"], seed_dev=seed_dev)
        g_gb[:, r2["off"]:r2["off"] + C] = dgam
        g_gb[:, r2["off"] + C:r2["off"] + 2 * C] = dbet
        g_w2 = _w4(conv.conv_wgrad(r2["x_in"], gy, TAPS3x3, C, C), 3, 3)
        wd, tapsd = conv.dgrad_pack(_taps_f32(blk.conv2.weight), TAPS3x3)
        g = conv.conv_fprop(gy, wd, tapsd, r2["Hin"], r2["Win"])
        g_nw2 = (dnw2 * sqrt(2.0 / C)).view(1, C, 1, 1)
        # ---------------- first half: conv1 (+blur) + noise1 + lrelu + adain1
        gy, dgam, dbet, dsum, dnw1 = ops.adain_lrelu_bwd(g, r1["a"], r1["save"], r1["coef"], 0.2, r1["nz"], seed,
                                                         r1["subseq"], row_subseq=(e["kind"] == "initial"),
                                                         seed_dev=seed_dev)
        g_gb[:, r1["off"]:r1["off"] + C] = dgam
        g_gb[:, r1["off"] + C:r1["off"] + 2 * C] = dbet
        g_nw1 = (dnw1 * sqrt(2.0 / C)).view(1, C, 1, 1)
        x_in, Hin, Win = r1["x_in"], r1["Hin"], r1["Win"]
        kind = e["kind"]
        if kind in ("vert_up", "fused_up"):
            # Blur is self-adjoint (symmetric stencil, zero padding); its per-channel sums are the conv bias gradient
            st = torch.zeros((B, C, 2), device=dev, dtype=torch.float32)
            gy = ops.blur_noise_act_stats(gy, None, None, st, ACT_NONE, 0.0)
            g_b1 = st[:, :, 0].sum(0)
        else:
            g_b1 = dsum
        if kind == "plain":
            Cin = blk.in_channel
            g_w1 = _w4(conv.conv_wgrad(x_in, gy, TAPS3x3, Cin, C), 3, 3)
            wd, tapsd = conv.dgrad_pack(_taps_f32(blk.conv1.weight), TAPS3x3)
            g = conv.conv_fprop(gy, wd, tapsd, Hin, Win)
        elif kind == "initial":
            w = blk.conv1.weight                      # [Cin, Cout, 4, 3]
            Cin, cin_pad = w.size(0), c["cin_pad"]
            g_w1 = torch.empty_like(w, dtype=torch.float32)
            for r in range(4):
                dw = conv.conv_wgrad(x_in, gy, e["taps1"], cin_pad, C, grid=(1, Win), gy_offset=(r, 0))  # [3,C,cin_pad]
                g_w1[:, :, r, :] = dw[:, :, :Cin].permute(2, 1, 0)
            # gradient w.r.t. the packed input: 12-tap convolution of gy [B,4,T,C] -> [B,1,T,Cin8]
            cin8 = ((Cin + 15) // 16) * 16
            mats = [torch.nn.functional.pad(w.detach().float()[:, :, r, kx], (0, 0, 0, cin8 - Cin))
                    for r in range(4) for kx in range(3)]
            taps = [(r, kx - 1) for r in range(4) for kx in range(3)]
            g = conv.conv_fprop(gy, conv.pack_taps(mats), taps, 1, Win)
        elif kind == "vert_up":
            w = blk.conv1[1].weight                   # [Cout, Cin, 3, 3]
            Cin = w.size(1)
            g_w1 = torch.zeros_like(w, dtype=torch.float32)
            for par, (taps, _) in enumerate(e["w1"]):
                dw = conv.conv_wgrad(x_in, gy, taps, Cin, C, grid=(Hin, Win), gy_stride=(2, 1), gy_offset=(par, 0))
                dhs = sorted({t[0] for t in taps})
                for kh in range(3):
                    j = dhs.index(_vert_src(par, kh))
                    g_w1[:, :, kh, :] += dw[3 * j:3 * j + 3].permute(1, 2, 0)
            wf = w.detach().float()
            comb = {-1: [2], 0: [1, 2], 1: [0, 1], 2: [0]}   # dh -> kh of the rows that land there (r - kh + 1 = dh)
            mats, taps = [], []
            for dh in (-1, 0, 1, 2):
                wk = sum(wf[:, :, kh, :] for kh in comb[dh])  # [Cout, Cin, 3]
                for kw in range(3):
                    taps.append((dh, 1 - kw))
                    mats.append(wk[:, :, kw].t())
            g = conv.conv_fprop(gy, conv.pack_taps(mats), taps, Hin, Win, in_stride=(2, 1))
        else:  # fused_up
            mod = blk.conv1[0]
            wl = mod.weight.detach().float().requires_grad_()
            with torch.enable_grad():
                w4 = _fused_w4(mod, wl)                # [Cin, Cout, 4, 4]
            Cin = wl.size(0)
            dw4 = torch.empty_like(w4)
            if Cin <= 32 and C <= 32:
                # all four parities in one launch (per-tap gy phase): gy and x are read once
                taps, phases, idx = [], [], []
                for py in (0, 1):
                    for px in (0, 1):
                        for dh, ky in _SEL[py]:
                            for dw_, kx in _SEL[px]:
                                taps.append((dh, dw_))
                                phases.append((py, px))
                                idx.append((ky, kx))
                dw = conv.conv_wgrad(x_in, gy, taps, Cin, C, grid=(Hin, Win), gy_stride=(2, 2), tap_phase=phases)
                # taps are ordered [py][px][a][b] with ky = 2a + (1-py), kx = 2b + (1-px)  (see _SEL)
                dw4 = dw.view(2, 2, 2, 2, C, Cin).flip(0, 1).permute(5, 4, 2, 0, 3, 1).reshape(Cin, C, 4, 4)
            else:
                for py in (0, 1):
                    for px in (0, 1):
                        taps, idx = [], []
                        for dh, ky in _SEL[py]:
                            for dw_, kx in _SEL[px]:
                                taps.append((dh, dw_))
                                idx.append((ky, kx))
                        dw = conv.conv_wgrad(x_in, gy, taps, Cin, C, grid=(Hin, Win), gy_stride=(2, 2), gy_offset=(py, px))
                        for t, (ky, kx) in enumerate(idx):
                            dw4[:, :, ky, kx] = dw[t].t()
            (g_w1,) = torch.autograd.grad(w4, wl, dw4)
            w4d = w4.detach()
            taps = [(ky - 1, kx - 1) for ky in range(4) for kx in range(4)]
            mats = [w4d[:, :, ky, kx] for ky in range(4) for kx in range(4)]   # [out=Cin][in=Cout]
            g = conv.conv_fprop(gy, conv.pack_taps(mats), taps, Hin, Win, in_stride=(2, 2))
        pgrads.insert(0, [g_w1, g_b1, g_nw1, g_w2, dbias2, g_nw2])
    # ---- packed input -> content and style
    S = m.style_size if m.append_style else 0
    gx0 = g[:, 0].float()                                # [B, T, Cin8]
    g_content = gx0[:, :, :ncls].permute(1, 0, 2).contiguous()
    g_s = gx0[:, :, ncls:ncls + S].sum(1) if S else None
    flat = [t for grp in pgrads for t in grp] + [g_wout, g_bout]
    return g_content, g_s, g_gb, flat


class _GenFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, noise, content, s, gb, *params):
        with torch.no_grad():
            out, saved = forward_train(module, content, s, gb, noise)
        ctx.module, ctx.saved = module, saved
        return out

    @staticmethod
    def backward(ctx, g):
        with torch.no_grad():
            g_content, g_s, g_gb, flat = backward_train(ctx.module, ctx.saved, g)
        ctx.saved = None
        return (None, None, g_content, g_s, g_gb) + tuple(flat)


def generator_apply(module, content, style, noise):
    s, gb = _style_path(module, style)
    return _GenFn.apply(module, noise, content, s, gb, *_param_list(module))
