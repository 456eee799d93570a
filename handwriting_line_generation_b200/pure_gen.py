"""Drop-in for the reference generator `SpacedGenerator` (model/pure_gen.py:12-50).

Same constructor signature, attribute names, construction order (so the same seed gives
the same random init) and `state_dict` keys/shapes as the reference; `forward` runs the
whole stack on libhwg_b200: tcgen05 implicit-GEMM convolutions with fused
bias/noise/LeakyReLU/InstanceNorm-statistics epilogues, and 16-byte-vector NHWC bf16 passes
for blur / AdaIN / output.  The torch modules held here are parameter containers only —
their own forward is never called.
"""
from math import sqrt

import torch
import torch.nn as nn

from . import _lib, conv, ops
from ._lib import ACT_LRELU, ACT_NONE


# ----------------------------------------------------------------------------------------------
# parameter containers with the reference's names (pure_gen.py:52-311)
# ----------------------------------------------------------------------------------------------
class PixelNorm(nn.Module):  # pure_gen.py:306
    pass


class AdaptiveInstanceNorm(nn.Module):  # pure_gen.py:52-69
    def __init__(self, in_channel, style_dim):
        super().__init__()
        self.norm = nn.InstanceNorm2d(in_channel)
        self.style = nn.Linear(style_dim, in_channel * 2)
        self.style.bias.data[:in_channel] = 1
        self.style.bias.data[in_channel:] = 0


class NoiseInjection(nn.Module):  # pure_gen.py:72-79, wrapped by equal_lr -> weight_orig
    def __init__(self, channel):
        super().__init__()
        self.weight_orig = nn.Parameter(torch.ones(1, channel, 1, 1) * 0.01)

    def effective_weight(self):
        c = self.weight_orig.size(1)  # EqualLR fan_in = size(1) * numel(weight[0][0]) = C
        return (self.weight_orig * sqrt(2.0 / c)).reshape(-1)


class Blur(nn.Module):  # pure_gen.py:119-137: buffers only
    def __init__(self, channel):
        super().__init__()
        weight = torch.tensor([[1, 2, 1], [2, 4, 2], [1, 2, 1]], dtype=torch.float32).view(1, 1, 3, 3)
        weight = weight / weight.sum()
        self.register_buffer("weight", weight.repeat(channel, 1, 1, 1))
        self.register_buffer("weight_flip", torch.flip(weight, [2, 3]).repeat(channel, 1, 1, 1))


class FusedUpsample(nn.Module):  # pure_gen.py:250-279
    def __init__(self, in_channel, out_channel, kernel_size, padding=0, only_vertical=False):
        super().__init__()
        if only_vertical:
            raise NotImplementedError("FusedUpsample(only_vertical=True) is not used by SpacedGenerator")
        self.stride = 2
        self.weight = nn.Parameter(torch.randn(in_channel, out_channel, kernel_size, kernel_size))
        self.bias = nn.Parameter(torch.zeros(out_channel))
        self.multiplier = sqrt(2 / (in_channel * kernel_size * kernel_size))
        self.pad = padding


class EqualConv2d(nn.Module):  # pure_gen.py:281-291 (equal_lr -> conv.weight_orig)
    def __init__(self, *args, **kwargs):
        super().__init__()
        conv_ = nn.Conv2d(*args, **kwargs)
        conv_.weight.data.normal_()
        conv_.bias.data.zero_()
        w = conv_.weight
        del conv_._parameters["weight"]
        conv_.register_parameter("weight_orig", nn.Parameter(w.data))
        self.conv = conv_

    def effective_weight(self):
        w = self.conv.weight_orig
        return w * sqrt(2.0 / (w.size(1) * w[0][0].numel()))


class StyledConvBlock(nn.Module):  # pure_gen.py:140-216
    def __init__(self, in_channel, out_channel, kernel_size=3, padding=1, style_dim=512, initial=False,
                 upsample=False, only_vertical=False, fused=False):
        super().__init__()
        if kernel_size != 3 or padding != 1:
            raise NotImplementedError("only the 3x3/pad 1 blocks SpacedGenerator builds are implemented")
        if initial == "1d":
            raise NotImplementedError("initial='1d' is not used by SpacedGenerator")
        if initial:
            self.kind = "initial"
            self.conv1 = nn.ConvTranspose2d(in_channel, out_channel, (4, 3), padding=(0, 1))
        elif upsample and fused:
            self.kind = "fused_up"
            self.conv1 = nn.Sequential(
                FusedUpsample(in_channel, out_channel, kernel_size, padding=padding, only_vertical=only_vertical),
                Blur(out_channel))
        elif upsample:
            if not only_vertical:
                raise NotImplementedError("nearest (2,2) upsampling is not used by SpacedGenerator")
            self.kind = "vert_up"
            self.conv1 = nn.Sequential(nn.Upsample(scale_factor=(2, 1), mode="nearest"),
                                       nn.Conv2d(in_channel, out_channel, kernel_size, padding=padding),
                                       Blur(out_channel))
        else:
            self.kind = "plain"
            self.conv1 = nn.Conv2d(in_channel, out_channel, kernel_size, padding=padding)
        self.noise1 = NoiseInjection(out_channel)
        self.adain1 = AdaptiveInstanceNorm(out_channel, style_dim)
        self.lrelu1 = nn.LeakyReLU(0.2)
        self.conv2 = nn.Conv2d(out_channel, out_channel, kernel_size, padding=padding)
        self.noise2 = NoiseInjection(out_channel)
        self.adain2 = AdaptiveInstanceNorm(out_channel, style_dim)
        self.lrelu2 = nn.LeakyReLU(0.2)
        self.in_channel, self.out_channel = in_channel, out_channel


# ----------------------------------------------------------------------------------------------
# weight re-layout (cached per parameter version)
# ----------------------------------------------------------------------------------------------
TAPS3x3 = conv.conv_taps(3, 3, 1, 1)


def _pack_initial(w, cin_pad):
    """ConvTranspose2d(Cin,Cout,(4,3),pad(0,1)) on H=1: out row r is a 1x3 correlation with flipped kx
    (pure_gen.py:161-163): out[r,x] = sum_kx in[x+1-kx] W[:, :, r, kx]."""
    taps = [(0, 1 - kx) for kx in range(3)]
    packs = [conv.pack_taps([w[:, :, r, kx].t() for kx in range(3)], cin_pad) for r in range(4)]
    return taps, packs


def _pack_vert_up(w):
    """nearest (2,1) upsample + 3x3 conv = two row-parity convolutions on the un-upsampled input with
    the taps that hit the same source row summed (pure_gen.py:176-186)."""
    out = []
    for par in (0, 1):
        rows = {}
        for kh in range(3):
            src = (par + kh - 1) // 2
            rows[src] = rows[src] + w[:, :, kh, :] if src in rows else w[:, :, kh, :]
        taps, mats = [], []
        for dh in sorted(rows):
            for kw in range(3):
                taps.append((dh, kw - 1))
                mats.append(rows[dh][:, :, kw])
        out.append((taps, conv.pack_taps(mats)))
    return out


def _pack_fused_up(mod):
    """FusedUpsample (pure_gen.py:259-279): 4x4 averaged kernel, conv_transpose2d stride 2 pad 1 =
    four output-parity 2x2 convolutions."""
    w = torch.nn.functional.pad(mod.weight * mod.multiplier, [1, 1, 1, 1])
    w4 = (w[:, :, 1:, 1:] + w[:, :, :-1, 1:] + w[:, :, 1:, :-1] + w[:, :, :-1, :-1]) / 4  # [Cin,Cout,4,4]
    sel = {0: [(0, 1), (-1, 3)], 1: [(1, 0), (0, 2)]}  # parity -> [(input offset, kernel index)]
    out = []
    for py in (0, 1):
        for px in (0, 1):
            taps, mats = [], []
            for dh, ky in sel[py]:
                for dw, kx in sel[px]:
                    taps.append((dh, dw))
                    mats.append(w4[:, :, ky, kx].t())
            out.append((py, px, taps, conv.pack_taps(mats)))
    return out


TAPS_UNION = [(dh, dw) for dh in (-1, 0, 1) for dw in (-1, 0, 1)]


def _pack_fused_up_folded(mod):
    """The four output parities of FusedUpsample as ONE launch: 9 union taps, Cout = 4 folds x C (fold = 2*py+px);
    (tap, parity) pairs the parity does not use carry zero weights.  The operand tile of a tap is then loaded once
    for all four parities (9 loads per tile instead of 16)."""
    w = torch.nn.functional.pad(mod.weight * mod.multiplier, [1, 1, 1, 1])
    w4 = (w[:, :, 1:, 1:] + w[:, :, :-1, 1:] + w[:, :, 1:, :-1] + w[:, :, :-1, :-1]) / 4  # [Cin,Cout,4,4]
    cin, cout = w4.shape[:2]
    sel = {0: {0: 1, -1: 3}, 1: {1: 0, 0: 2}}  # parity -> {input offset: kernel index}
    mats = []
    for dh, dw in TAPS_UNION:
        rows = []
        for py in (0, 1):
            for px in (0, 1):
                if dh in sel[py] and dw in sel[px]:
                    rows.append(w4[:, :, sel[py][dh], sel[px][dw]].t())
                else:
                    rows.append(torch.zeros((cout, cin), device=w4.device, dtype=w4.dtype))
        mats.append(torch.cat(rows, 0))
    return conv.pack_taps(mats)


def conv1_forward(x, e, B, H, W, st, nz, k, seed, seed_dev):
    """First convolution of a StyledConvBlock (+ Blur) + noise + LeakyReLU + statistics (pure_gen.py:205-208).
    x [B,H,W,Cin] bf16 -> (a [B,Ho,Wo,C] bf16, Ho, Wo); shared by the inference and the training forward."""
    C, dev = e["C"], x.device
    if e["kind"] == "initial":
        Ho, Wo = 4, W
        a = torch.empty((B, Ho, Wo, C), device=dev, dtype=torch.bfloat16)
        # the four output rows are four channel folds of one launch (same input, same taps)
        conv.conv_fprop(x, e["w1f"], e["taps1"], 1, Wo, bias=e["b1f"], act=ACT_LRELU, slope=0.2,
                        out_view=(a, Ho * Wo * C, Wo * C, C, 0), fold=(C, 1, Wo * C, 0),
                        noise_view=None if nz is None else (nz, Ho * Wo * C, Wo * C, C, 0),
                        noise_w=e["nw1f"], noise_seed=seed, noise_subseq=16 * k, noise_seed_dev=seed_dev, stats=st)
        return a, Ho, Wo
    if e["kind"] == "plain":
        a = conv.conv_fprop(x, e["w1"], TAPS3x3, H, W, bias=e["b1"], act=ACT_LRELU, slope=0.2, noise=nz,
                            noise_w=e["nw1"], noise_seed=seed, noise_subseq=16 * k, noise_seed_dev=seed_dev, stats=st)
        return a, H, W
    if e["kind"] == "vert_up":
        Ho, Wo = 2 * H, W
        raw = torch.empty((B, Ho, Wo, C), device=dev, dtype=torch.bfloat16)
        for par, (taps, wp) in enumerate(e["w1"]):
            conv.conv_fprop(x, wp, taps, H, W, bias=e["b1"], out_view=(raw, Ho * Wo * C, 2 * Wo * C, C, par * Wo * C))
    else:
        Ho, Wo = 2 * H, 2 * W
        raw = torch.empty((B, Ho, Wo, C), device=dev, dtype=torch.bfloat16)
        # the four output parities are four folds of one launch, four taps each
        conv.conv_fprop(x, e["w1f"], e["taps1f"], H, W, bias=e["b1f"], out_view=(raw, Ho * Wo * C, 2 * Wo * C, 2 * C, 0),
                        fold=(C, 2, Wo * C, C), fold_taps=4)
    a = ops.blur_noise_act_stats(raw, nz, e["nw1"], st, ACT_LRELU, 0.2, seed or 0, 16 * k, seed_dev)
    return a, Ho, Wo


class SpacedGenerator(nn.Module):
    """model/pure_gen.py:12-50.  content [T,B,n_class] + style [B,style_size] -> [B,1,64,4T]."""

    def __init__(self, n_class, style_size, dim=256, output_dim=1, n_style_trans=6, emb_dropout=False,
                 append_style=False, small=False):
        super().__init__()
        if output_dim != 1:
            raise NotImplementedError("output_dim != 1")
        if dim % 256 != 0:
            raise NotImplementedError("dim must be a multiple of 256 (channel counts are multiples of 16)")
        fused = True
        self.append_style = append_style
        in_ch = n_class + style_size if append_style else n_class
        self.conv = nn.Sequential(
            StyledConvBlock(in_ch, dim, upsample=False, style_dim=style_size, initial=True),
            StyledConvBlock(dim, dim // 2, upsample=True, only_vertical=True, fused=False, style_dim=style_size),
            StyledConvBlock(dim // 2, dim // 4, upsample=True, only_vertical=True, fused=False, style_dim=style_size),
            StyledConvBlock(dim // 4, dim // 8, upsample=True, only_vertical=False, fused=fused, style_dim=style_size),
            StyledConvBlock(dim // 8, dim // 16, upsample=not small, only_vertical=False, fused=fused,
                            style_dim=style_size),
        )
        self.out = nn.Sequential(EqualConv2d(dim // 16, output_dim, 1), nn.Tanh())
        layers = [PixelNorm()]
        drop = emb_dropout if type(emb_dropout) is float else 0.5
        for i in range(n_style_trans):
            layers.append(nn.Linear(style_size, style_size))
            if emb_dropout and i < n_style_trans - 1:
                layers.append(nn.Dropout(drop, True))
            layers.append(nn.LeakyReLU(0.2, True))
        self.style_emb = nn.Sequential(*layers)
        self.gen = self.conv
        self.in_ch, self.n_class, self.style_size = in_ch, n_class, style_size
        self._cache_key, self._cache = None, None
        # CUDA-graph mode (graphs.py): a device-side counter is added to the (captured, hence constant) host seed
        # so that every replay draws fresh noise; off by default so that torch.manual_seed reproduces a call
        self.register_buffer("_noise_step", torch.zeros(1, dtype=torch.int64), persistent=False)
        self.device_noise_counter = False

    def _noise_seed(self):
        """(host seed, device counter snapshot or None) for one forward."""
        seed = int(torch.empty((), dtype=torch.int64).random_().item()) & ((1 << 62) - 1)
        if not self.device_noise_counter:
            return seed, None
        self._noise_step += 1
        return seed, self._noise_step.clone()

    # -- derived weights ------------------------------------------------------------------------
    def _packed(self):
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._cache_key == key:
            return self._cache
        with torch.no_grad():
            c = {"blocks": []}
            cin_pad = ((self.in_ch + 63) // 64) * 64
            c["cin_pad"] = cin_pad
            gb_w, gb_b = [], []
            for blk in self.conv:
                e = {"kind": blk.kind, "C": blk.out_channel}
                if blk.kind == "initial":
                    e["taps1"], e["w1"] = _pack_initial(blk.conv1.weight, cin_pad)
                    e["w1f"] = torch.cat(e["w1"], 1).contiguous()      # [3, 4*C, cin_pad]: rows as channel folds
                    e["b1"] = blk.conv1.bias.detach().float().contiguous()
                    e["b1f"] = e["b1"].repeat(4)
                elif blk.kind == "vert_up":
                    e["w1"] = _pack_vert_up(blk.conv1[1].weight)
                    e["b1"] = blk.conv1[1].bias.detach().float().contiguous()
                elif blk.kind == "fused_up":
                    e["w1"] = _pack_fused_up(blk.conv1[0])
                    e["w1f"] = torch.cat([wp for _, _, _, wp in e["w1"]], 0).contiguous()   # [16, C, Cin], parity-major
                    e["taps1f"] = [t for _, _, taps, _ in e["w1"] for t in taps]
                    e["b1"] = blk.conv1[0].bias.detach().float().contiguous()
                    e["b1f"] = e["b1"].repeat(4)
                else:
                    e["w1"] = conv.pack_conv2d_weight(blk.conv1.weight)
                    e["b1"] = blk.conv1.bias.detach().float().contiguous()
                e["w2"] = conv.pack_conv2d_weight(blk.conv2.weight)
                e["b2"] = blk.conv2.bias.detach().float().contiguous()
                e["nw1"] = blk.noise1.effective_weight().detach().float().contiguous()
                e["nw2"] = blk.noise2.effective_weight().detach().float().contiguous()
                if blk.kind == "initial":
                    e["nw1f"] = e["nw1"].repeat(4)
                for ad in (blk.adain1, blk.adain2):
                    gb_w.append(ad.style.weight)
                    gb_b.append(ad.style.bias)
                c["blocks"].append(e)
            c["gb_w"] = torch.cat(gb_w, 0).detach().float().contiguous()
            c["gb_b"] = torch.cat(gb_b, 0).detach().float().contiguous()
            c["mlp"] = [(m.weight.detach().float().contiguous(), m.bias.detach().float().contiguous())
                        for m in self.style_emb if isinstance(m, nn.Linear)]
            c["w_out"] = self.out[0].effective_weight().detach().float().reshape(-1).contiguous()
            c["b_out"] = self.out[0].conv.bias.detach().float().reshape(1).contiguous()
        self._cache_key, self._cache = key, c
        return c

    # -- forward ----------------------------------------------------------------------------------
    def forward(self, content, style, return_intermediate=False, noise=None):
        """content [T,B,n_class] fp32 (one-hot or dense), style [B,style_size].
        `noise`: optional list of the ten [B,C,H,W] tensors the reference would have drawn with
        torch.randn_like (pure_gen.py:206,212), in call order — for parity tests.  By default the
        noise is generated inside the kernels (counter-based hash + Box-Muller), seeded from torch's CPU generator."""
        _lib.require_cuda(content, style)
        if torch.is_grad_enabled() and (content.requires_grad or style.requires_grad
                                        or any(p.requires_grad for p in self.parameters())):
            from .autograd_gen import generator_apply  # backward pass lives there
            return generator_apply(self, content, style, noise)
        return self._forward_impl(content, style, noise)[0]

    def _style_vectors(self, c, style):
        s = ops.pixelnorm(style.float().contiguous())
        li = 0
        for m in self.style_emb:
            if isinstance(m, nn.Linear):
                w, b = c["mlp"][li]
                li += 1
                s = ops.linear(s, w, b, ACT_LRELU, 0.2)
            elif isinstance(m, nn.Dropout) and self.training:
                s = torch.nn.functional.dropout(s, m.p, True)
        gb = ops.linear(s, c["gb_w"], c["gb_b"])  # all ten AdaIN projections at once: [B, sum 2C]
        return s, gb

    def _forward_impl(self, content, style, noise=None, keep=False):
        c = self._packed()
        T, B, ncls = content.shape
        dev = content.device
        s, gb = self._style_vectors(c, style)
        x = ops.gen_pack_input(content.float(), s if self.append_style else None, c["cin_pad"])
        seed, seed_dev = None, None
        if noise is None:
            seed, seed_dev = self._noise_seed()
        else:
            noise = [z.permute(0, 2, 3, 1).contiguous().float() for z in noise]  # NHWC for the kernels
        # one zero-fill for all ten statistics buffers ([B,C,2] each)
        stats_all = torch.zeros(sum(2 * B * e["C"] * 2 for e in c["blocks"]), device=dev, dtype=torch.float32)
        soff = 0

        def new_stats(C):
            nonlocal soff
            st = stats_all[soff:soff + B * C * 2].view(B, C, 2)
            soff += B * C * 2
            return st

        gbs = gb.stride(0)
        saved = []
        off = 0       # running offset into gb (each AdaIN projection is 2C wide: gamma | beta)
        H, W = 1, T
        k = 0         # AdaIN / noise index
        out = None
        nblk = len(c["blocks"])
        for bi, e in enumerate(c["blocks"]):
            C = e["C"]
            # ---------------- conv1 (+ blur) + noise + lrelu + stats ----------------
            st = new_stats(C)
            nz = None if noise is None else noise[k]
            y, Ho, Wo = conv1_forward(x, e, B, H, W, st, nz, k, seed, seed_dev)
            H, W = Ho, Wo
            coef = ops.adain_coeffs(st, gb[:, off:], gb[:, off + C:], gbs, B, C, H * W)
            if keep:
                saved.append((x, y, st, coef))
            x = ops.scale_shift_act(y, coef, True)
            off += 2 * C
            k += 1
            # ---------------- conv2 + noise + lrelu + stats ----------------
            st = new_stats(C)
            nz = None if noise is None else noise[k]
            y = conv.conv_fprop(x, e["w2"], TAPS3x3, H, W, bias=e["b2"], act=ACT_LRELU, slope=0.2, noise=nz,
                                noise_w=e["nw2"], noise_seed=seed, noise_subseq=16 * k, noise_seed_dev=seed_dev, stats=st)
            coef = ops.adain_coeffs(st, gb[:, off:], gb[:, off + C:], gbs, B, C, H * W)
            if keep:
                saved.append((x, y, st, coef))
            if bi == nblk - 1:
                out = ops.gen_output(y, coef, c["w_out"], c["b_out"])  # AdaIN + 1x1 conv + tanh in one pass
            else:
                x = ops.scale_shift_act(y, coef, True)
            off += 2 * C
            k += 1
        return out, saved
