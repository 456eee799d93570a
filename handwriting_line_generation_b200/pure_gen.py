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

from . import _lib, conv, ops, weightmap
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
        self._plan, self._plan_ptrs, self._bwd_plans = None, None, {}
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
    def _build_plan(self):
        """Allocates the kernel-side operand buffers once and the hwg_linear_map job table that (re)fills all of them
        from the fp32 parameters in ONE launch (weightmap.py): tap-major bf16 forward and dgrad operands of every
        convolution, EqualLR-scaled noise/output weights, folded biases, the concatenated AdaIN projections."""
        dev = self.out[0].conv.weight_orig.device
        bf = dict(device=dev, dtype=torch.bfloat16)
        f32 = dict(device=dev, dtype=torch.float32)
        t = weightmap.JobTable()
        c = {"blocks": []}
        cin_pad = ((self.in_ch + 63) // 64) * 64
        c["cin_pad"] = cin_pad
        one = [[1.0]]

        def vec_job(src, dst, C, scale=1.0, reps=1):
            t.add(src, dst, R=1, C=C, s_r=0, s_c=1, d_r=0, d_c=1, M=[[1.0] * reps], out_off=[f * C for f in range(reps)],
                  scale=scale)

        n_gb = sum(2 * blk.out_channel for blk in self.conv) * 2
        c["gb_w"] = torch.empty((n_gb, self.style_size), **f32)
        c["gb_b"] = torch.empty(n_gb, **f32)
        row = 0
        for blk in self.conv:
            C = blk.out_channel
            e = {"kind": blk.kind, "C": C}
            if blk.kind == "initial":
                w, b = blk.conv1.weight, blk.conv1.bias
                m1 = weightmap.map_initial(self.in_ch, C)
                e["taps1"] = [(0, 1 - kx) for kx in range(3)]
                e["w1f"] = torch.empty((3, 4 * C, cin_pad), **bf)      # output rows as channel folds
                m1.add_pack_fwd(t, w, e["w1f"], Cip=cin_pad,
                                out_off=[kx * 4 * C * cin_pad + r * C * cin_pad for r in range(4) for kx in range(3)])
                e["m1_cip"] = cin_pad
            elif blk.kind == "vert_up":
                w, b = blk.conv1[1].weight, blk.conv1[1].bias
                m1 = weightmap.map_vert_up(C, blk.in_channel)
                buf = torch.empty((12, C, m1.Cip), **bf)
                m1.add_pack_fwd(t, w, buf)
                e["w1"] = [(weightmap.vert_taps(par), buf[6 * par:6 * par + 6]) for par in (0, 1)]
            elif blk.kind == "fused_up":
                w, b = blk.conv1[0].weight, blk.conv1[0].bias
                m1 = weightmap.map_fused_up(blk.in_channel, C, blk.conv1[0].multiplier)
                e["w1f"] = torch.empty((16, C, m1.Cip), **bf)          # output parities as channel folds
                m1.add_pack_fwd(t, w, e["w1f"])
                e["taps1f"] = weightmap.fused_taps()
            else:
                w, b = blk.conv1.weight, blk.conv1.bias
                m1 = weightmap.map_conv3x3(C, blk.in_channel)
                e["w1"] = torch.empty((9, C, m1.Cip), **bf)
                m1.add_pack_fwd(t, w, e["w1"])
            e["d1"] = torch.empty(m1.dgrad_shape(), **bf)
            m1.add_pack_dgrad(t, w, e["d1"])
            e["m1"], e["p_w1"], e["p_b1"] = m1, w, b
            e["b1"] = b.detach()
            if blk.kind in ("initial", "fused_up"):
                e["b1f"] = torch.empty(4 * C, **f32)
                vec_job(b, e["b1f"], C, reps=4)
            m2 = weightmap.map_conv3x3(C, C)
            e["w2"] = torch.empty((9, C, C), **bf)
            e["d2"] = torch.empty(m2.dgrad_shape(), **bf)
            m2.add_pack_fwd(t, blk.conv2.weight, e["w2"])
            m2.add_pack_dgrad(t, blk.conv2.weight, e["d2"])
            e["m2"] = m2
            e["b2"] = blk.conv2.bias.detach()
            nscale = sqrt(2.0 / C)                   # EqualLR on NoiseInjection: fan_in = C (pure_gen.py:222-226)
            e["nw1"], e["nw2"] = torch.empty(C, **f32), torch.empty(C, **f32)
            vec_job(blk.noise1.weight_orig, e["nw1"], C, nscale)
            vec_job(blk.noise2.weight_orig, e["nw2"], C, nscale)
            if blk.kind == "initial":
                e["nw1f"] = torch.empty(4 * C, **f32)
                vec_job(blk.noise1.weight_orig, e["nw1f"], C, nscale, reps=4)
            for ad in (blk.adain1, blk.adain2):
                t.add(ad.style.weight, c["gb_w"], R=2 * C, C=self.style_size, s_r=self.style_size, s_c=1,
                      d_r=self.style_size, d_c=1, M=one, out_off=[row * self.style_size])
                t.add(ad.style.bias, c["gb_b"], R=1, C=2 * C, s_r=0, s_c=1, d_r=0, d_c=1, M=one, out_off=[row])
                row += 2 * C
            c["blocks"].append(e)
        c["mlp"] = [(m.weight.detach(), m.bias.detach()) for m in self.style_emb if isinstance(m, nn.Linear)]
        wo = self.out[0].conv.weight_orig
        c["out_scale"] = sqrt(2.0 / (wo.size(1) * wo[0][0].numel()))
        c["w_out"] = torch.empty(wo.size(1), **f32)
        vec_job(wo, c["w_out"], wo.size(1), c["out_scale"])
        c["b_out"] = self.out[0].conv.bias.detach().reshape(1)
        for p in self.parameters():
            assert p.is_contiguous() and p.dtype == torch.float32, "generator parameters must be contiguous fp32"
        t.finalize(dev)
        return {"table": t, "c": c}

    def _packed(self):
        key = tuple((p.data_ptr(), p._version) for p in _lib.params(self))
        if self._cache_key == key:
            return self._cache
        ptrs = tuple(k[0] for k in key)
        if self._plan is None or self._plan_ptrs != ptrs:
            self._plan, self._plan_ptrs = self._build_plan(), ptrs
            self._bwd_plans = {}
        self._plan["table"].run()
        self._cache_key, self._cache = key, self._plan["c"]
        return self._cache

    # -- forward ----------------------------------------------------------------------------------
    def forward(self, content, style, return_intermediate=False, noise=None):
        """content [T,B,n_class] fp32 (one-hot or dense), style [B,style_size].
        `noise`: optional list of the ten [B,C,H,W] tensors the reference would have drawn with
        torch.randn_like (pure_gen.py:206,212), in call order — for parity tests.  By default the
        noise is generated inside the kernels (counter-based hash + Box-Muller), seeded from torch's CPU generator."""
        _lib.require_cuda(content, style)
        if torch.is_grad_enabled() and (content.requires_grad or style.requires_grad
                                        or any(p.requires_grad for p in _lib.params(self))):
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
