"""Differentiable building blocks on NHWC bf16 tensors for the modules that are composed from the library's entry points
without a hand-written backward pass of their own (the style extractor, char_style.py): a convolution [+ GroupNorm] [+ ReLU]
as ONE torch.autograd.Function whose forward, input gradient, weight gradient and normalisation passes run on libhwg_b200
(`hwg_conv_fprop` with its statistics epilogue, `hwg_scale_shift_act`, `hwg_norm_bwd_reduce` / `hwg_norm_bwd_apply`,
`hwg_act_bwd`, `hwg_conv_wgrad`, `hwg_channel_sum`), plus the layout glue between such blocks (replicate padding,
space-to-depth for strided convolutions) as plain differentiable tensor indexing.

`conv_block(x, weight, bias, taps, ...)` takes the convolution weight in the reference's own [Cout, Cin, kh, kw] layout as
an autograd input, packs it to the kernels' tap-major bf16 operands on the fly and returns the weight gradient in the same
layout — so a caller can put any differentiable re-arrangement in front of it (e.g. the space-to-depth view of a stride-2
kernel) and autograd carries the gradient back to the parameter.

GROUPED mode (`counts`): the images of x are sorted into G consecutive groups and `weight` / `bias` / `gamma` / `beta` carry a
leading group dimension — the 79 per-character heads of the style extractor; the convolution launches (forward, input
gradient, weight gradient) loop over the groups, every other pass handles all images at once.

The per-(image, channel) GroupNorm coefficients are a few tensor operations on [N, C, 2] statistics (host-side composition,
no image-sized pass): forward  a = gamma * rstd, b = beta - mean * a;  backward  gz = sc * gy' + P * z + Q."""
import torch

from . import _lib, conv, ops
from ._lib import ACT_NONE, ACT_RELU


def _pack(weight, cin_pad):
    """[Co, Ci, kh, kw] fp32 -> (forward operand [K, Co, cin_pad], dgrad operand [K, cin_pad, Cop]) bf16."""
    co, ci, kh, kw = weight.shape
    K = kh * kw
    w = weight.detach().float()
    f = torch.zeros((K, co, cin_pad), device=w.device, dtype=torch.bfloat16)
    f[:, :, :ci] = w.permute(2, 3, 0, 1).reshape(K, co, ci)
    cop = -(-co // 16) * 16
    d = torch.zeros((K, cin_pad, cop), device=w.device, dtype=torch.bfloat16)
    d[:, :ci, :co] = w.permute(2, 3, 1, 0).reshape(K, ci, co)
    return f, d


def _gn_coeffs(stats, gamma_n, beta_n, groups, hw, eps):
    """stats [N,C,2] (sum, sum of squares over the image) -> coef [N,C,2] (a, b), save (mean [N,G,1], rstd [N,G,1])."""
    N, C, _ = stats.shape
    g = stats.view(N, groups, C // groups, 2).sum(2)                         # [N,G,2]
    m = float(hw * (C // groups))
    mean = g[..., 0] / m
    var = (g[..., 1] / m - mean * mean).clamp_min(0.0)
    rstd = torch.rsqrt(var + eps)
    mean_c = mean.repeat_interleave(C // groups, 1)
    rstd_c = rstd.repeat_interleave(C // groups, 1)
    a = gamma_n * rstd_c
    coef = torch.stack((a, beta_n - mean_c * a), 2).contiguous()
    return coef, (mean_c, rstd_c)


def _gn_bwd_coeffs(sums, save, gamma_n, groups, hw):
    """sums [N,C,2] = (sum gy', sum gy' * z) -> spq [N,C,3] with gz = sc * gy' + P * z + Q, and the per-image
    (d gamma, d beta) [N,C] contributions."""
    mean_c, rstd_c = save
    N, C, _ = sums.shape
    cg = C // groups
    s_g, s_gz = sums[..., 0], sums[..., 1]
    s_gx = (s_gz - mean_c * s_g) * rstd_c                                    # sum gy' * xhat per (n, c)
    m = float(hw * cg)
    S1 = (gamma_n * s_g).view(N, groups, cg).sum(2).repeat_interleave(cg, 1) / m
    S2 = (gamma_n * s_gx).view(N, groups, cg).sum(2).repeat_interleave(cg, 1) / m
    sc = gamma_n * rstd_c
    P = -rstd_c * rstd_c * S2
    Q = -rstd_c * S1 + rstd_c * rstd_c * mean_c * S2
    return torch.stack((sc, P, Q), 2).contiguous(), s_gx, s_g


class _ConvBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, taps, Ho, Wo, groups, eps, relu, counts):
        N, H, W, Cx = x.shape
        grouped = counts is not None
        wl = list(weight) if grouped else [weight]
        bl = list(bias) if grouped else [bias]
        co = wl[0].size(0)
        packs = [_pack(w, Cx) for w in wl]
        bounds = [0]
        for c in (counts if grouped else [N]):
            bounds.append(bounds[-1] + int(c))
        assert bounds[-1] == N, "conv_block: counts must add up to the number of images"
        gn = gamma is not None
        dev = x.device
        z = torch.empty((N, Ho, Wo, co), device=dev, dtype=torch.bfloat16)
        stats = torch.zeros((N, co, 2), device=dev, dtype=torch.float32) if gn else None
        act = ACT_RELU if (relu and not gn) else ACT_NONE
        for gi, (f, _) in enumerate(packs):
            s, e = bounds[gi], bounds[gi + 1]
            if e > s:
                conv.conv_fprop(x[s:e], f, taps, Ho, Wo, bias=bl[gi].detach().float().contiguous(), act=act, out=z[s:e],
                                stats=None if stats is None else stats[s:e], force_tcgen05=True)
        coef = save = gamma_n = None
        y = z
        if gn:
            idx = None
            if grouped:
                idx = torch.repeat_interleave(torch.arange(len(wl), device=dev),
                                              torch.tensor([bounds[i + 1] - bounds[i] for i in range(len(wl))], device=dev))
            gamma_n = gamma.detach().float()[idx] if grouped else gamma.detach().float()[None].expand(N, co)
            beta_n = beta.detach().float()[idx] if grouped else beta.detach().float()[None].expand(N, co)
            coef, save = _gn_coeffs(stats, gamma_n, beta_n, groups, Ho * Wo, eps)
            y = ops.scale_shift_act(z, coef, True, ACT_RELU if relu else ACT_NONE, 0.0, out=torch.empty_like(z))
            ctx.idx = idx
        ctx.geom = (taps, H, W, Ho, Wo, groups, relu, gn, grouped, bounds)
        ctx.saved = dict(x=x, z=z if gn else None, y=None if gn else (y if relu else None), coef=coef, save=save,
                         gamma_n=gamma_n, packs=packs, shapes=[tuple(w.shape) for w in wl])
        return y.detach() if y is z else y

    @staticmethod
    def backward(ctx, g):
        taps, H, W, Ho, Wo, groups, relu, gn, grouped, bounds = ctx.geom
        sv = _lib.saved_state(ctx.saved)
        x, packs = sv["x"], sv["packs"]
        dev = g.device
        g = g.contiguous()
        N, co = g.size(0), g.size(3)
        Cx = x.size(3)
        need = ctx.needs_input_grad
        dgamma = dbeta = None
        s_ = _lib.stream
        if gn:
            z, coef = sv["z"], sv["coef"]
            sums = torch.zeros((N, co, 2), device=dev, dtype=torch.float32)
            _lib.call("hwg_norm_bwd_reduce", g.data_ptr(), z.data_ptr(), coef.data_ptr(), 0.0 if relu else 1.0, N, Ho, Wo, co,
                      1, 1, sums.data_ptr(), s_())
            spq, dg_n, db_n = _gn_bwd_coeffs(sums, sv["save"], sv["gamma_n"], groups, Ho * Wo)
            gz = torch.empty_like(z)
            _lib.call("hwg_norm_bwd_apply", g.data_ptr(), z.data_ptr(), coef.data_ptr(), spq.data_ptr(), 0.0 if relu else 1.0,
                      N, Ho, Wo, co, 1, 1, gz.data_ptr(), s_())
            if grouped:
                G = len(packs)
                dgamma = torch.zeros((G, co), device=dev, dtype=torch.float32).index_add_(0, ctx.idx, dg_n)
                dbeta = torch.zeros((G, co), device=dev, dtype=torch.float32).index_add_(0, ctx.idx, db_n)
            else:
                dgamma, dbeta = dg_n.sum(0), db_n.sum(0)
        elif relu:
            y = sv["y"]
            gz = torch.empty_like(y)
            _lib.call("hwg_act_bwd", g.data_ptr(), y.data_ptr(), None, 0.0, N, Ho, Wo, co, 1, 1, gz.data_ptr(), s_())
        else:
            gz = g
        gws, gbs = [], []
        gx = torch.zeros_like(x) if need[0] else None
        taps_d = [(-dh, -dw) for dh, dw in taps]
        cop = -(-co // 16) * 16
        for gi, (f, d) in enumerate(packs):
            s, e = bounds[gi], bounds[gi + 1]
            co_, ci_, kh, kw = sv["shapes"][gi]
            K = kh * kw
            if e > s:
                dw = torch.zeros((K, co, Cx), device=dev, dtype=torch.float32)
                conv.conv_wgrad(x[s:e], gz[s:e], taps, Cx, co, out=dw)
                db = torch.zeros(co, device=dev, dtype=torch.float32)
                _lib.call("hwg_channel_sum", gz[s:e].data_ptr(), (e - s) * Ho * Wo, co, db.data_ptr(), s_())
                gws.append(dw[:, :, :ci_].permute(1, 2, 0).reshape(co_, ci_, kh, kw))
                gbs.append(db)
                if need[0]:
                    # d[:, :, :cop] is [K, Cx, Cop]: the input gradient is the convolution of gz with the transposed taps
                    conv.conv_fprop(gz[s:e], d, taps_d, H, W, out=gx[s:e], force_tcgen05=True)
            else:
                gws.append(torch.zeros((co_, ci_, kh, kw), device=dev, dtype=torch.float32))
                gbs.append(torch.zeros(co, device=dev, dtype=torch.float32))
        gw = torch.stack(gws, 0) if grouped else gws[0]
        gb = torch.stack(gbs, 0) if grouped else gbs[0]
        if not _lib.RETAIN_SAVED:
            ctx.saved = None
        return (gx, gw if need[1] else None, gb if need[2] else None, dgamma if (gn and need[3]) else None,
                dbeta if (gn and need[4]) else None, None, None, None, None, None, None, None)


def conv_block(x, weight, bias, taps, Ho, Wo, gamma=None, beta=None, groups=8, eps=1e-5, relu=False, counts=None):
    """y = [ReLU]([GroupNorm_groups](conv_taps(x; weight) + bias)).

    x      [N,H,W,Cx] bf16 NHWC, Cx in {16, 32, 64k} (channels beyond the weight's Cin must be zero)
    weight [Cout, Cin, kh, kw] fp32 with kh*kw == len(taps) (grouped: [G, Cout, Cin, kh, kw]), Cin <= Cx, Cout in {16, 32} or a
           multiple of 8 >= 64;  bias [Cout] (grouped [G, Cout]);  gamma / beta: GroupNorm affine or None
    taps   input-pixel offsets (dh, dw) per kernel position: out[ho, wo] += w[:, :, k] . x[ho + dh_k, wo + dw_k] (zero outside)
    counts grouped mode: images per group (host ints, consecutive groups)."""
    _lib.require_cuda(x)
    assert x.dtype == torch.bfloat16 and x.dim() == 4
    x = x.contiguous()
    return _ConvBlockFn.apply(x, weight, bias, gamma, beta, list(taps), int(Ho), int(Wo), int(groups), float(eps), bool(relu),
                              None if counts is None else [int(c) for c in counts])


# ---- layout glue (differentiable tensor indexing) ---------------------------------------------------------------
def replicate_pad(x, left, right, top, bottom):
    """nn.ReplicationPad2d((left, right, top, bottom)) on an NHWC tensor."""
    N, H, W, C = x.shape
    if left or right:
        iw = torch.arange(-left, W + right, device=x.device).clamp_(0, W - 1)
        x = x.index_select(2, iw)
    if top or bottom:
        ih = torch.arange(-top, H + bottom, device=x.device).clamp_(0, H - 1)
        x = x.index_select(1, ih)
    return x


def space_to_depth(x, sh, sw):
    """[N,H,W,C] -> [N,ceil(H/sh),ceil(W/sw),sh*sw*C] with channel order (row phase, column phase, c): a stride-(sh,sw)
    convolution of x becomes a stride-1 convolution of the result (see `strided_weight`)."""
    N, H, W, C = x.shape
    ph, pw = (-H) % sh, (-W) % sw
    if ph or pw:
        x = torch.nn.functional.pad(x, (0, 0, 0, pw, 0, ph))
        H, W = H + ph, W + pw
    return x.view(N, H // sh, sh, W // sw, sw, C).permute(0, 1, 3, 2, 4, 5).reshape(N, H // sh, W // sw, sh * sw * C)


def strided_weight(weight, sh, sw):
    """[Co, Ci, kh, kw] kernel of a stride-(sh,sw) convolution -> [Co, sh*sw*Ci, kh/sh, kw/sw] kernel of the stride-1
    convolution on `space_to_depth(x, sh, sw)` (kh % sh == 0 and kw % sw == 0)."""
    co, ci, kh, kw = weight.shape
    assert kh % sh == 0 and kw % sw == 0
    w = weight.view(co, ci, kh // sh, sh, kw // sw, sw).permute(0, 3, 5, 1, 2, 4)
    return w.reshape(co, sh * sw * ci, kh // sh, kw // sw)


def valid_taps(kh, kw):
    return [(i, j) for i in range(kh) for j in range(kw)]
