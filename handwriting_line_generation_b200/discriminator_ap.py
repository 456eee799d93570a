"""Drop-in for the reference discriminator `DiscriminatorAP` (model/discriminator_ap.py:68-161) with its
`SpectralNorm` wrapper (:11-65) — SURVEY.md §8 row f1.

Same constructor signature, module names, construction order (same seed -> same initial weights) and `state_dict` keys
as the reference (`convs1.0.module.weight_bar/_u/_v/bias`, ...).  The torch sub-modules are parameter containers; the
forward runs on libhwg_b200:

* every spectral-normalised layer's power iteration in ONE launch (`hwg_spectral_norm`), every weight re-layout —
  tap-major bf16 forward and dgrad operands scaled by the fresh 1/sigma — in ONE `hwg_linear_map` launch;
* `in_conv` (7x7, one input channel) by `hwg_stem_conv` straight from the fp32 image while the discriminator is frozen
  (the 'gen' lessons); as a 7-tap implicit GEMM over the 16-channel shift expansion of the image (`hwg_shift_expand`)
  when its weight gradient is needed; GroupNorm statistics from the convolution epilogue, GroupNorm + LeakyReLU as one scale-shift
  pass, all other convolutions on the tcgen05 kernel with bias + LeakyReLU epilogues, `hwg_avgpool_nhwc`,
  Dropout2d as a per-(sample, channel) scale fused with the LeakyReLU pass;
* backward to the INPUT image (what the generator's adversarial loss needs, trainer/hw_with_style_trainer.py:810-821):
  dgrad on the tensor cores, `hwg_act_bwd` (Dropout2d + LeakyReLU + AvgPool2d backward in one pass), the three-launch
  GroupNorm backward, `hwg_shift_collapse`;
* backward to the discriminator's own weights (the 'disc' lesson, trainer :785-804) when its parameters require
  gradients: `hwg_conv_wgrad` per layer into one zero-filled arena, `hwg_channel_sum` for the biases, GroupNorm
  gamma/beta sums from `hwg_gn_bwd_coeffs`, ONE `hwg_linear_map` launch that unpacks everything into the parameters'
  layouts (spectral layers scaled by 1/sigma) and `hwg_spectral_norm_bwd` for the rank-one term of
  weight = w_bar / sigma(w_bar).

u / v are updated in place on every forward, as the reference does, and the packed operands are shared between forward
and backward: the backward of a forward must run before the next forward of the same module (it raises otherwise).
There is no PyTorch fallback."""


import numpy as np
import torch
import torch.nn as nn

from . import _lib, conv, ops, weightmap
from ._lib import ACT_LRELU, ACT_NONE

LEAK = 0.1
_NO_STEM = bool(__import__("os").environ.get("HWG_NO_STEM_CONV"))   # development A/B switch: shift-expansion route for in_conv


def get_group_size(channels):
    """utils/util.py:391 for the channel counts that occur here (multiples of 8 >= 32)."""
    if channels >= 32 and channels % 8 == 0:
        return 8
    raise NotImplementedError(f"GroupNorm group size for {channels} channels")


class SpectralNorm(nn.Module):
    """Parameter container of the reference's SpectralNorm (:11-65): `module.weight` is replaced by `weight_bar`,
    `weight_u`, `weight_v` (same names, shapes and initialisation order); the power iteration itself runs in
    hwg_spectral_norm."""

    def __init__(self, module, name='weight', power_iterations=1):
        super().__init__()
        if name != 'weight' or power_iterations != 1:
            raise NotImplementedError("SpectralNorm: name='weight', power_iterations=1 only")
        self.module = module
        w = module.weight
        height = w.shape[0]
        width = w.reshape(height, -1).shape[1]
        u = nn.Parameter(w.data.new(height).normal_(0, 1), requires_grad=False)
        v = nn.Parameter(w.data.new(width).normal_(0, 1), requires_grad=False)
        u.data = u.data / (u.data.norm() + 1e-12)
        v.data = v.data / (v.data.norm() + 1e-12)
        w_bar = nn.Parameter(w.data)
        del module._parameters['weight']
        module.register_parameter('weight_u', u)
        module.register_parameter('weight_v', v)
        module.register_parameter('weight_bar', w_bar)


# (key, attribute path, taps, spectral) in forward order; Dropout2d p of the site that follows the conv
_T31 = conv.conv_taps(3, 3, 0, 1)
_T13 = conv.conv_taps(1, 3, 0, 1)
_DROP_P = {"convs1.3": 0.05, "convs3.4": 0.05, "convs4.0": 0.025, "convs4.4": 0.025, "convs4.7": 0.025, "convs4.11": 0.025}


def _tile_w(taps, Ho, Wo):
    """Output-tile width of a launch on a TALL activation (0 = the kernel's own choice).  A 128 x 1 pixel tile re-reads
    every input row once per vertical tap from HBM (the 121 MB activations of in_conv / convs1 do not stay in L2: ncu
    showed 348 MB read for a 121 MB operand); 32 x 4 (3 row taps: 1.5x) and 16 x 8 (7 row taps: 1.75x) tiles keep the
    vertical halo inside the tile."""
    rows = len({dh for dh, _ in taps})
    if rows >= 7 and Ho >= 8 and Wo >= 16:
        return 16
    if rows >= 3 and Ho >= 4 and Wo >= 32:
        return 32
    return 0


class DiscriminatorAP(nn.Module):
    def __init__(self, dim=64, use_low=False, use_med=True, small=False):
        super().__init__()
        if small:
            raise NotImplementedError("small=True is not used by the reference configs")
        if dim % 16 != 0:
            raise NotImplementedError("dim must be a multiple of 16")
        self.use_low, self.use_med, self.dim = use_low, use_med, dim
        leak = LEAK
        self.in_conv = nn.Sequential(nn.Conv2d(1, dim, 7, stride=1, padding=(0, 3)),
                                     nn.GroupNorm(get_group_size(dim), dim), nn.LeakyReLU(leak, True))
        self.convs1 = nn.Sequential(SpectralNorm(nn.Conv2d(dim, dim, 3, stride=1, padding=(0, 1))), nn.LeakyReLU(leak, True),
                                    nn.AvgPool2d(2),
                                    SpectralNorm(nn.Conv2d(dim, 2 * dim, 3, stride=1, padding=(0, 1))),
                                    nn.Dropout2d(0.05, True), nn.LeakyReLU(leak, True))
        self.convs2 = nn.Sequential(SpectralNorm(nn.Conv2d(2 * dim, 2 * dim, 3, stride=1, padding=(0, 1))),
                                    nn.LeakyReLU(leak, True), nn.AvgPool2d(2))
        self.convs3 = nn.Sequential(nn.Conv2d(2 * dim, 2 * dim, 3, stride=1, padding=(0, 1)),
                                    nn.GroupNorm(get_group_size(2 * dim), 2 * dim), nn.LeakyReLU(leak, True),
                                    nn.AvgPool2d(2),
                                    SpectralNorm(nn.Conv2d(2 * dim, 4 * dim, 3, stride=1, padding=(0, 1))),
                                    nn.Dropout2d(0.05, True), nn.LeakyReLU(leak, True))
        if use_med:
            self.finalMed = nn.Sequential(SpectralNorm(nn.Conv2d(4 * dim, 1, 3, stride=1, padding=(0, 1))))
        if use_low:
            self.convs4 = nn.Sequential(
                SpectralNorm(nn.Conv2d(4 * dim, 2 * dim, 3, stride=1, padding=(0, 1))), nn.Dropout2d(0.025, True),
                nn.LeakyReLU(leak, True), nn.AvgPool2d((1, 2)),
                SpectralNorm(nn.Conv2d(2 * dim, 4 * dim, (1, 3), stride=1, padding=(0, 1))), nn.Dropout2d(0.025, True),
                nn.LeakyReLU(leak, True),
                SpectralNorm(nn.Conv2d(4 * dim, 4 * dim, (1, 3), stride=1, padding=(0, 1))), nn.Dropout2d(0.025, True),
                nn.LeakyReLU(leak, True), nn.AvgPool2d((1, 2)),
                SpectralNorm(nn.Conv2d(4 * dim, 4 * dim, (1, 3), stride=1, padding=(0, 1))), nn.Dropout2d(0.025, True),
                nn.LeakyReLU(leak, True),
                SpectralNorm(nn.Conv2d(4 * dim, 1, 1, stride=1, padding=(0, 0))))
        self._plan, self._plan_ptrs = None, None
        self._site_channels = {site: m.bias.numel() for site, m, _, _ in self.conv_layers() if site in _DROP_P}
        self._drop_const = None
        self._generation = 0           # forwards so far: a backward must use the operands its own forward packed
        self.dropout_masks = None      # tests: dict site -> [B,C] 0/1 keep-mask instead of drawing one

    # -- layers ---------------------------------------------------------------------------------------------------
    def conv_layers(self):
        """(site, conv module, taps, spectral-normalised) of every tensor-core convolution, forward order."""
        out = [("in_conv.0", self.in_conv[0], [(dy, 0) for dy in range(7)], False),
               ("convs1.0", self.convs1[0].module, _T31, True), ("convs1.3", self.convs1[3].module, _T31, True),
               ("convs2.0", self.convs2[0].module, _T31, True), ("convs3.0", self.convs3[0], _T31, False),
               ("convs3.4", self.convs3[4].module, _T31, True)]
        if self.use_med:
            out.append(("finalMed.0", self.finalMed[0].module, _T31, True))
        if self.use_low:
            out += [("convs4.0", self.convs4[0].module, _T31, True), ("convs4.4", self.convs4[4].module, _T13, True),
                    ("convs4.7", self.convs4[7].module, _T13, True), ("convs4.11", self.convs4[11].module, _T13, True),
                    ("convs4.14", self.convs4[14].module, [(0, 0)], True)]
        return out

    def _build_plan(self):
        dev = self.in_conv[0].weight.device
        layers = self.conv_layers()
        sn = [(site, m) for site, m, _, spectral in layers if spectral]
        inv_sigma = torch.ones(len(sn), device=dev, dtype=torch.float32)
        jobs = np.zeros((len(sn), 4), np.int64)                     # {w, u, v, (h | wd << 32)} = 32 bytes per layer
        sn_index = {}
        for i, (site, m) in enumerate(sn):
            w = m.weight_bar
            assert w.is_contiguous() and w.dtype == torch.float32
            h, wd = w.size(0), w[0].numel()
            jobs[i] = (w.data_ptr(), m.weight_u.data_ptr(), m.weight_v.data_ptr(), h | (wd << 32))
            sn_index[site] = i
        t = weightmap.JobTable()
        c = {"dgrad": {}, "bias": {}}
        for site, m, taps, spectral in layers:
            w = m.weight_bar if spectral else m.weight
            co, ci = w.size(0), w.size(1)
            scale = inv_sigma[sn_index[site]:sn_index[site] + 1] if spectral else None
            if site == "in_conv.0":
                # [co,1,7,7]: rows dy are the taps, columns dx the 16 (7 used) channels of the shift expansion
                # one job per tap (kernel row)
                c[site] = torch.empty((7, co, 16), device=dev, dtype=torch.bfloat16)
                d = torch.empty((7, 16, co), device=dev, dtype=torch.bfloat16)
                for dy in range(7):
                    t.add(w[:, 0, dy, :], c[site][dy], R=co, C=7, Cp=16, s_r=49, s_c=1, d_r=16, d_c=1,
                          M=np.eye(1), dst_bf16=True)
                    t.add(w[:, 0, dy, :], d[dy], R=7, C=co, Rp=16, Cp=co, s_r=1, s_c=49, d_r=co, d_c=1,
                          M=np.eye(1), dst_bf16=True)
                c["dgrad"][site] = (d, [(-dy, 0) for dy in range(7)])
            else:
                mp = weightmap.map_conv_taps(co, ci, taps)
                cop = -(-co // 16) * 16
                c[site] = torch.empty((mp.Tf, cop, mp.Cip), device=dev, dtype=torch.bfloat16)
                d = torch.empty(mp.dgrad_shape(), device=dev, dtype=torch.bfloat16)
                mp.add_pack_fwd(t, w, c[site], scale_dev=scale, Cop=cop)
                mp.add_pack_dgrad(t, w, d, scale_dev=scale)
                c["dgrad"][site] = (d, mp.taps_d)
            b = m.bias.detach()
            if b.numel() % 16:                                       # 1-channel heads run as 16-channel launches
                bp = torch.zeros(-(-b.numel() // 16) * 16, device=dev, dtype=torch.float32)
                t.add(b, bp, R=1, C=b.numel(), s_r=0, s_c=1, d_r=0, d_c=1, M=np.eye(1))   # refreshed with the weights
                c["bias"][site] = (bp, b)
            else:
                c["bias"][site] = (b, None)
        t.finalize(dev)
        return {"table": t, "c": c, "inv_sigma": inv_sigma, "sn_jobs": torch.from_numpy(jobs).to(dev), "n_sn": len(sn),
                "sn_max": (int(max(j[3] & 0xffffffff for j in jobs)), int(max(j[3] >> 32 for j in jobs))),
                "sn_norms": torch.zeros(2 * len(sn), device=dev, dtype=torch.float32)}

    def _prepare(self):
        """Per forward (the reference updates u, v and re-derives weight = w_bar / sigma on EVERY forward, :62-64):
        hwg_spectral_norm (three small launches for all layers) + one hwg_linear_map launch."""
        ptrs = tuple(p.data_ptr() for p in _lib.params(self))
        if self._plan is None or self._plan_ptrs != ptrs:
            self._plan, self._plan_ptrs = self._build_plan(), ptrs
        p = self._plan
        self._generation += 1
        _lib.call("hwg_spectral_norm", p["sn_jobs"].data_ptr(), p["n_sn"], p["sn_max"][0], p["sn_max"][1],
                  p["sn_norms"].data_ptr(), p["inv_sigma"].data_ptr(), _lib.stream())
        p["table"].run()
        return p["c"]

    def _drop_all(self, B, dev):
        """Dropout2d (:89 etc., inplace, training only) of ALL sites in a handful of launches: per-(sample, channel)
        keep-mask / (1-p) as scale-shift coefficients [B,C,2] = (scale, 0) for the fused LeakyReLU pass and as the
        scale [B,C] the backward reads.  Returns {site: (coef, scale)}; {} in eval mode."""
        if not self.training:
            return {}
        sites = [(s, p, self._site_channels[s]) for s, p in _DROP_P.items() if s in self._site_channels]
        key = (B, str(dev))
        if getattr(self, "_drop_const", None) is None or self._drop_const[0] != key:
            p_flat = torch.cat([torch.full((B * C,), p, dtype=torch.float32) for _, p, C in sites]).to(dev)
            self._drop_const = (key, p_flat, 1.0 / (1.0 - p_flat))
        _, p_flat, inv_flat = self._drop_const
        if self.dropout_masks is not None:
            keep = torch.cat([self.dropout_masks[s].to(dev).float().reshape(-1) for s, _, _ in sites])
            scale = keep * inv_flat
        else:
            scale = (torch.rand(p_flat.numel(), device=dev) >= p_flat).float().mul_(inv_flat)
        coef = torch.zeros((p_flat.numel(), 2), device=dev, dtype=torch.float32)
        coef[:, 0] = scale
        out, off = {}, 0
        for s, _, C in sites:
            out[s] = (coef[off:off + B * C].view(B, C, 2), scale[off:off + B * C].view(B, C))
            off += B * C
        return out

    # -- forward ----------------------------------------------------------------------------------------------------
    def forward(self, x, return_features=False):
        _lib.require_cuda(x)
        if return_features:
            raise NotImplementedError("return_features=True is not used by the training step")
        if torch.is_grad_enabled():
            named = [(n, p) for n, p in _lib.named_params(self) if p.requires_grad]
            if named or x.requires_grad:
                return list(_DiscFn.apply(self, tuple(n for n, _ in named), x, *[p for _, p in named]))
        outs, _ = self._forward_impl(x, keep=False)
        return outs

    def _forward_impl(self, x, keep, need_wgrad=False):
        c = self._prepare()
        x = x.float().contiguous()
        B, _, H, W = x.shape
        dev = x.device
        dim = self.dim
        ctx = {"shape": (B, H, W)}
        drops = self._drop_all(B, dev)

        def cv(a, site, taps, Ho, Wo, act=ACT_NONE, stats=None, out_dtype=torch.bfloat16):
            return conv.conv_fprop(a, c[site], taps, Ho, Wo, bias=c["bias"][site][0], act=act, slope=LEAK, stats=stats,
                                   out_dtype=out_dtype, tile_w=_tile_w(taps, Ho, Wo))

        def gn_lrelu(z, st, gn, tag):
            N, Hh, Ww, C = z.shape
            coef = torch.empty((N, C, 2), device=dev, dtype=torch.float32)
            save = torch.empty((N, C, 2), device=dev, dtype=torch.float32)
            _lib.call("hwg_gn_coeffs", st.data_ptr(), gn.weight.data_ptr(), gn.bias.data_ptr(), N, C, gn.num_groups,
                      Hh * Ww, gn.eps, coef.data_ptr(), save.data_ptr(), _lib.stream())
            a = ops.scale_shift_act(z, coef, True, ACT_LRELU, LEAK, out=torch.empty_like(z) if keep else None)
            if keep:
                ctx[tag] = (z, coef, save)
            return a

        def conv_drop_lrelu(a, site, taps, Ho, Wo):
            """SN conv -> Dropout2d -> LeakyReLU: the channel scale rides on the activation pass (training), or the
            activation on the conv epilogue (eval)."""
            coef, scale = drops.get(site, (None, None))
            if coef is None:
                y = cv(a, site, taps, Ho, Wo, act=ACT_LRELU)
            else:
                y = ops.scale_shift_act(cv(a, site, taps, Ho, Wo), coef, True, ACT_LRELU, LEAK)
            if keep:
                ctx[site] = (a, y, scale)
            return y

        def pool(a, kh, kw):
            N, Hh, Ww, C = a.shape
            y = torch.empty((N, Hh // kh, Ww // kw, C), device=dev, dtype=torch.bfloat16)
            _lib.call("hwg_avgpool_nhwc", a.data_ptr(), y.data_ptr(), N, Hh, Ww, C, kh, kw, _lib.stream())
            return y

        st = torch.zeros((B, dim, 2), device=dev, dtype=torch.float32)
        if need_wgrad or dim != 64 or _NO_STEM:
            # the 'disc' lesson needs the shift expansion as the x operand of in_conv's weight gradient
            x7 = torch.empty((B, H, W, 16), device=dev, dtype=torch.bfloat16)
            _lib.call("hwg_shift_expand", x.data_ptr(), x7.data_ptr(), B, H, W, 7, 3, _lib.stream())
            if keep:
                ctx["x7"] = x7
            z0 = cv(x7, "in_conv.0", [(dy, 0) for dy in range(7)], H - 6, W, stats=st)
        else:
            # frozen discriminator ('gen' lessons, inference): the stem straight from the fp32 image (hwg_stem_conv reads
            # the same packed [7][64][16] operand)
            z0 = torch.empty((B, H - 6, W, dim), device=dev, dtype=torch.bfloat16)
            _lib.call("hwg_stem_conv", x.data_ptr(), c["in_conv.0"].data_ptr(), c["bias"]["in_conv.0"][0].data_ptr(),
                      B, H, W, 7, 7, 0, 3, dim, z0.data_ptr(), st.data_ptr(), _lib.stream())
        a = gn_lrelu(z0, st, self.in_conv[1], "gn0")                                  # [B,58,W,64]
        y1 = cv(a, "convs1.0", _T31, a.size(1) - 2, W, act=ACT_LRELU)               # [B,56,W,64]
        if keep:
            ctx["convs1.0"] = (a, y1, None)
        a = pool(y1, 2, 2)
        a = conv_drop_lrelu(a, "convs1.3", _T31, a.size(1) - 2, a.size(2))             # [B,26,W/2,128]
        ain = a
        y3 = cv(a, "convs2.0", _T31, a.size(1) - 2, a.size(2), act=ACT_LRELU)       # [B,24,W/2,128]
        if keep:
            ctx["convs2.0"] = (ain, y3, None)
        a = pool(y3, 2, 2)                                                             # [B,12,W/4,128]
        st = torch.zeros((B, 2 * dim, 2), device=dev, dtype=torch.float32)
        ain = a
        z4 = cv(a, "convs3.0", _T31, a.size(1) - 2, a.size(2), stats=st)               # [B,10,W/4,128]
        a = gn_lrelu(z4, st, self.convs3[1], "gn4")
        if keep:
            ctx["convs3.0"] = ain
        a = pool(a, 2, 2)                                                              # [B,5,W/8,128]
        mL = conv_drop_lrelu(a, "convs3.4", _T31, a.size(1) - 2, a.size(2))            # [B,3,W/8,256]
        outs = []
        if self.use_med:
            pM = cv(mL, "finalMed.0", _T31, mL.size(1) - 2, mL.size(2), out_dtype=torch.float32)   # [B,1,W/8,16]
            outs.append(pM[..., 0].reshape(B, -1))
        if self.use_low:
            a = conv_drop_lrelu(mL, "convs4.0", _T31, mL.size(1) - 2, mL.size(2))      # [B,1,W/8,128]
            a = pool(a, 1, 2)
            a = conv_drop_lrelu(a, "convs4.4", _T13, a.size(1), a.size(2))             # [B,1,W/16,256]
            a = conv_drop_lrelu(a, "convs4.7", _T13, a.size(1), a.size(2))
            a = pool(a, 1, 2)
            a = conv_drop_lrelu(a, "convs4.11", _T13, a.size(1), a.size(2))            # [B,1,W/32,256]
            pL = cv(a, "convs4.14", [(0, 0)], a.size(1), a.size(2), out_dtype=torch.float32)
            if keep:
                ctx["low_in"] = a
            outs.append(pL[..., 0].reshape(B, -1))
        if keep:
            ctx["mL"] = mL
            ctx["dgrad"], ctx["generation"] = c["dgrad"], self._generation
            ctx["out_shapes"] = [tuple(o.shape) for o in outs]
        return outs, ctx

    # -- backward ---------------------------------------------------------------------------------------------------
    def _wgrad_plan(self, dev):
        """Arena layout of one backward's accumulators (tap-major wgrad outputs, bias / GroupNorm sums, spectral-norm
        dots), the persistent flat buffer the parameter gradients are unpacked into, ONE hwg_linear_map table for the
        unpack (scaled by 1/sigma for the spectral-normalised layers) and the hwg_spectral_norm_bwd job table."""
        plan = self._plan.get("wgrad")
        if plan is not None:
            return plan
        layers = self.conv_layers()
        params = dict(_lib.named_params(self))
        inv_sigma, sn_sites = self._plan["inv_sigma"], [s for s, _, _, sp in layers if sp]
        off, slots = 0, {}

        def take(name, n):
            nonlocal off
            slots[name] = (off, n)
            off += -(-n // 4) * 4

        goff, gn_el = {}, 0

        def gslot(pname):
            nonlocal gn_el
            goff[pname] = gn_el
            gn_el += -(-params[pname].numel() // 4) * 4

        meta = {}
        for site, m, taps, spectral in layers:
            wname = site + (".module.weight_bar" if spectral else ".weight")
            bname = site + (".module.bias" if spectral else ".bias")
            co, ci = m.bias.numel(), (16 if site == "in_conv.0" else params[wname].size(1))
            cop = -(-co // 16) * 16
            take(("w", site), len(taps) * cop * ci)
            take(("b", site), cop)
            gslot(wname)
            gslot(bname)
            meta[site] = (wname, bname, co, cop, ci, taps, spectral)
        for gname in ("in_conv.1", "convs3.1"):
            C = params[gname + ".weight"].numel()
            take(("gamma", gname), C)
            take(("beta", gname), C)
            gslot(gname + ".weight")
            gslot(gname + ".bias")
        take("dots", len(sn_sites))
        gflat = torch.zeros(gn_el, device=dev, dtype=torch.float32)
        t = weightmap.JobTable()
        sn_jobs = np.zeros((len(sn_sites), 5), np.int64)
        for site, (wname, bname, co, cop, ci, taps, spectral) in meta.items():
            gw = gflat[goff[wname]:goff[wname] + params[wname].numel()]
            src = 4 * slots[("w", site)][0]
            if site == "in_conv.0":
                for dy in range(7):           # dw [7][64][16] -> gW[co, 0, dy, dx]
                    t.add(src + 4 * dy * cop * 16, gw[dy * 7:], R=co, C=7, s_r=16, s_c=1, d_r=49, d_c=1, M=np.eye(1))
            else:
                scale = inv_sigma[sn_sites.index(site):sn_sites.index(site) + 1] if spectral else None
                weightmap.map_conv_taps(co, ci, taps).add_unpack_wgrad(t, src, gw, co_rows=cop, scale_dev=scale)
            t.add(4 * slots[("b", site)][0], gflat[goff[bname]:goff[bname] + co], R=1, C=co, s_r=0, s_c=1, d_r=0, d_c=1,
                  M=np.eye(1))
            if spectral:
                i = sn_sites.index(site)
                mod = dict((s_, m_) for s_, m_, _, _ in layers)[site]
                w = mod.weight_bar
                sn_jobs[i] = (w.data_ptr(), gw.data_ptr(), mod.weight_u.data_ptr(), mod.weight_v.data_ptr(),
                              w.size(0) | (w[0].numel() << 32))
        for gname in ("in_conv.1", "convs3.1"):
            for kind, suffix in (("gamma", ".weight"), ("beta", ".bias")):
                C = params[gname + suffix].numel()
                t.add(4 * slots[(kind, gname)][0], gflat[goff[gname + suffix]:goff[gname + suffix] + C], R=1, C=C, s_r=0,
                      s_c=1, d_r=0, d_c=1, M=np.eye(1))
        t.finalize(dev)
        plan = self._plan["wgrad"] = dict(
            slots=slots, arena_floats=off, meta=meta, gflat=gflat, goff=goff, table=t,
            sn_jobs=torch.from_numpy(sn_jobs).to(dev), n_sn=len(sn_sites),
            sn_max=int(max((j[4] & 0xffffffff) * (j[4] >> 32) for j in sn_jobs)))
        return plan

    def _backward(self, ctx, grads, want_input, names):
        """grads: prediction gradients.  Returns (image gradient or None, {parameter name: gradient} for `names`)."""
        B, H, W = ctx["shape"]
        if ctx["generation"] != self._generation:
            raise RuntimeError("DiscriminatorAP: backward after another forward of the same module — the spectral-norm "
                               "operands of this graph have been re-packed (run backward before the next forward)")
        dg = ctx["dgrad"]
        dev = ctx["mL"].device
        want_w = bool(names)
        if want_w:
            wp = self._wgrad_plan(dev)
            arena = torch.zeros(wp["arena_floats"], device=dev, dtype=torch.float32)      # one memset

            def slot(key):
                o, n = wp["slots"][key]
                return arena[o:o + n]

        def collect(site, x_in, gz):
            """Weight and bias gradient of one convolution from its input and its output gradient."""
            if not want_w:
                return
            wname, bname, co, cop, ci, taps, _ = wp["meta"][site]
            conv.conv_wgrad(x_in, gz, taps, ci, cop, out=slot(("w", site)).view(len(taps), cop, ci))
            _lib.call("hwg_channel_sum", gz.data_ptr(), gz.numel() // gz.size(-1), gz.size(-1),
                      slot(("b", site)).data_ptr(), _lib.stream())

        def dgrad(g, site, Ho, Wo):
            wd, taps = dg[site]
            return conv.conv_fprop(g, wd, taps, Ho, Wo, tile_w=_tile_w(taps, Ho, Wo))

        def head_grad(g, like_w):
            """[B, Wp] fp32 prediction gradient -> [B,1,Wp,16] bf16 (channel 0 carries it)."""
            g16 = torch.zeros((B, 1, like_w, 16), device=dev, dtype=torch.bfloat16)
            g16[:, 0, :, 0] = g.reshape(B, like_w)
            return g16

        def act_bwd(g, site, kh=1, kw=1):
            a_in, y, scale = ctx[site]
            N, Hh, Ww, C = y.shape
            gz = torch.empty_like(y)
            _lib.call("hwg_act_bwd", g.data_ptr(), y.data_ptr(), _lib.ptr(scale), LEAK, N, Hh, Ww, C, kh, kw,
                      gz.data_ptr(), _lib.stream())
            return gz, a_in

        def gn_bwd(g, tag, gn, kh, kw, gname):
            z, coef, save = ctx[tag]
            N, Hh, Ww, C = z.shape
            sums = torch.zeros((N, C, 2), device=dev, dtype=torch.float32)
            spq = torch.empty((N, C, 3), device=dev, dtype=torch.float32)
            gz = torch.empty_like(z)
            s = _lib.stream()
            _lib.call("hwg_norm_bwd_reduce", g.data_ptr(), z.data_ptr(), coef.data_ptr(), LEAK, N, Hh, Ww, C, kh, kw,
                      sums.data_ptr(), s)
            _lib.call("hwg_gn_bwd_coeffs", sums.data_ptr(), save.data_ptr(), gn.weight.data_ptr(), N, C, gn.num_groups,
                      Hh * Ww, spq.data_ptr(), slot(("gamma", gname)).data_ptr() if want_w else None,
                      slot(("beta", gname)).data_ptr() if want_w else None, s)
            _lib.call("hwg_norm_bwd_apply", g.data_ptr(), z.data_ptr(), coef.data_ptr(), spq.data_ptr(), LEAK, N, Hh, Ww,
                      C, kh, kw, gz.data_ptr(), s)
            return gz

        mL = ctx["mL"]
        g_mL = None
        gi = 0
        if self.use_med:
            g16 = head_grad(grads[gi], mL.size(2))
            collect("finalMed.0", mL, g16)
            g_mL = dgrad(g16, "finalMed.0", mL.size(1), mL.size(2))
            gi += 1
        if self.use_low:
            a = ctx["low_in"]
            g16 = head_grad(grads[gi], a.size(2))
            collect("convs4.14", a, g16)
            g = dgrad(g16, "convs4.14", 1, a.size(2))
            gz, a_in = act_bwd(g, "convs4.11")
            collect("convs4.11", a_in, gz)
            g = dgrad(gz, "convs4.11", 1, a_in.size(2))                    # gradient of the pooled tensor
            gz, a_in = act_bwd(g, "convs4.7", 1, 2)
            collect("convs4.7", a_in, gz)
            g = dgrad(gz, "convs4.7", 1, a_in.size(2))
            gz, a_in = act_bwd(g, "convs4.4")
            collect("convs4.4", a_in, gz)
            g = dgrad(gz, "convs4.4", 1, a_in.size(2))                     # pooled again
            gz, a_in = act_bwd(g, "convs4.0", 1, 2)
            collect("convs4.0", a_in, gz)
            g_low = dgrad(gz, "convs4.0", mL.size(1), mL.size(2))
            g_mL = g_low if g_mL is None else g_mL.add_(g_low)
        gz, a_in = act_bwd(g_mL, "convs3.4")
        collect("convs3.4", a_in, gz)
        g = dgrad(gz, "convs3.4", a_in.size(1), a_in.size(2))              # [B,5,W/8,128], gradient of pool(a4)
        gz = gn_bwd(g, "gn4", self.convs3[1], 2, 2, "convs3.1")
        a_in = ctx["convs3.0"]
        collect("convs3.0", a_in, gz)
        g = dgrad(gz, "convs3.0", a_in.size(1), a_in.size(2))              # [B,12,W/4,128], gradient of pool(y3)
        gz, a_in = act_bwd(g, "convs2.0", 2, 2)
        collect("convs2.0", a_in, gz)
        g = dgrad(gz, "convs2.0", a_in.size(1), a_in.size(2))
        gz, a_in = act_bwd(g, "convs1.3")
        collect("convs1.3", a_in, gz)
        g = dgrad(gz, "convs1.3", a_in.size(1), a_in.size(2))              # [B,28,W/2,64], gradient of pool(y1)
        gz, a_in = act_bwd(g, "convs1.0", 2, 2)
        collect("convs1.0", a_in, gz)
        g = dgrad(gz, "convs1.0", a_in.size(1), a_in.size(2))              # [B,58,W,64]
        gz = gn_bwd(g, "gn0", self.in_conv[1], 1, 1, "in_conv.1")
        collect("in_conv.0", ctx.get("x7"), gz)
        dimg = None
        if want_input:
            g7 = dgrad(gz, "in_conv.0", H, W)                              # [B,64,W,16]
            dimg = torch.empty((B, 1, H, W), device=dev, dtype=torch.float32)
            _lib.call("hwg_shift_collapse", g7.data_ptr(), dimg.data_ptr(), B, H, W, 7, 3, 0, _lib.stream())
        pgrads = {}
        if want_w:
            # every parameter gradient in its own layout by ONE launch (spectral layers scaled by 1/sigma), then the
            # rank-one spectral-norm term for all layers
            wp["table"].run(src_base=arena)
            _lib.call("hwg_spectral_norm_bwd", wp["sn_jobs"].data_ptr(), wp["n_sn"], wp["sn_max"],
                      self._plan["inv_sigma"].data_ptr(), slot("dots").data_ptr(), _lib.stream())
            params = dict(self.named_parameters())
            gout = wp["gflat"].clone()      # autograd may adopt a returned gradient as .grad: never hand out the workspace
            for n in names:
                o = wp["goff"][n]
                pgrads[n] = gout[o:o + params[n].numel()].view_as(params[n])
        return dimg, pgrads


class _DiscFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, m, names, x, *params):
        with torch.no_grad():
            outs, saved = m._forward_impl(x, keep=True, need_wgrad=bool(names))
        ctx.m, ctx.names, ctx.saved = m, names, saved
        ctx.x_needs_grad = x.requires_grad
        return tuple(o.detach() for o in outs)       # detached aliases: no output -> node -> ctx -> output cycle

    @staticmethod
    def backward(ctx, *grads):
        _lib.saved_state(ctx.saved)
        dev = ctx.saved["mL"].device
        grads = [torch.zeros(shp, device=dev) if g is None else g.contiguous().float()
                 for g, shp in zip(grads, ctx.saved["out_shapes"])]
        with torch.no_grad():
            dimg, pg = ctx.m._backward(ctx.saved, grads, ctx.x_needs_grad, ctx.names)
        if not _lib.RETAIN_SAVED:
            ctx.saved = None
        return (None, None, dimg) + tuple(pg[n] for n in ctx.names)
