"""Drop-in for the reference perceptual encoder `Encoder2` (model/autoencoder.py:341-410; `encoder_type: "2tight"` ->
`Encoder2(32)`, trainer/hw_with_style_trainer.py:148-149) and the perceptual loss of the 'auto' lessons built on it
(trainer :724-748) — SURVEY.md §8 row f1, second half.

GPU parity against the goldens of the unmodified reference: tests/test_enc_gpu.py; the module runs inside bench.py's
headline step (the perceptual loss of the 'auto' lesson).

Same constructor signature, module names, construction order (same seed -> same initial weights) and `state_dict` keys
as the reference.  The torch sub-modules are parameter containers; the forward runs on libhwg_b200:

* `down_conv1[0]` (5x5, ONE input channel, padding 2) by `hwg_stem_conv` straight from the fp32 image (round 1: a 5-tap
  implicit GEMM over the 16-channel shift expansion of the image; its input gradient still is, with `hwg_shift_collapse`), the other ten convolutions on
  `hwg_conv_fprop` (staged-tile kernel for the 16/32-channel layers, tcgen05 for >= 64), the (6,3) head as two 9-tap
  launches (HWG_MAX_TAPS = 16);
* GroupNorm statistics from the producing kernel's epilogue (`stats`) or from `hwg_add_stats` for the two residual sums,
  GroupNorm [+ Dropout2d] + ReLU as ONE scale-shift pass (`hwg_gn_coeffs`, the Dropout2d keep-mask / (1-p) folded into the
  per-(sample, channel) coefficients: relu(s * v) = s * relu(v) for s >= 0), `hwg_avgpool_nhwc`;
* the in-place ReLU that opens `conv1` acts on the tensor the residual aliases (autoencoder.py:399-401), so the residual
  is the ReLU'd tensor: the ReLU rides on `down_conv1[4]`'s epilogue;
* backward to the INPUT image only (the encoder is never optimised: it is in no optimizer of the trainer): dgrad on the
  same convolution kernels, `hwg_norm_bwd_reduce` / `hwg_gn_bwd_coeffs` / `hwg_norm_bwd_apply` (GroupNorm + ReLU
  [+ AvgPool2d] backward; a Dropout2d scale enters as s on the (n, c) sums and on the direct term), `hwg_act_bwd`,
  `hwg_shift_collapse`;
* `perceptual_loss(image, recon)` = trainer :727-748 as one call: both images through the encoder as one batch, the L1
  between the halves of both returned feature tensors by `hwg_l1_halves` (loss and its sign gradient in one pass), and a
  backward that only walks the RECON half of the batch (the other half is data).

There is no PyTorch fallback."""
import numpy as np
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, conv, ops, weightmap
from ._lib import ACT_NONE, ACT_RELU
from .discriminator_ap import get_group_size

# Dropout2d sites in forward order: (name of the GroupNorm they follow, channels, p)
DROPOUT_SITES = (("conv1.2", 32, 0.1), ("conv2.0", 64, 0.1), ("conv2.4", 64, 0.1), ("down_conv3.4", 128, 0.1))
_T33 = conv.conv_taps(3, 3, 1, 1)          # 3x3, padding 1
_T33V = conv.conv_taps(3, 3, 0, 0)         # 3x3, no padding
_T11 = [(0, 0)]
_NO_STEM = bool(os.environ.get("HWG_NO_STEM_CONV"))
_T5 = [(dy - 2, 0) for dy in range(5)]     # rows of the 5x5 kernel; the columns are channels of the shift expansion


class Encoder2(nn.Module):
    def __init__(self, out_dim=256):
        super().__init__()
        if out_dim % 16 != 0:
            raise NotImplementedError("out_dim must be a multiple of 16")
        gn = lambda c: nn.GroupNorm(get_group_size(c), c)                                  # noqa: E731
        self.out_dim = out_dim
        self.down_conv1 = nn.Sequential(nn.Conv2d(1, 32, 5, padding=2), gn(32), nn.ReLU(True), nn.AvgPool2d(2),
                                        nn.Conv2d(32, 32, 1))
        self.conv1 = nn.Sequential(nn.ReLU(True), nn.Conv2d(32, 32, 3, padding=1), gn(32), nn.Dropout2d(0.1, True),
                                   nn.ReLU(True), nn.Conv2d(32, 32, 3, padding=1))
        self.down_conv2 = nn.Sequential(gn(32), nn.ReLU(True), nn.AvgPool2d(2), nn.Conv2d(32, 64, 1))
        self.conv2 = nn.Sequential(gn(64), nn.Dropout2d(0.1, True), nn.ReLU(True), nn.Conv2d(64, 64, 3, padding=1), gn(64),
                                   nn.Dropout2d(0.1, True), nn.ReLU(True), nn.Conv2d(64, 64, 3, padding=1))
        self.down_conv3 = nn.Sequential(gn(64), nn.ReLU(True), nn.AvgPool2d(2), nn.Conv2d(64, 128, 3), gn(128),
                                        nn.Dropout2d(0.1, True), nn.ReLU(True), nn.Conv2d(128, out_dim, (6, 3)))
        self._plan, self._plan_key = None, None
        self.dropout_masks = None      # tests: list of four [N,C] 0/1 keep-masks (DROPOUT_SITES order) instead of drawing

    # -- layers ---------------------------------------------------------------------------------------------------
    def conv_layers(self):
        """(site, conv module, forward taps) of the convolutions that are ONE hwg_conv_fprop launch, forward order.
        `down_conv1.0` (shift expansion) and `down_conv3.7` (two launches) are planned separately."""
        return [("down_conv1.4", self.down_conv1[4], _T11), ("conv1.1", self.conv1[1], _T33),
                ("conv1.5", self.conv1[5], _T33), ("down_conv2.3", self.down_conv2[3], _T11),
                ("conv2.3", self.conv2[3], _T33), ("conv2.7", self.conv2[7], _T33),
                ("down_conv3.3", self.down_conv3[3], _T33V)]

    def build_table(self, dev):
        """ONE hwg_linear_map job table for every packed operand (forward [taps][Cout][Cin] and dgrad [taps][Cin][Cout],
        bf16).  Returns (table, operands) with operands[site] = (forward operand, dgrad operand, dgrad taps)."""
        t = weightmap.JobTable()
        c = {}
        w = self.down_conv1[0].weight                                  # [32,1,5,5]
        co = w.size(0)
        f = torch.empty((5, co, 16), device=dev, dtype=torch.bfloat16)
        d = torch.empty((5, 16, co), device=dev, dtype=torch.bfloat16)
        for dy in range(5):   # kernel row dy = tap (dy-2, 0); kernel column dx = channel dx of the shift expansion
            t.add(w[:, 0, dy, :], f[dy], R=co, C=5, Cp=16, s_r=25, s_c=1, d_r=16, d_c=1, M=np.eye(1), dst_bf16=True)
            t.add(w[:, 0, dy, :], d[dy], R=5, C=co, Rp=16, Cp=co, s_r=1, s_c=25, d_r=co, d_c=1, M=np.eye(1), dst_bf16=True)
        c["down_conv1.0"] = (f, d, [(-dh, 0) for dh, _ in _T5])
        for site, m, taps in self.conv_layers():
            co, ci = m.weight.size(0), m.weight.size(1)
            mp = weightmap.map_conv_taps(co, ci, taps)
            f = torch.empty((mp.Tf, co, mp.Cip), device=dev, dtype=torch.bfloat16)
            d = torch.empty(mp.dgrad_shape(), device=dev, dtype=torch.bfloat16)
            mp.add_pack_fwd(t, m.weight, f)
            mp.add_pack_dgrad(t, m.weight, d)
            c[site] = (f, d, mp.taps_d)
        # (6,3) head: kernel rows 0-2 and 3-5 as two 9-tap operands; within the [co,ci,6,3] parameter each half is nine
        # contiguous kernel positions with channel strides 18*ci / 18
        w = self.down_conv3[7].weight
        co, ci = w.size(0), w.size(1)
        flat = w.detach().view(-1)
        halves = []
        for half in range(2):
            mp = weightmap.map_conv_taps(co, ci, _T33V)
            mp.s_co, mp.s_ci = ci * 18, 18
            f = torch.empty((9, co, mp.Cip), device=dev, dtype=torch.bfloat16)
            d = torch.empty(mp.dgrad_shape(), device=dev, dtype=torch.bfloat16)
            src = flat[9 * half:]
            mp.add_pack_fwd(t, src, f)
            mp.add_pack_dgrad(t, src, d)
            halves.append((f, d, mp.taps_d))
        c["down_conv3.7"] = halves
        t.finalize(dev)
        return t, c

    def _prepare(self):
        """Packed operands, re-derived by one launch when a parameter changed (data_ptr / _version)."""
        key = tuple((p.data_ptr(), p._version) for p in _lib.params(self))
        if self._plan is None or self._plan_key != key:
            dev = self.down_conv1[0].weight.device
            if self._plan is None or self._plan["ptrs"] != tuple(k[0] for k in key):
                table, c = self.build_table(dev)
                self._plan = {"table": table, "c": c, "ptrs": tuple(k[0] for k in key)}
            self._plan["table"].run()
            self._plan_key = key
        return self._plan["c"]

    def _drop_scales(self, N, dev):
        """Dropout2d(0.1, inplace) keep-mask / (1-p) per (sample, channel) for the four sites; None in eval mode."""
        if not self.training:
            return [None] * len(DROPOUT_SITES)
        if self.dropout_masks is not None:
            return [m.to(dev).float().reshape(N, C) / (1.0 - p) for m, (_, C, p) in zip(self.dropout_masks, DROPOUT_SITES)]
        tot = sum(C for _, C, _ in DROPOUT_SITES)
        p = DROPOUT_SITES[0][2]                                           # the same p at every site
        s = (torch.rand((N, tot), device=dev) >= p).float().mul_(1.0 / (1.0 - p))
        out, off = [], 0
        for _, C, _ in DROPOUT_SITES:
            out.append(s[:, off:off + C].contiguous())
            off += C
        return out

    # -- forward ----------------------------------------------------------------------------------------------------
    def forward(self, x):
        """x [N,1,64,W] -> (features [N,out_dim,1,W/8-4], mid_features [N,64,16,W/4]) fp32, as the reference."""
        _lib.require_cuda(x)
        if torch.is_grad_enabled() and x.requires_grad:
            return _EncFn.apply(self, x)
        feat, mid, _ = self._forward_impl(x, keep=False)
        return feat.permute(0, 3, 1, 2), mid.permute(0, 3, 1, 2).float()

    def _forward_impl(self, x, keep):
        c = self._prepare()
        x = x.float().contiguous()
        N, one, H, W = x.shape
        if one != 1 or H != 64 or W % 8 != 0 or W < 40:
            raise NotImplementedError(f"Encoder2: input {tuple(x.shape)}: expected [N,1,64,W], W >= 40 a multiple of 8")
        dev = x.device
        drops = self._drop_scales(N, dev)
        ctx = {"shape": (N, H, W)}
        s = _lib.stream

        def zstats(C):
            return torch.zeros((N, C, 2), device=dev, dtype=torch.float32)

        def cv(a, site, taps, Ho, Wo, m, act=ACT_NONE, stats=None):
            return conv.conv_fprop(a, c[site][0], taps, Ho, Wo, bias=m.bias.detach(), act=act, stats=stats)

        def gn_relu(z, st, gn, tag, drop=None):
            """GroupNorm [-> Dropout2d] -> ReLU as one scale-shift pass; keeps (z, coef', save, drop) for backward."""
            Nn, Hh, Ww, C = z.shape
            coef = torch.empty((Nn, C, 2), device=dev, dtype=torch.float32)
            save = torch.empty((Nn, C, 2), device=dev, dtype=torch.float32)
            _lib.call("hwg_gn_coeffs", st.data_ptr(), gn.weight.data_ptr(), gn.bias.data_ptr(), Nn, C, gn.num_groups,
                      Hh * Ww, gn.eps, coef.data_ptr(), save.data_ptr(), s())
            if drop is not None:
                coef.mul_(drop[:, :, None])
            a = ops.scale_shift_act(z, coef, True, ACT_RELU, 0.0, out=torch.empty_like(z))
            if keep:
                ctx[tag] = (z, coef, save, drop)
            return a

        def pool(a):
            Nn, Hh, Ww, C = a.shape
            y = torch.empty((Nn, Hh // 2, Ww // 2, C), device=dev, dtype=torch.bfloat16)
            _lib.call("hwg_avgpool_nhwc", a.data_ptr(), y.data_ptr(), Nn, Hh, Ww, C, 2, 2, s())
            return y

        def add(a, b, stats):
            Nn, Hh, Ww, C = a.shape
            y = torch.empty_like(a)
            _lib.call("hwg_add_stats", a.data_ptr(), b.data_ptr(), y.data_ptr(), Nn, Hh * Ww, C, _lib.ptr(stats), s())
            return y

        # down_conv1 (:344-351): 5x5 conv -> GroupNorm -> ReLU -> AvgPool2d(2) -> 1x1 conv [-> the ReLU opening conv1]
        st = zstats(32)
        if _NO_STEM:       # development A/B switch: the round-1 route (5-tap implicit GEMM over the shift expansion)
            x5 = torch.empty((N, H, W, 16), device=dev, dtype=torch.bfloat16)
            _lib.call("hwg_shift_expand", x.data_ptr(), x5.data_ptr(), N, H, W, 5, 2, s())
            z0 = cv(x5, "down_conv1.0", _T5, H, W, self.down_conv1[0], stats=st)
        else:
            z0 = torch.empty((N, H, W, 32), device=dev, dtype=torch.bfloat16)                  # [N,64,W,32]
            _lib.call("hwg_stem_conv", x.data_ptr(), c["down_conv1.0"][0].data_ptr(), self.down_conv1[0].bias.data_ptr(),
                      N, H, W, 5, 5, 2, 2, 32, z0.data_ptr(), st.data_ptr(), s())
        p0 = pool(gn_relu(z0, st, self.down_conv1[1], "gn0"))                                  # [N,32,W/2,32]
        r1 = cv(p0, "down_conv1.4", _T11, H // 2, W // 2, self.down_conv1[4], act=ACT_RELU)    # res = ReLU(x), aliased
        # conv1 (:354-362) + residual (:399-401)
        st = zstats(32)
        z1 = cv(r1, "conv1.1", _T33, H // 2, W // 2, self.conv1[1], stats=st)
        a1 = gn_relu(z1, st, self.conv1[2], "gn1", drops[0])
        y1 = cv(a1, "conv1.5", _T33, H // 2, W // 2, self.conv1[5])
        st = zstats(32)
        x1 = add(y1, r1, st)                                                                   # [N,32,W/2,32]
        # down_conv2 (:364-369)
        p1 = pool(gn_relu(x1, st, self.down_conv2[0], "gn2"))                                  # [N,16,W/4,32]
        st = zstats(64)
        r2 = cv(p1, "down_conv2.3", _T11, H // 4, W // 4, self.down_conv2[3], stats=st)        # res (no aliasing ReLU here)
        # conv2 (:371-380) + residual (:403-405)
        a2 = gn_relu(r2, st, self.conv2[0], "gn3", drops[1])
        st = zstats(64)
        z2 = cv(a2, "conv2.3", _T33, H // 4, W // 4, self.conv2[3], stats=st)
        a3 = gn_relu(z2, st, self.conv2[4], "gn4", drops[2])
        y2 = cv(a3, "conv2.7", _T33, H // 4, W // 4, self.conv2[7])
        st = zstats(64)
        mid = add(y2, r2, st)                                                                  # [N,16,W/4,64]
        # down_conv3 (:382-394)
        p2 = pool(gn_relu(mid, st, self.down_conv3[0], "gn5"))                                 # [N,8,W/8,64]
        st = zstats(128)
        z3 = cv(p2, "down_conv3.3", _T33V, H // 8 - 2, W // 8 - 2, self.down_conv3[3], stats=st)   # [N,6,W/8-2,128]
        a4 = gn_relu(z3, st, self.down_conv3[4], "gn6", drops[3])
        (fa, _, _), (fb, _, _) = c["down_conv3.7"]
        Wf = W // 8 - 4
        head = self.down_conv3[7]
        feat = conv.conv_fprop(a4, fa, _T33V, 1, Wf, bias=head.bias.detach(), out_dtype=torch.float32)
        feat.add_(conv.conv_fprop(a4, fb, [(dh + 3, dw) for dh, dw in _T33V], 1, Wf, out_dtype=torch.float32))
        if keep:
            ctx.update(p0=p0, r1=r1, a1=a1, p1=p1, a2=a2, a3=a3, p2=p2, a4=a4, c=c)
        return feat, mid, ctx                                            # [N,1,Wf,out_dim] fp32, [N,16,W/4,64] bf16

    # -- backward ---------------------------------------------------------------------------------------------------
    def _backward(self, ctx, g_feat, g_mid, lo=0):
        """g_feat [n,1,Wf,out_dim] bf16 and g_mid [n,16,W/4,64] bf16 (or None): gradients of the two outputs for samples
        lo..lo+n of the forward's batch.  Returns the image gradient [n,1,64,W] fp32."""
        N, H, W = ctx["shape"]
        c = ctx["c"]
        n = g_feat.size(0)
        hi = lo + n
        dev = g_feat.device
        s = _lib.stream

        def sl(t):
            return None if t is None else t[lo:hi]

        def dgrad(g, site, Ho, Wo):
            _, wd, taps = c[site]
            return conv.conv_fprop(g, wd, taps, Ho, Wo)

        def gn_bwd(g, tag, gn, k):
            """Backward of a = ReLU(s * GroupNorm(z)) [-> AvgPool2d(k)]; g is the gradient of the (pooled) output."""
            z, coef, save, drop = (sl(t) for t in ctx[tag])
            Nn, Hh, Ww, C = z.shape
            sums = torch.zeros((Nn, C, 2), device=dev, dtype=torch.float32)
            spq = torch.empty((Nn, C, 3), device=dev, dtype=torch.float32)
            gz = torch.empty_like(z)
            _lib.call("hwg_norm_bwd_reduce", g.data_ptr(), z.data_ptr(), coef.data_ptr(), 0.0, Nn, Hh, Ww, C, k, k,
                      sums.data_ptr(), s())
            if drop is not None:                 # d/d(GroupNorm output) = s * gy': s enters the sums ...
                sums.mul_(drop[:, :, None])
            _lib.call("hwg_gn_bwd_coeffs", sums.data_ptr(), save.data_ptr(), gn.weight.data_ptr(), Nn, C, gn.num_groups,
                      Hh * Ww, spq.data_ptr(), None, None, s())
            if drop is not None:                 # ... and the direct term sc * gy'
                spq[:, :, 0].mul_(drop)
            _lib.call("hwg_norm_bwd_apply", g.data_ptr(), z.data_ptr(), coef.data_ptr(), spq.data_ptr(), 0.0, Nn, Hh, Ww,
                      C, k, k, gz.data_ptr(), s())
            return gz

        def add(a, b):
            Nn, Hh, Ww, C = a.shape
            _lib.call("hwg_add_stats", a.data_ptr(), b.data_ptr(), a.data_ptr(), Nn, Hh * Ww, C, None, s())
            return a

        a4, p2 = sl(ctx["a4"]), sl(ctx["p2"])
        Wp = a4.size(2)
        # (6,3) head: input rows 0-2 from the first half of the kernel rows, rows 3-5 from the second
        g_a4 = torch.empty_like(a4)                                                          # [n,6,Wp,128]
        C4 = a4.size(3)
        for half, (_, wd, taps) in enumerate(c["down_conv3.7"]):
            conv.conv_fprop(g_feat, wd, taps, 3, Wp,
                            out_view=(g_a4, 6 * Wp * C4, Wp * C4, C4, 3 * half * Wp * C4))
        gz = gn_bwd(g_a4, "gn6", self.down_conv3[4], 1)
        g = dgrad(gz, "down_conv3.3", p2.size(1), p2.size(2))                                # [n,8,W/8,64]
        g_m = gn_bwd(g, "gn5", self.down_conv3[0], 2)                                        # [n,16,W/4,64]
        if g_mid is not None:
            g_m = add(g_m, g_mid)
        # conv2 + residual: mid = conv2(r2) + r2
        g = dgrad(g_m, "conv2.7", g_m.size(1), g_m.size(2))
        gz = gn_bwd(g, "gn4", self.conv2[4], 1)
        g = dgrad(gz, "conv2.3", g_m.size(1), g_m.size(2))
        g_r2 = add(gn_bwd(g, "gn3", self.conv2[0], 1), g_m)
        p1 = sl(ctx["p1"])
        g = dgrad(g_r2, "down_conv2.3", p1.size(1), p1.size(2))                              # [n,16,W/4,32]
        g_x1 = gn_bwd(g, "gn2", self.down_conv2[0], 2)                                       # [n,32,W/2,32]
        # conv1 + residual: x1 = conv1(r1) + r1, r1 = ReLU(down_conv1[4] output)
        g = dgrad(g_x1, "conv1.5", g_x1.size(1), g_x1.size(2))
        gz = gn_bwd(g, "gn1", self.conv1[2], 1)
        g_r1 = add(dgrad(gz, "conv1.1", g_x1.size(1), g_x1.size(2)), g_x1)
        r1 = sl(ctx["r1"])
        gz = torch.empty_like(r1)
        _lib.call("hwg_act_bwd", g_r1.data_ptr(), r1.data_ptr(), None, 0.0, n, r1.size(1), r1.size(2), r1.size(3), 1, 1,
                  gz.data_ptr(), s())
        g = dgrad(gz, "down_conv1.4", r1.size(1), r1.size(2))                                # [n,32,W/2,32]
        gz = gn_bwd(g, "gn0", self.down_conv1[1], 2)                                         # [n,64,W,32]
        g5 = dgrad(gz, "down_conv1.0", H, W)                                                 # [n,64,W,16]
        dimg = torch.empty((n, 1, H, W), device=dev, dtype=torch.float32)
        _lib.call("hwg_shift_collapse", g5.data_ptr(), dimg.data_ptr(), n, H, W, 5, 2, 0, s())
        return dimg

    # -- the trainer's perceptual loss ------------------------------------------------------------------------------
    def perceptual_loss(self, image, recon):
        """trainer :727-748: pads the narrower of the two images (and both to >= 40 px), runs both through the encoder as
        one batch and sums the L1 losses between the halves of the two feature tensors.  Differentiable w.r.t. recon."""
        _lib.require_cuda(image, recon)
        if image.size(3) > recon.size(3):
            diff = image.size(3) - recon.size(3)
            recon = F.pad(recon, (diff // 2, diff // 2 + diff % 2))
        elif image.size(3) < recon.size(3):
            diff = recon.size(3) - image.size(3)
            image = F.pad(image, (diff // 2, diff // 2 + diff % 2))
        if image.size(3) < 40:
            diff = 40 - image.size(3)
            image, recon = (F.pad(t, (diff // 2, diff // 2 + diff % 2)) for t in (image, recon))
        return _PerceptualFn.apply(self, image.detach(), recon)


class _EncFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, m, x):
        with torch.no_grad():
            feat, mid, saved = m._forward_impl(x, keep=True)
        ctx.m, ctx.saved = m, saved
        return feat.permute(0, 3, 1, 2), mid.permute(0, 3, 1, 2).float()

    @staticmethod
    def backward(ctx, g_feat, g_mid):
        N, H, W = _lib.saved_state(ctx.saved)["shape"]
        dev = ctx.saved["a4"].device

        def nhwc(g, shape):
            if g is None:
                return torch.zeros(shape, device=dev, dtype=torch.bfloat16)
            return g.permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()

        with torch.no_grad():
            gf = nhwc(g_feat, (N, 1, W // 8 - 4, ctx.m.out_dim))
            gm = None if g_mid is None else nhwc(g_mid, None)
            dimg = ctx.m._backward(ctx.saved, gf, gm, 0)
        if not _lib.RETAIN_SAVED:
            ctx.saved = None
        return None, dimg


class _PerceptualFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, m, image, recon):
        B = image.size(0)
        with torch.no_grad():
            feat, mid, saved = m._forward_impl(torch.cat((image.float(), recon.detach().float()), 0),
                                               keep=recon.requires_grad)
            dev = feat.device
            loss = torch.zeros((), device=dev, dtype=torch.float32)
            g_feat = g_mid = None
            if recon.requires_grad:
                g_feat = torch.empty((B,) + tuple(feat.shape[1:]), device=dev, dtype=torch.bfloat16)
                g_mid = torch.empty((B,) + tuple(mid.shape[1:]), device=dev, dtype=torch.bfloat16)
            for f, g, dt in ((feat, g_feat, _lib.DT_F32), (mid, g_mid, _lib.DT_BF16)):
                half = f.numel() // 2
                _lib.call("hwg_l1_halves", f.data_ptr(), dt, half, 1.0 / half, 1.0 / half, loss.data_ptr(), _lib.ptr(g),
                          _lib.stream())
        ctx.m, ctx.saved, ctx.B, ctx.g = m, saved, B, (g_feat, g_mid)
        return loss

    @staticmethod
    def backward(ctx, g_loss):
        with torch.no_grad():
            dimg = ctx.m._backward(_lib.saved_state(ctx.saved), ctx.g[0], ctx.g[1], ctx.B)      # the recon half only
            dimg.mul_(g_loss)                                                # the chain is linear in the loss gradient
        if not _lib.RETAIN_SAVED:
            ctx.saved = ctx.g = None
        return None, None, dimg
