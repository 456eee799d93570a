"""Drop-in for the reference spacer `CountCNN` (model/count_cnn.py:7-45; built by model/hw_with_style.py:200-204 as
CountCNN(num_class, style_dim, 128, 2) for `spacer: "CNN duplicates"`) — SURVEY.md §8 row f4 / a1: the text -> spacing
front end of `HWWithStyle.forward` (hw_with_style.py:236-237), trained by the 'count' lesson.

Same constructor signature, module names, construction order (same seed -> same initial weights) and `state_dict` keys as
the reference.  The torch sub-modules are parameter containers; forward and backward run on libhwg_b200:

* the input `cat(onehot text, style broadcast over the positions)` (count_cnn.py:35-37) is packed by `hwg_gen_pack_input`
  (the generator's input kernel) into one NHWC bf16 line [B,1,L,Cp];
* the three `Conv1d(k=3, padding=1)` are 3-tap launches of `hwg_conv_fprop` whose epilogue also emits the per-(sample,
  channel) sums of the GroupNorm that follows; GroupNorm -> Dropout2d -> ReLU is ONE scale-shift pass (`hwg_gn_coeffs`,
  the keep-mask / (1-p) folded into the per-(sample, channel) coefficients), the 1x1 head writes fp32;
* `output * std + mean` (:45) stays two broadcast ops of torch autograd on the [L,B,n_out] result (they carry the
  gradients of the `mean` / `std` parameters);
* backward: input gradients by the same convolution kernel on the transposed operands, `hwg_norm_bwd_reduce` /
  `hwg_gn_bwd_coeffs` / `hwg_norm_bwd_apply` for GroupNorm + Dropout2d + ReLU (they also emit the GroupNorm affine
  gradients), `hwg_conv_wgrad` + `hwg_channel_sum` for weights and biases, ONE `hwg_linear_map` launch that unpacks every
  parameter gradient into its own layout, and the gradients of the text input and of the style vector (the column sums of
  the packed-input gradient).

The reference's `assert(not torch.isnan(...))` host synchronisations (:41-43) are not reproduced.
There is no PyTorch fallback."""
import numpy as np
import torch
import torch.nn as nn

from . import _lib, conv, ops, weightmap
from ._lib import ACT_RELU
from .discriminator_ap import get_group_size

_T3 = conv.conv_taps(1, 3, 0, 1)          # Conv1d(k=3, padding=1) on a [B,1,L,C] line
_T1 = [(0, 0)]
DROPOUT_P = 0.1


class CountCNN(nn.Module):
    def __init__(self, class_size, style_size, hidden_size=128, n_out=1, emb_style=0):
        super().__init__()
        if hidden_size % 64 != 0 or hidden_size // 4 < 32:
            raise NotImplementedError("CountCNN: hidden_size must be a multiple of 128")
        if n_out > 16:
            raise NotImplementedError("CountCNN: n_out <= 16")
        h = hidden_size
        gn = lambda c: nn.GroupNorm(get_group_size(c), c)                                  # noqa: E731
        self.cnn = nn.Sequential(
            nn.Conv1d(class_size + style_size, h, kernel_size=3, stride=1, padding=1), gn(h), nn.Dropout2d(DROPOUT_P),
            nn.ReLU(inplace=True),
            nn.Conv1d(h, h // 2, kernel_size=3, stride=1, padding=1), gn(h // 2), nn.Dropout2d(DROPOUT_P),
            nn.ReLU(inplace=True),
            nn.Conv1d(h // 2, h // 4, kernel_size=3, stride=1, padding=1), gn(h // 4), nn.ReLU(inplace=True),
            nn.Conv1d(h // 4, n_out, kernel_size=1, stride=1, padding=0))
        if n_out == 1 or n_out > 2:
            self.mean = nn.Parameter(torch.FloatTensor(1, n_out).fill_(2))
            self.std = nn.Parameter(torch.FloatTensor(1, n_out).fill_(1))
        else:
            self.mean = nn.Parameter(torch.FloatTensor([2.0, 0.0]))
            self.std = nn.Parameter(torch.FloatTensor([1.5, 0.5]))
        self.class_size, self.style_size, self.n_out = class_size, style_size, n_out
        self.cin_pad = -(-(class_size + style_size) // 64) * 64
        self._plan, self._plan_key = None, None
        self.dropout_masks = None      # tests: two [B,C] 0/1 keep-masks (after cnn.1 and cnn.5) instead of drawing

    # (site, conv index, GroupNorm index or None, dropout?, taps)
    LAYERS = (("cnn.0", 0, 1, True, _T3), ("cnn.4", 4, 5, True, _T3), ("cnn.8", 8, 9, False, _T3), ("cnn.11", 11, None, False, _T1))

    # -- derived operands ------------------------------------------------------------------------------------------
    def _prepare(self):
        """Packed bf16 operands (forward [taps][Cout][Cin] and dgrad [taps][Cin][Cout]) of the four convolutions, re-derived
        by ONE hwg_linear_map launch when a parameter changed (data_ptr / _version)."""
        convs = [self.cnn[i] for _, i, _, _, _ in self.LAYERS]
        key = tuple((m.weight.data_ptr(), m.weight._version) for m in convs)
        if self._plan is None or self._plan["ptrs"] != tuple(k[0] for k in key):
            dev = convs[0].weight.device
            t = weightmap.JobTable()
            c = {}
            for (site, i, _, _, taps), m in zip(self.LAYERS, convs):
                co, ci = m.weight.size(0), m.weight.size(1)
                mp = weightmap.map_conv_taps(co, ci, taps)
                cip = self.cin_pad if site == "cnn.0" else mp.Cip
                f = torch.zeros((mp.Tf, co, cip), device=dev, dtype=torch.bfloat16)
                d = torch.zeros((mp.Td, cip, mp.Cop), device=dev, dtype=torch.bfloat16)
                mp.add_pack_fwd(t, m.weight, f, Cip=cip)
                # dgrad operand [Td][cip][Cop]: rows past Ci stay zero (channel padding of the packed input)
                t.add(m.weight, d, R=ci, C=co, Rp=ci, Cp=mp.Cop, s_r=mp.s_ci, s_c=mp.s_co, d_r=mp.Cop, d_c=1, M=mp.Ad,
                      out_off=[k * cip * mp.Cop for k in range(mp.Td)], dst_bf16=True)
                c[site] = (f, d, mp.taps_d, mp, cip)
            t.finalize(dev)
            self._plan = {"table": t, "c": c, "ptrs": tuple(k[0] for k in key), "wgrad": None}
            self._plan_key = None
        if self._plan_key != key:
            self._plan["table"].run()
            self._plan_key = key
        return self._plan["c"]

    def _drop_scales(self, B, dev):
        if not self.training:
            return [None, None]
        chans = [self.cnn[1].num_channels, self.cnn[5].num_channels]
        if self.dropout_masks is not None:
            return [m.to(dev).float().reshape(B, C) / (1.0 - DROPOUT_P) for m, C in zip(self.dropout_masks, chans)]
        s = (torch.rand((B, sum(chans)), device=dev) >= DROPOUT_P).float().mul_(1.0 / (1.0 - DROPOUT_P))
        return [s[:, :chans[0]].contiguous(), s[:, chans[0]:].contiguous()]

    # -- forward ---------------------------------------------------------------------------------------------------
    def forward(self, input, style):
        """input [L,B,class_size] (the one-hot text), style [B,style_size] -> [L,B,n_out] (count_cnn.py:34-45)."""
        _lib.require_cuda(input, style)
        named = [(n, p) for n, p in _lib.named_params(self) if n.startswith("cnn.")]
        if torch.is_grad_enabled() and (input.requires_grad or style.requires_grad or any(p.requires_grad for _, p in named)):
            raw = _CountFn.apply(self, tuple(n for n, _ in named), input, style, *[p for _, p in named])
        else:
            raw, _ = self._forward_impl(input, style, keep=False)
        return raw * self.std + self.mean

    def _forward_impl(self, content, style, keep):
        c = self._prepare()
        L, B, C = content.shape
        if C != self.class_size or style.size(1) != self.style_size:
            raise RuntimeError(f"CountCNN: input {tuple(content.shape)} / style {tuple(style.shape)} do not match the module")
        dev = content.device
        s = _lib.stream
        x = ops.gen_pack_input(content.float(), style.float().contiguous(), self.cin_pad)        # [B,1,L,Cp] bf16
        drops = self._drop_scales(B, dev)
        ctx = {"shape": (L, B), "x0": x, "c": c, "gn": []}
        a = x
        di = 0
        for site, i, gi, has_drop, taps in self.LAYERS[:3]:
            m, gn = self.cnn[i], self.cnn[gi]
            co = m.weight.size(0)
            st = torch.zeros((B, co, 2), device=dev, dtype=torch.float32)
            z = conv.conv_fprop(a, c[site][0], taps, 1, L, bias=m.bias.detach(), stats=st)
            coef = torch.empty((B, co, 2), device=dev, dtype=torch.float32)
            save = torch.empty((B, co, 2), device=dev, dtype=torch.float32)
            _lib.call("hwg_gn_coeffs", st.data_ptr(), gn.weight.data_ptr(), gn.bias.data_ptr(), B, co, gn.num_groups, L,
                      gn.eps, coef.data_ptr(), save.data_ptr(), s())
            drop = None
            if has_drop:
                drop = drops[di]
                di += 1
                if drop is not None:
                    coef.mul_(drop[:, :, None])
            a_in = a
            a = ops.scale_shift_act(z, coef, True, ACT_RELU, 0.0, out=torch.empty_like(z))
            if keep:
                ctx["gn"].append((site, a_in, z, coef, save, drop, gn))
        head = self.cnn[11]
        out = conv.conv_fprop(a, c["cnn.11"][0], _T1, 1, L, bias=head.bias.detach(), out_dtype=torch.float32,
                              force_tcgen05=True)                                                # [B,1,L,n_out] fp32
        if keep:
            ctx["a3"] = a
        return out[:, 0].permute(1, 0, 2).contiguous(), ctx

    # -- backward --------------------------------------------------------------------------------------------------
    def _wgrad_plan(self, dev):
        plan = self._plan.get("wgrad")
        if plan is not None:
            return plan
        params = dict(self.named_parameters())
        off, slots = 0, {}

        def take(name, n):
            nonlocal off
            slots[name] = (off, n)
            off += -(-n // 4) * 4

        goff, n_el = {}, 0
        for n, p in params.items():
            if n.startswith("cnn."):
                goff[n] = n_el
                n_el += -(-p.numel() // 4) * 4
        gflat = torch.zeros(n_el, device=dev, dtype=torch.float32)
        t = weightmap.JobTable()
        meta = {}
        for site, i, gi, _, taps in self.LAYERS:
            m = self.cnn[i]
            co, ci = m.weight.size(0), m.weight.size(1)
            _, _, _, mp, cip = self._plan["c"][site]
            cop = mp.Cop if co < 64 else co
            take(("w", site), len(taps) * cop * cip)
            take(("b", site), cop)
            meta[site] = (co, cop, cip, taps)
            wn, bn = f"{site}.weight", f"{site}.bias"
            mp.add_unpack_wgrad(t, 4 * slots[("w", site)][0], gflat[goff[wn]:goff[wn] + params[wn].numel()], Cip=cip, co_rows=cop)
            t.add(4 * slots[("b", site)][0], gflat[goff[bn]:goff[bn] + co], R=1, C=co, s_r=0, s_c=1, d_r=0, d_c=1, M=np.eye(1))
            if gi is not None:
                for kind, suffix in (("gamma", ".weight"), ("beta", ".bias")):
                    take((kind, site), co)
                    pn = f"cnn.{gi}{suffix}"
                    t.add(4 * slots[(kind, site)][0], gflat[goff[pn]:goff[pn] + co], R=1, C=co, s_r=0, s_c=1, d_r=0, d_c=1,
                          M=np.eye(1))
        t.finalize(dev)
        plan = self._plan["wgrad"] = dict(slots=slots, arena_floats=off, meta=meta, gflat=gflat, goff=goff, table=t)
        return plan

    def _backward(self, ctx, g_raw, names, want_content, want_style):
        """g_raw [L,B,n_out] fp32.  Returns (content gradient [L,B,C] or None, style gradient [B,S] or None,
        {parameter name: gradient} for `names`)."""
        L, B = ctx["shape"]
        c = ctx["c"]
        dev = g_raw.device
        s = _lib.stream
        want_w = bool(names)
        if want_w:
            wp = self._wgrad_plan(dev)
            arena = torch.zeros(wp["arena_floats"], device=dev, dtype=torch.float32)      # one memset

            def slot(key):
                o, n = wp["slots"][key]
                return arena[o:o + n]

        def collect(site, x_in, gz):
            if not want_w:
                return
            co, cop, cip, taps = wp["meta"][site]
            conv.conv_wgrad(x_in, gz, taps, cip, cop, out=slot(("w", site)).view(len(taps), cop, cip))
            _lib.call("hwg_channel_sum", gz.data_ptr(), gz.numel() // gz.size(-1), gz.size(-1), slot(("b", site)).data_ptr(),
                      s())

        def dgrad(g, site):
            _, wd, taps, _, _ = c[site]
            return conv.conv_fprop(g, wd, taps, 1, L, force_tcgen05=(site == "cnn.11"))

        g16 = torch.zeros((B, 1, L, 16), device=dev, dtype=torch.bfloat16)
        g16[:, 0, :, :self.n_out] = g_raw.permute(1, 0, 2)
        collect("cnn.11", ctx["a3"], g16)
        g = dgrad(g16, "cnn.11")
        for site, a_in, z, coef, save, drop, gn in reversed(ctx["gn"]):
            C = z.size(3)
            sums = torch.zeros((B, C, 2), device=dev, dtype=torch.float32)
            spq = torch.empty((B, C, 3), device=dev, dtype=torch.float32)
            gz = torch.empty_like(z)
            _lib.call("hwg_norm_bwd_reduce", g.data_ptr(), z.data_ptr(), coef.data_ptr(), 0.0, B, 1, L, C, 1, 1,
                      sums.data_ptr(), s())
            if drop is not None:                 # d/d(GroupNorm output) = s * gy': s enters the sums ...
                sums.mul_(drop[:, :, None])
            _lib.call("hwg_gn_bwd_coeffs", sums.data_ptr(), save.data_ptr(), gn.weight.data_ptr(), B, C, gn.num_groups, L,
                      spq.data_ptr(), slot(("gamma", site)).data_ptr() if want_w else None,
                      slot(("beta", site)).data_ptr() if want_w else None, s())
            if drop is not None:                 # ... and the direct term sc * gy'
                spq[:, :, 0].mul_(drop)
            _lib.call("hwg_norm_bwd_apply", g.data_ptr(), z.data_ptr(), coef.data_ptr(), spq.data_ptr(), 0.0, B, 1, L, C, 1,
                      1, gz.data_ptr(), s())
            collect(site, a_in, gz)
            if site != "cnn.0" or want_content or want_style:
                g = dgrad(gz, site)
        g_content = g_style = None
        if want_content or want_style:
            g0 = g[:, 0].float()                                                            # [B,L,Cp]
            C, S = self.class_size, self.style_size
            if want_content:
                g_content = g0[:, :, :C].permute(1, 0, 2).contiguous()
            if want_style:
                g_style = g0[:, :, C:C + S].sum(1)
        pgrads = {}
        if want_w:
            wp["table"].run(src_base=arena)
            params = dict(self.named_parameters())
            gout = wp["gflat"].clone()      # autograd may adopt a returned gradient as .grad: never hand out the workspace
            for n in names:
                o = wp["goff"][n]
                pgrads[n] = gout[o:o + params[n].numel()].view_as(params[n])
        return g_content, g_style, pgrads


class _CountFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, m, names, content, style, *params):
        with torch.no_grad():
            out, saved = m._forward_impl(content, style, keep=True)
        ctx.m, ctx.names, ctx.saved = m, names, saved
        return out

    @staticmethod
    def backward(ctx, g):
        need = ctx.needs_input_grad
        names = [n for n, nd in zip(ctx.names, need[4:]) if nd]
        with torch.no_grad():
            gc, gs, pg = ctx.m._backward(_lib.saved_state(ctx.saved), g.contiguous().float(), names, need[2], need[3])
        if not _lib.RETAIN_SAVED:
            ctx.saved = None
        return (None, None, gc, gs) + tuple(pg.get(n) for n in ctx.names)
