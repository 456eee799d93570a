"""Training path of the recognizer: one torch.autograd.Function around the fused forward and the
backward kernels of libhwg_b200 (reference: the autograd graph PyTorch records for
model/cnn_only_hwr.py:96-107; backward = cuDNN dgrad/wgrad + ATen batch_norm/max_pool/relu/
log_softmax backward kernels).

Backward of every convolution = dgrad (hwg_conv_fprop on the output gradient with negated taps and
transposed weights) + wgrad (hwg_conv_wgrad, tcgen05 with MN-major operands); the memory-bound passes
between them (log-softmax, BatchNorm+ReLU, ReLU+MaxPool, stem) emit the bias gradients on the way.
"""
import torch

from . import _lib, conv, ops, weightmap
from ._lib import ACT_LOGSOFTMAX, ACT_NONE, ACT_RELU

_T3 = conv.conv_taps(3, 3, 1, 1)
_T3P0 = conv.conv_taps(3, 3, 0, 0)
_CNN1D = [(0, 1, 2, 2), (3, 4, 4, 4), (6, 7, 0, 1), (9, 10, 8, 8)]
_POOL22 = ((2, 2), (2, 2), (0, 0))
_POOL21 = ((2, 2), (2, 1), (0, 1))


def _dgrad_packs(m):
    """Transposed tap operands of every tensor-core convolution (filled by the module's hwg_linear_map table), plus
    the 9-tap operand of the two-step image-gradient path (only built when that path is used)."""
    return m._packed()["dgrad"]


def _w0_dgrad(m, c):
    if "w0" not in c["dgrad"]:
        # conv0's weights as a [Cout=1 (padded to 16)] x [Cin=64] 9-tap dgrad convolution
        w0 = m.cnn.conv0.weight.detach().float()                       # [64,1,3,3]
        mats = [torch.nn.functional.pad(w0[:, 0, i, j].view(1, 64), (0, 0, 0, 15)) for i in range(3) for j in range(3)]
        c["dgrad"]["w0"] = (conv.pack_taps(mats), [(1 - i, 1 - j) for i in range(3) for j in range(3)])
    return c["dgrad"]["w0"]


def _bn_train(m, y, stats, bn):
    N, H, W, C = y.shape
    momentum = 0.1 if bn.momentum is None else bn.momentum
    use_batch = m.training or not bn.track_running_stats
    if use_batch and m.sync_bn_group is not None:
        coef, save = ops.bn_coeffs_synced(stats, N, C, H * W, bn.weight.detach(), bn.bias.detach(), bn.running_mean,
                                          bn.running_var, momentum, bn.eps, m.sync_bn_group, key=("f", id(bn)))
    else:
        coef, save = ops.bn_coeffs(stats, N, C, H * W, bn.weight.detach(), bn.bias.detach(), bn.running_mean,
                                   bn.running_var, momentum, bn.eps, use_batch)
    if m.training and bn.num_batches_tracked is not None:
        m._bn_counters.append(bn.num_batches_tracked)     # bumped together at the end of the forward
    if not use_batch:  # eval-mode BN under autograd: (mean, rstd) from the running statistics
        save = torch.stack([bn.running_mean, torch.rsqrt(bn.running_var + bn.eps)], 1).contiguous()
    a = ops.scale_shift_act(y, coef, False, ACT_RELU, out=torch.empty_like(y))
    return a, coef, save


def forward_train(m, x):
    """Forward that keeps what the backward needs.  Returns (log_probs [T,B,C], ctx dict)."""
    if m.pad is not None:
        x = m.pad(x)
    if m.small:
        raise NotImplementedError("small=True is not implemented")
    c = m._packed()
    x = x.float().contiguous()
    B = x.size(0)
    dev = x.device
    # eval mode under autograd (a frozen `hwr.eval()` guiding the generator): the running statistics are constants of the
    # backward, so no batch-statistics terms and no exchange
    ctx = {"x": x, "sync_bn_group": m.sync_bn_group if m.training else None, "bn_use_batch": bool(m.training)}
    m._bn_counters = []

    arena = ops.ZeroArena(B * 2 * (256 + 6 * 512) + 64, dev)     # all BatchNorm statistics of this pass: one memset

    def stats_for(C):
        return arena.take(B, C, 2)

    a0 = ops.hwr_stem(x, c["w0"], c["b0"])
    c1 = conv.conv_fprop(a0, c["w1"], _T3, a0.size(1), a0.size(2), bias=c["b1"], act=ACT_RELU)
    a1 = ops.maxpool_nhwc(c1, *_POOL22)
    st = stats_for(256)
    z2 = conv.conv_fprop(a1, c["w2"], _T3, a1.size(1), a1.size(2), bias=c["b2"], stats=st)
    a2, coef2, save2 = _bn_train(m, z2, st, m.cnn.batchnorm2)
    c3 = conv.conv_fprop(a2, c["w3"], _T3, a2.size(1), a2.size(2), bias=c["b3"], act=ACT_RELU)
    a3 = ops.maxpool_nhwc(c3, *_POOL21)
    st = stats_for(512)
    z4 = conv.conv_fprop(a3, c["w4"], _T3, a3.size(1), a3.size(2), bias=c["b4"], stats=st)
    a4, coef4, save4 = _bn_train(m, z4, st, m.cnn.batchnorm4)
    c5 = conv.conv_fprop(a4, c["w5"], _T3P0, a4.size(1) - 2, a4.size(2) - 2, bias=c["b5"], act=ACT_RELU)
    if m._save:
        m.saved_features[0] = c5.permute(0, 3, 1, 2).float()
    a5 = ops.maxpool_nhwc(c5, *_POOL21)
    st = stats_for(512)
    z6 = conv.conv_fprop(a5, c["w6"], _T3P0, a5.size(1) - 2, a5.size(2) - 2, bias=c["b6"], stats=st)
    a6, coef6, save6 = _bn_train(m, z6, st, m.cnn.batchnorm6)
    if a6.size(1) != 1:
        raise RuntimeError(f"CNNOnlyHWR expects 64-px-high images (conv height {a6.size(1)} != 1)")
    ctx.update(a0=a0, c1=c1, a1=a1, z2=z2, a2=a2, bn2=(coef2, save2), c3=c3, a3=a3, z4=z4, a4=a4, bn4=(coef4, save4),
               c5=c5, a5=a5, z6=z6, bn6=(coef6, save6))
    a = a6
    ctx["head_in"] = []
    for ci, bi, pad, dil in _CNN1D:
        Wo = a.size(2) + 2 * pad - 2 * dil
        st = stats_for(512)
        z = conv.conv_fprop(a, c[f"v{ci}"], conv.conv_taps(1, 3, 0, pad, 1, dil), 1, Wo, bias=c[f"c{ci}"], stats=st)
        an, coef, save = _bn_train(m, z, st, m.cnn1d[bi])
        ctx["head_in"].append((a, z, coef, save))
        a = an
    T = a.size(2) - 2
    C = m.nclass
    out = torch.empty((T, B, C), device=dev, dtype=torch.float32)
    conv.conv_fprop(a, c["v12"], conv.conv_taps(1, 3, 0, 0), 1, T, bias=c["c12"], act=ACT_LOGSOFTMAX,
                    out_view=(out, C, 0, B * C, 0))
    ctx["a10"] = a
    ctx["lp"] = out
    if m._bn_counters:
        torch._foreach_add_(m._bn_counters, 1)
    m._bn_counters = []
    return out, ctx


def _w4(dw, kh, kw):
    """wgrad output [kh*kw, Cout, Cin] -> parameter layout [Cout, Cin, kh, kw]."""
    t, co, ci = dw.shape
    return dw.view(kh, kw, co, ci).permute(2, 3, 0, 1).contiguous()


class _Grads(dict):
    """Collects the parameter gradients of one backward pass.

    Every accumulator (per-channel sums, tap-major wgrad outputs) is a slice of ONE zero-filled arena.  Small vectors
    are stored as views; convolution weights are registered with their ConvMap and turned into the parameter layout
    by one hwg_linear_map launch at the end (`finish`).  With a flat optimizer attached (module._grad_sink) every
    gradient is ADDED straight into its flat buffer by that launch and nothing is returned to autograd.
    wgrad launches of parameters whose gradient is not needed are skipped (frozen recognizer in the GAN lessons:
    the reference computes those gradients and throws them away)."""

    def __init__(self, m, needed, arena):
        super().__init__()
        self.m, self.needed, self.arena = m, needed, arena
        self.conv = []        # (name, ConvMap, dW view)

    def want(self, name):
        return self.needed is None or name in self.needed

    def wgrad(self, name, key, x, gy, taps, cin, cout):
        if not self.want(name):
            self[name] = None
            return
        cm = self.m._packed()["maps"][key]
        dw = self.arena.take(cm.Tf, cout, cm.Cip)
        conv.conv_wgrad(x, gy, taps, cin, cout, out=dw)
        self.conv.append((name, cm, dw, cout))
        self[name] = None     # filled by finish()

    def finish(self, params):
        """params: {name: parameter}.  Returns {name: gradient or None}."""
        m, base = self.m, self.arena.buf
        sink = getattr(m, "_grad_sink", None)
        if sink is not None and not all(sink.owns(p) for p in params.values()):
            sink = None
        vecs = [(n, v) for n, v in self.items() if v is not None and self.want(n)] if sink is not None else []
        if not self.conv and not vecs:
            return self
        def off(t):
            assert t.untyped_storage().data_ptr() == base.untyped_storage().data_ptr(), "accumulator outside the arena"
            return 4 * (t.storage_offset() - base.storage_offset())

        sig = (id(sink), tuple((n, off(dw), co) for n, _, dw, co in self.conv),
               tuple((n, off(v), v.stride()) for n, v in vecs))
        plan = m._bwd_plans.get("hwr_unpack")
        if plan is None or plan["sig"] != sig:
            t = weightmap.JobTable()
            goff, n_el = {}, 0
            for name, cm, dw, co_rows in self.conv:
                if sink is not None:
                    dst = sink.grad_view(params[name])
                else:
                    goff[name] = n_el
                    dst = 4 * n_el
                    n_el += -(-params[name].numel() // 4) * 4
                cm.add_unpack_wgrad(t, off(dw), dst, accumulate=sink is not None, co_rows=co_rows)
            for name, v in vecs:
                assert v.dim() == 1
                t.add(off(v), sink.grad_view(params[name]), R=1, C=v.numel(), s_r=0, s_c=v.stride(0), d_r=0, d_c=1,
                      M=[[1.0]], accumulate=True)
            t.finalize(base.device)
            plan = m._bwd_plans["hwr_unpack"] = dict(sig=sig, table=t, goff=goff, numel=n_el)
        gflat = None if sink is not None else torch.empty(max(plan["numel"], 4), device=base.device, dtype=torch.float32)
        plan["table"].run(src_base=base, dst_base=gflat)
        if sink is not None:
            for n in list(self):
                self[n] = None                         # already accumulated into the flat gradient buffer
            ready = getattr(m, "_grad_ready_cb", None)
            if ready is not None:
                for p in params.values():
                    ready(p)
        else:
            for name, _, _, _ in self.conv:
                o = plan["goff"][name]
                self[name] = gflat[o:o + params[name].numel()].view_as(params[name])
        return self


def backward_train(m, ctx, g_lp, want_input_grad=False, needed=None):
    """Returns ({parameter name: gradient or None} for every parameter of the module, image gradient or None)."""
    c = m._packed()
    dg = c["dgrad"]
    lp = ctx["lp"]
    T, B, C = lp.shape
    Cp = ((C + 15) // 16) * 16
    params = dict(_lib.named_params(m))
    # one zero-filled arena: every per-channel accumulator of this pass and the tap-major wgrad outputs
    wnames = {"w%d" % i: "cnn.conv%d.weight" % i for i in range(1, 7)}
    wnames.update({"v%d" % ci: "cnn1d.%d.weight" % ci for ci in (0, 3, 6, 9, 12)})
    dw_floats = sum(cm.Tf * (-(-cm.Co // 16) * 16) * cm.Cip + 4 for key, cm in c["maps"].items()
                    if needed is None or wnames[key] in needed)
    arena = ops.ZeroArena(dw_floats + 32768, lp.device)
    grads = _Grads(m, needed, arena)
    sync = ctx.get("sync_bn_group")        # the forward normalised with joint statistics of this group (SyncBN)
    ub = ctx.get("bn_use_batch", True)

    def dgrad(gz, key, H, W):
        wd, tapsd = dg[key]
        return conv.conv_fprop(gz, wd, tapsd, H, W)

    # ---- head: log-softmax + Conv1d(512, C, 3)
    gz, db = ops.logsoftmax_bwd(g_lp.contiguous().float(), lp, Cp, arena=arena)
    a10 = ctx["a10"]
    grads.wgrad("cnn1d.12.weight", "v12", a10, gz, conv.conv_taps(1, 3, 0, 0), 512, Cp)
    grads["cnn1d.12.bias"] = db
    g = dgrad(gz, "v12", 1, a10.size(2))
    # ---- dilated 1-D blocks, last to first
    for (ci, bi, pad, dil), (a_in, z, coef, save) in zip(reversed(_CNN1D), reversed(ctx["head_in"])):
        bn = m.cnn1d[bi]
        gz, dgam, dbet, dcb = ops.bn_bwd(g, z, coef, save, bn.weight.detach(), arena=arena, sync_group=sync, sync_key=("b", id(bn)), use_batch=ub)
        grads[f"cnn1d.{bi}.weight"], grads[f"cnn1d.{bi}.bias"] = dgam, dbet
        grads.wgrad(f"cnn1d.{ci}.weight", f"v{ci}", a_in, gz, conv.conv_taps(1, 3, 0, pad, 1, dil), 512, 512)
        grads[f"cnn1d.{ci}.bias"] = dcb
        g = dgrad(gz, f"v{ci}", 1, a_in.size(2))
    # ---- conv6 + BN + ReLU
    coef, save = ctx["bn6"]
    gz, dgam, dbet, dcb = ops.bn_bwd(g, ctx["z6"], coef, save, m.cnn.batchnorm6.weight.detach(), arena=arena, sync_group=sync,
                                     sync_key=("b", id(m.cnn.batchnorm6)), use_batch=ub)
    grads["cnn.batchnorm6.weight"], grads["cnn.batchnorm6.bias"] = dgam, dbet
    a5 = ctx["a5"]
    grads.wgrad("cnn.conv6.weight", "w6", a5, gz, _T3P0, 512, 512)
    grads["cnn.conv6.bias"] = dcb
    g = dgrad(gz, "w6", a5.size(1), a5.size(2))
    # ---- pool + ReLU + conv5
    gc, db = ops.relu_maxpool_bwd(g, ctx["c5"], *_POOL21, arena=arena)
    a4 = ctx["a4"]
    grads.wgrad("cnn.conv5.weight", "w5", a4, gc, _T3P0, 512, 512)
    grads["cnn.conv5.bias"] = db
    g = dgrad(gc, "w5", a4.size(1), a4.size(2))
    # ---- conv4 + BN + ReLU
    coef, save = ctx["bn4"]
    gz, dgam, dbet, dcb = ops.bn_bwd(g, ctx["z4"], coef, save, m.cnn.batchnorm4.weight.detach(), arena=arena, sync_group=sync,
                                     sync_key=("b", id(m.cnn.batchnorm4)), use_batch=ub)
    grads["cnn.batchnorm4.weight"], grads["cnn.batchnorm4.bias"] = dgam, dbet
    a3 = ctx["a3"]
    grads.wgrad("cnn.conv4.weight", "w4", a3, gz, _T3, 256, 512)
    grads["cnn.conv4.bias"] = dcb
    g = dgrad(gz, "w4", a3.size(1), a3.size(2))
    # ---- pool + ReLU + conv3
    gc, db = ops.relu_maxpool_bwd(g, ctx["c3"], *_POOL21, arena=arena)
    a2 = ctx["a2"]
    grads.wgrad("cnn.conv3.weight", "w3", a2, gc, _T3, 256, 256)
    grads["cnn.conv3.bias"] = db
    g = dgrad(gc, "w3", a2.size(1), a2.size(2))
    # ---- conv2 + BN + ReLU
    coef, save = ctx["bn2"]
    gz, dgam, dbet, dcb = ops.bn_bwd(g, ctx["z2"], coef, save, m.cnn.batchnorm2.weight.detach(), arena=arena, sync_group=sync,
                                     sync_key=("b", id(m.cnn.batchnorm2)), use_batch=ub)
    grads["cnn.batchnorm2.weight"], grads["cnn.batchnorm2.bias"] = dgam, dbet
    a1 = ctx["a1"]
    grads.wgrad("cnn.conv2.weight", "w2", a1, gz, _T3, 128, 256)
    grads["cnn.conv2.bias"] = dcb
    g = dgrad(gz, "w2", a1.size(1), a1.size(2))
    # ---- pool + ReLU + conv1
    gc, db = ops.relu_maxpool_bwd(g, ctx["c1"], *_POOL22, arena=arena)
    a0 = ctx["a0"]
    grads.wgrad("cnn.conv1.weight", "w1", a0, gc, _T3, 64, 128)
    grads["cnn.conv1.bias"] = db
    g = dgrad(gc, "w1", a0.size(1), a0.size(2))
    # ---- stem
    if needed is None or "cnn.conv0.weight" in needed or "cnn.conv0.bias" in needed:
        dw0, db0 = ops.hwr_stem_bwd(ctx["x"], c["w0"], c["b0"], g, arena=arena)
        grads["cnn.conv0.weight"] = dw0.view(-1)
        grads["cnn.conv0.bias"] = db0
    else:
        grads["cnn.conv0.weight"] = grads["cnn.conv0.bias"] = None
    grads.finish(params)
    if grads.get("cnn.conv0.weight") is not None:
        grads["cnn.conv0.weight"] = grads["cnn.conv0.weight"].view(64, 1, 3, 3)
    g_img = None
    if want_input_grad:
        # image gradient (GAN lessons): conv0 + ReLU + MaxPool backward to the image
        if c["w0"].size(0) == 64 and ctx["x"].size(2) % 2 == 0 and ctx["x"].size(3) % 2 == 0:
            g_img = ops.hwr_stem_bwd_image(ctx["x"], c["w0"], c["b0"], g)       # one fused pass
        else:
            gc0 = ops.hwr_stem_bwd_expand(ctx["x"], c["w0"], c["b0"], g)
            w0d, taps = _w0_dgrad(m, c)
            gi = conv.conv_fprop(gc0, w0d, taps, gc0.size(1), gc0.size(2), out_dtype=torch.float32)
            g_img = gi[..., 0].unsqueeze(1).contiguous()
        if m.pad is not None:
            p = m.pad.padding
            g_img = g_img[:, :, :, p[0]:g_img.size(3) - p[1]].contiguous()
    return grads, g_img


class _HWRFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, names, x, *params):
        with torch.no_grad():
            out, saved = forward_train(module, x)
        ctx.module, ctx.names, ctx.saved = module, names, saved
        ctx.x_needs_grad = x.requires_grad
        return out.detach()          # detached alias of the saved log-probs: no output -> node -> ctx -> output cycle

    @staticmethod
    def backward(ctx, g):
        needed = {n for n, need in zip(ctx.names, ctx.needs_input_grad[3:]) if need}
        with torch.no_grad():
            grads, g_img = backward_train(ctx.module, _lib.saved_state(ctx.saved), g, ctx.x_needs_grad, needed)
        if not _lib.RETAIN_SAVED:
            ctx.saved = None
        return (None, None, g_img) + tuple(grads[n] for n in ctx.names)


def hwr_apply(module, input):
    named = _lib.named_params(module)
    names = tuple(n for n, _ in named)
    return _HWRFn.apply(module, names, input, *[p for _, p in named])
