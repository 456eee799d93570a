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

from . import _lib, conv, ops, weightmap
from ._lib import ACT_LRELU, ACT_NONE
from .pure_gen import TAPS3x3, conv1_forward



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
    # one zero-fill for all ten statistics accumulators ([B,C,2] each)
    stats_all = torch.zeros(sum(2 * B * e["C"] * 2 for e in c["blocks"]), device=dev, dtype=torch.float32)
    soff = 0

    def new_stats(C):
        nonlocal soff
        st = stats_all[soff:soff + B * C * 2].view(B, C, 2)
        soff += B * C * 2
        return st

    for bi, e in enumerate(c["blocks"]):
        C = e["C"]
        st = new_stats(C)
        nz = None if noise is None else noise[k]
        x_in, Hin, Win = x, H, W
        a, Ho, Wo = conv1_forward(x, e, B, H, W, st, nz, k, seed, seed_dev)
        H, W = Ho, Wo
        coef, save = ops.adain_coeffs(st, gb[:, off:], gb[:, off + C:], gbs, B, C, H * W, save=True)
        recs.append(dict(x_in=x_in, Hin=Hin, Win=Win, a=a, coef=coef, save=save, nz=nz, subseq=16 * k, off=off))
        x = ops.scale_shift_act(a, coef, True, out=torch.empty_like(a))
        off += 2 * C
        k += 1
        st = new_stats(C)
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
def _bwd_plan(m, c, B, gbw, sink=None):
    """Workspace layout + hwg_linear_map job table of one backward pass, built once per (module, batch).

    Workspace `ws` (fp32, zero-filled by ONE memset per backward): every accumulator the backward kernels add into —
    the per-(n,c) AdaIN sums, the per-channel bias / noise-weight sums, the blur-adjoint statistics, the output-head
    sums and the tap-major wgrad outputs of all convolutions.  The job table then writes, in ONE launch, every
    conv-side parameter gradient in the parameter's own layout (adjoint of the forward pack, EqualLR scales) and the
    gradient of the AdaIN projections `g_gb`, into one flat buffer `gflat`."""
    off = {"n": 0}
    lay = {}

    def ws(name, numel):
        lay[name] = (off["n"], numel)
        off["n"] += -(-numel // 4) * 4           # 16-byte aligned slots

    params = _param_list(m)
    acc = sink is not None         # gradients are ADDED straight into the optimizer's flat gradient buffer
    goff, n = [], 0
    if not acc:
        for p in params:
            goff.append(n)
            n += -(-p.numel() // 4) * 4
    g_gb_off = n
    n += B * gbw
    t = weightmap.JobTable()
    one = [[1.0]]

    def dst(i):
        return sink.grad_view(params[i]) if acc else 4 * goff[i]

    def vec(src_name, src_elem, dst_param_idx, C, scale=1.0, s_c=2):
        t.add(4 * (lay[src_name][0] + src_elem), dst(dst_param_idx), R=1, C=C, s_r=0, s_c=s_c, d_r=0, d_c=1,
              M=one, scale=scale, accumulate=acc)

    gb_off = 0
    for bi, (blk, e) in enumerate(zip(m.conv, c["blocks"])):
        C = e["C"]
        m1, m2 = e["m1"], e["m2"]
        cip1 = e.get("m1_cip", m1.Cip)
        for h in (1, 2):
            ws(f"sums{bi}{h}", B * C * 2)
            ws(f"dch{bi}{h}", C * 2)
        ws(f"dw{bi}1", m1.Tf * m1.Co * cip1)
        ws(f"dw{bi}2", m2.Tf * m2.Co * m2.Cip)
        if e["kind"] in ("vert_up", "fused_up"):
            ws(f"st{bi}", B * C * 2)
        base = 6 * bi
        m1.add_unpack_wgrad(t, 4 * lay[f"dw{bi}1"][0], dst(base + 0), Cip=cip1, accumulate=acc)
        if e["kind"] in ("vert_up", "fused_up"):
            # Blur is self-adjoint; the per-channel sums of the blurred gradient are the conv bias gradient
            t.add(4 * lay[f"st{bi}"][0], dst(base + 1), R=1, C=C, s_r=0, s_c=2, d_r=0, d_c=1, M=None, nin=B,
                  in_stride=2 * C, accumulate=acc)
        else:
            vec(f"dch{bi}1", 0, base + 1, C)
        vec(f"dch{bi}1", 1, base + 2, C, sqrt(2.0 / C))
        m2.add_unpack_wgrad(t, 4 * lay[f"dw{bi}2"][0], dst(base + 3), accumulate=acc)
        vec(f"dch{bi}2", 0, base + 4, C)
        vec(f"dch{bi}2", 1, base + 5, C, sqrt(2.0 / C))
        for h in (1, 2):
            # sums[n,c] = (dbeta, dgamma)  ->  g_gb[n, off + c] = dgamma, g_gb[n, off + C + c] = dbeta
            t.add(4 * lay[f"sums{bi}{h}"][0], 4 * (g_gb_off + gb_off), R=B, C=C, s_r=2 * C, s_c=2, d_r=gbw, d_c=1,
                  M=[[1.0, 0.0], [0.0, 1.0]], in_off=[0, 1], out_off=[C, 0])
            gb_off += 2 * C
    Cl = c["blocks"][-1]["C"]
    ws("dwb", Cl + 1)
    vec("dwb", 0, len(params) - 2, Cl, c["out_scale"], s_c=1)
    vec("dwb", Cl, len(params) - 1, 1, 1.0, s_c=1)
    t.finalize(params[0].device)
    return dict(lay=lay, ws_numel=off["n"], goff=goff, g_gb_off=g_gb_off, gflat_numel=n, table=t,
                shapes=[p.shape for p in params], numels=[p.numel() for p in params])


def backward_train(m, ctx, g_out):
    """Returns (g_content [T,B,ncls], g_s [B,S], g_gb [B,sum 2C], [conv-side parameter grads in _param_list order])."""
    recs, seed, seed_dev = ctx["recs"], ctx["seed"], ctx["seed_dev"]
    c = m._packed()
    B, T, ncls, gbw = ctx["B"], ctx["T"], ctx["ncls"], ctx["gb_width"]
    dev = g_out.device
    sink = getattr(m, "_grad_sink", None)
    if sink is not None and not all(sink.owns(p) for p in _param_list(m)):
        sink = None
    key = (B, gbw, id(sink))
    plan = m._bwd_plans.get(key)
    if plan is None:
        plan = m._bwd_plans[key] = _bwd_plan(m, c, B, gbw, sink)
    ws = torch.zeros(plan["ws_numel"], device=dev, dtype=torch.float32)
    gflat = torch.empty(plan["gflat_numel"], device=dev, dtype=torch.float32)

    def W(name, *shape):
        o, n = plan["lay"][name]
        return ws[o:o + n].view(*shape)

    # The weight gradients feed nothing but the unpack launch at the end, while the input-gradient chain
    # (AdaIN backward -> dgrad -> next layer) is serial: the wgrad launches go to a side stream and run next to the
    # chain.  Each one is ordered after the kernel that produced its gy (wait_stream) and the chain's tensors it reads
    # are kept alive until the streams join, so the caching allocator cannot hand them out again early.  All of it is
    # stream work, so a captured step holds the two branches.
    main = torch.cuda.current_stream()
    side = None
    if getattr(m, "parallel_wgrad", True):
        side = getattr(m, "_wgrad_stream", None)
        if side is None or side.device != dev:
            side = m._wgrad_stream = torch.cuda.Stream(device=dev)
    keep = []

    def wgrad(x_in, gy, *a, **k):
        if side is None:
            return conv.conv_wgrad(x_in, gy, *a, **k)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            conv.conv_wgrad(x_in, gy, *a, **k)
        keep.append(gy)

    # ---- output head
    last = recs[-1]
    g, _, _ = ops.gen_output_bwd(g_out.contiguous().float(), ctx["out"], last["a"], last["coef"], c["w_out"],
                                 dwb=W("dwb", -1))
    blocks = list(m.conv)
    for bi in range(len(blocks) - 1, -1, -1):
        blk, e = blocks[bi], c["blocks"][bi]
        C = e["C"]
        m1, m2 = e["m1"], e["m2"]
        r1, r2 = recs[2 * bi], recs[2 * bi + 1]
        # ---------------- second half: conv2 + noise2 + lrelu + adain2
        gy = ops.adain_lrelu_bwd(g, r2["a"], r2["save"], r2["coef"], 0.2, r2["nz"], seed, r2["subseq"],
                                 seed_dev=seed_dev, sums=W(f"sums{bi}2", B, C, 2), dch=W(f"dch{bi}2", C, 2))[0]
        wgrad(r2["x_in"], gy, TAPS3x3, C, C, out=W(f"dw{bi}2", 9, C, C))
        g = conv.conv_fprop(gy, e["d2"], m2.taps_d, r2["Hin"], r2["Win"])
        # ---------------- first half: conv1 (+blur) + noise1 + lrelu + adain1
        gy = ops.adain_lrelu_bwd(g, r1["a"], r1["save"], r1["coef"], 0.2, r1["nz"], seed, r1["subseq"],
                                 row_subseq=(e["kind"] == "initial"), seed_dev=seed_dev,
                                 sums=W(f"sums{bi}1", B, C, 2), dch=W(f"dch{bi}1", C, 2))[0]
        x_in, Hin, Win = r1["x_in"], r1["Hin"], r1["Win"]
        kind = e["kind"]
        if kind in ("vert_up", "fused_up"):
            gy = ops.blur_noise_act_stats(gy, None, None, W(f"st{bi}", B, C, 2), ACT_NONE, 0.0)
        Cin = m1.Ci
        if kind == "plain":
            wgrad(x_in, gy, TAPS3x3, Cin, C, out=W(f"dw{bi}1", 9, C, Cin))
            g = conv.conv_fprop(gy, e["d1"], m1.taps_d, Hin, Win)
        elif kind == "initial":
            cin_pad = c["cin_pad"]
            dw = W(f"dw{bi}1", 12, C, cin_pad)
            for r in range(4):
                wgrad(x_in, gy, e["taps1"], cin_pad, C, grid=(1, Win), gy_offset=(r, 0), out=dw[3 * r:3 * r + 3])
            # gradient w.r.t. the packed input: 12-tap convolution of gy [B,4,T,C] -> [B,1,T,Cin16]
            g = conv.conv_fprop(gy, e["d1"], m1.taps_d, 1, Win)
        elif kind == "vert_up":
            dw = W(f"dw{bi}1", 12, C, Cin)
            for par in (0, 1):
                wgrad(x_in, gy, weightmap.vert_taps(par), Cin, C, grid=(Hin, Win), gy_stride=(2, 1),
                                gy_offset=(par, 0), out=dw[6 * par:6 * par + 6])
            g = conv.conv_fprop(gy, e["d1"], m1.taps_d, Hin, Win, in_stride=m1.d_in_stride)
        else:  # fused_up
            dw = W(f"dw{bi}1", 16, C, Cin)
            taps = weightmap.fused_taps()
            if Cin <= 32 and C <= 32:
                # all four parities in one launch (per-tap gy phase): gy and x are read once
                wgrad(x_in, gy, taps, Cin, C, grid=(Hin, Win), gy_stride=(2, 2),
                                tap_phase=weightmap.fused_phases(), out=dw)
            else:
                for q, (py, px) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
                    wgrad(x_in, gy, taps[4 * q:4 * q + 4], Cin, C, grid=(Hin, Win), gy_stride=(2, 2),
                                    gy_offset=(py, px), out=dw[4 * q:4 * q + 4])
            g = conv.conv_fprop(gy, e["d1"], m1.taps_d, Hin, Win, in_stride=m1.d_in_stride)
    # ---- every parameter gradient + g_gb in one launch
    if side is not None:
        main.wait_stream(side)
    keep.clear()
    plan["table"].run(src_base=ws, dst_base=gflat)
    if sink is None:
        flat = [gflat[o:o + n].view(sh) for o, n, sh in zip(plan["goff"], plan["numels"], plan["shapes"])]
    else:
        flat = [None] * len(plan["shapes"])      # already accumulated into the flat gradient buffer
        ready = getattr(m, "_grad_ready_cb", None)
        if ready is not None:
            for p in _param_list(m):
                ready(p)
    g_gb = gflat[plan["g_gb_off"]:plan["g_gb_off"] + B * gbw].view(B, gbw)
    # ---- packed input -> content and style
    S = m.style_size if m.append_style else 0
    gx0 = g[:, 0].float()                                # [B, T, Cin16]
    g_content = gx0[:, :, :ncls].permute(1, 0, 2).contiguous()
    g_s = gx0[:, :, ncls:ncls + S].sum(1) if S else None
    return g_content, g_s, g_gb, flat


def _style_params(m):
    ps = []
    for mod in m.style_emb:
        if isinstance(mod, torch.nn.Linear):
            ps += [mod.weight, mod.bias]
    for blk in m.conv:
        for ad in (blk.adain1, blk.adain2):
            ps += [ad.style.weight, ad.style.bias]
    return ps


def _style_scatter_table(m, c, sink):
    """Jobs that add the gradient of the concatenated AdaIN projection ([sum 2C, S] weights | [sum 2C] biases, one
    temporary) into the twenty separate parameters' slots of the flat gradient buffer."""
    S, n_gb = m.style_size, c["gb_w"].size(0)
    t = weightmap.JobTable()
    row = 0
    for blk in m.conv:
        for ad in (blk.adain1, blk.adain2):
            C2 = ad.style.weight.size(0)
            t.add(4 * row * S, sink.grad_view(ad.style.weight), R=C2, C=S, s_r=S, s_c=1, d_r=S, d_c=1, M=[[1.0]],
                  accumulate=True)
            t.add(4 * (n_gb * S + row), sink.grad_view(ad.style.bias), R=1, C=C2, s_r=0, s_c=1, d_r=0, d_c=1, M=[[1.0]],
                  accumulate=True)
            row += C2
    return t.finalize(c["gb_w"].device)


class _StyleFn(torch.autograd.Function):
    """PixelNorm'ed style -> (s, gb): the six Linear+LeakyReLU(0.2) layers and the ten AdaIN projections as one
    concatenated Linear (pure_gen.py:31-39,57,63) on hwg_linear_f32 / hwg_linear_bwd_f32."""

    @staticmethod
    def forward(ctx, module, s0, *params):
        c = module._packed()
        acts = [s0.contiguous()]
        for w, b in c["mlp"]:
            acts.append(ops.linear(acts[-1], w, b, ACT_LRELU, 0.2))
        gb = ops.linear(acts[-1], c["gb_w"], c["gb_b"])
        ctx.module, ctx.acts = module, acts
        return acts[-1].detach(), gb               # detached alias: see _GenFn.forward

    @staticmethod
    def backward(ctx, g_s, g_gb):
        m, acts = ctx.module, _lib.saved_state(ctx.acts)
        c = m._packed()
        sink = getattr(m, "_grad_sink", None)
        sp = _style_params(m)
        if sink is not None and not all(sink.owns(p) for p in sp):
            sink = None
        s = acts[-1]
        S, n_gb, L = m.style_size, c["gb_w"].size(0), len(c["mlp"])
        grads = [None] * len(sp)
        g = g_s
        if g_gb is not None:
            tmp = torch.empty(n_gb * S + n_gb, device=s.device, dtype=torch.float32)
            tw, tb = tmp[:n_gb * S].view(n_gb, S), tmp[n_gb * S:]
            gx, _, _ = ops.linear_bwd(s, None, g_gb.contiguous(), c["gb_w"], gW=tw, gb=tb)
            g = gx if g is None else g + gx
            if sink is not None:
                key = ("style_scatter", id(sink))
                t = m._bwd_plans.get(key)
                if t is None:
                    t = m._bwd_plans[key] = _style_scatter_table(m, c, sink)
                t.run(src_base=tmp)
            else:
                row = 0
                for i in range(2 * L, len(sp), 2):
                    C2 = sp[i].size(0)
                    grads[i], grads[i + 1] = tw[row:row + C2], tb[row:row + C2]
                    row += C2
        for li in range(L - 1, -1, -1):
            w, _ = c["mlp"][li]
            need_gx = li > 0 or ctx.needs_input_grad[1]
            if sink is not None:
                g, _, _ = ops.linear_bwd(acts[li], acts[li + 1], g.contiguous(), w, ACT_LRELU, 0.2, need_gx,
                                         gW=sink.grad_view(sp[2 * li]), gb=sink.grad_view(sp[2 * li + 1]), accumulate=True)
            else:
                g, grads[2 * li], grads[2 * li + 1] = ops.linear_bwd(acts[li], acts[li + 1], g.contiguous(), w,
                                                                     ACT_LRELU, 0.2, need_gx)
        if sink is not None:
            ready = getattr(m, "_grad_ready_cb", None)
            if ready is not None:
                for p in sp:
                    ready(p)
        if not _lib.RETAIN_SAVED:
            ctx.acts = None
        return (None, g) + tuple(grads)


class _GenFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, noise, content, s, gb, *params):
        with torch.no_grad():
            out, saved = forward_train(module, content, s, gb, noise)
        ctx.module, ctx.saved = module, saved
        # a detached alias: the saved state keeps the tensor itself, and an output that is ALSO reachable from ctx would
        # close a reference cycle through its grad_fn (output -> node -> ctx -> saved -> output) that is only broken when
        # backward drops the state — never, with set_retain_graph(True) or when no backward runs
        return out.detach()

    @staticmethod
    def backward(ctx, g):
        with torch.no_grad():
            g_content, g_s, g_gb, flat = backward_train(ctx.module, _lib.saved_state(ctx.saved), g)
        if not _lib.RETAIN_SAVED:
            ctx.saved = None
        return (None, None, g_content, g_s, g_gb) + tuple(flat)


def generator_apply(module, content, style, noise):
    if module.training and any(isinstance(mod, torch.nn.Dropout) for mod in module.style_emb):
        s, gb = _style_path(module, style)       # style dropout (unused by the shipped configs): torch autograd
    else:
        st = style.float()
        if st.requires_grad:                      # PixelNorm (pure_gen.py:306-311) under autograd
            s0 = st / torch.sqrt((st * st).mean(1, keepdim=True) + 1e-8)
        else:
            s0 = ops.pixelnorm(st.contiguous())
        s, gb = _StyleFn.apply(module, s0, *_style_params(module))
    return _GenFn.apply(module, noise, content, s, gb, *_param_list(module))
