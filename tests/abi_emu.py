"""TEST INFRASTRUCTURE — CPU interpreter of the C-ABI entry points the host modules call (include/hwg_b200.h semantics, plain
torch/numpy on raw addresses), so that the HOST-side composition of the drop-in modules — tap lists, operand packing, shapes,
residual / Dropout2d / GroupNorm / BatchNorm bookkeeping, noise seeds and subsequences, accumulator arenas, the backward
chains, the optimizer plumbing — can be checked against the oracle and the reference goldens without a GPU, up to running
the UNMODIFIED reference trainer's curriculum on the drop-ins (tests/test_trainer_dropin_cpu.py).

It is a checker, never a fallback: it is installed by monkeypatching `_lib.call` inside a test (`installed()`), nothing in
the package imports it, and the GPU tests compare the real kernels with the same oracle.  Storage types follow the kernels
(NHWC bf16 activations and gradients, fp32 statistics); arithmetic is fp32.  The in-kernel NoiseInjection RNG
(csrc/noise_rng.cuh) is ported bit for bit in its integer part (exact log / sin / cos instead of the MUFU approximations).
The generator, recognizer and discriminator modules pass the same assertions through this interpreter as on the B200
(tests/test_*_emulated_cpu.py vs tests/test_*_gpu.py), which is what gives the not-yet-run modules' CPU checks their
weight."""
import contextlib
import ctypes

import numpy as np
import torch

from handwriting_line_generation_b200 import _lib, weightmap

from . import ref_map


def _view(ptr, n, dtype):
    """Tensor view (shared memory) of n elements of `dtype` at address ptr."""
    if ptr is None or ptr == 0:
        return None
    if dtype == torch.bfloat16:
        a = np.frombuffer((ctypes.c_uint16 * n).from_address(ptr), dtype=np.uint16)
        return torch.from_numpy(a).view(torch.bfloat16)
    a = np.frombuffer((ctypes.c_float * n).from_address(ptr), dtype=np.float32)
    return torch.from_numpy(a)


# ---- the in-kernel NoiseInjection RNG (csrc/noise_rng.cuh): counter-based, any element independently ----------------
GENERATED_NOISE = []          # forward launches append the N(0,1) tensors they drew (logical NHWC), in launch order
_M32 = 0xFFFFFFFF


def _fmix32(h):
    h = np.asarray(h, np.uint64) & _M32
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85EBCA6B)) & _M32
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE35)) & _M32
    h ^= h >> np.uint64(16)
    return h


def _noise_key(seed, subseq):
    seed, subseq = int(seed) & 0xFFFFFFFFFFFFFFFF, int(subseq) & 0xFFFFFFFFFFFFFFFF
    a = int(_fmix32((seed & _M32) ^ 0x9E3779B9)) ^ int(_fmix32(((subseq & _M32) + 0x7F4A7C15) & _M32))
    b = int(_fmix32(((seed >> 32) + 0x94D049BB) & _M32)) ^ int(_fmix32(((subseq >> 32) & _M32) ^ 0xBF58476D))
    return a, b


def _normals(key, idx):
    """normal_one(key, idx) for an int64 array of element indices (exact log / sin / cos instead of the MUFU approximations:
    differences of ~1e-6, irrelevant for what the interpreter checks — that forward and backward address the SAME noise)."""
    idx = np.asarray(idx, np.uint64)
    pair = idx >> np.uint64(1)
    x = _fmix32((((pair & _M32) ^ np.uint64(key[0])) + (pair >> np.uint64(32)) * np.uint64(0x9E3779B1)) & _M32)
    y = ((x ^ np.uint64(key[1])) * np.uint64(0x2C1B3C6D)) & _M32
    y ^= y >> np.uint64(16)
    u1 = ((x >> np.uint64(8)).astype(np.float64) + 0.5) / 2 ** 24
    ang = (y >> np.uint64(8)).astype(np.float64) * (2 * np.pi / 2 ** 24)
    r = np.sqrt(-2.0 * np.log(u1))
    z = np.where((idx & np.uint64(1)) == 0, r * np.cos(ang), r * np.sin(ang))
    return torch.from_numpy(z.astype(np.float32))


def _seed_total(seed, seed_dev):
    extra = int(np.frombuffer((ctypes.c_uint64 * 1).from_address(seed_dev), dtype=np.uint64)[0]) if seed_dev else 0
    return (int(seed) + extra) & 0xFFFFFFFFFFFFFFFF


def _act(v, act, slope):
    if act == _lib.ACT_RELU:
        return torch.relu(v)
    if act == _lib.ACT_LRELU:
        return torch.where(v > 0, v, slope * v)
    if act == _lib.ACT_LOGSOFTMAX:
        return torch.log_softmax(v, -1)
    assert act == _lib.ACT_NONE
    return v


def hwg_conv_fprop(d_addr, x, w, bias, noise, noise_w, stats, y, stream):
    """The full tap-list convolution of include/hwg_b200.h: input strides, channel folding (union taps or per-fold taps),
    bias -> noise tensor -> activation -> statistics -> (strided, folded) store.  In-kernel noise is not interpreted."""
    d = _lib.ConvDesc.from_address(d_addr)
    N, H, W, Ci, Cp, Co, Ho, Wo, T = d.N, d.H, d.W, d.Cin, d.x_pitch, d.Cout, d.Ho, d.Wo, d.ntaps
    rng = noise_w is not None and noise is None          # in-kernel noise: fold f draws from subsequence noise_subseq + f
    sh, sw = max(d.in_stride_h, 1), max(d.in_stride_w, 1)
    fc = d.fold_c if d.fold_c else Co
    F = Co // fc
    fw = d.fold_w if d.fold_w > 0 else 1
    rows = fc if (d.fold_c and d.fold_taps > 0) else Co                      # rows of one tap matrix
    xv = _view(x, N * H * W * Cp, torch.bfloat16).view(N, H, W, Cp)[..., :Ci].float()
    wv = _view(w, T * rows * Ci, torch.bfloat16).view(T, rows, Ci).float()
    taps = [(d.tap_dh[t], d.tap_dw[t]) for t in range(T)]
    P = max(max(abs(a), abs(b)) for a, b in taps) + max(Ho * sh, Wo * sw, H, W)   # out-of-bounds reads are zero
    xp = torch.nn.functional.pad(xv, (0, 0, P, P, P, P))

    def window(dh, dw):
        return xp[:, P + dh:P + dh + Ho * sh:sh, P + dw:P + dw + Wo * sw:sw, :]

    acc = torch.zeros((N, Ho, Wo, Co))
    for t, (dh, dw) in enumerate(taps):
        if d.fold_c and d.fold_taps > 0:
            f = t // d.fold_taps
            acc[..., f * fc:(f + 1) * fc] += window(dh, dw) @ wv[t].t()
        else:
            acc += window(dh, dw) @ wv[t].t()
    if bias is not None:
        acc += _view(bias, Co, torch.float32)
    ydt = torch.float32 if d.y_dtype == _lib.DT_F32 else torch.bfloat16

    def placed(ptr, dt, sn_, sh_, sw_, f):
        """[N,Ho,Wo,fc] view of fold f's pixels inside a tensor with these pixel strides."""
        disp = (f // fw) * d.fold_stride_h + (f % fw) * d.fold_stride_w if d.fold_c else 0
        span = disp + (N - 1) * sn_ + (Ho - 1) * sh_ + (Wo - 1) * sw_ + fc
        return torch.as_strided(_view(ptr, span, dt), (N, Ho, Wo, fc), (sn_, sh_, sw_, 1), disp)

    if noise is not None:
        nw = _view(noise_w, Co, torch.float32)
        for f in range(F):
            z = placed(noise, torch.float32, d.nz_stride_n, d.nz_stride_h, d.nz_stride_w, f)
            acc[..., f * fc:(f + 1) * fc] += nw[f * fc:(f + 1) * fc] * z
    elif rng:
        nw = _view(noise_w, Co, torch.float32)
        seed = _seed_total(d.noise_seed, d.noise_seed_dev)
        drawn = []
        for f in range(F):                                 # element index over the fold's logical [N,Ho,Wo,fold_c] output
            key = _noise_key(seed, d.noise_subseq + (f if d.fold_c else 0))
            z = _normals(key, np.arange(N * Ho * Wo * fc)).view(N, Ho, Wo, fc)
            acc[..., f * fc:(f + 1) * fc] += nw[f * fc:(f + 1) * fc] * z
            drawn.append(z)
        GENERATED_NOISE.append(torch.cat(drawn, 1) if F > 1 else drawn[0])     # folds of the initial conv = output rows
    out = _act(acc, d.act, d.slope).to(ydt)
    if stats is not None:
        st = _view(stats, N * fc * 2, torch.float32).view(N, fc, 2)
        of = out.float().view(N, Ho * Wo, F, fc)
        st[:, :, 0] += of.sum((1, 2))
        st[:, :, 1] += (of * of).sum((1, 2))
    for f in range(F):
        placed(y, ydt, d.y_stride_n, d.y_stride_h, d.y_stride_w, f).copy_(out[..., f * fc:(f + 1) * fc])
    return 0


def hwg_shift_expand(img, out, N, H, W, kw, pad, stream):
    iv = _view(img, N * H * W, torch.float32).view(N, H, W)
    ov = _view(out, N * H * W * 16, torch.bfloat16).view(N, H, W, 16)
    ip = torch.nn.functional.pad(iv, (16, 16))
    ov.zero_()
    for j in range(kw):
        ov[..., j] = ip[:, :, 16 + j - pad:16 + j - pad + W].to(torch.bfloat16)
    return 0


def hwg_stem_conv(img, w, bias, N, H, W, kh, kw, pad_h, pad_w, Cout, y, stats, stream):
    iv = _view(img, N * H * W, torch.float32).view(N, 1, H, W).to(torch.bfloat16).float()
    wv = _view(w, kh * Cout * 16, torch.bfloat16).view(kh, Cout, 16).float()[:, :, :kw]        # [kh][Cout][kw]
    wt = wv.permute(1, 0, 2).reshape(Cout, 1, kh, kw)
    b = _view(bias, Cout, torch.float32) if bias else None
    out = torch.nn.functional.conv2d(iv, wt, b, padding=(pad_h, pad_w))                        # [N,Cout,Ho,Wo] fp32
    Ho, Wo = out.shape[2], out.shape[3]
    if stats:
        st = _view(stats, N * Cout * 2, torch.float32).view(N, Cout, 2)
        st[:, :, 0] += out.sum((2, 3))
        st[:, :, 1] += (out * out).sum((2, 3))
    _view(y, N * Ho * Wo * Cout, torch.bfloat16).view(N, Ho, Wo, Cout).copy_(out.permute(0, 2, 3, 1).to(torch.bfloat16))
    return 0


def hwg_shift_collapse(g, dimg, N, H, W, kw, pad, accumulate, stream):
    gv = _view(g, N * H * W * 16, torch.bfloat16).view(N, H, W, 16).float()
    dv = _view(dimg, N * H * W, torch.float32).view(N, H, W)
    gp = torch.nn.functional.pad(gv, (0, 0, 16, 16))
    acc = torch.zeros((N, H, W))
    for j in range(kw):
        acc += gp[:, :, 16 - j + pad:16 - j + pad + W, j]
    dv.copy_(dv + acc if accumulate else acc)
    return 0


def hwg_gn_coeffs(stats, gamma, beta, N, C, groups, HW, eps, coef, save, stream):
    st = _view(stats, N * C * 2, torch.float32).view(N, groups, C // groups, 2)
    cnt = (C // groups) * HW
    mean = st[..., 0].sum(2) / cnt
    var = st[..., 1].sum(2) / cnt - mean * mean
    rstd = 1.0 / torch.sqrt(var + eps)
    mean_c, rstd_c = (t.repeat_interleave(C // groups, 1) for t in (mean, rstd))
    gm = _view(gamma, C, torch.float32) if gamma else torch.ones(C)
    bt = _view(beta, C, torch.float32) if beta else torch.zeros(C)
    a = gm * rstd_c
    _view(coef, N * C * 2, torch.float32).view(N, C, 2).copy_(torch.stack((a, bt - mean_c * a), -1))
    if save:
        _view(save, N * C * 2, torch.float32).view(N, C, 2).copy_(torch.stack((mean_c, rstd_c), -1))
    return 0


def hwg_scale_shift_act(x, out, coef, per_sample, N, HW, C, act, slope, stream):
    xv = _view(x, N * HW * C, torch.bfloat16).view(N, HW, C).float()
    cf = _view(coef, (N if per_sample else 1) * C * 2, torch.float32).view(-1, 1, C, 2)
    _view(out, N * HW * C, torch.bfloat16).view(N, HW, C).copy_(_act(cf[..., 0] * xv + cf[..., 1], act, slope))
    return 0


def hwg_avgpool_nhwc(x, y, N, H, W, C, kh, kw, stream):
    xv = _view(x, N * H * W * C, torch.bfloat16).view(N, H, W, C).float()
    Ho, Wo = H // kh, W // kw
    p = xv[:, :Ho * kh, :Wo * kw].reshape(N, Ho, kh, Wo, kw, C).mean((2, 4))
    _view(y, N * Ho * Wo * C, torch.bfloat16).view(N, Ho, Wo, C).copy_(p)
    return 0


def hwg_add_stats(a, b, y, N, HW, C, stats, stream):
    s = (_view(a, N * HW * C, torch.bfloat16).float() + _view(b, N * HW * C, torch.bfloat16).float()).to(torch.bfloat16)
    _view(y, N * HW * C, torch.bfloat16).copy_(s)
    if stats:
        st = _view(stats, N * C * 2, torch.float32).view(N, C, 2)
        sf = s.float().view(N, HW, C)
        st[:, :, 0] += sf.sum(1)
        st[:, :, 1] += (sf * sf).sum(1)
    return 0


def hwg_l1_halves(f, dtype, half, loss_scale, grad_scale, loss, g, stream):
    fv = _view(f, 2 * half, torch.float32 if dtype == _lib.DT_F32 else torch.bfloat16).float()
    d = fv[half:] - fv[:half]
    _view(loss, 1, torch.float32).add_(loss_scale * d.abs().sum())
    if g:
        _view(g, half, torch.bfloat16).copy_(grad_scale * torch.sign(d))
    return 0


def _up(g, N, H, W, C, kh, kw):
    """Gradient of the AvgPool2d(kh,kw) output spread back over the [N,H,W,C] grid (zero where the floor drops rows)."""
    Ho, Wo = H // kh, W // kw
    gv = _view(g, N * Ho * Wo * C, torch.bfloat16).view(N, Ho, Wo, C).float() / (kh * kw)
    full = torch.zeros((N, H, W, C))
    full[:, :Ho * kh, :Wo * kw] = gv.repeat_interleave(kh, 1).repeat_interleave(kw, 2)
    return full


def _gyp(g, z, coef, slope, N, H, W, C, kh, kw):
    zv = _view(z, N * H * W * C, torch.bfloat16).view(N, H, W, C).float()
    cf = _view(coef, N * C * 2, torch.float32).view(N, 1, 1, C, 2)
    pre = cf[..., 0] * zv + cf[..., 1]
    return _up(g, N, H, W, C, kh, kw) * torch.where(pre > 0, torch.ones(()), torch.full((), slope)), zv


def hwg_norm_bwd_reduce(g, z, coef, slope, N, H, W, C, kh, kw, sums, stream):
    gy, zv = _gyp(g, z, coef, slope, N, H, W, C, kh, kw)
    sv = _view(sums, N * C * 2, torch.float32).view(N, C, 2)
    sv[:, :, 0] += gy.sum((1, 2))
    sv[:, :, 1] += (gy * zv).sum((1, 2))
    return 0


def hwg_gn_bwd_coeffs(sums, save, gamma, N, C, groups, HW, spq, dgamma, dbeta, stream):
    cg = C // groups
    sv = _view(sums, N * C * 2, torch.float32).view(N, C, 2)
    sa = _view(save, N * C * 2, torch.float32).view(N, C, 2)
    gm = _view(gamma, C, torch.float32) if gamma else torch.ones(C)
    mu, r = sa[..., 0], sa[..., 1]
    T = r * (sv[..., 1] - mu * sv[..., 0])
    cnt = cg * HW
    A = (gm * sv[..., 0]).view(N, groups, cg).sum(2).repeat_interleave(cg, 1) / cnt
    B = (gm * T).view(N, groups, cg).sum(2).repeat_interleave(cg, 1) / cnt
    _view(spq, N * C * 3, torch.float32).view(N, C, 3).copy_(torch.stack((r * gm, -r * r * B, -r * A + r * r * mu * B), -1))
    if dgamma:
        _view(dgamma, C, torch.float32).add_(T.sum(0))
    if dbeta:
        _view(dbeta, C, torch.float32).add_(sv[..., 0].sum(0))
    return 0


def hwg_norm_bwd_apply(g, z, coef, spq, slope, N, H, W, C, kh, kw, gz, stream):
    gy, zv = _gyp(g, z, coef, slope, N, H, W, C, kh, kw)
    sp = _view(spq, N * C * 3, torch.float32).view(N, 1, 1, C, 3)
    _view(gz, N * H * W * C, torch.bfloat16).view(N, H, W, C).copy_(sp[..., 0] * gy + sp[..., 1] * zv + sp[..., 2])
    return 0


def hwg_act_bwd(g, y, scale, slope, N, H, W, C, kh, kw, gz, stream):
    yv = _view(y, N * H * W * C, torch.bfloat16).view(N, H, W, C).float()
    out = _up(g, N, H, W, C, kh, kw) * torch.where(yv > 0, torch.ones(()), torch.full((), slope))
    if scale:
        out = out * _view(scale, N * C, torch.float32).view(N, 1, 1, C)
    _view(gz, N * H * W * C, torch.bfloat16).view(N, H, W, C).copy_(out)
    return 0


def hwg_balance(g_main, sets_host, K, x_dev, seg_off, seg_len, nseg, block_tab, nblocks, sums, mult, stream):
    """Header semantics (include/hwg_b200.h): D_s += x_k * R_k,s * (mean|D_s| / mean|R_k,s|), mean|D_s| taken first, exact
    zeros replaced by the average of the non-zero means, sets with mean|R_k,s| == 0 skipped.  Also checks that the
    (segment, chunk) table covers every segment exactly once per chunk."""
    chunk = _lib.load().hwg_balance_chunk()
    off = np.frombuffer((ctypes.c_int64 * nseg).from_address(seg_off), dtype=np.int64)
    ln = np.frombuffer((ctypes.c_int64 * nseg).from_address(seg_len), dtype=np.int64)
    tab = np.frombuffer((ctypes.c_int32 * (2 * nblocks)).from_address(block_tab), dtype=np.int32).reshape(-1, 2)
    want = [(s_, c) for s_ in range(nseg) for c in range(-(-int(ln[s_]) // chunk))]
    assert sorted(map(tuple, tab.tolist())) == want, "block table does not tile the segments"
    total = int((off + ln).max())
    g = _view(g_main, total, torch.float32)
    ptrs = (ctypes.c_void_p * K).from_address(sets_host)
    sets = [_view(ptrs[k], total, torch.float32) for k in range(K)]
    x = _view(x_dev, K, torch.float32)
    means = [g[int(o):int(o + n)].abs().mean() for o, n in zip(off, ln)]
    nz = [m for m in means if m != 0]
    fill = sum(nz) / len(nz) if nz else None
    for s_, (o, n) in enumerate(zip(off, ln)):
        o, n = int(o), int(n)
        mD = means[s_] if (means[s_] != 0 or fill is None) else fill
        D = g[o:o + n]
        add = torch.zeros(n)
        for k in range(K):
            R = sets[k][o:o + n]
            r = R.abs().mean()
            if r != 0:
                add += x[k] * R * (mD / r)
        D += add
    return 0


# ---- recognizer (cnn_only_hwr.py) ---------------------------------------------------------------------------------
def hwg_bn_coeffs(stats, N, C, count_per_n, weight, bias, rmean, rvar, momentum, eps, use_batch_stats, coef, save, stream):
    w = _view(weight, C, torch.float32) if weight else torch.ones(C)
    b = _view(bias, C, torch.float32) if bias else torch.zeros(C)
    if use_batch_stats:
        tot = _view(stats, N * C * 2, torch.float32).view(N, C, 2).sum(0)
        cnt = float(count_per_n * N)
        mean = tot[:, 0] / cnt
        var = (tot[:, 1] / cnt - mean * mean).clamp_min(0)
        if rmean:
            rm, rv = _view(rmean, C, torch.float32), _view(rvar, C, torch.float32)
            rm.copy_((1 - momentum) * rm + momentum * mean)
            rv.copy_((1 - momentum) * rv + momentum * var * (cnt / max(cnt - 1, 1)))
    else:
        mean, var = _view(rmean, C, torch.float32).clone(), _view(rvar, C, torch.float32).clone()
    rstd = 1.0 / torch.sqrt(var + eps)
    a = w * rstd
    _view(coef, 2 * C, torch.float32).view(C, 2).copy_(torch.stack((a, b - mean * a), -1))
    if save:
        _view(save, 2 * C, torch.float32).view(C, 2).copy_(torch.stack((mean, rstd), -1))
    return 0


def _stem(img, w, b, N, H, W, Cout):
    iv = _view(img, N * H * W, torch.float32).view(N, 1, H, W).clone().requires_grad_()
    wv = _view(w, Cout * 9, torch.float32).view(Cout, 1, 3, 3).clone().requires_grad_()
    bv = _view(b, Cout, torch.float32).clone().requires_grad_()
    with torch.enable_grad():                        # the modules call their backward entry points under no_grad
        y = torch.nn.functional.max_pool2d(torch.relu(torch.nn.functional.conv2d(iv, wv, bv, padding=1)), 2, 2)
    return iv, wv, bv, y


def hwg_hwr_stem(img, w, b, N, H, W, Cout, y, stream):
    with torch.no_grad():
        out = _stem(img, w, b, N, H, W, Cout)[3]
    _view(y, out.numel(), torch.bfloat16).view(N, H // 2, W // 2, Cout).copy_(out.permute(0, 2, 3, 1))
    return 0


def _stem_grads(img, w, b, ga, N, H, W, Cout):
    iv, wv, bv, y = _stem(img, w, b, N, H, W, Cout)
    g = _view(ga, y.numel(), torch.bfloat16).view(N, H // 2, W // 2, Cout).float().permute(0, 3, 1, 2)
    return torch.autograd.grad(y, (iv, wv, bv), g)


def hwg_hwr_stem_bwd(img, w, b, ga, N, H, W, Cout, dw, db, stream):
    _, gw, gb = _stem_grads(img, w, b, ga, N, H, W, Cout)
    _view(dw, Cout * 9, torch.float32).add_(gw.reshape(-1))
    _view(db, Cout, torch.float32).add_(gb)
    return 0


def hwg_hwr_stem_bwd_image(img, w, b, ga, N, H, W, Cout, gimg, stream):
    gi, _, _ = _stem_grads(img, w, b, ga, N, H, W, Cout)
    _view(gimg, N * H * W, torch.float32).add_(gi.reshape(-1))
    return 0


def hwg_maxpool_nhwc(x, y, N, H, W, C, kh, kw, sh, sw, ph, pw, Ho, Wo, stream):
    xv = _view(x, N * H * W * C, torch.bfloat16).view(N, H, W, C).float().permute(0, 3, 1, 2)
    out = torch.nn.functional.max_pool2d(xv, (kh, kw), (sh, sw), (ph, pw))
    assert tuple(out.shape[2:]) == (Ho, Wo)
    _view(y, N * Ho * Wo * C, torch.bfloat16).view(N, Ho, Wo, C).copy_(out.permute(0, 2, 3, 1))
    return 0


def hwg_relu_maxpool_bwd(ga, c, N, H, W, C, kh, kw, sh, sw, ph, pw, Ho, Wo, gc, dbias, stream):
    cv = _view(c, N * H * W * C, torch.bfloat16).view(N, H, W, C).float().permute(0, 3, 1, 2).clone().requires_grad_()
    with torch.enable_grad():
        out = torch.nn.functional.max_pool2d(cv, (kh, kw), (sh, sw), (ph, pw))   # first maximum wins, as in the kernel
    g = _view(ga, N * Ho * Wo * C, torch.bfloat16).view(N, Ho, Wo, C).float().permute(0, 3, 1, 2)
    (gx,) = torch.autograd.grad(out, cv, g)
    gx = gx * (cv.detach() > 0)
    _view(gc, N * H * W * C, torch.bfloat16).view(N, H, W, C).copy_(gx.permute(0, 2, 3, 1))
    _view(dbias, C, torch.float32).add_(gx.sum((0, 2, 3)))
    return 0


def hwg_logsoftmax_bwd(g, lp, T, B, C, Cp, gz, dbias, stream):
    gv = _view(g, T * B * C, torch.float32).view(T, B, C)
    lv = _view(lp, T * B * C, torch.float32).view(T, B, C)
    z = gv - torch.exp(lv) * gv.sum(2, keepdim=True)
    out = _view(gz, B * T * Cp, torch.bfloat16).view(B, 1, T, Cp)
    out.zero_()
    out[:, 0, :, :C] = z.permute(1, 0, 2).to(torch.bfloat16)
    _view(dbias, C, torch.float32).add_(z.sum((0, 1)))
    return 0


def _bn_gy(g, z, coef, save, rows, C, relu):
    gv = _view(g, rows * C, torch.bfloat16).view(rows, C).float()
    zv = _view(z, rows * C, torch.bfloat16).view(rows, C).float()
    cf = _view(coef, 2 * C, torch.float32).view(C, 2)
    sv = _view(save, 2 * C, torch.float32).view(C, 2)
    gy = gv * ((cf[:, 0] * zv + cf[:, 1]) > 0) if relu else gv
    return gy, (zv - sv[:, 0]) * sv[:, 1], sv[:, 1]


def hwg_bn_bwd_reduce(g, z, coef, save, rows, C, relu, sums, stream):
    gy, xhat, _ = _bn_gy(g, z, coef, save, rows, C, relu)
    sv = _view(sums, 2 * C, torch.float32).view(C, 2)
    sv[:, 0] += gy.sum(0)
    sv[:, 1] += (gy * xhat).sum(0)
    return 0


def hwg_bn_bwd_apply(g, z, coef, save, weight, sums, rows, norm_rows, C, relu, gz, dconv_bias, stream):
    gy, xhat, rstd = _bn_gy(g, z, coef, save, rows, C, relu)
    sv = _view(sums, 2 * C, torch.float32).view(C, 2)
    M = float(norm_rows if norm_rows > 0 else rows)
    out = _view(weight, C, torch.float32) * rstd * (gy - sv[:, 0] / M - xhat * sv[:, 1] / M)
    _view(gz, rows * C, torch.bfloat16).view(rows, C).copy_(out)
    _view(dconv_bias, C, torch.float32).add_(out.sum(0))
    return 0


# ---- CTC (model/loss.py:28-30) through torch's own F.ctc_loss ---------------------------------------------------------
def _i32(ptr, n):
    return torch.from_numpy(np.frombuffer((ctypes.c_int32 * n).from_address(ptr), dtype=np.int32))


def _ctc_args(lp, T, B, C, tg, ts_b, ts_s, S, il, tl):
    lpv = _view(lp, T * B * C, torch.float32).view(T, B, C)
    if S:
        span = (B - 1) * ts_b + (S - 1) * ts_s + 1
        tgt = torch.as_strided(_i32(tg, span), (B, S), (ts_b, ts_s)).long()
    else:
        tgt = torch.zeros((B, 0), dtype=torch.long)
    return lpv, tgt, _i32(il, B).long(), _i32(tl, B).long()


def hwg_ctc_forward(lp, T, B, C, tg, ts_b, ts_s, S, il, tl, blank, nll, log_alpha, log_beta, stream):
    lpv, tgt, ilv, tlv = _ctc_args(lp, T, B, C, tg, ts_b, ts_s, S, il, tl)
    _view(nll, B, torch.float32).copy_(torch.nn.functional.ctc_loss(lpv, tgt, ilv, tlv, blank=blank, reduction="none"))
    return 0


def hwg_ctc_reduce_mean(nll, tl, B, loss, unit, stream):
    n = _view(nll, B, torch.float32)
    s = _i32(tl, B).float().clamp_min(1)
    v = (n / s).mean()
    bad = bool(torch.isinf(v))
    _view(loss, 1, torch.float32).fill_(0.0 if bad else float(v))
    _view(unit, B, torch.float32).copy_(torch.zeros(B) if bad else 1.0 / (B * s))
    return 0


def hwg_ctc_backward(go, unit, lp, T, B, C, tg, ts_b, ts_s, S, il, tl, blank, nll, la, lb, beta_ready, grad, stream):
    lpv, tgt, ilv, tlv = _ctc_args(lp, T, B, C, tg, ts_b, ts_s, S, il, tl)
    x = lpv.clone().requires_grad_()
    with torch.enable_grad():
        per = torch.nn.functional.ctc_loss(x, tgt, ilv, tlv, blank=blank, reduction="none")
        tot = (per * _view(unit, B, torch.float32)).sum() * _view(go, 1, torch.float32)[0]
    (g,) = torch.autograd.grad(tot, x)
    _view(grad, T * B * C, torch.float32).view(T, B, C).copy_(g)
    return 0


def hwg_dtw_align(pred, T, B, C, label, ls_s, ls_b, S, hist, out, out_len, scratch, stream):
    """model/hw_with_style.py:18-74 through the oracle's restatement (bit-exact against the reference goldens); checks the
    host wrapper's buffers: out [T+L,B] zero-filled by the caller, out_len [B]."""
    from oracle import style as ostyle
    pv = _view(pred, T * B * C, torch.float32).view(T, B, C).numpy()
    span = (S - 1) * ls_s + (B - 1) * ls_b + 1
    lab = torch.as_strided(_i32(label, span), (S, B), (ls_s, ls_b)).numpy()
    L = 2 * S + 1
    ov = _i32(out, (T + L) * B).view(T + L, B)
    assert int(ov.abs().max()) == 0, "out must be zero-filled by the caller"
    lv = _i32(out_len, B)
    for b in range(B):
        path = ostyle.correct_pred(pv[:, b:b + 1], lab[:, b:b + 1])[:, 0]      # per line: its own (unpadded) path
        ov[:len(path), b] = torch.from_numpy(path.astype(np.int32))
        lv[b] = len(path)
    return 0


def hwg_adam_flat(p, g, m, v, n, lr, beta1, beta2, eps, clip_value, grad_scale, step_dev, zero_grad, stream):
    """include/hwg_b200.h: clip_grad_value_ + Adam (bias-corrected, as torch.optim.Adam) + zero_grad over flat buffers."""
    pv, gv, mv, vv = (_view(x, n, torch.float32) for x in (p, g, m, v))
    st = _view(step_dev, 1, torch.float32)
    st += 1
    t = float(st[0])
    gg = gv * grad_scale
    if clip_value > 0:
        gg = gg.clamp(-clip_value, clip_value)
    mv.copy_(mv + (gg - mv) * (1 - beta1))
    vv.copy_(beta2 * vv + (1 - beta2) * gg * gg)
    pv.sub_(lr / (1 - beta1 ** t) * mv / (vv.sqrt() / (1 - beta2 ** t) ** 0.5 + eps))
    if zero_grad:
        gv.zero_()
    return 0


def _i64(ptr, n):
    return np.frombuffer((ctypes.c_int64 * n).from_address(ptr), dtype=np.int64)


def hwg_spectral_norm(jobs_dev, njobs, max_h, max_wd, norms, inv_sigma, stream):
    """SpectralNorm._update_u_v (discriminator_ap.py:19-32), one power iteration: v = l2n(W^T u), u = l2n(W v) in place,
    inv_sigma = 1 / (u . W v); l2n(x) = x / (|x| + 1e-12)."""
    jobs = _i64(jobs_dev, 4 * njobs).reshape(njobs, 4)
    inv = _view(inv_sigma, njobs, torch.float32)
    for i, (w, u, v, hw) in enumerate(jobs.tolist()):
        h, wd = hw & 0xffffffff, hw >> 32
        W, uu, vv = _view(w, h * wd, torch.float32).view(h, wd), _view(u, h, torch.float32), _view(v, wd, torch.float32)
        nv = W.t().mv(uu)
        vv.copy_(nv / (nv.norm() + 1e-12))
        nu = W.mv(vv)
        uu.copy_(nu / (nu.norm() + 1e-12))
        inv[i] = 1.0 / uu.dot(W.mv(vv))
    return 0


def hwg_spectral_norm_bwd(jobs_dev, njobs, max_elems, inv_sigma, dots, stream):
    """In place gw -= (<gw, w_bar> * inv_sigma[layer]) * u v^T."""
    jobs = _i64(jobs_dev, 5 * njobs).reshape(njobs, 5)
    inv = _view(inv_sigma, njobs, torch.float32)
    for i, (w, gw, u, v, hw) in enumerate(jobs.tolist()):
        h, wd = hw & 0xffffffff, hw >> 32
        W, G = _view(w, h * wd, torch.float32).view(h, wd), _view(gw, h * wd, torch.float32).view(h, wd)
        G -= ((G * W).sum() * inv[i]) * torch.outer(_view(u, h, torch.float32), _view(v, wd, torch.float32))
    return 0


def hwg_channel_sum(x, rows, C, out, stream):
    _view(out, C, torch.float32).add_(_view(x, rows * C, torch.bfloat16).view(rows, C).float().sum(0))
    return 0


def hwg_conv_wgrad(d_addr, x, gy, dw, stream):
    """dw[t][co][ci] += sum over the iteration grid (i, j) of gy[n, i*gs_h + go_h + ph_t, j*gs_w + go_w + pw_t, co] *
    x[n, i + dh_t, j + dw_t, ci]  (grid 0 = the gy extent; strides / offsets / per-tap phases: the up-sampling layers)."""
    d = _lib.WgradDesc.from_address(d_addr)
    N, H, W, Ci, Cp, Ho, Wo, Co, Gp, T = d.N, d.H, d.W, d.Cin, d.x_pitch, d.Ho, d.Wo, d.Cout, d.gy_pitch, d.ntaps
    Hi, Wi = (d.Hi or Ho), (d.Wi or Wo)
    gsh, gsw = max(d.gy_stride_h, 1), max(d.gy_stride_w, 1)
    xv = _view(x, N * H * W * Cp, torch.bfloat16).view(N, H, W, Cp)[..., :Ci].float()
    gv = _view(gy, N * Ho * Wo * Gp, torch.bfloat16).view(N, Ho, Wo, Gp)[..., :Co].float()
    out = _view(dw, T * Co * Ci, torch.float32).view(T, Co, Ci)
    P = max(max(abs(d.tap_dh[t]), abs(d.tap_dw[t])) for t in range(T)) + max(Hi, Wi, H, W)
    xp = torch.nn.functional.pad(xv, (0, 0, P, P, P, P))
    Q = max(Hi * gsh, Wi * gsw)
    gp = torch.nn.functional.pad(gv, (0, 0, 0, Q, 0, Q))                         # grid points beyond gy contribute zero
    for t in range(T):
        xs = xp[:, P + d.tap_dh[t]:P + d.tap_dh[t] + Hi, P + d.tap_dw[t]:P + d.tap_dw[t] + Wi, :]
        oh, ow = d.gy_off_h + d.tap_gy_h[t], d.gy_off_w + d.tap_gy_w[t]
        gs = gp[:, oh:oh + Hi * gsh:gsh, ow:ow + Wi * gsw:gsw, :]
        out[t] += torch.einsum("nhwo,nhwi->oi", gs, xs)
    return 0


# ---- generator (pure_gen.py) ----------------------------------------------------------------------------------------
def hwg_linear_f32(inp, W, bias, out, B, K, O, act, slope, stream):
    v = _view(inp, B * K, torch.float32).view(B, K) @ _view(W, O * K, torch.float32).view(O, K).t()
    if bias:
        v = v + _view(bias, O, torch.float32)
    _view(out, B * O, torch.float32).view(B, O).copy_(_act(v, act, slope))
    return 0


def hwg_linear_bwd_f32(x, y, gy, W, B, K, O, act, slope, gx, gW, gb, accumulate, stream):
    g = _view(gy, B * O, torch.float32).view(B, O).clone()
    if act == _lib.ACT_LRELU:
        g = g * torch.where(_view(y, B * O, torch.float32).view(B, O) > 0, torch.ones(()), torch.full((), slope))
    xv, Wv = _view(x, B * K, torch.float32).view(B, K), _view(W, O * K, torch.float32).view(O, K)
    if gx:
        _view(gx, B * K, torch.float32).view(B, K).add_(g @ Wv)
    for ptr, val, n in ((gW, g.t() @ xv, O * K), (gb, g.sum(0), O)):
        if ptr:
            dst = _view(ptr, n, torch.float32)
            dst.copy_(dst + val.reshape(-1) if accumulate else val.reshape(-1))
    return 0


def hwg_pixelnorm_f32(inp, out, B, K, stream):
    v = _view(inp, B * K, torch.float32).view(B, K)
    _view(out, B * K, torch.float32).view(B, K).copy_(v / torch.sqrt((v * v).mean(1, keepdim=True) + 1e-8))
    return 0


def hwg_gen_pack_input(content, cs_t, cs_b, cs_c, style, T, B, C, S, Cp, x, stream):
    span = (T - 1) * cs_t + (B - 1) * cs_b + (C - 1) * cs_c + 1
    cv = torch.as_strided(_view(content, span, torch.float32), (T, B, C), (cs_t, cs_b, cs_c))
    xv = _view(x, B * T * Cp, torch.bfloat16).view(B, 1, T, Cp)
    xv.zero_()
    xv[:, 0, :, :C] = cv.permute(1, 0, 2).to(torch.bfloat16)
    if style and S:
        xv[:, 0, :, C:C + S] = _view(style, B * S, torch.float32).view(B, 1, S).to(torch.bfloat16)
    return 0


def hwg_adain_coeffs(stats, gamma, beta, gb_stride, N, C, HW, eps, coef, save, stream):
    st = _view(stats, N * C * 2, torch.float32).view(N, C, 2)
    span = (N - 1) * gb_stride + C
    gm = torch.as_strided(_view(gamma, span, torch.float32), (N, C), (gb_stride, 1))
    bt = torch.as_strided(_view(beta, span, torch.float32), (N, C), (gb_stride, 1))
    mean = st[..., 0] / HW
    var = (st[..., 1] / HW - mean * mean).clamp_min(0)
    rstd = 1.0 / torch.sqrt(var + eps)
    a = gm * rstd
    _view(coef, N * C * 2, torch.float32).view(N, C, 2).copy_(torch.stack((a, bt - mean * a), -1))
    if save:
        _view(save, N * C * 2, torch.float32).view(N, C, 2).copy_(torch.stack((mean, rstd), -1))
    return 0


def hwg_blur_noise_act_stats(x, y, N, H, W, C, noise, noise_w, seed, subseq, seed_dev, act, slope, stats, stream):
    xv = _view(x, N * H * W * C, torch.bfloat16).view(N, H, W, C).float().permute(0, 3, 1, 2)
    k = torch.tensor([1.0, 2.0, 1.0])
    k = (k[:, None] * k[None, :] / 16.0).expand(C, 1, 3, 3)
    v = torch.nn.functional.conv2d(xv, k, padding=1, groups=C).permute(0, 2, 3, 1)
    if noise_w is not None:
        if noise is not None:
            z = _view(noise, N * H * W * C, torch.float32).view(N, H, W, C)
        else:
            z = _normals(_noise_key(_seed_total(seed, seed_dev), subseq), np.arange(N * H * W * C)).view(N, H, W, C)
            GENERATED_NOISE.append(z)
        v = v + _view(noise_w, C, torch.float32) * z
    out = _act(v, act, slope).to(torch.bfloat16)
    _view(y, N * H * W * C, torch.bfloat16).view(N, H, W, C).copy_(out)
    if stats:
        st = _view(stats, N * C * 2, torch.float32).view(N, C, 2)
        of = out.float()
        st[:, :, 0] += of.sum((1, 2))
        st[:, :, 1] += (of * of).sum((1, 2))
    return 0


def hwg_gen_output(x, coef, w, b0, N, HW, C, out, stream):
    xv = _view(x, N * HW * C, torch.bfloat16).view(N, HW, C).float()
    cf = _view(coef, N * C * 2, torch.float32).view(N, 1, C, 2)
    v = ((cf[..., 0] * xv + cf[..., 1]) * _view(w, C, torch.float32)).sum(2) + _view(b0, 1, torch.float32)
    _view(out, N * HW, torch.float32).view(N, HW).copy_(torch.tanh(v))
    return 0


def hwg_gen_output_bwd(g_out, out, a, coef, w, N, HW, C, gx, dwb, stream):
    dv = _view(g_out, N * HW, torch.float32).view(N, HW) * (1 - _view(out, N * HW, torch.float32).view(N, HW) ** 2)
    av = _view(a, N * HW * C, torch.bfloat16).view(N, HW, C).float()
    cf = _view(coef, N * C * 2, torch.float32).view(N, 1, C, 2)
    wv = _view(w, C, torch.float32)
    _view(gx, N * HW * C, torch.bfloat16).view(N, HW, C).copy_(dv[..., None] * wv)      # w.r.t. the AdaIN output
    acc = _view(dwb, C + 1, torch.float32)
    acc[:C] += (dv[..., None] * (cf[..., 0] * av + cf[..., 1])).sum((0, 1))
    acc[C] += dv.sum()
    return 0


def _adain_ahat(a, save, N, HW, C):
    av = _view(a, N * HW * C, torch.bfloat16).view(N, HW, C).float()
    sv = _view(save, N * C * 2, torch.float32).view(N, 1, C, 2)
    return av, (av - sv[..., 0]) * sv[..., 1]


def hwg_adain_bwd_reduce(g, a, save, N, HW, C, sums, stream):
    gv = _view(g, N * HW * C, torch.bfloat16).view(N, HW, C).float()
    _, ahat = _adain_ahat(a, save, N, HW, C)
    sv = _view(sums, N * C * 2, torch.float32).view(N, C, 2)
    sv[:, :, 0] += gv.sum(1)
    sv[:, :, 1] += (gv * ahat).sum(1)
    return 0


def hwg_adain_bwd_apply(g, a, save, coef, sums, N, H, W, C, slope, noise, seed, subseq, seed_dev, row_subseq, gy, dch,
                        stream):
    HW = H * W
    gv = _view(g, N * HW * C, torch.bfloat16).view(N, HW, C).float()
    av, ahat = _adain_ahat(a, save, N, HW, C)
    A = _view(coef, N * C * 2, torch.float32).view(N, 1, C, 2)[..., 0]
    sv = _view(sums, N * C * 2, torch.float32).view(N, 1, C, 2)
    ga = A * (gv - sv[..., 0] / HW - ahat * sv[..., 1] / HW)
    out = ga * torch.where(av > 0, torch.ones(()), torch.full((), slope))
    _view(gy, N * HW * C, torch.bfloat16).view(N, HW, C).copy_(out)
    dc = _view(dch, 2 * C, torch.float32).view(C, 2)
    dc[:, 0] += out.sum((0, 1))
    if noise is not None:
        z = _view(noise, N * HW * C, torch.float32).view(N, HW, C)
    else:                                               # regenerated from (seed, subsequence), in the FORWARD's numbering
        st = _seed_total(seed, seed_dev)
        if row_subseq:                                  # the initial conv's four output rows were four folds
            rows = [_normals(_noise_key(st, subseq + h), np.arange(N * W * C)).view(N, 1, W, C) for h in range(H)]
            z = torch.cat(rows, 1).reshape(N, HW, C)
        else:
            z = _normals(_noise_key(st, subseq), np.arange(N * HW * C)).view(N, HW, C)
    dc[:, 1] += (out * z).sum((0, 1))
    return 0


# ---- text spacing (spacing.py) -------------------------------------------------------------------------------------
def _iview(ptr, n, ctype, npdtype):
    return np.frombuffer((ctype * n).from_address(ptr), dtype=npdtype)


def hwg_insert_spaces_plan(lengths, counts, n_out, z, z_off, L, B, count_std, dup_std, reps, offsets, info, stream):
    """Header semantics: per line the rounded normal draws (doubles, half to even), exclusive prefix sums, line length,
    ceil(max counts) — plain Python loops, as the reference itself walks them."""
    ln = _iview(lengths, B, ctypes.c_int32, np.int32)
    c = _view(counts, L * B * n_out, torch.float32).view(L, B, n_out).numpy()
    zoff = _iview(z_off, B, ctypes.c_int64, np.int64)
    nz = int(sum(int(v) for v in ln) * n_out)
    zz = _iview(z, max(nz, 1), ctypes.c_double, np.float64)
    rp = _iview(reps, B * L * 2, ctypes.c_int32, np.int32).reshape(B, L, 2)
    off = _iview(offsets, B * (L + 1), ctypes.c_int32, np.int32).reshape(B, L + 1)
    inf = _iview(info, B + 1, ctypes.c_int32, np.int32)
    for b in range(B):
        n = min(max(int(ln[b]), 0), L)
        tot = 0
        for i in range(n):
            cnt = max(0, int(round(np.float64(c[i, b, 0]) + np.float64(count_std) * zz[zoff[b] + i * n_out])))
            dup = max(0, int(round(np.float64(c[i, b, 1]) + np.float64(dup_std) * zz[zoff[b] + i * n_out + 1]))) if n_out > 1 else 1
            rp[b, i] = (cnt, dup)
            off[b, i] = tot
            tot += cnt + dup
        off[b, n:] = tot
        inf[b] = tot
    inf[B] = int(np.ceil(c.max()))
    return 0


def hwg_insert_spaces_fill(label, is_i64, ls_l, ls_b, lengths, reps, offsets, L, B, T, C, spaced, stream):
    ln = _iview(lengths, B, ctypes.c_int32, np.int32)
    rp = _iview(reps, B * L * 2, ctypes.c_int32, np.int32).reshape(B, L, 2)
    out = _view(spaced, T * B * C, torch.float32).view(T, B, C)
    ct, dt = (ctypes.c_int64, np.int64) if is_i64 else (ctypes.c_int32, np.int32)
    span = (L - 1) * ls_l + (B - 1) * ls_b + 1
    lab = _iview(label, span, ct, dt)
    out.zero_()
    for b in range(B):
        line = []
        for i in range(min(max(int(ln[b]), 0), L)):
            line += [0] * int(rp[b, i, 0]) + [int(lab[i * ls_l + b * ls_b])] * int(rp[b, i, 1])
        line += [0] * (T - len(line))
        out[torch.arange(T), b, torch.tensor(line[:T])] = 1.0
    return 0


_TABLE = {f.__name__: f for f in (hwg_insert_spaces_plan, hwg_insert_spaces_fill, hwg_dtw_align, hwg_adam_flat, hwg_ctc_forward, hwg_ctc_reduce_mean, hwg_ctc_backward, hwg_linear_f32, hwg_linear_bwd_f32, hwg_pixelnorm_f32, hwg_gen_pack_input, hwg_adain_coeffs,
                                  hwg_blur_noise_act_stats, hwg_gen_output, hwg_gen_output_bwd, hwg_adain_bwd_reduce,
                                  hwg_adain_bwd_apply, hwg_bn_coeffs, hwg_hwr_stem, hwg_hwr_stem_bwd, hwg_hwr_stem_bwd_image, hwg_maxpool_nhwc,
                                  hwg_relu_maxpool_bwd, hwg_logsoftmax_bwd, hwg_bn_bwd_reduce, hwg_bn_bwd_apply,
                                  hwg_balance, hwg_spectral_norm, hwg_spectral_norm_bwd, hwg_channel_sum, hwg_conv_wgrad,
                                  hwg_conv_fprop, hwg_shift_expand, hwg_stem_conv, hwg_shift_collapse, hwg_gn_coeffs, hwg_scale_shift_act,
                                  hwg_avgpool_nhwc, hwg_add_stats, hwg_l1_halves, hwg_norm_bwd_reduce, hwg_gn_bwd_coeffs,
                                  hwg_norm_bwd_apply, hwg_act_bwd)}


@contextlib.contextmanager
def installed(monkeypatch):
    """Routes `_lib.call` to the interpreter above, `JobTable.run` to the CPU job-table interpreter and lifts the
    CUDA-tensor check, for the duration of one test.  Records the entry points that were called."""
    calls = []

    def call(name, *args):
        if name not in _TABLE:
            raise NotImplementedError(f"abi_emu: {name} is not interpreted")
        calls.append(name)
        rc = _TABLE[name](*args)
        assert rc == 0

    monkeypatch.setattr(_lib, "call", call)
    monkeypatch.setattr(_lib, "stream", lambda: 0)
    monkeypatch.setattr(_lib, "require_cuda", lambda *t: None)
    monkeypatch.setattr(weightmap.JobTable, "run",
                        lambda self, src_base=None, dst_base=None: ref_map.run_jobs_cpu(self, src_base, dst_base))

    class _Stream:                       # the generator's backward forks its wgrad launches onto a side stream
        device = torch.device("cpu")
        cuda_stream = 0

        def wait_stream(self, other):
            pass

    main = _Stream()
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: main)
    monkeypatch.setattr(torch.cuda, "Stream", lambda *a, **k: _Stream())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    yield calls
