"""Host side of the tcgen05 implicit-GEMM convolution (hwg_conv_fprop).

Activations cross the kernels as NHWC bf16 tensors ([N,H,W,C], C contiguous); weights as
tap-major packed bf16 [ntaps, Cout, Cin] (`pack_taps`).  A "tap" is an input-pixel offset
(dh, dw); every convolution flavour of the hot path is a tap list (see include/hwg_b200.h).
"""
import ctypes

import torch

from . import _lib

# bench.py sets this to a list to time every launch with CUDA events on the launching stream:
# entries are (start_event, end_event, issued_flops, kernel, activation bytes, (N, Ho, Wo, Cin, Cout, ntaps))
PROFILE = None


def conv_taps(kh, kw, pad_h, pad_w, dil_h=1, dil_w=1):
    """Tap list of a stride-1 cross-correlation (nn.Conv2d): tap (i,j) reads x[ho+i*dil-pad, ...]."""
    return [(i * dil_h - pad_h, j * dil_w - pad_w) for i in range(kh) for j in range(kw)]


def pack_taps(mats, cin_pad=None):
    """mats: list of [Cout, Cin] fp32 matrices, one per tap -> bf16 [ntaps, Cout, Cin_pad]."""
    w = torch.stack(list(mats), 0)
    cin = w.size(2)
    cin_pad = cin_pad or ((cin + 15) // 16) * 16
    if cin_pad != cin:
        w = torch.nn.functional.pad(w, (0, cin_pad - cin))
    return w.to(torch.bfloat16).contiguous()


def pack_conv2d_weight(weight, cin_pad=None):
    """nn.Conv2d weight [Cout, Cin, kh, kw] -> packed taps in conv_taps() order."""
    co, ci, kh, kw = weight.shape
    return pack_taps([weight[:, :, i, j] for i in range(kh) for j in range(kw)], cin_pad)


def conv_fprop(x, w_packed, taps, Ho, Wo, *, bias=None, act=_lib.ACT_NONE, slope=0.0, out=None,
               out_dtype=torch.bfloat16, out_view=None, noise=None, noise_w=None, noise_view=None,
               noise_seed=None, noise_subseq=0, noise_seed_dev=None, stats=None, cin=None, tile_w=0,
               in_stride=(1, 1), fold=None, fold_taps=0, force_tcgen05=False):
    """y[n,ho,wo,co] = epi(sum_t sum_ci x[n,ho+dh_t,wo+dw_t,ci] * w[t,co,ci]).

    x         [N,H,W,Cp] bf16 NHWC contiguous; `cin` (default w_packed.size(2)) channels are read
    out_view  optional (tensor, stride_n, stride_h, stride_w, element_offset): write into an existing
              tensor with custom pixel strides (used by the parity/phase launches of the
              up-sampling convolutions); otherwise a fresh [N,Ho,Wo,Cout] tensor is returned
    noise     optional fp32 [N,Ho,Wo,Cout] NHWC tensor added as noise_w[c]*noise before the activation
              (or noise_view=(tensor, sn, sh, sw, element_offset) for a strided one); with noise_w and
              noise_seed but no tensor, N(0,1) is drawn inside the kernel (counter-based hash + Box-Muller)
    stats     optional zeroed fp32 [N,Cout,2]; receives per-(n,c) sum and sum of squares of the output
    fold      optional (fold_c, fold_w, stride_h, stride_w): Cout = F*fold_c, channel f*fold_c+ch is channel ch of
              the output pixel displaced by (f//fold_w)*stride_h + (f%fold_w)*stride_w elements; bias/noise_w are
              [Cout], stats [N,fold_c,2] (see include/hwg_b200.h)
    """
    _lib.require_cuda(x, w_packed)
    assert x.dtype == torch.bfloat16 and x.dim() == 4 and x.is_contiguous()
    assert w_packed.dtype == torch.bfloat16 and w_packed.is_contiguous()
    N, H, W, Cp = x.shape
    ntaps, Cout, Cin = w_packed.shape
    cin = cin or Cin
    assert cin == Cin and Cin <= Cp and ntaps == len(taps)
    if fold is not None and not fold_taps:
        assert Cout % fold[0] == 0
    elif fold is not None:
        Cout = Cout * (ntaps // fold_taps)      # w is [ntaps][fold_c][Cin]: the launch's channel count is F * fold_c
    d = _lib.ConvDesc()
    d.N, d.H, d.W, d.Cin, d.x_pitch, d.Cout, d.Ho, d.Wo, d.ntaps = N, H, W, Cin, Cp, Cout, Ho, Wo, ntaps
    for i, (dh, dw) in enumerate(taps):
        d.tap_dh[i], d.tap_dw[i] = dh, dw
    if out_view is None:
        if out is None:
            out = torch.empty((N, Ho, Wo, Cout), device=x.device, dtype=out_dtype)
        y, sn, sh, sw, off = out, Ho * Wo * Cout, Wo * Cout, Cout, 0
    else:
        y, sn, sh, sw, off = out_view
        out = y
    d.y_stride_n, d.y_stride_h, d.y_stride_w = sn, sh, sw
    d.y_dtype = _lib.DT_F32 if y.dtype == torch.float32 else _lib.DT_BF16
    assert y.dtype in (torch.float32, torch.bfloat16)
    d.act, d.slope, d.tile_w = act, slope, tile_w
    d.in_stride_h, d.in_stride_w = in_stride
    nz_ptr = None
    if noise is not None:
        assert noise.dtype == torch.float32 and noise.is_contiguous() and tuple(noise.shape) == (N, Ho, Wo, Cout)
        d.nz_stride_n, d.nz_stride_h, d.nz_stride_w = Ho * Wo * Cout, Wo * Cout, Cout
        nz_ptr = noise.data_ptr()
    elif noise_view is not None:
        nz, d.nz_stride_n, d.nz_stride_h, d.nz_stride_w, nz_off = noise_view
        assert nz.dtype == torch.float32
        nz_ptr = nz.data_ptr() + 4 * nz_off
    if noise_w is not None:
        assert noise_w.dtype == torch.float32 and noise_w.numel() == Cout
        if nz_ptr is None:
            assert noise_seed is not None, "noise_w without a noise tensor needs noise_seed"
            d.noise_seed, d.noise_subseq = noise_seed, noise_subseq
            d.noise_seed_dev = 0 if noise_seed_dev is None else noise_seed_dev.data_ptr()
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == Cout
    if fold is not None:
        d.fold_c, d.fold_w, d.fold_stride_h, d.fold_stride_w = fold
        d.fold_taps = fold_taps
    d.force_tcgen05 = int(force_tcgen05)
    if stats is not None:
        assert stats.dtype == torch.float32 and stats.numel() == N * (fold[0] if fold else Cout) * 2
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _lib.call("hwg_conv_fprop", ctypes.addressof(d), x.data_ptr(), w_packed.data_ptr(), _lib.ptr(bias),
              nz_ptr, _lib.ptr(noise_w), _lib.ptr(stats), y.data_ptr() + off * y.element_size(),
              _lib.stream())
    if PROFILE is not None:
        e1.record()
        kind = "conv_small_kernel" if _lib.load().hwg_last_conv_kernel() == 2 else "conv_fprop_kernel"
        by = N * H * W * Cin * 2 + N * Ho * Wo * Cout * y.element_size()    # activations read once + written once
        flops = 2.0 * N * Ho * Wo * Cin * (w_packed.size(0) * w_packed.size(1))   # MACs actually issued
        PROFILE.append((e0, e1, flops, kind, by, (N, Ho, Wo, Cin, Cout, ntaps)))
    return out


def to_nhwc_bf16(x, c_pad=None):
    """[N,C,H,W] float -> [N,H,W,Cp] bf16 (plumbing used at the module boundary and in tests)."""
    n, c, h, w = x.shape
    c_pad = c_pad or ((c + 15) // 16) * 16
    y = x.permute(0, 2, 3, 1).to(torch.bfloat16)
    if c_pad != c:
        y = torch.nn.functional.pad(y, (0, c_pad - c))
    return y.contiguous()


def dgrad_pack(w_taps_f32, taps):
    """Forward tap matrices [ntaps][Cout][Cin] (fp32) -> (packed bf16 [ntaps][Cin][Cout_pad], negated taps):
    the input gradient is hwg_conv_fprop of the output gradient with these."""
    mats = [w_taps_f32[t].t() for t in range(w_taps_f32.size(0))]
    return pack_taps(mats), [(-dh, -dw) for dh, dw in taps]


def conv_wgrad(x, gy, taps, cin, cout, out=None, grid=None, gy_stride=(1, 1), gy_offset=(0, 0), tap_phase=None):
    """dw[t][co][ci] = sum_pixels gy[n,ho,wo,co] * x[n,ho+dh_t,wo+dw_t,ci]  (fp32 [ntaps,cout,cin]).
    x [N,H,W,Cp>=cin] bf16 NHWC, gy [N,Ho,Wo,Gp>=cout] bf16 NHWC."""
    _lib.require_cuda(x, gy)
    assert x.dtype == torch.bfloat16 and gy.dtype == torch.bfloat16 and x.is_contiguous() and gy.is_contiguous()
    N, H, W, Cp = x.shape
    N2, Ho, Wo, Gp = gy.shape
    assert N == N2
    d = _lib.WgradDesc()
    d.N, d.H, d.W, d.Cin, d.x_pitch, d.Ho, d.Wo, d.Cout, d.gy_pitch, d.ntaps = N, H, W, cin, Cp, Ho, Wo, cout, Gp, len(taps)
    for i, (dh, dw) in enumerate(taps):
        d.tap_dh[i], d.tap_dw[i] = dh, dw
    if grid is not None:   # phase of an up-sampling conv: iterate the input grid, gy read with a stride
        d.Hi, d.Wi = grid
        d.gy_stride_h, d.gy_stride_w = gy_stride
        d.gy_off_h, d.gy_off_w = gy_offset
    if tap_phase is not None:   # per-tap gy phase: all parities of an up-sampling conv in one launch (small C only)
        for i, (ph, pw) in enumerate(tap_phase):
            d.tap_gy_h[i], d.tap_gy_w[i] = ph, pw
    if out is None:
        out = torch.zeros((len(taps), cout, cin), device=x.device, dtype=torch.float32)
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _lib.call("hwg_conv_wgrad", ctypes.addressof(d), x.data_ptr(), gy.data_ptr(), out.data_ptr(), _lib.stream())
    if PROFILE is not None:
        e1.record()
        PROFILE.append((e0, e1, 2.0 * N * Ho * Wo * cout * cin * len(taps), "conv_wgrad_kernel",
                        (x.numel() + gy.numel()) * 2, (N, Ho, Wo, cin, cout, len(taps))))
    return out
