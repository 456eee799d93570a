"""GPU parity of the convolution backward: dgrad (= hwg_conv_fprop on the output gradient with negated taps and
transposed weights) and wgrad (hwg_conv_wgrad, tcgen05 with MN-major operands) against torch autograd evaluated in
fp64 on the same bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = 3e-3


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-20)).item()


CASES = [
    # N, Cin, Cout, H, W, k, pad, dil
    (2, 64, 128, 16, 96, (3, 3), (1, 1), (1, 1)),
    (2, 256, 256, 8, 130, (3, 3), (1, 1), (1, 1)),
    (1, 512, 512, 8, 70, (3, 3), (0, 0), (1, 1)),
    (3, 512, 512, 1, 254, (1, 3), (0, 4), (1, 4)),
    (2, 512, 80, 1, 100, (1, 3), (0, 0), (1, 1)),
    (2, 128, 256, 5, 9, (3, 3), (1, 1), (1, 1)),
    (2, 512, 80, 1, 28, (1, 3), (0, 0), (1, 1)),      # head at W=128: 26 frames, 2-row pixel chunks on H=1
    (2, 512, 512, 1, 26, (1, 3), (0, 8), (1, 8)),     # dilation wider than a chunk
    (2, 256, 512, 3, 32, (3, 3), (0, 0), (1, 1)),
    # halo-mode tiles of conv_wgrad_kernel: 8x16 / 4x32 pixel chunks, several tap groups, ragged edges
    (2, 64, 64, 16, 40, (3, 3), (1, 1), (1, 1)),
    (3, 128, 64, 4, 50, (3, 3), (1, 1), (1, 1)),
    (2, 64, 192, 11, 23, (3, 3), (1, 1), (1, 1)),
    (1, 128, 128, 2, 70, (3, 3), (1, 1), (1, 1)),
]


@pytest.mark.parametrize("case", CASES)
def test_dgrad_and_wgrad_match_autograd(case):
    from handwriting_line_generation_b200 import conv
    N, Cin, Cout, H, W, k, pad, dil = case
    g = torch.Generator().manual_seed(hash(case) % (1 << 31))
    x = torch.randn(N, Cin, H, W, generator=g).to(torch.bfloat16).double().requires_grad_()
    w = (torch.randn(Cout, Cin, *k, generator=g) / (Cin * k[0] * k[1]) ** 0.5).to(torch.bfloat16).double().requires_grad_()
    y = F.conv2d(x, w, padding=pad, dilation=dil)
    gy = torch.randn(y.shape, generator=g).to(torch.bfloat16).double()
    gx_ref, gw_ref = torch.autograd.grad(y, (x, w), gy)
    Ho, Wo = y.shape[2], y.shape[3]
    taps = conv.conv_taps(k[0], k[1], pad[0], pad[1], dil[0], dil[1])
    w_taps = torch.stack([w.detach().float()[:, :, i, j] for i in range(k[0]) for j in range(k[1])], 0).cuda()
    gyc = conv.to_nhwc_bf16(gy.float().cuda())
    xc = conv.to_nhwc_bf16(x.detach().float().cuda())
    # dgrad
    wd, tapsd = conv.dgrad_pack(w_taps, taps)
    gx = conv.conv_fprop(gyc, wd, tapsd, H, W, out_dtype=torch.float32)
    assert _rel(gx.permute(0, 3, 1, 2).cpu().double(), gx_ref) <= TOL
    # wgrad
    dw = conv.conv_wgrad(xc, gyc, taps, Cin, Cout)                     # [ntaps, Cout, Cin]
    got = dw.view(k[0], k[1], Cout, Cin).permute(2, 3, 0, 1).cpu().double()
    assert _rel(got, gw_ref) <= TOL, _rel(got, gw_ref)


def test_wgrad_kernel_selection():
    """Which kernel serves which shape (hwg_last_wgrad_kernel): staged tiles for 16/32 channels, the tcgen05 kernel with a
    halo box of x for Cin % 64 == 0, one x box per tap otherwise."""
    import os
    from handwriting_line_generation_b200 import conv, _lib
    want_halo = 1 if os.environ.get("HWG_WGRAD_HALO") == "1" else 0
    for Cin, Cout, mode in [(16, 16, 0), (32, 32, 0), (64, 64, 1 + want_halo), (16, 64, 1), (128, 32, 1 + want_halo)]:
        x = torch.randn(1, 12, 40, Cin, device="cuda").to(torch.bfloat16)
        gy = torch.randn(1, 12, 40, Cout, device="cuda").to(torch.bfloat16)
        conv.conv_wgrad(x, gy, conv.conv_taps(3, 3, 1, 1), Cin, Cout)
        assert _lib.load().hwg_last_wgrad_kernel() == mode, (Cin, Cout)


@pytest.mark.parametrize("case", [(2, 64, 32, 12, 40, (2, 2)), (1, 32, 16, 16, 64, (2, 2)), (2, 128, 64, 8, 50, (2, 1))])
def test_strided_input_conv(case):
    """in_stride: y[ho,wo] = sum_t x[s*ho+dh, s*wo+dw] w_t — a stride-s convolution (TMA element strides), which is
    the input gradient of the generator's stride-2 transposed / up-sampling convolutions."""
    from handwriting_line_generation_b200 import conv
    N, Cin, Cout, H, W, st = case
    g = torch.Generator().manual_seed(sum(case[:5]))
    x = torch.randn(N, Cin, H, W, generator=g).to(torch.bfloat16).double()
    w = (torch.randn(Cout, Cin, 4, 4, generator=g) / (Cin * 16) ** 0.5).to(torch.bfloat16).double()
    ref = F.conv2d(x, w, stride=st, padding=1)
    Ho, Wo = ref.shape[2], ref.shape[3]
    taps = [(i - 1, j - 1) for i in range(4) for j in range(4)]
    y = conv.conv_fprop(conv.to_nhwc_bf16(x.float().cuda()), conv.pack_conv2d_weight(w.float().cuda()), taps, Ho, Wo,
                        out_dtype=torch.float32, in_stride=st)
    assert _rel(y.permute(0, 3, 1, 2).cpu().double(), ref) <= TOL


@pytest.mark.parametrize("Cin,Cout", [(16, 16), (32, 16), (32, 32), (64, 32), (16, 64), (256, 128)])
def test_wgrad_small_channels(Cin, Cout):
    """16/32-channel operands: 32/64-byte swizzled MN-major boxes, output-channel blocks past Cout zero-filled."""
    from handwriting_line_generation_b200 import conv
    N, H, W = 2, 20, 72
    g = torch.Generator().manual_seed(Cin * 100 + Cout)
    x = torch.randn(N, Cin, H, W, generator=g).to(torch.bfloat16).double()
    w = torch.zeros(Cout, Cin, 3, 3, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x, w, padding=1)
    gy = torch.randn(y.shape, generator=g).to(torch.bfloat16).double()
    (gw_ref,) = torch.autograd.grad(y, w, gy)
    dw = conv.conv_wgrad(conv.to_nhwc_bf16(x.float().cuda()), conv.to_nhwc_bf16(gy.float().cuda()),
                         conv.conv_taps(3, 3, 1, 1), Cin, Cout)
    got = dw.view(3, 3, Cout, Cin).permute(2, 3, 0, 1).cpu().double()
    assert _rel(got, gw_ref) <= TOL, _rel(got, gw_ref)


@pytest.mark.parametrize("Cin,Cout", [(64, 32), (32, 16)])
def test_wgrad_of_transposed_conv_phases(Cin, Cout):
    """Weight gradient of conv_transpose2d(stride 2, pad 1, 4x4) (FusedUpsample, pure_gen.py:277) as four
    output-parity launches reading gy with stride 2."""
    from handwriting_line_generation_b200 import conv
    N, H, W = 2, 10, 36
    g = torch.Generator().manual_seed(Cin + Cout)
    x = torch.randn(N, Cin, H, W, generator=g).to(torch.bfloat16).double()
    w4 = torch.zeros(Cin, Cout, 4, 4, dtype=torch.float64, requires_grad=True)
    y = F.conv_transpose2d(x, w4, stride=2, padding=1)
    gy = torch.randn(y.shape, generator=g).to(torch.bfloat16).double()
    (gw_ref,) = torch.autograd.grad(y, w4, gy)          # [Cin, Cout, 4, 4]
    xc, gyc = conv.to_nhwc_bf16(x.float().cuda()), conv.to_nhwc_bf16(gy.float().cuda())
    sel = {0: [(0, 1), (-1, 3)], 1: [(1, 0), (0, 2)]}   # parity -> [(input offset, kernel index)]
    got = torch.zeros_like(gw_ref)
    for py in (0, 1):
        for px in (0, 1):
            taps, idx = [], []
            for dh, ky in sel[py]:
                for dw, kx in sel[px]:
                    taps.append((dh, dw))
                    idx.append((ky, kx))
            dw_ = conv.conv_wgrad(xc, gyc, taps, Cin, Cout, grid=(H, W), gy_stride=(2, 2), gy_offset=(py, px)).cpu().double()
            for t, (ky, kx) in enumerate(idx):
                got[:, :, ky, kx] = dw_[t].t()
    assert _rel(got, gw_ref) <= TOL, _rel(got, gw_ref)


@pytest.mark.parametrize("Cin,Cout,H,W", [(32, 16, 10, 36), (32, 32, 9, 130), (16, 16, 33, 64)])
def test_wgrad_all_phases_in_one_launch(Cin, Cout, H, W):
    """Same gradient with per-tap gy phases: the four parities (16 taps) in ONE launch of the staged-tile kernel."""
    from handwriting_line_generation_b200 import conv
    N = 2
    g = torch.Generator().manual_seed(Cin + Cout + W)
    x = torch.randn(N, Cin, H, W, generator=g).to(torch.bfloat16).double()
    w4 = torch.zeros(Cin, Cout, 4, 4, dtype=torch.float64, requires_grad=True)
    y = F.conv_transpose2d(x, w4, stride=2, padding=1)
    gy = torch.randn(y.shape, generator=g).to(torch.bfloat16).double()
    (gw_ref,) = torch.autograd.grad(y, w4, gy)
    xc, gyc = conv.to_nhwc_bf16(x.float().cuda()), conv.to_nhwc_bf16(gy.float().cuda())
    sel = {0: [(0, 1), (-1, 3)], 1: [(1, 0), (0, 2)]}
    taps, phases, idx = [], [], []
    for py in (0, 1):
        for px in (0, 1):
            for dh, ky in sel[py]:
                for dw, kx in sel[px]:
                    taps.append((dh, dw)); phases.append((py, px)); idx.append((ky, kx))
    dw_ = conv.conv_wgrad(xc, gyc, taps, Cin, Cout, grid=(H, W), gy_stride=(2, 2), tap_phase=phases).cpu().double()
    got = torch.zeros_like(gw_ref)
    for t, (ky, kx) in enumerate(idx):
        got[:, :, ky, kx] = dw_[t].t()
    assert _rel(got, gw_ref) <= TOL, _rel(got, gw_ref)
