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
