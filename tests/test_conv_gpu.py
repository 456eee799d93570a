"""GPU parity of the tcgen05 implicit-GEMM convolution (hwg_conv_fprop) against torch's fp32
convolution evaluated on the same bf16-rounded operands (so only the accumulation order
differs): tolerance 2e-3 relative to the tensor's max, far inside the 2e-2 bf16 budget."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = 2e-3


def _ref_conv(x, w, b, pad, dil=(1, 1)):
    xr = x.to(torch.bfloat16).float()
    wr = w.to(torch.bfloat16).float()
    return F.conv2d(xr.double(), wr.double(), None if b is None else b.double(), padding=pad, dilation=dil).float()


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-20)).item()


CASES = [
    # N, Cin, Cout, H, W, k, pad, dil
    (2, 64, 128, 32, 128, (3, 3), (1, 1), (1, 1)),    # CK=64, BN=128
    (1, 128, 256, 16, 256, (3, 3), (1, 1), (1, 1)),   # BN=256
    (2, 256, 512, 8, 257, (3, 3), (1, 1), (1, 1)),    # two N tiles, ragged width
    (1, 512, 512, 8, 130, (3, 3), (0, 0), (1, 1)),    # no padding (HWR conv5)
    (3, 512, 512, 1, 254, (1, 3), (0, 4), (1, 4)),    # dilated 1-D conv (HWR cnn1d)
    (2, 512, 80, 1, 252, (1, 3), (0, 0), (1, 1)),     # Cout=80 head
    (2, 32, 32, 32, 64, (3, 3), (1, 1), (1, 1)),      # CK=32 (64-byte swizzle)
    (2, 16, 16, 64, 128, (3, 3), (1, 1), (1, 1)),     # CK=16 (32-byte swizzle), BN=16
    (1, 64, 78, 4, 40, (3, 3), (1, 1), (1, 1)),       # Cout not a multiple of 8/16
    (1, 64, 64, 5, 9, (3, 3), (1, 1), (1, 1)),        # tile larger than the image
]


@pytest.mark.parametrize("case", CASES)
def test_conv_matches_torch(case):
    from handwriting_line_generation_b200 import conv, _lib
    N, Cin, Cout, H, W, k, pad, dil = case
    g = torch.Generator().manual_seed(hash(case) % (1 << 31))
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, *k, generator=g) / (Cin * k[0] * k[1]) ** 0.5
    b = torch.randn(Cout, generator=g)
    ref = _ref_conv(x, w, b, pad, dil)
    Ho, Wo = ref.shape[2], ref.shape[3]
    taps = conv.conv_taps(k[0], k[1], pad[0], pad[1], dil[0], dil[1])
    n0 = _lib.launch_count()
    y = conv.conv_fprop(conv.to_nhwc_bf16(x.cuda()), conv.pack_conv2d_weight(w.cuda()), taps, Ho, Wo,
                        bias=b.cuda(), out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert _lib.launch_count() == n0 + 1
    got = y.permute(0, 3, 1, 2).cpu()
    assert _rel(got, ref) <= TOL, _rel(got, ref)


@pytest.mark.parametrize("tile_w", [8, 16, 32, 64, 128])
def test_tile_shapes(tile_w):
    from handwriting_line_generation_b200 import conv
    g = torch.Generator().manual_seed(tile_w)
    x = torch.randn(2, 64, 20, 150, generator=g)
    w = torch.randn(64, 64, 3, 3, generator=g) / 24
    ref = _ref_conv(x, w, None, (1, 1))
    y = conv.conv_fprop(conv.to_nhwc_bf16(x.cuda()), conv.pack_conv2d_weight(w.cuda()), conv.conv_taps(3, 3, 1, 1),
                        20, 150, out_dtype=torch.float32, tile_w=tile_w)
    assert _rel(y.permute(0, 3, 1, 2).cpu(), ref) <= TOL


def test_epilogue_bias_noise_lrelu_stats_bf16_out():
    from handwriting_line_generation_b200 import conv, _lib
    g = torch.Generator().manual_seed(3)
    N, C, H, W = 2, 64, 16, 96
    x = torch.randn(N, C, H, W, generator=g)
    w = torch.randn(C, C, 3, 3, generator=g) / 24
    b = torch.randn(C, generator=g)
    nz = torch.randn(N, C, H, W, generator=g)
    nw = torch.rand(C, generator=g)
    pre = _ref_conv(x, w, b, (1, 1)) + nw.view(1, C, 1, 1) * nz
    ref = F.leaky_relu(pre, 0.2)
    stats = torch.zeros(N, C, 2, device="cuda")
    y = conv.conv_fprop(conv.to_nhwc_bf16(x.cuda()), conv.pack_conv2d_weight(w.cuda()), conv.conv_taps(3, 3, 1, 1),
                        H, W, bias=b.cuda(), act=_lib.ACT_LRELU, slope=0.2,
                        noise=nz.permute(0, 2, 3, 1).contiguous().cuda(), noise_w=nw.cuda(), stats=stats)
    assert y.dtype == torch.bfloat16
    got = y.float().permute(0, 3, 1, 2).cpu()
    assert _rel(got, ref) <= 1e-2  # bf16 output rounding
    s = stats.cpu()
    assert _rel(s[:, :, 0], ref.sum((2, 3))) <= 2e-3 * (H * W) ** 0.5
    assert _rel(s[:, :, 1], (ref * ref).sum((2, 3))) <= 5e-3


def test_logsoftmax_head_writes_tbc():
    """cnn1d.12 + LogSoftmax + permute(2,0,1): fp32 [T,B,C] straight from the epilogue."""
    from handwriting_line_generation_b200 import conv, _lib
    g = torch.Generator().manual_seed(4)
    B, Cin, C, Wi = 3, 512, 80, 70
    x = torch.randn(B, Cin, 1, Wi, generator=g)
    w = torch.randn(C, Cin, 1, 3, generator=g) / 39
    b = torch.randn(C, generator=g)
    ref = F.log_softmax(_ref_conv(x, w, b, (0, 0)), dim=1)[:, :, 0, :].permute(2, 0, 1)  # [T,B,C]
    T = Wi - 2
    out = torch.empty(T, B, C, device="cuda")
    conv.conv_fprop(conv.to_nhwc_bf16(x.cuda()), conv.pack_conv2d_weight(w.cuda()), conv.conv_taps(1, 3, 0, 0), 1, T,
                    bias=b.cuda(), act=_lib.ACT_LOGSOFTMAX, out_view=(out, C, 0, B * C, 0))
    assert (out.cpu() - ref).abs().max().item() <= 5e-3
    assert torch.allclose(out.exp().sum(2).cpu(), torch.ones(T, B), atol=1e-4)


def test_phase_launch_strided_output():
    """Row-parity launches of nearest-upsample(2,1)+conv3x3 (pure_gen.py:176-186) with pre-summed taps."""
    from handwriting_line_generation_b200 import conv
    g = torch.Generator().manual_seed(5)
    N, Cin, Cout, H, W = 2, 64, 32, 6, 50
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / 24
    xr = x.to(torch.bfloat16).float()
    up = F.interpolate(xr, scale_factor=(2, 1), mode="nearest")
    ref = F.conv2d(up.double(), w.double(), padding=1).float()
    out = torch.empty(N, 2 * H, W, Cout, device="cuda", dtype=torch.float32)
    xc, wc = conv.to_nhwc_bf16(x.cuda()), w.cuda()
    for par in (0, 1):
        # output row 2i+par reads upsampled rows 2i+par-1..2i+par+1 = source rows i-1+ (par+kh)//2 ...
        rows = {}
        for kh in range(3):
            src = (par + kh - 1) // 2 if (par + kh - 1) >= 0 else -1
            rows.setdefault(src, 0)
            rows[src] = rows[src] + wc[:, :, kh, :]
        taps, mats = [], []
        for dh, wk in sorted(rows.items()):
            for kw in range(3):
                taps.append((dh, kw - 1))
                mats.append(wk[:, :, kw])
        conv.conv_fprop(xc, conv.pack_taps(mats), taps, H, W,
                        out_view=(out, 2 * H * W * Cout, 2 * W * Cout, Cout, par * W * Cout))
    got = out.permute(0, 3, 1, 2).cpu()
    assert _rel(got, ref) <= 1e-2  # summed taps are rounded to bf16 once more than the reference


@pytest.mark.parametrize("shape", [(2, 64, 64, 16, 96), (3, 16, 16, 40, 200), (2, 256, 256, 4, 130), (1, 32, 32, 33, 70)])
def test_specialised_noise_stats_kernel_with_zero_noise_weight(shape):
    """The generator's epilogue specialisation (LeakyReLU + in-kernel noise + statistics, bf16 out) with the noise
    weight set to zero must equal the plain convolution: checks the persistent tile loop, the TMEM double
    buffering and the shared-memory statistics accumulation across tiles and images."""
    from handwriting_line_generation_b200 import conv, _lib
    N, Cin, Cout, H, W = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5
    b = torch.randn(Cout, generator=g)
    ref = F.leaky_relu(_ref_conv(x, w, b, (1, 1)), 0.2)
    stats = torch.zeros(N, Cout, 2, device="cuda")
    y = conv.conv_fprop(conv.to_nhwc_bf16(x.cuda()), conv.pack_conv2d_weight(w.cuda()), conv.conv_taps(3, 3, 1, 1),
                        H, W, bias=b.cuda(), act=_lib.ACT_LRELU, slope=0.2, noise_w=torch.zeros(Cout, device="cuda"),
                        noise_seed=7, stats=stats)
    got = y.float().permute(0, 3, 1, 2).cpu()
    assert _rel(got, ref) <= 1e-2
    s = stats.cpu()
    assert _rel(s[:, :, 0], ref.sum((2, 3))) <= 3e-3 * (H * W) ** 0.5
    assert _rel(s[:, :, 1], (ref * ref).sum((2, 3))) <= 5e-3


def test_many_tiles_per_cta_and_two_n_tiles():
    """More tiles than CTAs (persistent loop wraps the smem ring and both TMEM buffers many times), Cout=512
    (two N tiles), statistics on."""
    from handwriting_line_generation_b200 import conv
    g = torch.Generator().manual_seed(11)
    N, Cin, Cout, H, W = 6, 64, 512, 24, 260
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / 24
    b = torch.randn(Cout, generator=g)
    ref = _ref_conv(x, w, b, (1, 1))
    stats = torch.zeros(N, Cout, 2, device="cuda")
    y = conv.conv_fprop(conv.to_nhwc_bf16(x.cuda()), conv.pack_conv2d_weight(w.cuda()), conv.conv_taps(3, 3, 1, 1),
                        H, W, bias=b.cuda(), stats=stats)
    got = y.float().permute(0, 3, 1, 2).cpu()
    assert _rel(got, ref) <= 1e-2
    s = stats.cpu()
    assert _rel(s[:, :, 1], (ref * ref).sum((2, 3))) <= 5e-3


def test_folded_launch_fused_upsample_matches_conv_transpose():
    """FusedUpsample (pure_gen.py:259-279) as ONE launch: 9 union taps, the four output parities as channel folds."""
    from handwriting_line_generation_b200 import conv
    from handwriting_line_generation_b200.pure_gen import FusedUpsample
    from tests.ref_pack import fused_up_folded, TAPS_UNION
    torch.manual_seed(5)
    Cin, C, N, H, W = 32, 16, 2, 12, 40
    mod = FusedUpsample(Cin, C, 3, padding=1)
    mod.bias.data.normal_()
    x = torch.randn(N, Cin, H, W)
    wpad = F.pad(mod.weight * mod.multiplier, [1, 1, 1, 1])
    w4 = (wpad[:, :, 1:, 1:] + wpad[:, :, :-1, 1:] + wpad[:, :, 1:, :-1] + wpad[:, :, :-1, :-1]) / 4
    ref = F.conv_transpose2d(x.to(torch.bfloat16).double(), w4.to(torch.bfloat16).double(), mod.bias.double(),
                             stride=2, padding=1).float()
    Ho, Wo = 2 * H, 2 * W
    raw = torch.zeros((N, Ho, Wo, C), device="cuda", dtype=torch.bfloat16)
    conv.conv_fprop(conv.to_nhwc_bf16(x.cuda()), fused_up_folded(mod.weight.detach().cuda(), mod.multiplier), TAPS_UNION, H, W,
                    bias=mod.bias.detach().float().repeat(4).cuda(), out_view=(raw, Ho * Wo * C, 2 * Wo * C, 2 * C, 0),
                    fold=(C, 2, Wo * C, C))
    got = raw.float().permute(0, 3, 1, 2).cpu()
    assert _rel(got, ref.detach()) <= 1e-2   # bf16 output rounding


def test_folded_launch_rows_with_noise_tensor_and_stats():
    """ConvTranspose2d (4,3) on H=1 (pure_gen.py:161-163): the four output rows as channel folds of one launch, with
    bias + explicit noise + LeakyReLU + per-(n,c) statistics accumulated over all four rows."""
    from handwriting_line_generation_b200 import conv, _lib
    from tests.ref_pack import initial_fwd
    torch.manual_seed(6)
    Cin, C, N, W = 64, 64, 2, 70
    wt = torch.randn(Cin, C, 4, 3) / (3 * Cin) ** 0.5
    b, nw = torch.randn(C), torch.rand(C)
    x = torch.randn(N, Cin, 1, W)
    nz = torch.randn(N, C, 4, W)
    pre = F.conv_transpose2d(x.to(torch.bfloat16).double(), wt.to(torch.bfloat16).double(), b.double(),
                             padding=(0, 1)).float() + nw.view(1, C, 1, 1) * nz
    ref = F.leaky_relu(pre, 0.2)
    taps = [(0, 1 - kx) for kx in range(3)]
    a = torch.zeros((N, 4, W, C), device="cuda", dtype=torch.bfloat16)
    st = torch.zeros((N, C, 2), device="cuda")
    nzh = nz.permute(0, 2, 3, 1).contiguous().cuda()
    conv.conv_fprop(conv.to_nhwc_bf16(x.cuda()), initial_fwd(wt.cuda(), Cin), taps, 1, W, bias=b.repeat(4).cuda(),
                    act=_lib.ACT_LRELU, slope=0.2, out_view=(a, 4 * W * C, W * C, C, 0), fold=(C, 1, W * C, 0),
                    noise_view=(nzh, 4 * W * C, W * C, C, 0), noise_w=nw.repeat(4).cuda(), stats=st)
    got = a.float().permute(0, 3, 1, 2).cpu()
    assert _rel(got, ref) <= 1e-2
    s1 = ref.sum((2, 3))
    s2 = (ref * ref).sum((2, 3))
    assert _rel(st[:, :, 0].cpu(), s1) <= 5e-3 and _rel(st[:, :, 1].cpu(), s2) <= 5e-3


# ---- staged-tile kernel for the small-channel layers (hwg_conv_small.cu) ------------------------------------
SMALL_CASES = [
    # N, Cin, Cout, H, W
    (2, 16, 16, 20, 150),
    (2, 32, 32, 9, 70),
    (1, 16, 32, 33, 64),
    (2, 32, 16, 5, 16),
    (1, 32, 64, 12, 40),
]


@pytest.mark.parametrize("case", SMALL_CASES)
def test_small_channel_kernel_matches_torch(case):
    """bf16-output 3x3 convolutions with Cin in {16,32}: routed to conv_small_kernel (TMA-staged halo tiles +
    mma.sync); bias + explicit noise + LeakyReLU + statistics epilogue; also checked against the tcgen05 kernel."""
    from handwriting_line_generation_b200 import conv, _lib
    N, Cin, Cout, H, W = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5
    b, nw = torch.randn(Cout, generator=g), torch.rand(Cout, generator=g)
    nz = torch.randn(N, Cout, H, W, generator=g)
    ref = F.leaky_relu(_ref_conv(x, w, b, (1, 1)) + nw.view(1, Cout, 1, 1) * nz, 0.2)
    xs, ws = conv.to_nhwc_bf16(x.cuda()), conv.pack_conv2d_weight(w.cuda())
    taps = conv.conv_taps(3, 3, 1, 1)
    outs = []
    for force in (False, True):
        st = torch.zeros((N, Cout, 2), device="cuda")
        y = conv.conv_fprop(xs, ws, taps, H, W, bias=b.cuda(), act=_lib.ACT_LRELU, slope=0.2,
                            noise=nz.permute(0, 2, 3, 1).contiguous().cuda(), noise_w=nw.cuda(), stats=st,
                            force_tcgen05=force)
        got = y.float().permute(0, 3, 1, 2).cpu()
        assert _rel(got, ref) <= 1e-2, (force, _rel(got, ref))
        assert _rel(st[:, :, 0].cpu(), ref.sum((2, 3))) <= 5e-3 and _rel(st[:, :, 1].cpu(), (ref * ref).sum((2, 3))) <= 5e-3
        outs.append(got)
    assert _rel(outs[0], outs[1]) <= 1e-2


@pytest.mark.parametrize("case", [(2, 32, 16, 12, 40), (1, 64, 32, 9, 70), (2, 64, 32, 4, 64)])
def test_small_channel_kernel_transposed_conv_one_launch(case):
    """FusedUpsample's conv_transpose2d(4x4, stride 2, pad 1) as one launch: 4 folds (output parities) x 4 taps."""
    from handwriting_line_generation_b200 import conv
    from handwriting_line_generation_b200.pure_gen import FusedUpsample
    from tests.ref_pack import fused_up_fwd
    N, Cin, C, H, W = case
    torch.manual_seed(sum(case))
    mod = FusedUpsample(Cin, C, 3, padding=1)
    mod.bias.data.normal_()
    x = torch.randn(N, Cin, H, W)
    wpad = F.pad(mod.weight * mod.multiplier, [1, 1, 1, 1])
    w4 = (wpad[:, :, 1:, 1:] + wpad[:, :, :-1, 1:] + wpad[:, :, 1:, :-1] + wpad[:, :, :-1, :-1]) / 4
    ref = F.conv_transpose2d(x.to(torch.bfloat16).double(), w4.to(torch.bfloat16).double(), mod.bias.double(),
                             stride=2, padding=1).float().detach()
    wf, taps = fused_up_fwd(mod.weight.detach().cuda(), mod.multiplier)
    Ho, Wo = 2 * H, 2 * W
    raw = torch.zeros((N, Ho, Wo, C), device="cuda", dtype=torch.bfloat16)
    conv.conv_fprop(conv.to_nhwc_bf16(x.cuda()), wf, taps, H, W, bias=mod.bias.detach().float().repeat(4).cuda(),
                    out_view=(raw, Ho * Wo * C, 2 * Wo * C, 2 * C, 0), fold=(C, 2, Wo * C, C), fold_taps=4)
    assert _rel(raw.float().permute(0, 3, 1, 2).cpu(), ref) <= 1e-2


@pytest.mark.parametrize("case", [(2, 16, 32, 16, 64), (1, 32, 64, 10, 36)])
def test_small_channel_kernel_strided_input(case):
    """Stride-2 input access (the input gradient of the stride-2 transposed convolutions) on the staged-tile kernel."""
    from handwriting_line_generation_b200 import conv
    N, Cin, Cout, H, W = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(N, Cin, H, W, generator=g).to(torch.bfloat16).double()
    w = (torch.randn(Cout, Cin, 4, 4, generator=g) / (Cin * 16) ** 0.5).to(torch.bfloat16).double()
    ref = F.conv2d(x, w, stride=2, padding=1).float()
    Ho, Wo = ref.shape[2], ref.shape[3]
    taps = [(i - 1, j - 1) for i in range(4) for j in range(4)]
    y = conv.conv_fprop(conv.to_nhwc_bf16(x.float().cuda()), conv.pack_conv2d_weight(w.float().cuda()), taps, Ho, Wo,
                        in_stride=(2, 2))
    assert _rel(y.float().permute(0, 3, 1, 2).cpu(), ref) <= 1e-2
