"""GPU parity of the generator's memory-bound backward passes against torch autograd (fp64) on identical
bf16-rounded inputs, and consistency of the noise regenerated in the backward with the noise the forward drew."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-20)).item()


def _bf(x):
    return x.to(torch.bfloat16).double()


def _nhwc(x, dt=torch.bfloat16):
    return x.permute(0, 2, 3, 1).contiguous().to(dt).cuda()


@pytest.mark.parametrize("N,C,H,W", [(2, 256, 4, 24), (3, 64, 16, 37), (2, 16, 64, 96), (2, 32, 32, 50)])
def test_adain_lrelu_noise_bwd(N, C, H, W):
    from handwriting_line_generation_b200 import ops
    g0 = torch.Generator().manual_seed(C + W)
    pre = _bf(torch.randn(N, C, H, W, generator=g0)).requires_grad_()           # conv(+blur) output
    z = torch.randn(N, C, H, W, generator=g0).double()
    nw = (torch.rand(C, generator=g0) + 0.2).double().requires_grad_()
    gamma = (torch.rand(N, C, generator=g0) + 0.5).double().requires_grad_()
    beta = torch.randn(N, C, generator=g0).double().requires_grad_()
    y = pre + nw.view(1, C, 1, 1) * z
    a = F.leaky_relu(y, 0.2)
    ab = a.detach().to(torch.bfloat16).double()                                  # what the forward stored
    a_st = a + (ab - a.detach())
    mean = ab.mean((2, 3), keepdim=True)
    var = ab.var((2, 3), keepdim=True, unbiased=False)
    xn = gamma[:, :, None, None] * (a_st - a_st.mean((2, 3), keepdim=True)) / torch.sqrt(
        a_st.var((2, 3), keepdim=True, unbiased=False) + 1e-5) + beta[:, :, None, None]
    g = _bf(torch.randn(xn.shape, generator=g0))
    gpre_ref, gnw_ref, ggam_ref, gbet_ref = torch.autograd.grad(xn, (pre, nw, gamma, beta), g)
    rstd = 1.0 / torch.sqrt(var + 1e-5)
    save = torch.stack([mean[:, :, 0, 0], rstd[:, :, 0, 0]], 2).float().contiguous().cuda()
    A = (gamma.detach() * rstd[:, :, 0, 0])
    coef = torch.stack([A, beta.detach() - mean[:, :, 0, 0] * A], 2).float().contiguous().cuda()
    gy, dgam, dbet, dsum, dnw = ops.adain_lrelu_bwd(_nhwc(g), _nhwc(ab), save, coef, 0.2, noise=_nhwc(z, torch.float32))
    assert _rel(gy.float().permute(0, 3, 1, 2).cpu(), gpre_ref) <= 1e-2
    assert _rel(dgam.cpu(), ggam_ref) <= 2e-3 and _rel(dbet.cpu(), gbet_ref) <= 2e-3
    assert _rel(dnw.cpu(), gnw_ref) <= 1e-2
    assert _rel(dsum.cpu(), gpre_ref.sum((0, 2, 3))) <= 1e-2


@pytest.mark.parametrize("N,H,W", [(2, 64, 96), (3, 64, 130)])
def test_gen_output_bwd(N, H, W):
    from handwriting_line_generation_b200 import ops
    C = 16
    g0 = torch.Generator().manual_seed(W)
    a = _bf(torch.randn(N, C, H, W, generator=g0)).requires_grad_()
    A = (torch.rand(N, C, generator=g0) + 0.5).double()
    Bc = torch.randn(N, C, generator=g0).double()
    w = torch.randn(C, generator=g0).double().requires_grad_()
    b0 = torch.tensor(0.3, dtype=torch.float64, requires_grad=True)
    x_last = A[:, :, None, None] * a + Bc[:, :, None, None]
    out = torch.tanh((x_last * w.view(1, C, 1, 1)).sum(1, keepdim=True) + b0)
    g = torch.randn(out.shape, generator=g0).double()
    gx_ref, gw_ref, gb_ref = torch.autograd.grad(out, (x_last, w, b0), g)
    coef = torch.stack([A, Bc], 2).float().contiguous().cuda()
    gx, dw, db0 = ops.gen_output_bwd(g.float().cuda(), out.detach().float().cuda(), _nhwc(a.detach()), coef,
                                     w.detach().float().cuda())
    assert _rel(gx.float().permute(0, 3, 1, 2).cpu(), gx_ref) <= 1e-2
    assert _rel(dw.cpu(), gw_ref) <= 2e-3 and abs(db0.item() - gb_ref.item()) <= 2e-3 * abs(gb_ref.item()) + 1e-4


def _identity_stats(N, C, H, W):
    save = torch.zeros(N, C, 2, device="cuda"); save[:, :, 1] = 1.0     # mean 0, rstd 1
    coef = torch.zeros(N, C, 2, device="cuda"); coef[:, :, 0] = 1.0     # A = 1
    return save, coef


@pytest.mark.parametrize("site", ["conv_epilogue", "blur", "initial_rows"])
def test_backward_regenerates_the_forward_noise(site):
    """dnoise_w computed with the regenerated in-kernel noise == computed from the noise the forward produced."""
    from handwriting_line_generation_b200 import conv, ops, _lib
    N, C, H, W = 2, 32, 8, 70
    seed, subseq = 12345, 48
    ones = torch.ones(C, device="cuda")
    if site == "conv_epilogue":     # zero conv + noise_w = 1, fp32 out: the output IS the noise
        x = torch.zeros(N, H, W, C, device="cuda", dtype=torch.bfloat16)
        wz = torch.zeros(9, C, C, device="cuda", dtype=torch.bfloat16)
        z = conv.conv_fprop(x, wz, conv.conv_taps(3, 3, 1, 1), H, W, noise_w=ones, noise_seed=seed,
                            noise_subseq=subseq, out_dtype=torch.float32)
        row = False
    elif site == "blur":            # blur of zeros + noise: bf16 copy of the noise
        x = torch.zeros(N, H, W, C, device="cuda", dtype=torch.bfloat16)
        z = ops.blur_noise_act_stats(x, None, ones, None, _lib.ACT_NONE, 0.0, seed, subseq).float()
        row = False
    else:                           # the initial block: one launch per output row, subsequence = base + row
        H = 4
        x = torch.zeros(N, 1, W, 64, device="cuda", dtype=torch.bfloat16)
        wz = torch.zeros(3, C, 64, device="cuda", dtype=torch.bfloat16)
        z = torch.empty(N, H, W, C, device="cuda", dtype=torch.float32)
        for r in range(4):
            conv.conv_fprop(x, wz, [(0, 1), (0, 0), (0, -1)], 1, W, noise_w=ones, noise_seed=seed,
                            noise_subseq=subseq + r, out_view=(z, H * W * C, W * C, C, r * W * C))
        row = True
    assert abs(z.mean().item()) < 0.05 and abs(z.var().item() - 1) < 0.1
    g0 = torch.Generator().manual_seed(1)
    g = torch.randn(N, H, W, C, generator=g0).to(torch.bfloat16).cuda()
    a = torch.randn(N, H, W, C, generator=g0).to(torch.bfloat16).cuda()
    save, coef = _identity_stats(N, C, H, W)
    _, _, _, _, dnw_tensor = ops.adain_lrelu_bwd(g, a, save, coef, 0.2, noise=z.contiguous())
    _, _, _, _, dnw_regen = ops.adain_lrelu_bwd(g, a, save, coef, 0.2, noise=None, seed=seed, subseq=subseq, row_subseq=row)
    assert _rel(dnw_regen.cpu(), dnw_tensor.cpu()) <= (2e-2 if site == "blur" else 1e-4)


@pytest.mark.parametrize("N,C,H,W", [(2, 16, 64, 200), (3, 128, 8, 37), (2, 64, 13, 70), (1, 32, 1, 9), (2, 16, 9, 1),
                                     (2, 256, 5, 40)])
@pytest.mark.parametrize("with_noise", [False, True])
def test_blur_noise_act_stats_matches_torch(N, C, H, W, with_noise):
    """`Blur` ([1,2,1] x [1,2,1] / 16 depthwise, zero padding; reference model/pure_gen.py Blur) + NoiseInjection with a
    given noise tensor + LeakyReLU + per-(n,c) sums, on tiles that are ragged in every direction (rows not a multiple
    of 8, row segments not a multiple of 256 items, one-row and one-column lines)."""
    from handwriting_line_generation_b200 import _lib, ops
    g = torch.Generator().manual_seed(N * 1000 + C + H + W)
    x = _bf(torch.randn(N, C, H, W, generator=g))
    k = torch.tensor([1.0, 2.0, 1.0], dtype=torch.float64)
    k = (k[:, None] * k[None, :] / 16.0).expand(C, 1, 3, 3).contiguous()
    ref = F.conv2d(x, k, padding=1, groups=C)
    noise = nw = None
    if with_noise:
        noise = torch.randn(N, C, H, W, generator=g)
        nw = torch.randn(C, generator=g)
        ref = ref + nw.double()[None, :, None, None] * noise.double()
    ref = F.leaky_relu(ref, 0.2)
    stats = torch.zeros(N, C, 2, device="cuda")
    y = ops.blur_noise_act_stats(_nhwc(x), None if noise is None else _nhwc(noise, torch.float32),
                                 None if nw is None else nw.cuda(), stats, _lib.ACT_LRELU, 0.2)
    y = y.float().permute(0, 3, 1, 2).cpu().double()
    assert _rel(y, ref) <= 6e-3                                          # one bf16 rounding of the output
    assert _rel(stats[..., 0].cpu(), ref.sum((2, 3))) <= 5e-3
    assert _rel(stats[..., 1].cpu(), (ref * ref).sum((2, 3))) <= 5e-3
