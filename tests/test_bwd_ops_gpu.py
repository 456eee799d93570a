"""GPU parity of the memory-bound backward passes, each against torch autograd (fp64) on identical bf16-rounded
inputs: log-softmax, BatchNorm+ReLU, ReLU+MaxPool (all three pooling geometries of cnn_only_hwr.py), stem."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-20)).item()


def _bf(x):
    return x.to(torch.bfloat16).double()


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()


@pytest.mark.parametrize("T,B,C", [(26, 2, 80), (59, 3, 78), (250, 8, 80)])
def test_logsoftmax_bwd(T, B, C):
    from handwriting_line_generation_b200 import ops
    g0 = torch.Generator().manual_seed(T)
    z = torch.randn(T, B, C, generator=g0, dtype=torch.float64, requires_grad=True)
    lp = F.log_softmax(z, 2)
    g = torch.randn(T, B, C, generator=g0, dtype=torch.float64)
    (gz_ref,) = torch.autograd.grad(lp, z, g)
    Cp = (C + 15) // 16 * 16
    gz, db = ops.logsoftmax_bwd(g.float().cuda(), lp.detach().float().cuda(), Cp)
    got = gz.float().cpu()[:, 0].permute(1, 0, 2)  # [B,1,T,Cp] -> [T,B,Cp]
    assert _rel(got[:, :, :C], gz_ref) <= 6e-3     # bf16 output
    if Cp > C:
        assert got[:, :, C:].abs().max() == 0
    assert _rel(db.cpu(), gz_ref.sum((0, 1))) <= 1e-4


@pytest.mark.parametrize("N,C,H,W", [(2, 512, 1, 26), (3, 256, 16, 65), (2, 512, 8, 33)])
def test_bn_relu_bwd(N, C, H, W):
    from handwriting_line_generation_b200 import ops
    g0 = torch.Generator().manual_seed(C + W)
    z = _bf(torch.randn(N, C, H, W, generator=g0) * 2 + 0.5).requires_grad_()
    wt = (torch.rand(C, generator=g0) + 0.5).double().requires_grad_()
    bs = (torch.randn(C, generator=g0) * 0.3).double().requires_grad_()
    y = F.relu(F.batch_norm(z, None, None, wt, bs, True, 0.1, 1e-5))
    g = _bf(torch.randn(y.shape, generator=g0))
    gz_ref, gw_ref, gb_ref = torch.autograd.grad(y, (z, wt, bs), g)
    zc = _nhwc(z.detach())
    # forward statistics exactly as the conv epilogue + bn_coeffs produce them
    zf = zc.float()
    stats = torch.stack([zf.sum((1, 2)), (zf * zf).sum((1, 2))], 2).contiguous()
    coef, save = ops.bn_coeffs(stats, N, C, H * W, wt.detach().float().cuda(), bs.detach().float().cuda(), None, None,
                               0.1, 1e-5, True)
    gz, dgam, dbet, dcb = ops.bn_bwd(_nhwc(g), zc, coef, save, wt.detach().float().cuda())
    assert _rel(gz.float().permute(0, 3, 1, 2).cpu(), gz_ref) <= 1e-2
    assert _rel(dgam.cpu(), gw_ref) <= 2e-3
    assert _rel(dbet.cpu(), gb_ref) <= 2e-3
    assert dcb.abs().max().item() <= 1e-2 * gz_ref.abs().sum((0, 2, 3)).max().item()


@pytest.mark.parametrize("geom", [((2, 2), (2, 2), (0, 0)), ((2, 2), (2, 1), (0, 1))])
@pytest.mark.parametrize("N,C,H,W", [(2, 128, 32, 64), (3, 256, 16, 65), (2, 512, 6, 31), (1, 64, 7, 9)])
def test_relu_maxpool_bwd(geom, N, C, H, W):
    from handwriting_line_generation_b200 import ops
    k, s, p = geom
    g0 = torch.Generator().manual_seed(C + W + s[1])
    pre = _bf(torch.randn(N, C, H, W, generator=g0)).requires_grad_()
    c = F.relu(pre)
    a = F.max_pool2d(c, k, s, p)
    g = _bf(torch.randn(a.shape, generator=g0))
    (gpre_ref,) = torch.autograd.grad(a, pre, g)
    gc, db = ops.relu_maxpool_bwd(_nhwc(g), _nhwc(c.detach()), k, s, p)
    assert _rel(gc.float().permute(0, 3, 1, 2).cpu(), gpre_ref) <= 1e-2
    # the bias gradient sums the bf16 gradient tensor the pass writes (the values the convolution's weight gradient reads
    # too): one 2^-9 rounding per element, which only averages out over many pixels — the 7x9 case keeps 4e-3
    assert _rel(db.cpu(), gpre_ref.sum((0, 2, 3))) <= (2e-3 if H * W >= 512 else 4e-3)


def test_relu_maxpool_bwd_ties_first_max_wins():
    from handwriting_line_generation_b200 import ops
    # all-equal positive window: gradient goes to the first element only (torch's max_pool2d rule)
    c = torch.ones(1, 64, 4, 4, dtype=torch.float64, requires_grad=True)
    a = F.max_pool2d(c, 2, 2)
    g = torch.arange(1, 5, dtype=torch.float64).view(1, 1, 2, 2).expand(1, 64, 2, 2).contiguous()
    (ref,) = torch.autograd.grad(a, c, g)
    gc, _ = ops.relu_maxpool_bwd(_nhwc(g), _nhwc(c.detach()), (2, 2), (2, 2), (0, 0))
    assert torch.equal(gc.float().permute(0, 3, 1, 2).cpu().double(), ref)


def test_relu_maxpool_bwd_ties_overlapping_windows():
    from handwriting_line_generation_b200 import ops
    # MaxPool2d((2,2),(2,1),(0,1)) on a constant positive map: every window's first in-image element takes the gradient
    c = torch.ones(1, 64, 4, 5, dtype=torch.float64, requires_grad=True)
    a = F.max_pool2d(c, (2, 2), (2, 1), (0, 1))
    g = torch.arange(1, a.numel() // 64 + 1, dtype=torch.float64).view(1, 1, a.size(2), a.size(3)).expand(1, 64, -1, -1).contiguous()
    (ref,) = torch.autograd.grad(a, c, g)
    gc, _ = ops.relu_maxpool_bwd(_nhwc(g), _nhwc(c.detach()), (2, 2), (2, 1), (0, 1))
    assert torch.equal(gc.float().permute(0, 3, 1, 2).cpu().double(), ref)


@pytest.mark.parametrize("geom", [((2, 2), (2, 2), (0, 0)), ((2, 2), (2, 1), (0, 1))])
@pytest.mark.parametrize("N,C,H,W", [(2, 64, 8, 21), (1, 256, 5, 12)])
def test_relu_maxpool_bwd_many_ties_bit_exact(geom, N, C, H, W):
    """Small-integer activations: most windows hold ties, and zero maxima meet the ReLU mask.  Integer-valued gradients keep
    every sum exact, so the CUDA result must EQUAL autograd's (first maximum in scan order takes the gradient)."""
    from handwriting_line_generation_b200 import ops
    k, s, p = geom
    g0 = torch.Generator().manual_seed(7 * C + W + s[1])
    pre = torch.randint(-1, 3, (N, C, H, W), generator=g0).double().requires_grad_()
    c = F.relu(pre)
    a = F.max_pool2d(c, k, s, p)
    g = torch.randint(-8, 9, a.shape, generator=g0).double()
    (ref,) = torch.autograd.grad(a, pre, g)
    gc, db = ops.relu_maxpool_bwd(_nhwc(g), _nhwc(c.detach()), k, s, p)
    assert torch.equal(gc.float().permute(0, 3, 1, 2).cpu().double(), ref)
    assert torch.equal(db.cpu().double(), ref.sum((0, 2, 3)))


@pytest.mark.parametrize("geom", [((2, 2), (2, 2), (0, 0)), ((2, 2), (2, 1), (0, 1))])
@pytest.mark.parametrize("N,C,H,W", [(2, 128, 32, 64), (3, 256, 16, 65), (1, 64, 7, 9)])
def test_maxpool_nhwc_forward_bit_exact(geom, N, C, H, W):
    from handwriting_line_generation_b200 import ops
    k, s, p = geom
    x = _bf(torch.randn(N, C, H, W, generator=torch.Generator().manual_seed(C + H)))
    ref = F.max_pool2d(x, k, s, p)
    y = ops.maxpool_nhwc(_nhwc(x), k, s, p)
    assert torch.equal(y.float().permute(0, 3, 1, 2).cpu().double(), ref)


@pytest.mark.parametrize("N,H,W", [(2, 64, 128), (3, 64, 260)])
def test_stem_fwd_bwd(N, H, W):
    from handwriting_line_generation_b200 import ops
    g0 = torch.Generator().manual_seed(W)
    img = (torch.rand(N, 1, H, W, generator=g0) * 2 - 1).double()
    w = (torch.randn(64, 1, 3, 3, generator=g0) / 3).float().double().requires_grad_()
    b = (torch.randn(64, generator=g0) * 0.1).float().double().requires_grad_()
    a = F.max_pool2d(F.relu(F.conv2d(img, w, b, padding=1)), 2, 2)
    g = _bf(torch.randn(a.shape, generator=g0))
    gw_ref, gb_ref = torch.autograd.grad(a, (w, b), g)
    wc, bc = w.detach().float().reshape(64, 9).contiguous().cuda(), b.detach().float().cuda()
    out = ops.hwr_stem(img.float().cuda(), wc, bc)
    assert _rel(out.float().permute(0, 3, 1, 2).cpu(), a.detach()) <= 6e-3
    dw, db = ops.hwr_stem_bwd(img.float().cuda(), wc, bc, _nhwc(g))
    assert _rel(dw.view(64, 1, 3, 3).cpu(), gw_ref) <= 2e-3
    assert _rel(db.cpu(), gb_ref) <= 2e-3


@pytest.mark.parametrize("N,H,W", [(2, 64, 128), (1, 64, 200), (3, 16, 72)])
def test_stem_image_gradient_fused(N, H, W):
    """hwg_hwr_stem_bwd_image == autograd of MaxPool(ReLU(conv0(img))) w.r.t. the image, and == the two-step path
    (expand + 9-tap transposed convolution) it replaces."""
    from handwriting_line_generation_b200 import ops
    g0 = torch.Generator().manual_seed(W + N)
    img = (torch.rand(N, 1, H, W, generator=g0) * 2 - 1).double().requires_grad_()
    w = (torch.randn(64, 1, 3, 3, generator=g0) / 3).float().double()
    b = (torch.randn(64, generator=g0) * 0.1).float().double()
    a = F.max_pool2d(F.relu(F.conv2d(img, w, b, padding=1)), 2, 2)
    g = _bf(torch.randn(a.shape, generator=g0))
    (ref,) = torch.autograd.grad(a, img, g.double())
    wc, bc = w.float().reshape(64, 9).contiguous().cuda(), b.float().cuda()
    got = ops.hwr_stem_bwd_image(img.detach().float().cuda(), wc, bc, _nhwc(g))
    assert got.shape == (N, 1, H, W)
    assert _rel(got.cpu().double(), ref) <= 2e-3
