"""GPU parity of the discriminator path: each memory-bound kernel against torch (fp64 autograd on
identical bf16-rounded inputs), then the drop-in DiscriminatorAP against the golden outputs / input gradients of the
unmodified reference (tests/golden/disc.npz) and against the oracle with bf16-storage emulation."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import disc as odisc
from oracle import synth
from oracle.make_golden import DISC_CASES, digest

pytestmark = pytest.mark.gpu
LEAK = 0.1


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-20)).item()


def _rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()


def _nchw(x):
    return x.float().cpu().permute(0, 3, 1, 2).double()


def _bf(x):
    return x.to(torch.bfloat16).double()


def test_shift_expand_and_collapse_are_adjoint_and_match_definition():
    from handwriting_line_generation_b200 import _lib
    g0 = torch.Generator().manual_seed(0)
    N, H, W = 2, 9, 37
    img = torch.randn(N, 1, H, W, generator=g0)
    out = torch.empty((N, H, W, 16), device="cuda", dtype=torch.bfloat16)
    img_d = img.cuda()
    _lib.call("hwg_shift_expand", img_d.data_ptr(), out.data_ptr(), N, H, W, 7, 3, _lib.stream())
    ref = torch.zeros(N, H, W, 16)
    pad = F.pad(img[:, 0], (3, 3))
    for j in range(7):
        ref[..., j] = pad[:, :, j:j + W]
    assert torch.equal(out.float().cpu(), ref.to(torch.bfloat16).float())
    g = torch.randn(N, H, W, 16, generator=g0).to(torch.bfloat16)
    dimg = torch.empty((N, 1, H, W), device="cuda")
    g_d = g.cuda()
    _lib.call("hwg_shift_collapse", g_d.data_ptr(), dimg.data_ptr(), N, H, W, 7, 3, 0, _lib.stream())
    # adjoint: <expand(img), g> == <img, collapse(g)>
    lhs = (ref.double() * g.double()).sum()
    rhs = (img.double() * dimg.cpu().double()).sum()
    assert abs(lhs - rhs) <= 1e-4 * abs(lhs)


@pytest.mark.parametrize("N,H,W,kh,kw,ph,pw,Cout", [
    (2, 64, 200, 7, 7, 0, 3, 64),        # DiscriminatorAP.in_conv (discriminator_ap.py:75), ragged width
    (3, 64, 128, 5, 5, 2, 2, 32),        # Encoder2.down_conv1[0] (autoencoder.py:345)
    (1, 9, 37, 7, 7, 0, 3, 64),          # fewer rows / columns than one tile
    (2, 21, 70, 3, 4, 1, 2, 32),         # even kernel width, odd tile remainders
])
def test_stem_conv_matches_conv2d(N, H, W, kh, kw, ph, pw, Cout):
    """hwg_stem_conv (one-input-channel convolution straight from the fp32 image) against F.conv2d in fp64 on the bf16-rounded
    operands, its statistics epilogue, and against the route it replaces (hwg_shift_expand + kh-tap hwg_conv_fprop)."""
    from handwriting_line_generation_b200 import _lib, conv
    g0 = torch.Generator().manual_seed(N * 1000 + W)
    img = torch.randn(N, 1, H, W, generator=g0)
    w = (torch.randn(Cout, 1, kh, kw, generator=g0) / (kh * kw) ** 0.5).to(torch.bfloat16)
    b = torch.randn(Cout, generator=g0)
    packed = torch.zeros(kh, Cout, 16, dtype=torch.bfloat16)
    packed[:, :, :kw] = w[:, 0].permute(1, 0, 2)
    packed[:, :, kw:] = 7.0                       # columns past kw must be ignored
    ref = F.conv2d(_bf(img), w.double(), b.double(), padding=(ph, pw))           # [N,Cout,Ho,Wo]
    Ho, Wo = ref.shape[2], ref.shape[3]
    img_d, pk_d, b_d = img.cuda(), packed.cuda(), b.cuda()
    y = torch.empty((N, Ho, Wo, Cout), device="cuda", dtype=torch.bfloat16)
    st = torch.zeros((N, Cout, 2), device="cuda")
    _lib.call("hwg_stem_conv", img_d.data_ptr(), pk_d.data_ptr(), b_d.data_ptr(), N, H, W, kh, kw, ph, pw, Cout,
              y.data_ptr(), st.data_ptr(), _lib.stream())
    got = _nchw(y)
    assert ((got - ref).abs().max() / ref.abs().max()).item() <= 6e-3
    s_ref = torch.stack([ref.sum((2, 3)), (ref * ref).sum((2, 3))], -1)
    assert ((st.cpu().double() - s_ref).abs().max() / s_ref.abs().max()).item() <= 1e-3
    if ph == 0 and kw <= 7:
        # the shift-expansion route on the tensor-core kernel: same operands, fp32 accumulation in another order
        x16 = torch.empty((N, H, W, 16), device="cuda", dtype=torch.bfloat16)
        _lib.call("hwg_shift_expand", img_d.data_ptr(), x16.data_ptr(), N, H, W, kw, pw, _lib.stream())
        pk0 = packed.clone(); pk0[:, :, kw:] = 0
        y2 = conv.conv_fprop(x16, pk0.cuda(), [(dy, 0) for dy in range(kh)], Ho, Wo, bias=b_d)
        assert ((_nchw(y2) - got).abs().max() / ref.abs().max()).item() <= 8e-3


@pytest.mark.parametrize("N,C,H,W", [(2, 64, 10, 33), (3, 128, 5, 16)])
def test_gn_coeffs(N, C, H, W):
    from handwriting_line_generation_b200 import _lib
    g0 = torch.Generator().manual_seed(C)
    z = torch.randn(N, C, H, W, generator=g0, dtype=torch.float64) * 1.3 + 0.2
    gamma, beta = torch.rand(C, generator=g0) + 0.5, torch.randn(C, generator=g0)
    stats = torch.stack([z.sum((2, 3)), (z * z).sum((2, 3))], -1).float().cuda().contiguous()
    coef = torch.empty((N, C, 2), device="cuda")
    save = torch.empty((N, C, 2), device="cuda")
    gam, bet = gamma.cuda(), beta.cuda()          # keep the device copies alive across the launch
    _lib.call("hwg_gn_coeffs", stats.data_ptr(), gam.data_ptr(), bet.data_ptr(), N, C, 8, H * W, 1e-5,
              coef.data_ptr(), save.data_ptr(), _lib.stream())
    ref = F.group_norm(z, 8, gamma.double(), beta.double(), 1e-5)
    got = coef[:, :, 0].cpu().double()[:, :, None, None] * z + coef[:, :, 1].cpu().double()[:, :, None, None]
    assert _rel(got, ref) <= 1e-4


@pytest.mark.parametrize("N,C,H,W,kh,kw", [(2, 64, 8, 12, 2, 2), (2, 128, 5, 33, 2, 2), (3, 256, 1, 17, 1, 2)])
def test_avgpool(N, C, H, W, kh, kw):
    from handwriting_line_generation_b200 import _lib
    x = torch.randn(N, C, H, W, generator=torch.Generator().manual_seed(W))
    xn = _nhwc(x)
    y = torch.empty((N, H // kh, W // kw, C), device="cuda", dtype=torch.bfloat16)
    _lib.call("hwg_avgpool_nhwc", xn.data_ptr(), y.data_ptr(), N, H, W, C, kh, kw, _lib.stream())
    ref = F.avg_pool2d(_bf(x), (kh, kw))
    assert _rel(_nchw(y), ref) <= 6e-3          # bf16 output


@pytest.mark.parametrize("N,C,H,W,kh,kw,scaled", [(2, 64, 8, 12, 2, 2, False), (2, 128, 5, 33, 1, 1, True),
                                                   (2, 256, 1, 17, 1, 2, True)])
def test_act_bwd(N, C, H, W, kh, kw, scaled):
    from handwriting_line_generation_b200 import _lib
    g0 = torch.Generator().manual_seed(H * W)
    x = _bf(torch.randn(N, C, H, W, generator=g0)).requires_grad_()
    scale = ((torch.rand(N, C, generator=g0) > 0.3).double() / 0.95) if scaled else torch.ones(N, C, dtype=torch.float64)
    y = F.leaky_relu(x * scale[:, :, None, None], LEAK)
    out = F.avg_pool2d(y, (kh, kw))
    g = _bf(torch.randn(out.shape, generator=g0))
    (gx,) = torch.autograd.grad(out, x, g)
    yn = _nhwc(y.detach().float())
    gz = torch.empty_like(yn)
    g_d, scale_d = _nhwc(g.float()), scale.float().cuda()
    _lib.call("hwg_act_bwd", g_d.data_ptr(), yn.data_ptr(), scale_d.data_ptr() if scaled else None, LEAK, N, H, W, C, kh, kw, gz.data_ptr(), _lib.stream())
    # the mask comes from the bf16-rounded y: identical signs except exact zeros of dropped channels (scale 0 -> gz 0)
    assert _rel(_nchw(gz), gx) <= 6e-3


@pytest.mark.parametrize("N,C,H,W,kh,kw", [(2, 64, 6, 20, 1, 1), (2, 128, 10, 13, 2, 2), (1, 16, 4, 8, 1, 1)])
def test_groupnorm_lrelu_pool_bwd(N, C, H, W, kh, kw):
    from handwriting_line_generation_b200 import _lib
    g0 = torch.Generator().manual_seed(C + W)
    groups = 8
    z = _bf(torch.randn(N, C, H, W, generator=g0) * 1.2 + 0.1).requires_grad_()
    gamma = (torch.rand(C, generator=g0) + 0.5).double()
    beta = torch.randn(C, generator=g0).double() * 0.3
    a = F.leaky_relu(F.group_norm(z, groups, gamma, beta, 1e-5), LEAK)
    out = F.avg_pool2d(a, (kh, kw))
    g = _bf(torch.randn(out.shape, generator=g0))
    (gz_ref,) = torch.autograd.grad(out, z, g)
    zd = z.detach()
    stats = torch.stack([zd.sum((2, 3)), (zd * zd).sum((2, 3))], -1).float().cuda().contiguous()
    coef = torch.empty((N, C, 2), device="cuda")
    save = torch.empty((N, C, 2), device="cuda")
    s = _lib.stream()
    gam, bet = gamma.float().cuda(), beta.float().cuda()
    _lib.call("hwg_gn_coeffs", stats.data_ptr(), gam.data_ptr(), bet.data_ptr(), N, C, groups, H * W, 1e-5,
              coef.data_ptr(), save.data_ptr(), s)
    zn, gn = _nhwc(zd.float()), _nhwc(g.float())
    sums = torch.zeros((N, C, 2), device="cuda")
    spq = torch.empty((N, C, 3), device="cuda")
    dgam, dbet = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    gz = torch.empty_like(zn)
    _lib.call("hwg_norm_bwd_reduce", gn.data_ptr(), zn.data_ptr(), coef.data_ptr(), LEAK, N, H, W, C, kh, kw,
              sums.data_ptr(), s)
    _lib.call("hwg_gn_bwd_coeffs", sums.data_ptr(), save.data_ptr(), gam.data_ptr(), N, C, groups, H * W, spq.data_ptr(),
              dgam.data_ptr(), dbet.data_ptr(), s)
    _lib.call("hwg_norm_bwd_apply", gn.data_ptr(), zn.data_ptr(), coef.data_ptr(), spq.data_ptr(), LEAK, N, H, W, C, kh,
              kw, gz.data_ptr(), s)
    assert _rel(_nchw(gz), gz_ref) <= 1e-2
    gg, gb = torch.autograd.grad(F.avg_pool2d(F.leaky_relu(F.group_norm(zd, groups, gamma.requires_grad_(),
                                                                         beta.requires_grad_(), 1e-5), LEAK), (kh, kw)),
                                 (gamma, beta), g)
    assert _rel(dgam.cpu(), gg) <= 2e-3 and _rel(dbet.cpu(), gb) <= 2e-3


def _module(seed, training):
    from handwriting_line_generation_b200 import DiscriminatorAP
    torch.manual_seed(seed)
    m = DiscriminatorAP(64, use_low=True, use_med=True)
    sd = synth.perturb_disc(m.state_dict(), seed + 1)
    sd_cpu = {k: v.clone() for k, v in sd.items()}
    m = m.cuda().train(training)
    for p in m.parameters():
        p.requires_grad_(False)
    return m, sd_cpu


def test_spectral_norm_and_packed_weights_match_oracle():
    m, sd = _module(7, True)
    upd = {}
    w_ref = {site: odisc.spectral_weight(sd, site, upd) for site in ("convs1.0", "convs3.4", "finalMed.0", "convs4.14")}
    c = m._prepare()
    torch.cuda.synchronize()
    for site, w in w_ref.items():
        mod = dict((s, mm) for s, mm, _, _ in m.conv_layers())[site]
        assert _rel(mod.weight_u.cpu(), upd[site + ".module.weight_u"]) <= 1e-4
        assert _rel(mod.weight_v.cpu(), upd[site + ".module.weight_v"]) <= 1e-4
        co, ci, kh, kw = w.shape
        packed = c[site].float().cpu()[:, :co, :ci]                       # [taps][co][ci]
        ref = w.permute(2, 3, 0, 1).reshape(kh * kw, co, ci).to(torch.bfloat16).float()
        assert _rel(packed, ref) <= 1e-2                                  # 1/sigma in fp32, then one bf16 rounding
        if c[site].size(1) > co:
            assert c[site][:, co:].abs().max() == 0
    # in_conv operand: taps = kernel rows, channels = kernel columns (7 of 16)
    w0 = sd["in_conv.0.weight"]
    assert torch.equal(c["in_conv.0"].float().cpu()[:, :, :7], w0[:, 0].permute(1, 0, 2).to(torch.bfloat16).float())
    assert c["in_conv.0"][:, :, 7:].abs().max() == 0


@pytest.mark.parametrize("name", sorted(DISC_CASES))
def test_discriminator_matches_reference_golden(name, golden_dir):
    """Predictions and the generator-loss input gradient against the unmodified reference; bf16 path tolerance:
    per-tensor rel-L2 <= 2e-2 (BASELINE north_star)."""
    gold = np.load(f"{golden_dir}/disc.npz")
    B, W, wseed, iseed, training = DISC_CASES[name]
    m, sd = _module(wseed, training)
    m.dropout_masks = {k: torch.from_numpy(v) for k, v in synth.disc_masks(B, iseed + 7).items()}
    img = torch.from_numpy(synth.hwr_case(B, W, iseed)).cuda().requires_grad_()
    preds = m(img)
    loss = odisc.gen_loss(preds)
    loss.backward()
    for i, p in enumerate(preds):
        ref = torch.from_numpy(gold[f"{name}/pred{i}"])
        assert list(p.shape) == list(ref.shape)
        assert _rel_l2(p.detach().cpu(), ref) <= 2e-2, (i, _rel_l2(p.detach().cpu(), ref))
    assert abs(loss.item() - float(gold[f"{name}/loss"])) <= 2e-2 * abs(float(gold[f"{name}/loss"]))
    # spectral-norm vectors are updated in place like the reference's
    for k in ("convs1.0.module.weight_u", "convs3.4.module.weight_v", "convs4.14.module.weight_u"):
        assert np.abs(m.state_dict()[k].cpu().numpy() - gold[f"{name}/{k}"]).max() <= 1e-4
    # input gradient: compare with the oracle at fp32 on the full tensor (the golden stores a strided sample)
    img2 = torch.from_numpy(synth.hwr_case(B, W, iseed)).requires_grad_()
    masks = {k: torch.from_numpy(v) for k, v in synth.disc_masks(B, iseed + 7).items()}
    odisc.gen_loss(odisc.disc_forward(sd, img2, masks, training=training)).backward()
    _, samp = digest(img2.grad.numpy())
    assert np.abs(samp - gold[f"{name}/grad_sample"]).max() <= 1e-4 * gold[f"{name}/grad_digest"][3]   # oracle == reference
    e = _rel_l2(img.grad.cpu(), img2.grad)
    cos = F.cosine_similarity(img.grad.cpu().double().flatten(), img2.grad.double().flatten(), dim=0).item()
    # the gradient crosses 12 bf16 layers with LeakyReLU masks: bounded against the bf16-emulating oracle
    img3 = torch.from_numpy(synth.hwr_case(B, W, iseed)).requires_grad_()
    odisc.gen_loss(odisc.disc_forward(sd, img3, masks, training=training, emulate_bf16=True)).backward()
    e_emul = _rel_l2(img3.grad, img2.grad)
    assert e <= 1.3 * e_emul + 2e-2 and cos >= 0.95, (e, e_emul, cos)


def test_discriminator_eval_no_grad():
    m, sd = _module(11, False)
    x = torch.from_numpy(synth.hwr_case(2, 96, 5)).cuda()
    with torch.no_grad():
        preds = m(x)
    ref = odisc.disc_forward(sd, torch.from_numpy(synth.hwr_case(2, 96, 5)), None, training=False)
    for p, r in zip(preds, ref):
        assert _rel_l2(p.cpu(), r) <= 2e-2


@pytest.mark.parametrize("rows,C", [(1000, 64), (37, 16), (513, 256)])
def test_channel_sum(rows, C):
    from handwriting_line_generation_b200 import _lib
    x = torch.randn(rows, C, generator=torch.Generator().manual_seed(rows)).to(torch.bfloat16).cuda()
    out = torch.zeros(C, device="cuda")
    _lib.call("hwg_channel_sum", x.data_ptr(), rows, C, out.data_ptr(), _lib.stream())
    assert _rel(out.cpu(), x.double().cpu().sum(0)) <= 1e-5


def test_spectral_norm_backward_matches_autograd():
    from handwriting_line_generation_b200 import _lib
    g0 = torch.Generator().manual_seed(3)
    shapes = [(64, 576), (1, 2304), (256, 1152)]
    jobs = np.zeros((len(shapes), 5), np.int64)
    keep, refs = [], []
    inv = torch.empty(len(shapes))
    for i, (h, wd) in enumerate(shapes):
        w = torch.randn(h, wd, generator=g0, dtype=torch.float64).requires_grad_()
        u = F.normalize(torch.randn(h, generator=g0, dtype=torch.float64), dim=0)
        v = F.normalize(torch.randn(wd, generator=g0, dtype=torch.float64), dim=0)
        G = torch.randn(h, wd, generator=g0, dtype=torch.float64)
        sigma = u.dot(w.mv(v))
        ((w / sigma) * G).sum().backward()
        refs.append(w.grad)
        inv[i] = 1.0 / sigma.item()
        wd_, gw, ud, vd = w.detach().float().cuda(), (G / sigma.item()).float().cuda(), u.float().cuda(), v.float().cuda()
        keep.append((wd_, gw, ud, vd))
        jobs[i] = (wd_.data_ptr(), gw.data_ptr(), ud.data_ptr(), vd.data_ptr(), h | (wd << 32))
    jd, invd = torch.from_numpy(jobs).cuda(), inv.cuda()
    dots = torch.zeros(len(shapes), device="cuda")
    _lib.call("hwg_spectral_norm_bwd", jd.data_ptr(), len(shapes), max(h * w for h, w in shapes), invd.data_ptr(),
              dots.data_ptr(), _lib.stream())
    for (_, gw, _, _), ref in zip(keep, refs):
        assert _rel(gw.cpu(), ref) <= 1e-4


@pytest.mark.parametrize("name", ["hinge_w128", "hinge_w200"])
def test_discriminator_lesson_parameter_gradients(name, golden_dir):
    """'disc' lesson: hinge loss on real || fake rows, gradient of every trainable parameter against the oracle (itself
    pinned to the reference's gradients on CPU); bf16 pipeline bound as in DESIGN §5."""
    from oracle.make_golden import DISC_LESSON_CASES
    from tests.test_disc_cpu import lesson_oracle_grads
    gold = np.load(f"{golden_dir}/disc.npz")
    B, W, wseed, iseed = DISC_LESSON_CASES[name]
    from handwriting_line_generation_b200 import DiscriminatorAP
    torch.manual_seed(wseed)
    m = DiscriminatorAP(64, use_low=True, use_med=True)
    sd = {k: v.clone() for k, v in synth.perturb_disc(m.state_dict(), wseed + 1).items()}
    m = m.cuda().train()
    m.dropout_masks = {k: torch.from_numpy(v) for k, v in synth.disc_masks(B, iseed + 7).items()}
    preds = m(torch.from_numpy(synth.hwr_case(B, W, iseed)).cuda())
    loss = odisc.hinge_loss(preds, B // 2)
    loss.backward()
    assert abs(loss.item() - float(gold[f"{name}/loss"])) <= 2e-2 * abs(float(gold[f"{name}/loss"]))
    _, ref = lesson_oracle_grads(sd, B, W, iseed)
    _, emu = lesson_oracle_grads(sd, B, W, iseed, emulate_bf16=True)
    got = {n: p.grad for n, p in m.named_parameters() if p.requires_grad}
    assert sorted(got) == sorted(ref)
    worst = {}
    for n in ref:
        assert got[n] is not None and got[n].shape == ref[n].shape, n
        if ref[n].abs().max() < 1e-7:      # e.g. the bias of a head when every hinge term is active: +1/n and -1/n cancel
            assert got[n].abs().max() <= 1e-4, n
            continue
        e, e_emu = _rel_l2(got[n].cpu(), ref[n]), _rel_l2(emu[n], ref[n])
        cos = F.cosine_similarity(got[n].cpu().double().flatten(), ref[n].double().flatten(), dim=0).item()
        worst[n] = (e, e_emu, cos)
        assert e <= 1.3 * e_emu + 2e-2 and cos >= 0.95, (n, e, e_emu, cos)
    # a second backward through a fresh forward must not alias the first gradients (the unpack workspace is reused)
    g_first = {n: g.clone() for n, g in got.items()}
    for p in m.parameters():
        p.grad = None
    odisc.hinge_loss(m(torch.from_numpy(synth.hwr_case(B, W, iseed)).cuda()), B // 2).backward()
    assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.requires_grad)
    assert all(torch.equal(got[n], g_first[n]) for n in g_first)
