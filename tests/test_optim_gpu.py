"""GPU: flat fused Adam (hwg_adam_flat) against torch.optim.Adam + clip_grad_value_ (the reference trainer's step,
trainer/hw_with_style_trainer.py:381-391); hwg_linear_bwd_f32 against autograd; the generator's direct-to-flat-buffer
gradient path against the autograd-returned gradients."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def test_flat_adam_matches_torch_adam_with_value_clipping():
    import handwriting_line_generation_b200 as pkg
    torch.manual_seed(0)
    shapes = [(7, 5), (3,), (16, 4, 3, 3), (1, 9, 1, 1)]
    ours = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    opt = pkg.FlatAdam(ours, lr=2e-4, betas=(0.5, 0.999), clip_value=2.0)
    ropt = torch.optim.Adam(ref, lr=2e-4, betas=(0.5, 0.999))
    v0 = [p._version for p in ours]
    for it in range(5):
        for p, r in zip(ours, ref):
            g = torch.randn(p.shape, device="cuda") * (4.0 if it % 2 else 0.5)
            p.grad.add_(g)                    # gradients accumulate into the flat buffer views
            r.grad = g.clone()
        torch.nn.utils.clip_grad_value_(ref, 2.0)
        ropt.step()
        opt.step()
        assert float(opt.flat_g.abs().max()) == 0.0          # zeroed by the step
        for p, r in zip(ours, ref):
            assert torch.allclose(p, r, rtol=1e-5, atol=1e-7), (it, (p - r).abs().max())
    assert all(p._version > v for p, v in zip(ours, v0))       # derived-weight caches see the update
    assert all(p.data_ptr() >= opt.flat_p.data_ptr() for p in ours)


@pytest.mark.parametrize("act", [0, 2])
def test_linear_bwd_matches_autograd(act):
    from handwriting_line_generation_b200 import ops
    torch.manual_seed(1)
    B, K, O = 6, 128, 200
    x = torch.randn(B, K, device="cuda", requires_grad=True)
    W = torch.randn(O, K, device="cuda", requires_grad=True)
    b = torch.randn(O, device="cuda", requires_grad=True)
    y = F.linear(x, W, b)
    if act:
        y = F.leaky_relu(y, 0.2)
    gy = torch.randn_like(y)
    y.backward(gy)
    gx, gW, gb = ops.linear_bwd(x.detach(), y.detach(), gy, W.detach(), act, 0.2)
    for a, r in ((gx, x.grad), (gW, W.grad), (gb, b.grad)):
        assert torch.allclose(a, r, rtol=1e-4, atol=1e-4), (a - r).abs().max()
    # accumulate into existing buffers
    accW, accb = torch.ones_like(gW), torch.ones_like(gb)
    ops.linear_bwd(x.detach(), y.detach(), gy, W.detach(), act, 0.2, need_gx=False, gW=accW, gb=accb, accumulate=True)
    assert torch.allclose(accW, W.grad + 1, rtol=1e-4, atol=1e-4) and torch.allclose(accb, b.grad + 1, rtol=1e-4, atol=1e-4)


def test_generator_gradients_direct_to_flat_buffer_equal_autograd_path():
    """The backward kernels can add their gradients straight into FlatAdam's buffer (module._grad_sink) instead of
    returning them to autograd.  Both modes are run on the SAME saved forward state (one forward_train context), so
    the only differences left are fp32 atomics orders inside the backward kernels (a few bf16 rounding flips in the
    propagated gradient): every tensor must agree to 2e-2 rel-L2, the style path (no bf16 anywhere) to 1e-4."""
    import handwriting_line_generation_b200 as pkg
    from handwriting_line_generation_b200 import autograd_gen as ag, ops
    from oracle import synth
    from tests.test_modules_gpu import rel_l2
    torch.manual_seed(2)
    T, B = 24, 2
    gen = pkg.SpacedGenerator(80, 128, 256, n_style_trans=6, emb_dropout=False, append_style=True).cuda().train()
    content, style = synth.gen_case(T, B, 80, 128, 4, dense=True)
    c, st = torch.from_numpy(content).cuda(), torch.from_numpy(style).cuda()
    noise = [torch.randn(sh, device="cuda") for sh in synth.gen_noise_shapes(T, B, 256)]
    g_out = torch.randn(B, 1, 64, 4 * T, device="cuda")
    plist, splist = ag._param_list(gen), ag._style_params(gen)
    with torch.no_grad():
        s0 = ops.pixelnorm(st)
        s_, gb = ag._style_path(gen, st)
        out, ctx = ag.forward_train(gen, c, s_, gb, noise)
        g_content, g_s, g_gb, flat = ag.backward_train(gen, ctx, g_out)
        ref = [t.clone() for t in flat]
        ref_c, ref_s, ref_gb = g_content.clone(), g_s.clone(), g_gb.clone()
    # style path reference through torch autograd
    s0r = s0.clone().requires_grad_()
    s2, gb2 = ag._StyleFn.apply(gen, s0r, *splist)
    torch.autograd.backward([s2, gb2], [ref_s, ref_gb])
    ref_style = [p.grad.clone() for p in splist]
    ref_s0 = s0r.grad.clone()
    for p in gen.parameters():
        p.grad = None
    # ---- flat-buffer mode
    opt = pkg.FlatAdam(gen.parameters(), lr=2e-4, betas=(0.5, 0.999))
    gen._grad_sink = opt
    with torch.no_grad():
        g_content, g_s, g_gb, flat = ag.backward_train(gen, ctx, g_out)
    assert all(t is None for t in flat)
    torch.cuda.synchronize()
    for p, r in zip(plist, ref):
        assert rel_l2(opt.grad_view(p).cpu().numpy(), r.cpu().numpy()) <= 2e-2, tuple(p.shape)
    assert rel_l2(g_content.cpu().numpy(), ref_c.cpu().numpy()) <= 2e-2
    assert rel_l2(g_gb.cpu().numpy(), ref_gb.cpu().numpy()) <= 2e-2
    s0r = s0.clone().requires_grad_()
    s2, gb2 = ag._StyleFn.apply(gen, s0r, *splist)
    torch.autograd.backward([s2, gb2], [ref_s, ref_gb])
    torch.cuda.synchronize()
    for p, r in zip(splist, ref_style):
        assert p.grad.data_ptr() == opt.grad_view(p).data_ptr()
        assert rel_l2(p.grad.cpu().numpy(), r.cpu().numpy()) <= 1e-4, tuple(p.shape)
    assert rel_l2(s0r.grad.cpu().numpy(), ref_s0.cpu().numpy()) <= 1e-4
    # a second backward accumulates
    before = opt.grad_view(plist[-2]).clone()
    with torch.no_grad():
        ag.backward_train(gen, ctx, g_out)
    assert rel_l2((opt.grad_view(plist[-2]) - before).cpu().numpy(), ref[-2].cpu().numpy()) <= 2e-2
    # the public path end to end: gradients land in the flat buffer, optimizer step consumes and clears it
    w0 = plist[0].detach().clone()
    (gen(c, st, noise=noise) * g_out).sum().backward()
    assert all(p.grad.data_ptr() == opt.grad_view(p).data_ptr() for p in gen.parameters())
    opt.step()
    assert float(opt.flat_g.abs().max()) == 0.0 and float((plist[0] - w0).abs().max()) > 0


def test_recognizer_gradients_direct_to_flat_buffer_equal_autograd_path():
    """CNNOnlyHWR backward on ONE saved forward state: gradients returned to autograd (job table -> fresh flat tensor)
    vs added straight into FlatAdam's buffer (module._grad_sink)."""
    import handwriting_line_generation_b200 as pkg
    from handwriting_line_generation_b200 import autograd_hwr as ah
    from oracle import synth
    from tests.test_modules_gpu import rel_l2
    torch.manual_seed(4)
    B, W, C = 2, 256, 78                               # 78 classes: the head's wgrad runs on 80 padded channels
    hwr = pkg.CNNOnlyHWR(C, norm='batch').cuda().train()
    x = torch.from_numpy(synth.hwr_case(B, W, 5)).cuda()
    with torch.no_grad():
        lp, ctx = ah.forward_train(hwr, x)
        g_lp = torch.randn_like(lp)
        ref, g_img = ah.backward_train(hwr, ctx, g_lp, want_input_grad=True)
        ref = {n: v.clone() for n, v in ref.items()}
    assert set(ref) == {n for n, _ in hwr.named_parameters()} and all(v is not None for v in ref.values())
    for n, p in hwr.named_parameters():
        assert ref[n].shape == p.shape, n
    opt = pkg.FlatAdam(hwr.parameters(), lr=1e-4)
    hwr._grad_sink = opt
    with torch.no_grad():
        got, g_img2 = ah.backward_train(hwr, ctx, g_lp, want_input_grad=True)
    assert all(v is None for v in got.values())
    torch.cuda.synchronize()
    for n, p in hwr.named_parameters():
        assert rel_l2(opt.grad_view(p).cpu().numpy(), ref[n].cpu().numpy()) <= 2e-2, n
    assert rel_l2(g_img2.cpu().numpy(), g_img.cpu().numpy()) <= 2e-2
    # frozen recognizer: no wgrad launches, no gradients, image gradient only
    for p in hwr.parameters():
        p.requires_grad_(False)
    hwr._grad_sink = None
    xg = x.clone().requires_grad_()
    hwr(xg).backward(g_lp)
    assert xg.grad is not None and all(p.grad is None or p.grad.data_ptr() == opt.grad_view(p).data_ptr() for p in hwr.parameters())
