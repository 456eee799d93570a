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
    """Same weights, inputs and noise: gradients accumulated by the backward kernels straight into FlatAdam's
    buffer (module._grad_sink) == the gradients the autograd.Functions return.  The statistics are summed with fp32
    atomics, so two runs of the SAME path differ at the bf16 noise floor (a few rounding flips cascade through ten
    layers); that floor is measured here (two autograd-path runs) and the flat-buffer path must sit within 3x of it."""
    import handwriting_line_generation_b200 as pkg
    from oracle import synth
    from tests.test_modules_gpu import rel_l2
    torch.manual_seed(2)
    T, B = 24, 2
    gen = pkg.SpacedGenerator(80, 128, 256, n_style_trans=6, emb_dropout=False, append_style=True).cuda().train()
    content, style = synth.gen_case(T, B, 80, 128, 4, dense=True)
    c = torch.from_numpy(content).cuda().requires_grad_()
    s = torch.from_numpy(style).cuda().requires_grad_()
    noise = [torch.randn(sh, device="cuda") for sh in synth.gen_noise_shapes(T, B, 256)]
    w = torch.randn(B, 1, 64, 4 * T, device="cuda")

    def run():
        for p in gen.parameters():
            if getattr(gen, "_grad_sink", None) is None:
                p.grad = None
        c.grad = s.grad = None
        (gen(c, s, noise=noise) * w).sum().backward()
        torch.cuda.synchronize()
        g = {n: p.grad.detach().cpu().numpy().copy() for n, p in gen.named_parameters()}
        g["<content>"], g["<style>"] = c.grad.cpu().numpy().copy(), s.grad.cpu().numpy().copy()
        return g

    ref, ref2 = run(), run()
    opt = pkg.FlatAdam(gen.parameters(), lr=2e-4, betas=(0.5, 0.999))
    gen._grad_sink = opt
    got = run()
    for n, p in gen.named_parameters():
        assert p.grad.data_ptr() == opt.grad_view(p).data_ptr(), n
    for n in ref:
        floor = rel_l2(ref2[n], ref[n])
        assert rel_l2(got[n], ref[n]) <= 3 * floor + 2e-3, (n, rel_l2(got[n], ref[n]), floor)
    # a second backward accumulates into the same buffer
    n0 = "out.0.conv.weight_orig"
    before = dict(gen.named_parameters())[n0].grad.clone()
    (gen(c, s, noise=noise) * w).sum().backward()
    after = dict(gen.named_parameters())[n0].grad
    assert rel_l2((after - before).cpu().numpy(), ref[n0]) <= 2e-2
