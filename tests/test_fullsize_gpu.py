"""GPU, BASELINE.json's full sizes: size-independent properties of the hot path where an oracle run would take
minutes (north_star ③): per-line independence of the generator (InstanceNorm / AdaIN are per sample), per-line
independence of the recognizer in eval mode, exact 2x-linearity of the generator backward on one saved forward,
and the long-line stress shapes of configs[4] against torch's own CUDA CTC on the same log-probs."""
import numpy as np
import pytest
import torch

from oracle import synth
from tests.test_modules_gpu import rel_l2

pytestmark = pytest.mark.gpu


def _gen(n_class=80):
    import handwriting_line_generation_b200 as pkg
    torch.manual_seed(0)
    return pkg.SpacedGenerator(n_class, 128, 256, n_style_trans=6, emb_dropout=False, append_style=True).cuda()


def test_generator_config2_lines_are_independent():
    """configs[1] shapes (batch 32, T_s = 256 -> 64x1024 px): lines 5..8 generated inside the batch == the same four
    lines generated alone with the same noise.  The fp32 atomics order of the statistics differs between the two
    launches, which flips a few bf16 roundings per layer: the bound is the bf16-path tolerance (2e-2; observed 9e-3)."""
    B, T = 32, 256
    gen = _gen().eval()
    content, style = synth.gen_case(T, B, 80, 128, 21)
    c, s = torch.from_numpy(content).cuda(), torch.from_numpy(style).cuda()
    g = torch.Generator(device="cuda").manual_seed(5)
    noise = [torch.randn(sh, device="cuda", generator=g) for sh in synth.gen_noise_shapes(T, B, 256)]
    with torch.no_grad():
        full = gen(c, s, noise=noise)
        part = gen(c[:, 5:9].contiguous(), s[5:9].contiguous(), noise=[z[5:9].contiguous() for z in noise])
    assert full.shape == (B, 1, 64, 4 * T)
    assert rel_l2(part.cpu().numpy(), full[5:9].cpu().numpy()) <= 2e-2
    assert float(full.abs().max()) <= 1.0 and torch.isfinite(full).all()


def test_recognizer_config1_eval_lines_are_independent():
    """configs[0] shapes (8 lines of 64x1024, 80 classes), eval-mode BatchNorm: per-line log-probs do not depend on
    the rest of the batch; every frame is a normalised distribution."""
    import handwriting_line_generation_b200 as pkg
    torch.manual_seed(1)
    hwr = pkg.CNNOnlyHWR(80, norm='batch').cuda().eval()
    x = torch.from_numpy(synth.hwr_case(8, 1024, 3)).cuda()
    with torch.no_grad():
        full = hwr(x)
        part = hwr(x[2:4].contiguous())
    assert full.shape == (250, 8, 80)
    assert rel_l2(part.cpu().numpy(), full[:, 2:4].cpu().numpy()) <= 5e-3
    assert torch.allclose(full.exp().sum(2), torch.ones(250, 8, device="cuda"), atol=1e-3)


def test_generator_backward_is_linear_in_the_upstream_gradient():
    """Train-step shapes (16 lines of 64x1024): on ONE saved forward, backward(2g) == 2 backward(g) for every
    parameter gradient (scaling by two is exact in bf16 and fp32, so only atomics orders differ)."""
    from handwriting_line_generation_b200 import autograd_gen as ag, ops
    B, T = 16, 256
    gen = _gen().train()
    content, style = synth.gen_case(T, B, 80, 128, 22, dense=True)
    c, s = torch.from_numpy(content).cuda(), torch.from_numpy(style).cuda()
    g_out = torch.randn(B, 1, 64, 4 * T, device="cuda")
    with torch.no_grad():
        s_, gb = ag._style_path(gen, s)
        torch.manual_seed(3)
        _, ctx = ag.forward_train(gen, c, s_, gb, None)
        gc1, gs1, ggb1, f1 = ag.backward_train(gen, ctx, g_out)
        f1 = [t.clone() for t in f1]
        gc1, ggb1 = gc1.clone(), ggb1.clone()
        gc2, gs2, ggb2, f2 = ag.backward_train(gen, ctx, 2 * g_out)
    for a, b, p in zip(f1, f2, ag._param_list(gen)):
        assert rel_l2(b.cpu().numpy(), 2 * a.cpu().numpy()) <= 2e-2, tuple(p.shape)
    assert rel_l2(gc2.cpu().numpy(), 2 * gc1.cpu().numpy()) <= 2e-2
    assert rel_l2(ggb2.cpu().numpy(), 2 * ggb1.cpu().numpy()) <= 2e-2


def test_long_line_stress_config5_chain():
    """configs[4]: width 2048, 120-char transcripts, RIMES charset (78 classes), batch 64: generator -> recognizer ->
    CTC forward/backward runs at these shapes, the loss equals torch's CUDA ctc_loss on the same log-probs (1e-4),
    gradients reach every generator parameter and are finite, decodes are bit-exact against numpy."""
    import handwriting_line_generation_b200 as pkg
    B, Ts, C, S = 64, 512, 78, 120
    gen = _gen(C).train()
    torch.manual_seed(2)
    hwr = pkg.CNNOnlyHWR(C, norm='batch').cuda().train()
    for p in hwr.parameters():
        p.requires_grad_(False)
    content, style = synth.gen_case(Ts, B, C, 128, 11)
    c, s = torch.from_numpy(content).cuda(), torch.from_numpy(style).cuda()
    T = Ts - 6
    tg = torch.from_numpy(np.random.RandomState(3).randint(1, C, (B, S)).astype(np.int32)).cuda()
    il = torch.full((B,), T, dtype=torch.int32, device="cuda")
    tl = torch.full((B,), S, dtype=torch.int32, device="cuda")
    img = gen(c, s)
    lp = hwr(img)
    loss = pkg.CTCLoss(lp, tg, il, tl)
    loss.backward()
    torch.cuda.synchronize()
    assert img.shape == (B, 1, 64, 4 * Ts) and lp.shape == (T, B, C)
    ref = torch.nn.functional.ctc_loss(lp.detach(), tg.long(), il.long(), tl.long(), blank=0, reduction='mean')
    assert abs(loss.item() - ref.item()) <= 1e-4 * abs(ref.item())
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in gen.parameters())
    raw, dec, dl = pkg.ctc_greedy_decode(lp.detach())
    am = lp.detach().cpu().numpy().argmax(2)                      # [T,B], first maximum wins
    assert np.array_equal(raw.cpu().numpy(), am)
    for b in (0, 17, 63):
        seq = [int(v) for t, v in enumerate(am[:, b]) if v != 0 and (t == 0 or v != am[t - 1, b])]
        assert dec[b, :int(dl[b])].cpu().tolist() == seq
