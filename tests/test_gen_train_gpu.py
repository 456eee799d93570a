"""GPU parity of the generator's training step (forward + backward through every layer, to every parameter, the
style vector and dense content) against torch autograd on the CPU oracle (oracle/gen.py).

As for the recognizer (tests/test_hwr_train_gpu.py) the end-to-end gradient of a bf16-activation pipeline cannot
sit within 2e-2 of the fp32 one: LeakyReLU slopes flip for the ~0.4 % of activations that lie within the bf16
rounding error of zero, each flip changes a whole gradient entry, and ten stacked layers carry that to ~15 % at
the input (torch shows the same with bf16 storage emulation).  Asserted here:
  * image: rel-L2 <= 2e-2 (bf16 path);
  * every gradient tensor: rel-L2(cuda, fp32) <= 1.6 * rel-L2(bf16-emulated torch, fp32) + 2e-2 (2.5x for
    tensors with fewer than 256 entries: 16-element bias sums with heavy cancellation fluctuate more).  The factor is the
    measured spread, not a wish: which LeakyReLU slopes flip depends on the last bits of the InstanceNorm statistics, which
    are summed with fp32 atomics (run to run) and in a kernel-dependent order (build to build).  tests/tools/
    gen_grad_spread.py on the B200, 60 runs over seven builds of round 2: the AdaIN projection of the 32-channel block sat
    anywhere in 0.09 ... 0.145 against 0.088 for the emulation (one build clustering at 0.14), the style-MLP weights in
    0.133 ... 0.174 against 0.108-0.117; a factor of 1.3 failed one run in two on some builds;
  * cosine(cuda, fp32) >= 0.95 on every gradient tensor (0.9 for tensors with fewer than 256 entries);
  * the tensors next to the output (out conv weight, last AdaIN projection) within 3e-2.
The backward kernels themselves are held to <= 1e-2 on identical inputs in test_gen_bwd_ops_gpu.py /
test_conv_bwd_gpu.py."""
import numpy as np
import pytest
import torch

from oracle import gen as ogen
from oracle import synth
from oracle.make_golden import GEN_CASES
from tests.test_modules_cpu import _gen_module
from tests.test_modules_gpu import rel_l2

pytestmark = pytest.mark.gpu
BF16_REL = 2e-2


def _oracle(sd, content, style, noise, R, emulate):
    p = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    c = torch.from_numpy(content).requires_grad_()
    s = torch.from_numpy(style).requires_grad_()
    img = ogen.generator_forward(p, c, s, [torch.from_numpy(z) for z in noise], emulate_bf16=emulate)
    (img * R).sum().backward()
    g = {k: v.grad for k, v in p.items() if v.requires_grad and v.grad is not None}
    g["<content>"], g["<style>"] = c.grad, s.grad
    return img.detach(), g


@pytest.mark.parametrize("name", ["small", "odd_T"])
def test_generator_backward_matches_oracle(name):
    from handwriting_line_generation_b200 import _lib
    T, B, _, wseed, iseed = GEN_CASES[name]
    m, sd = _gen_module(wseed)
    sd = {k: v.clone() for k, v in sd.items()}
    m = m.cuda().train()
    content, style = synth.gen_case(T, B, 80, 128, iseed, True)      # dense content, so that it has a gradient
    noise = synth.gen_noise(synth.gen_noise_shapes(T, B), iseed + 7)
    R = torch.randn(B, 1, 64, 4 * T, generator=torch.Generator().manual_seed(2))
    c = torch.from_numpy(content).cuda().requires_grad_()
    s = torch.from_numpy(style).cuda().requires_grad_()
    n0 = _lib.launch_count()
    img = m(c, s, noise=[torch.from_numpy(z).cuda() for z in noise])
    (img * R.cuda()).sum().backward()
    torch.cuda.synchronize()
    assert _lib.launch_count() - n0 >= 90, "the training step did not run on the CUDA extension"
    img32, g32 = _oracle(sd, content, style, noise, R, False)
    _, gemu = _oracle(sd, content, style, noise, R, True)
    assert rel_l2(img.detach().cpu().numpy(), img32.numpy()) <= BF16_REL
    got = {n: p.grad.cpu() for n, p in m.named_parameters() if not n.startswith("gen.")}
    got["<content>"], got["<style>"] = c.grad.cpu(), s.grad.cpu()
    assert set(got) == set(g32)
    for n, g in g32.items():
        if g.numel() == 1:
            continue       # out.0.conv.bias: one heavily cancelling sum; held to an absolute bound below
        ours, emu = rel_l2(got[n].numpy(), g.numpy()), rel_l2(gemu[n].numpy(), g.numpy())
        cos = float((got[n].double() * g.double()).sum() / (got[n].double().norm() * g.double().norm()))
        # tensors with a handful of entries (per-channel bias sums with heavy cancellation) fluctuate more
        k = 1.6 if g.numel() >= 256 else 2.5
        assert ours <= k * emu + BF16_REL, f"{n}: cuda-vs-fp32 {ours:.3f}, bf16-emulated-torch-vs-fp32 {emu:.3f}"
        # (16-entry bias sums: 0.9 — B200, round 2: conv.4.conv1.0.bias at 0.9499 in one of five suite runs)
        assert cos >= (0.95 if g.numel() >= 256 else 0.9), f"{n}: cosine {cos:.3f}"
    for n in ("out.0.conv.weight_orig", "conv.4.adain2.style.weight"):
        assert rel_l2(got[n].numpy(), g32[n].numpy()) <= 3e-2, n
    # out.0.conv.bias is ONE number: a sum over every pixel of +/- terms; compare against the size of the terms
    scale = float((R.abs() * (1 - img32 ** 2)).sum())
    assert abs(got["out.0.conv.bias"].item() - g32["out.0.conv.bias"].item()) <= 1e-3 * scale


def test_generator_backward_with_inkernel_noise_runs():
    """Default (in-kernel) noise: backward runs, every parameter gets a finite gradient, and the same seed gives
    the same noise-weight gradients up to the statistics' summation order."""
    m, _ = _gen_module(100)
    m = m.cuda().train()
    content, style = synth.gen_case(16, 2, 80, 128, 3, True)
    c, s = torch.from_numpy(content).cuda(), torch.from_numpy(style).cuda()
    R = torch.randn(2, 1, 64, 64, generator=torch.Generator().manual_seed(5)).cuda()
    grads = []
    for _ in range(2):
        m.zero_grad()
        torch.manual_seed(11)
        (m(c, s) * R).sum().backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
        grads.append(torch.cat([b.noise2.weight_orig.grad.flatten() for b in m.conv]).cpu().numpy())
    assert rel_l2(grads[1], grads[0]) <= 0.2


def test_gen_lesson_chain_generator_hwr_ctc():
    """The 'gen' lesson's recognition branch (trainer/hw_with_style_trainer.py:760-764): generated lines are read by
    the (frozen) recognizer, the CTC loss flows back through the recognizer's input into the generator.
    22 bf16 layers deep, so only direction is asserted: cosine with the fp32 oracle's gradient and descent."""
    from handwriting_line_generation_b200 import CTCLoss, _lib
    from oracle import hwr as ohwr
    from tests.test_modules_cpu import _hwr_module
    T, B, S = 32, 2, 5
    g, gsd = _gen_module(100)
    h, hsd = _hwr_module(200)
    gsd = {k: v.clone() for k, v in gsd.items()}
    hsd = {k: v.clone() for k, v in hsd.items()}
    g, h = g.cuda().train(), h.cuda().train()
    for p in h.parameters():
        p.requires_grad_(False)                        # hwr_frozen: not optimised, no wgrad needed
    content, style = synth.gen_case(T, B, 80, 128, 9)
    noise = synth.gen_noise(synth.gen_noise_shapes(T, B), 10)
    tg = np.random.RandomState(1).randint(1, 80, (B, S)).astype(np.int32)
    Tc = 4 * T // 4 - 6
    il, tl = np.full(B, Tc, np.int32), np.full(B, S, np.int32)
    n0 = _lib.launch_count()
    img = g(torch.from_numpy(content).cuda(), torch.from_numpy(style).cuda(), noise=[torch.from_numpy(z).cuda() for z in noise])
    loss = CTCLoss(h(img), torch.from_numpy(tg).cuda(), torch.from_numpy(il), torch.from_numpy(tl))
    loss.backward()
    torch.cuda.synchronize()
    assert all(p.grad is None for p in h.parameters())
    # oracle chain
    gp = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in gsd.items()}
    oimg = ogen.generator_forward(gp, torch.from_numpy(content), torch.from_numpy(style), [torch.from_numpy(z) for z in noise])
    lp = ohwr.hwr_forward(hsd, oimg, True, None)
    ol = torch.nn.functional.ctc_loss(lp, torch.from_numpy(tg), torch.from_numpy(il), torch.from_numpy(tl))
    ol.backward()
    assert abs(loss.item() - ol.item()) <= 5e-2 * abs(ol.item())
    num = den1 = den2 = 0.0
    for n, p in g.named_parameters():
        if n.startswith("gen.") or gp[n].grad is None:
            continue
        a, b = p.grad.cpu().double(), gp[n].grad.double()
        num += float((a * b).sum()); den1 += float((a * a).sum()); den2 += float((b * b).sum())
    cos = num / (den1 * den2) ** 0.5
    assert cos >= 0.6, cos
    with torch.no_grad():
        step = 0.02 * ol.item() / den2
        sd2 = {k: (v - step * dict(g.named_parameters())[k].grad.cpu() if k in dict(g.named_parameters()) and dict(g.named_parameters())[k].grad is not None else v)
               for k, v in gsd.items()}
        oimg2 = ogen.generator_forward(sd2, torch.from_numpy(content), torch.from_numpy(style), [torch.from_numpy(z) for z in noise])
        l2 = torch.nn.functional.ctc_loss(ohwr.hwr_forward(hsd, oimg2, True, None), torch.from_numpy(tg), torch.from_numpy(il), torch.from_numpy(tl))
    assert l2.item() < ol.item(), (l2.item(), ol.item(), cos)
