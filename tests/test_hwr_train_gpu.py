"""GPU parity of the recognizer's training step (forward + CTC loss + backward through every layer) against
torch autograd on the CPU oracle (oracle/hwr.py) with the same weights and inputs.

What can be asserted.  The per-kernel backward tests (test_conv_bwd_gpu.py, test_bwd_ops_gpu.py) hold every
backward kernel to <= 1e-2 against autograd on identical inputs.  End to end, the gradient of this network is
discontinuous in its activations (ReLU masks, max-pool arg-maxes): a forward that differs from the fp32 one by
bf16 rounding (~0.5 % rms) flips ~0.4 % of the masks per layer, and every flip moves a whole gradient entry, so
the parameter gradients of ANY bf16-activation implementation sit 6 % (head) to ~45 % (stem) in rel-L2 from the
fp32 gradient — torch itself shows exactly that when the oracle is run with bf16 storage emulation
(oracle/hwr.py emulate_bf16=True; even a 1e-4 input perturbation in pure fp32 moves the stem gradient by 10 %).
The end-to-end assertions are therefore:
  * loss (a smooth function of the forward): rel 2e-2 (bf16 path);
  * every gradient tensor: rel-L2(cuda, fp32) <= 1.3 * rel-L2(bf16-emulated torch, fp32) + 2e-2, i.e. the CUDA
    path is as close to the fp32 reference as a plain bf16 emulation of the reference is;
  * cosine(cuda, fp32) >= 0.85 on every weight tensor, and the fp32 oracle's loss decreases along -grad_cuda.
"""
import numpy as np
import pytest
import torch

from oracle import hwr as ohwr
from oracle import synth
from tests.test_modules_cpu import _hwr_module
from tests.test_modules_gpu import rel_l2

pytestmark = pytest.mark.gpu
BF16_REL = 2e-2
ZERO_GRAD = {"cnn.conv2.bias", "cnn.conv4.bias", "cnn.conv6.bias", "cnn1d.0.bias", "cnn1d.3.bias", "cnn1d.6.bias",
             "cnn1d.9.bias"}  # bias of a conv that feeds BatchNorm: true gradient is identically zero


def _oracle(sd, img, tg, il, tl, emulate):
    p = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    lp = ohwr.hwr_forward(p, torch.from_numpy(img), True, None, emulate_bf16=emulate)
    loss = torch.nn.functional.ctc_loss(lp, torch.from_numpy(tg), torch.from_numpy(il), torch.from_numpy(tl))
    loss.backward()
    return loss.item(), {k: v.grad for k, v in p.items() if v.requires_grad and v.grad is not None}


def _oracle_loss(sd, img, tg, il, tl):
    with torch.no_grad():
        lp = ohwr.hwr_forward(sd, torch.from_numpy(img), True, None)
        return torch.nn.functional.ctc_loss(lp, torch.from_numpy(tg), torch.from_numpy(il), torch.from_numpy(tl)).item()


@pytest.mark.parametrize("B,W,S", [(2, 128, 6), (3, 260, 12)])
def test_hwr_ctc_train_step_matches_oracle(B, W, S):
    from handwriting_line_generation_b200 import CTCLoss, _lib
    m, sd = _hwr_module(200)
    sd = {k: v.clone() for k, v in sd.items()}
    m = m.cuda().train()
    img = synth.hwr_case(B, W, 31)
    T = W // 4 - 6
    r = np.random.RandomState(5)
    tg = r.randint(1, 80, (B, S)).astype(np.int32)
    il, tl = np.full(B, T, np.int32), np.full(B, S, np.int32)
    n0 = _lib.launch_count()
    lp = m(torch.from_numpy(img).cuda())
    loss = CTCLoss(lp, torch.from_numpy(tg).cuda(), torch.from_numpy(il), torch.from_numpy(tl))
    loss.backward()
    torch.cuda.synchronize()
    assert _lib.launch_count() - n0 >= 60, "the training step did not run on the CUDA extension"
    loss32, g32 = _oracle(sd, img, tg, il, tl, False)
    _, gemu = _oracle(sd, img, tg, il, tl, True)
    assert abs(loss.item() - loss32) <= BF16_REL * abs(loss32)
    got = {n: p.grad.cpu() for n, p in m.named_parameters()}
    assert set(got) == set(g32)
    report = {}
    for n, g in g32.items():
        if n in ZERO_GRAD:
            assert got[n].abs().max() <= 1e-2 * g32[n.replace("bias", "weight")].abs().max(), n
            continue
        ours, emu = rel_l2(got[n].numpy(), g.numpy()), rel_l2(gemu[n].numpy(), g.numpy())
        cos = float((got[n].double() * g.double()).sum() / (got[n].double().norm() * g.double().norm()))
        report[n] = (ours, emu, cos)
        assert ours <= 1.3 * emu + BF16_REL, f"{n}: cuda-vs-fp32 {ours:.3f}, bf16-emulated-torch-vs-fp32 {emu:.3f}"
        if n.endswith("weight"):
            assert cos >= 0.85, f"{n}: cosine {cos:.3f}"
    # the tensor next to the loss sees no mask flips, only the CTC occupancies' sensitivity to the ~0.5 % log-prob
    # error (observed 1.5-2.2 %, varying with the atomics' summation order)
    assert report["cnn1d.12.bias"][0] <= 5e-2
    # descent direction for the fp32 reference
    step = 0.05 * loss32 / sum(float((g.double() ** 2).sum()) for g in g32.values())
    sd2 = {k: (v - step * got[k] if k in got else v) for k, v in sd.items()}
    assert _oracle_loss(sd2, img, tg, il, tl) < loss32


def test_hwr_backward_without_input_grad_and_frozen_parameters():
    """Parameters with requires_grad=False get no .grad; asking for the image gradient fails loudly."""
    from handwriting_line_generation_b200 import _lib
    m, _ = _hwr_module(200)
    m = m.cuda().train()
    for n, p in m.named_parameters():
        if n.startswith("cnn.conv0"):
            p.requires_grad_(False)
    img = torch.from_numpy(synth.hwr_case(2, 128, 3)).cuda()
    m(img).sum().backward()
    assert m.cnn.conv0.weight.grad is None and m.cnn.conv1.weight.grad is not None


def test_hwr_input_gradient_for_gan_lessons():
    """Gradient w.r.t. the image (the generated line in the 'gen' lessons, trainer :760-764): the stem backward
    itself is checked on identical inputs; end to end it must be a descent direction for the fp32 oracle."""
    import torch.nn.functional as F
    from handwriting_line_generation_b200 import ops, conv
    # (1) stem only, exact: conv0 + ReLU + MaxPool backward to the image
    g0 = torch.Generator().manual_seed(3)
    img = (torch.rand(2, 1, 64, 96, generator=g0) * 2 - 1).double().requires_grad_()
    w = (torch.randn(64, 1, 3, 3, generator=g0) / 3).float().double()
    b = (torch.randn(64, generator=g0) * 0.1).float().double()
    a = F.max_pool2d(F.relu(F.conv2d(img, w, b, padding=1)), 2, 2)
    g = torch.randn(a.shape, generator=g0).to(torch.bfloat16).double()
    (gi_ref,) = torch.autograd.grad(a, img, g)
    gc0 = ops.hwr_stem_bwd_expand(img.detach().float().cuda(), w.float().reshape(64, 9).contiguous().cuda(),
                                  b.float().cuda(), g.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda())
    mats = [F.pad(w.float()[:, 0, i, j].view(1, 64), (0, 0, 0, 15)).cuda() for i in range(3) for j in range(3)]
    taps = [(1 - i, 1 - j) for i in range(3) for j in range(3)]
    gi = conv.conv_fprop(gc0, conv.pack_taps(mats), taps, 64, 96, out_dtype=torch.float32)[..., 0].cpu().double()
    assert ((gi - gi_ref[:, 0]).abs().max() / gi_ref.abs().max()).item() <= 1e-2
    # (2) whole recognizer + CTC: cosine with the fp32 oracle and descent
    from handwriting_line_generation_b200 import CTCLoss
    m, sd = _hwr_module(200)
    sd = {k: v.clone() for k, v in sd.items()}
    m = m.cuda().train()
    B, W, S = 2, 128, 6
    x = synth.hwr_case(B, W, 31)
    T = W // 4 - 6
    tg = np.random.RandomState(5).randint(1, 80, (B, S)).astype(np.int32)
    il, tl = np.full(B, T, np.int32), np.full(B, S, np.int32)
    xc = torch.from_numpy(x).cuda().requires_grad_()
    CTCLoss(m(xc), torch.from_numpy(tg).cuda(), torch.from_numpy(il), torch.from_numpy(tl)).backward()
    xo = torch.from_numpy(x).requires_grad_()
    lp = ohwr.hwr_forward(sd, xo, True, None)
    lo = torch.nn.functional.ctc_loss(lp, torch.from_numpy(tg), torch.from_numpy(il), torch.from_numpy(tl))
    lo.backward()
    gq, gr = xc.grad.cpu().double(), xo.grad.double()
    cos = float((gq * gr).sum() / (gq.norm() * gr.norm()))
    assert cos >= 0.8, cos
    with torch.no_grad():
        step = 0.05 * lo.item() / float((gr ** 2).sum())
        lp2 = ohwr.hwr_forward(sd, torch.from_numpy(x) - step * gq.float(), True, None)
        l2 = torch.nn.functional.ctc_loss(lp2, torch.from_numpy(tg), torch.from_numpy(il), torch.from_numpy(tl))
    assert l2.item() < lo.item()


def test_hwr_eval_mode_gradient_uses_running_statistics_as_constants():
    """A frozen `hwr.eval()` under autograd: BatchNorm's running statistics are constants of the backward (gz = sc*gy),
    against torch autograd on the fp32 oracle in eval mode."""
    from handwriting_line_generation_b200 import CTCLoss
    from tests.test_hwr_emulated_cpu import _eval_case
    m, sd = _hwr_module(200)
    sd = _eval_case(m, {k: v.clone() for k, v in sd.items()})
    m = m.cuda().eval()
    B, W, S = 2, 128, 6
    img = synth.hwr_case(B, W, 31)
    T = W // 4 - 6
    tg = np.random.RandomState(5).randint(1, 80, (B, S)).astype(np.int32)
    il, tl = np.full(B, T, np.int32), np.full(B, S, np.int32)
    xc = torch.from_numpy(img).cuda().requires_grad_()
    lp = m(xc)
    CTCLoss(lp, torch.from_numpy(tg).cuda(), torch.from_numpy(il), torch.from_numpy(tl)).backward()
    p = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    xo = torch.from_numpy(img).requires_grad_()
    lpo = ohwr.hwr_forward(p, xo, False, None)
    torch.nn.functional.ctc_loss(lpo, torch.from_numpy(tg), torch.from_numpy(il), torch.from_numpy(tl)).backward()
    assert rel_l2(lp.detach().cpu().numpy(), lpo.detach().numpy()) <= BF16_REL
    gq, gr = xc.grad.cpu().double(), xo.grad.double()
    assert float((gq * gr).sum() / (gq.norm() * gr.norm())) >= 0.9
    for n in ("cnn1d.12.weight", "cnn1d.9.weight", "cnn.conv6.weight", "cnn.batchnorm6.weight"):
        a, b = dict(m.named_parameters())[n].grad.cpu().double(), p[n].grad.double()
        assert float((a * b).sum() / (a.norm() * b.norm())) >= 0.9, n
