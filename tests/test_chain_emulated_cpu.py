"""CPU: the 'gen' lesson the way the reference trainer runs it with `balance_loss` (trainer :300-338) — ONE generator
forward, the adversarial and the recognition loss back-propagated one after the other through the same graph
(`retain_graph=True`), each gradient set stashed — through the CPU interpreter of the C-ABI, against the oracle chain
(oracle/gen.py -> oracle/disc.py / oracle/hwr.py, pinned to the reference trainer by tests/test_trainer_gen_cpu.py).
Checks that a second backward over the retained state gives the gradient of ITS loss (not a stale or doubled one) and that
FlatAdam.stash() separates the sets."""
import numpy as np
import torch

from oracle import disc as odisc
from oracle import gen as ogen
from oracle import hwr as ohwr
from oracle import synth

from . import abi_emu
from .test_modules_cpu import _gen_module, _hwr_module


def _cos(a, b):
    num = sum(float((x.double() * y.double()).sum()) for x, y in zip(a, b))
    return num / (sum(float((x.double() ** 2).sum()) for x in a) * sum(float((y.double() ** 2).sum()) for y in b)) ** 0.5


def test_two_losses_through_one_retained_generator_graph(hwg_lib, monkeypatch):
    import handwriting_line_generation_b200 as pkg
    T, B, S = 32, 2, 5
    g, gsd = _gen_module(100)
    h, hsd = _hwr_module(200)
    torch.manual_seed(300)
    d = pkg.DiscriminatorAP(64, use_low=True, use_med=True)
    dsd = {k: v.clone() for k, v in synth.perturb_disc(d.state_dict(), 301).items()}
    gsd = {k: v.clone() for k, v in gsd.items()}
    hsd = {k: v.clone() for k, v in hsd.items()}
    g.train(), h.train(), d.train()
    for p in list(h.parameters()) + list(d.parameters()):
        p.requires_grad_(False)
    content, style = synth.gen_case(T, B, 80, 128, 9)
    noise = [torch.from_numpy(z) for z in synth.gen_noise(synth.gen_noise_shapes(T, B), 10)]
    masks = {k: torch.from_numpy(v) for k, v in synth.disc_masks(B, 11).items()}
    d.dropout_masks = masks
    tg = torch.from_numpy(np.random.RandomState(1).randint(1, 80, (B, S)).astype(np.int32))
    il, tl = torch.full((B,), T - 6, dtype=torch.int32), torch.full((B,), S, dtype=torch.int32)
    names = [n for n, _ in g.named_parameters()]
    with abi_emu.installed(monkeypatch):
        params = list(g.parameters())
        opt = pkg.FlatAdam(params, lr=2e-4, betas=(0.5, 0.999), clip_value=2.0)
        g._grad_sink = opt
        pkg.set_retain_graph(True)
        try:
            img = g(torch.from_numpy(content), torch.from_numpy(style), noise=noise)
            adv = odisc.gen_loss(d(img))
            recog = torch.nn.functional.ctc_loss(h(img), tg, il, tl)
            adv.backward(retain_graph=True)
            opt.stash()
            recog.backward()
            opt.stash()
        finally:
            pkg.set_retain_graph(False)
        sets = [{n: s[opt.offsets[id(p)][0]:opt.offsets[id(p)][0] + p.numel()].view_as(p).clone()
                 for n, p in zip(names, params)} for s in opt._stash]
        assert float(opt.flat_g.abs().max()) == 0.0
    # oracle chain, one loss at a time
    ref = []
    for which in ("adv", "recog"):
        gp = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in gsd.items()}
        oimg = ogen.generator_forward(gp, torch.from_numpy(content), torch.from_numpy(style), noise)
        if which == "adv":
            loss = odisc.gen_loss(odisc.disc_forward(dsd, oimg, masks, training=True))
        else:
            loss = torch.nn.functional.ctc_loss(ohwr.hwr_forward(hsd, oimg, True, None), tg, il, tl)
        loss.backward()
        ref.append(({k: v.grad for k, v in gp.items() if v.requires_grad and v.grad is not None}, loss.item()))
    assert abs(adv.item() - ref[0][1]) <= 2e-2 * abs(ref[0][1]) + 2e-3
    assert abs(recog.item() - ref[1][1]) <= 5e-2 * abs(ref[1][1])
    keys = [k for k in ref[0][0] if k in sets[0]]
    assert len(keys) >= 60
    own = [_cos([sets[i][k] for k in keys], [ref[i][0][k] for k in keys]) for i in range(2)]
    cross = _cos([sets[1][k] for k in keys], [ref[0][0][k] for k in keys])
    # 22-34 bf16 layers deep: direction, as tests/test_gen_train_gpu.py asserts for the chain on the real kernels
    assert min(own) >= 0.6, own
    assert abs(cross) < min(own) - 0.2, (own, cross)         # the second set is the recognition gradient, not a stale copy
