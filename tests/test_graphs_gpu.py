"""GPU: CUDA-graph capture of the generator forward and of a full recognizer train step (fwd + CTC + bwd + Adam)."""
import numpy as np
import pytest
import torch

from oracle import synth
from tests.test_modules_cpu import _gen_module, _hwr_module
from tests.test_modules_gpu import rel_l2

pytestmark = pytest.mark.gpu


def test_graphed_generator_forward_draws_fresh_noise_each_replay():
    from handwriting_line_generation_b200 import graphs, _lib
    m, _ = _gen_module(100)
    m = m.cuda().eval()
    content, style = synth.gen_case(32, 4, 80, 128, 5)
    c, s = torch.from_numpy(content).cuda(), torch.from_numpy(style).cuda()

    def fwd(c_, s_):
        with torch.no_grad():
            return m(c_, s_)

    g = graphs.GraphedStep(fwd, [c, s], modules=[m])
    n0 = _lib.launch_count()
    a = g(c, s).clone()
    b = g(c, s).clone()
    torch.cuda.synchronize()
    assert _lib.launch_count() == n0, "replays must not go through the host launch path"
    assert a.shape == (4, 1, 64, 128) and torch.isfinite(a).all()
    assert rel_l2(b.cpu().numpy(), a.cpu().numpy()) > 2e-2, "each replay must draw new noise"
    # new inputs are picked up
    content2, style2 = synth.gen_case(32, 4, 80, 128, 6)
    d = g(torch.from_numpy(content2).cuda(), torch.from_numpy(style2).cuda()).clone()
    assert rel_l2(d.cpu().numpy(), a.cpu().numpy()) > 0.2


def test_graphed_hwr_train_step_matches_eager():
    from handwriting_line_generation_b200 import CTCLoss, graphs
    B, W, S = 2, 128, 6
    T = W // 4 - 6
    img = torch.from_numpy(synth.hwr_case(B, W, 31)).cuda()
    tg = torch.from_numpy(np.random.RandomState(5).randint(1, 80, (B, S)).astype(np.int32)).cuda()
    il = torch.full((B,), T, dtype=torch.int32, device="cuda")
    tl = torch.full((B,), S, dtype=torch.int32, device="cuda")

    def make():
        m, _ = _hwr_module(200)
        m = m.cuda().train()
        opt = torch.optim.Adam(m.parameters(), lr=1e-3, capturable=True)
        return m, opt

    def step_fn(m, opt):
        def step(x, t):
            loss = CTCLoss(m(x), t, il, tl)
            loss.backward()
            opt.step()
            opt.zero_grad(set_to_none=False)
            return loss
        return step

    # eager: warmup(3) + capture-equivalent(1) + 3 = 7 steps
    m1, o1 = make()
    s1 = step_fn(m1, o1)
    losses_eager = [s1(img, tg).item() for _ in range(7)]
    m2, o2 = make()
    o2.zero_grad(set_to_none=True)
    g = graphs.GraphedStep(step_fn(m2, o2), [img, tg], modules=[m2], warmup=3)   # 3 warmup + 1 captured (not run)
    losses_graph = [g(img, tg).item() for _ in range(4)]
    # capture itself does not execute, so replay i corresponds to eager step 3+i
    assert np.allclose(losses_graph, losses_eager[3:7], rtol=1e-1), (losses_graph, losses_eager)
    assert losses_graph[-1] < losses_graph[0]     # it trains
    # both runs moved the weights the same way (Adam turns the bf16/atomics noise of tiny gradient entries into
    # full-size +-lr steps, so the two trajectories agree in direction, not digit for digit)
    w0 = _hwr_module(200)[0].cnn.conv3.weight.detach()
    d1 = (m1.cnn.conv3.weight.detach().cpu() - w0).flatten().double()
    d2 = (m2.cnn.conv3.weight.detach().cpu() - w0).flatten().double()
    assert float((d1 * d2).sum() / (d1.norm() * d2.norm())) > 0.5
