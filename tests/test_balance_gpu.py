"""FlatAdam.stash() / FlatAdam.balance() (hwg_balance: three launches over the flat gradient buffers) against
oracle/balance.py, the restatement of trainer/hw_with_style_trainer.py:340-377 that tests/test_balance_cpu.py pins to the
unmodified reference trainer."""
import os
import sys

import pytest
import torch

from oracle import balance as obal   # noqa: E402

pytestmark = pytest.mark.gpu


def test_flat_balance_matches_the_oracle():
    import handwriting_line_generation_b200 as pkg
    g0 = torch.Generator().manual_seed(0)
    shapes = [(64, 32, 3, 3), (64,), (5000,), (1,), (3, 7), (128, 128, 3, 3), (17,)]
    params = [torch.nn.Parameter(torch.randn(s, generator=g0).cuda()) for s in shapes]
    opt = pkg.FlatAdam(params, lr=1e-3)
    K, mult = 4, [0.6, 0.5, 0.4, 0.75]
    sets_cpu = []
    for k in range(K):
        grads = [torch.randn(s, generator=g0) * (0.1 + k) for s in shapes]
        if k == 1:
            grads[2] = torch.zeros(shapes[2])            # a set without a gradient for this tensor: skipped (:373)
        sets_cpu.append(grads)
        for p, g in zip(params, grads):
            opt.grad_view(p).copy_(g.cuda())
        opt.stash()
        assert float(opt.flat_g.abs().max()) == 0.0
    main_cpu = [torch.randn(s, generator=g0) * 0.01 for s in shapes]
    main_cpu[3] = torch.zeros(shapes[3])                 # mean|D| == 0: takes the fill value (:354-359)
    for p, g in zip(params, main_cpu):
        opt.grad_view(p).copy_(g.cuda())
    ref = obal.balance([g.clone() for g in main_cpu], sets_cpu, mult)
    opt.balance(mult)
    torch.cuda.synchronize()
    assert opt._stash == []
    for p, r in zip(params, ref):
        got = opt.grad_view(p).cpu()
        assert float((got - r).abs().max()) <= 2e-5 * float(r.abs().max()) + 1e-9      # fp32: atomics vs pairwise means
