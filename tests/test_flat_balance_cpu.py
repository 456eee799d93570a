"""CPU: the HOST side of `FlatAdam.stash()` / `FlatAdam.balance()` (trainer/hw_with_style_trainer.py:300-377):
the stash bookkeeping, the segment / block tables and the pointer array handed to `hwg_balance`, run through the CPU
interpreter of the C-ABI (tests/abi_emu.py, which also checks that the block table tiles every segment) and compared with
oracle/balance.py — the restatement tests/test_balance_cpu.py pins to the unmodified reference trainer.
tests/test_balance_gpu.py is the same scenario on the real launches."""
import torch

from oracle import balance as obal

from . import abi_emu


def test_flat_stash_and_balance_match_the_oracle(hwg_lib, monkeypatch):
    import handwriting_line_generation_b200 as pkg
    g0 = torch.Generator().manual_seed(0)
    shapes = [(64, 32, 3, 3), (64,), (5000,), (1,), (3, 7), (128, 128, 3, 3), (17,)]
    with abi_emu.installed(monkeypatch) as calls:
        params = [torch.nn.Parameter(torch.randn(s, generator=g0)) for s in shapes]
        opt = pkg.FlatAdam(params, lr=1e-3)
        K, mult = 4, [0.6, 0.5, 0.4, 0.75]
        sets_cpu = []
        for k in range(K):
            grads = [torch.randn(s, generator=g0) * (0.1 + k) for s in shapes]
            if k == 1:
                grads[2] = torch.zeros(shapes[2])            # a set without a gradient for this tensor: skipped (:373)
            sets_cpu.append(grads)
            for p, g in zip(params, grads):
                opt.grad_view(p).copy_(g)
            opt.stash()
            assert float(opt.flat_g.abs().max()) == 0.0
        main_cpu = [torch.randn(s, generator=g0) * 0.01 for s in shapes]
        main_cpu[3] = torch.zeros(shapes[3])                 # mean|D| == 0: takes the fill value (:354-359)
        for p, g in zip(params, main_cpu):
            opt.grad_view(p).copy_(g)
        ref = obal.balance([g.clone() for g in main_cpu], sets_cpu, mult)
        opt.balance(mult)
        assert opt._stash == [] and calls == ["hwg_balance"]
        for p, r in zip(params, ref):
            got = opt.grad_view(p)
            assert float((got - r).abs().max()) <= 2e-6 * float(r.abs().max()) + 1e-12
        opt.balance(mult)                                    # nothing stashed: a no-op, no launch
        assert calls == ["hwg_balance"]


def test_stash_slots_equal_the_stash_path(hwg_lib, monkeypatch):
    """FlatAdam.sink(k): backward passes that write their gradient set straight into stash slot k (bench step: one slot
    per loss, no clone + zero) balance to the same result as the stash() path, and the slots come back zeroed."""
    import handwriting_line_generation_b200 as pkg
    g0 = torch.Generator().manual_seed(2)
    shapes = [(32, 16, 3, 3), (32,), (700,), (5, 3)]
    with abi_emu.installed(monkeypatch) as calls:
        params = [torch.nn.Parameter(torch.randn(s, generator=g0)) for s in shapes]
        opt = pkg.FlatAdam(params, lr=1e-3)
        mult = [0.6, 0.5]
        sets_cpu = [[torch.randn(s, generator=g0) * (0.2 + k) for s in shapes] for k in range(2)]
        main_cpu = [torch.randn(s, generator=g0) * 0.05 for s in shapes]
        for rep in range(2):                                 # twice: the slots are re-zeroed by balance()
            for k in range(2):
                sink = opt.sink(k)
                assert sink is opt.sink(k) and sink.owns(params[0]) and float(sink.buf.abs().max()) == 0.0
                for p, g in zip(params, sets_cpu[k]):
                    sink.grad_view(p).add_(g)                # what a backward kernel does
            assert opt.sink() is opt
            for p, g in zip(params, main_cpu):
                opt.grad_view(p).copy_(g)
            ref = obal.balance([g.clone() for g in main_cpu], sets_cpu, mult)
            opt.balance(mult)
            for p, r in zip(params, ref):
                assert float((opt.grad_view(p) - r).abs().max()) <= 2e-6 * float(r.abs().max()) + 1e-12
            opt.flat_g.zero_()
        assert calls == ["hwg_balance", "hwg_balance"]


def test_flat_adam_step_equals_clip_plus_torch_adam(hwg_lib, monkeypatch):
    """FlatAdam's host side (flat slots, .data / .grad re-pointing, version bump) through the interpreter of
    hwg_adam_flat against `clip_grad_value_` + `torch.optim.Adam` (trainer :381-391), five steps."""
    import handwriting_line_generation_b200 as pkg
    g0 = torch.Generator().manual_seed(1)
    shapes = [(16, 8, 3, 3), (16,), (5,), (3, 7)]
    ref = [torch.nn.Parameter(torch.randn(s, generator=g0)) for s in shapes]
    ours = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    topt = torch.optim.Adam(ref, lr=2e-4, betas=(0.5, 0.999))
    with abi_emu.installed(monkeypatch) as calls:
        opt = pkg.FlatAdam(ours, lr=2e-4, betas=(0.5, 0.999), clip_value=2.0)
        for step in range(5):
            grads = [torch.randn(s, generator=g0) * (3.0 if step % 2 else 0.3) for s in shapes]     # some beyond the clip
            for p, q, g in zip(ref, ours, grads):
                p.grad = g.clone()
                opt.grad_view(q).copy_(g)
            torch.nn.utils.clip_grad_value_(ref, 2)
            topt.step()
            v0 = [q._version for q in ours]
            opt.step()
            assert all(q._version > v for q, v in zip(ours, v0))            # derived-weight caches see the update
            assert float(opt.flat_g.abs().max()) == 0.0                      # zero_grad is part of the launch
            for p, q in zip(ref, ours):
                assert torch.allclose(q, p, rtol=1e-5, atol=1e-7)
    assert calls.count("hwg_adam_flat") == 5
