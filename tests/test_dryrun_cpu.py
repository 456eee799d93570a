"""CPU: the HOST code of the whole default train step runs end to end with the C-ABI replaced by a recorder (no arithmetic):
every Python-level name, attribute, shape and autograd plumbing of the GPU-only paths is exercised in the build container,
where the kernels themselves cannot run.  Values are garbage by construction — numerics are the GPU parity tests' job
(and, for host-side composition, tests/test_encoder2_cpu.py with the arithmetic interpreter)."""
import contextlib

import numpy as np
import pytest
import torch

from oracle import synth


class _Stream:
    device = torch.device("cpu")
    cuda_stream = 0

    def wait_stream(self, other):
        pass


@pytest.fixture
def recorder(monkeypatch, hwg_lib):
    from handwriting_line_generation_b200 import _lib, weightmap
    calls = []
    monkeypatch.setattr(_lib, "call", lambda name, *a: calls.append(name))
    monkeypatch.setattr(_lib, "require_cuda", lambda *t: None)
    monkeypatch.setattr(weightmap.JobTable, "run", lambda self, src_base=None, dst_base=None: calls.append("hwg_linear_map"))
    main = _Stream()
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: main)
    monkeypatch.setattr(torch.cuda, "Stream", lambda *a, **k: _Stream())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    return calls


def _modules():
    import handwriting_line_generation_b200 as pkg
    torch.manual_seed(0)
    gen = pkg.SpacedGenerator(80, 128, 256, n_style_trans=6, emb_dropout=False, append_style=True, small=False).train()
    hwr = pkg.CNNOnlyHWR(80, norm='batch').train()
    disc = pkg.DiscriminatorAP(64, use_low=True, use_med=True).train()
    return pkg, gen, hwr, disc


def test_gen_lesson_step_host_code_runs(recorder):
    """bench_gan_train.train(): generator -> {recognizer -> CTC, discriminator -> adversarial loss} -> backward ->
    FlatAdam.step, with the gradients written straight into the flat buffer."""
    pkg, gen, hwr, disc = _modules()
    for p in list(hwr.parameters()) + list(disc.parameters()):
        p.requires_grad_(False)
    opt = pkg.FlatAdam(gen.parameters(), lr=2e-4, betas=(0.5, 0.999), clip_value=2.0)
    gen._grad_sink = opt
    T, B, S = 32, 2, 5
    content, style = synth.gen_case(T, B, 80, 128, 9)
    tg = torch.randint(1, 80, (B, S), dtype=torch.int32)
    il, tl = torch.full((B,), T - 6, dtype=torch.int32), torch.full((B,), S, dtype=torch.int32)
    for _ in range(2):
        img = gen(torch.from_numpy(content), torch.from_numpy(style))
        assert tuple(img.shape) == (B, 1, 64, 4 * T)
        lp = hwr(img)
        assert tuple(lp.shape) == (T - 6, B, 80)
        preds = disc(img)
        loss = 1e-4 * pkg.CTCLoss(lp, tg, il, tl) - sum(p.mean() for p in preds) / len(preds)
        loss.backward()
        opt.step()
    need = {"hwg_conv_fprop", "hwg_conv_wgrad", "hwg_ctc_forward", "hwg_ctc_backward", "hwg_adam_flat", "hwg_linear_map",
            "hwg_spectral_norm", "hwg_stem_conv", "hwg_shift_collapse", "hwg_adain_bwd_apply", "hwg_bn_bwd_apply"}
    assert need <= set(recorder), need - set(recorder)


def test_disc_lesson_and_recognizer_training_host_code_runs(recorder):
    """The 'disc' lesson (real || generated lines, hinge loss, all 28 parameter gradients), recognizer training
    (configs[0]) and the eval-mode forwards."""
    pkg, gen, hwr, disc = _modules()
    T, B = 32, 2
    content, style = synth.gen_case(T, B, 80, 128, 9)
    with torch.no_grad():
        fake = gen(torch.from_numpy(content), torch.from_numpy(style))
    real = torch.from_numpy(synth.hwr_case(B, 4 * T, 3))
    preds = disc(torch.cat((real, fake), 0))
    loss = sum(torch.relu(1.0 - p[:B]).mean() + torch.relu(1.0 + p[B:]).mean() for p in preds)
    loss.backward()
    assert all(p.grad is not None for n, p in disc.named_parameters() if not n.endswith(("weight_u", "weight_v")))
    lp = hwr(real)
    S = 5
    tg = torch.randint(1, 80, (B, S), dtype=torch.int32)
    pkg.CTCLoss(lp, tg, torch.full((B,), T - 6, dtype=torch.int32), torch.full((B,), S, dtype=torch.int32)).backward()
    assert all(p.grad is not None for p in hwr.parameters())
    gen.eval(), hwr.eval(), disc.eval()
    with torch.no_grad():
        img = gen(torch.from_numpy(content), torch.from_numpy(style))
        hwr(img)
        disc(img)
    raw, dec, dl = pkg.ctc_greedy_decode(lp.detach())
    assert tuple(raw.shape) == (T - 6, B)
    assert {"hwg_spectral_norm_bwd", "hwg_channel_sum", "hwg_hwr_stem", "hwg_ctc_greedy_decode"} <= set(recorder)


def test_bench_headline_step_host_code_runs(recorder):
    """bench_gan_train.GanStep (the default bench step, also used by tools/step_runner.py): both step kinds, with the
    discriminator's forward on its side stream."""
    import bench_gan_train as bg
    import handwriting_line_generation_b200 as pkg
    T, B, S = 32, 2, bg.GAN["S"]
    content, style = (torch.from_numpy(a) for a in synth.gen_case(T, B, 80, 128, 9))
    real = torch.from_numpy(synth.hwr_case(B, 4 * T, 3))
    tg = torch.randint(1, 80, (B, S), dtype=torch.int32)
    try:
        for kind in ("balanced", "gen_only"):
            del recorder[:]
            st = bg.GanStep(torch.device("cpu"), B, kind=kind, Ts=T)
            for _ in range(2):
                loss = st.train(content, style, tg, real)
                assert loss.dim() == 0
            assert recorder.count("hwg_adam_flat") == 2
            assert recorder.count("hwg_balance") == (2 if kind == "balanced" else 0)
            gf = st.conv_gflop(bg.gen_layers(T)[0])
            assert gf["conv_fprop_kernel"]["gflop"] > 0 and gf["conv_wgrad_kernel"]["gflop"] > 0
    finally:
        pkg.set_retain_graph(False)


def test_balanced_two_lesson_step_host_code_runs(recorder):
    """The balanced step spelled out (bench_gan_train.GanStep.train_balanced): two losses back-propagated separately through ONE
    generator graph and stashed, the perceptual lesson, FlatAdam.balance, FlatAdam.step."""
    pkg, gen, hwr, disc = _modules()
    enc = pkg.Encoder2(32).train()
    for p in list(hwr.parameters()) + list(disc.parameters()):
        p.requires_grad_(False)
    opt = pkg.FlatAdam(gen.parameters(), lr=2e-4, betas=(0.5, 0.999), clip_value=2.0)
    gen._grad_sink = opt
    T, B, S = 32, 2, 5
    content, style = (torch.from_numpy(a) for a in synth.gen_case(T, B, 80, 128, 9))
    real = torch.from_numpy(synth.hwr_case(B, 4 * T, 3))
    tg = torch.randint(1, 80, (B, S), dtype=torch.int32)
    il, tl = torch.full((B,), T - 6, dtype=torch.int32), torch.full((B,), S, dtype=torch.int32)
    pkg.set_retain_graph(True)
    try:
        for _ in range(2):
            img = gen(content, style)
            preds = disc(img)
            adv = -sum(p.mean() for p in preds) / len(preds)
            recog = 1e-4 * pkg.CTCLoss(hwr(img), tg, il, tl)
            recog.backward(retain_graph=True)
            opt.stash()
            adv.backward()
            opt.stash()
            assert len(opt._stash) == 2
            (0.5 * enc.perceptual_loss(real, gen(content, style))).backward()
            opt.balance([0.6, 0.5])
            assert opt._stash == []
            opt.step()
    finally:
        pkg.set_retain_graph(False)
    assert recorder.count("hwg_balance") == 2 and recorder.count("hwg_adam_flat") == 2
    assert {"hwg_l1_halves", "hwg_add_stats"} <= set(recorder)


def test_module_outputs_do_not_keep_their_graph_alive(recorder):
    """An output tensor that is also held by the Function's saved state closes a cycle output -> grad_fn -> ctx -> output
    that Python cannot collect (it runs through the C++ node): with set_retain_graph(True), or whenever no backward runs,
    every forward leaked its activations and left stale AccumulateGrad nodes behind (which is what broke the CUDA-graph
    capture of the balanced step).  The outputs are detached aliases; the graph dies with the last reference."""
    import gc
    import weakref
    pkg, gen, hwr, disc = _modules()
    pkg.set_retain_graph(True)
    try:
        T, B = 32, 2
        content, style = (torch.from_numpy(a) for a in synth.gen_case(T, B, 80, 128, 9))
        img = gen(content, style)
        lp = hwr(img)
        preds = disc(img)
        refs = [weakref.ref(t.grad_fn) for t in (img, lp, preds[0])]
        assert all(r() is not None for r in refs)
        del img, lp, preds
        gc.collect()
        assert all(r() is None for r in refs), "a module output keeps its own autograd graph alive"
    finally:
        pkg.set_retain_graph(False)
