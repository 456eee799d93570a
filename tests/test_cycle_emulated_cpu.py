"""CPU: bench_cycle.GanCycle — the reference curriculum's 7-lesson cycle with every module on the drop-ins (generator,
recognizer, CTC, discriminator, Encoder2, DTW, spacer + insert_spaces, style extractor) — at toy size through the CPU
interpreter of the C-ABI: every lesson runs, losses are finite, each optimizer group is stepped by the lessons that own it."""
import numpy as np
import torch

from . import abi_emu


def test_seven_lesson_cycle_host_code(hwg_lib, monkeypatch):
    import bench_cycle
    import handwriting_line_generation_b200 as pkg
    try:
        with abi_emu.installed(monkeypatch) as calls:
            np.random.seed(0)
            cyc = bench_cycle.GanCycle(torch.device("cpu"), B=2, a_batch=2, W=128, L=6)
            snap = lambda m: {n: p.detach().clone() for n, p in m.named_parameters()}        # noqa: E731
            changed = lambda m, s: sum(int(not torch.equal(p.detach(), s[n])) for n, p in m.named_parameters())   # noqa: E731
            g0, s0, sp0, d0 = snap(cyc.gen), snap(cyc.style), snap(cyc.spacer), snap(cyc.disc)
            loss = cyc.run_lesson("count")
            assert np.isfinite(float(loss)) and changed(cyc.spacer, sp0) >= 14 and changed(cyc.style, s0) >= 40
            assert changed(cyc.gen, g0) == 0 and changed(cyc.disc, d0) == 0
            g0, s0 = snap(cyc.gen), snap(cyc.style)
            assert np.isfinite(float(cyc.run_lesson("gen"))) and len(cyc.opt._stash) == 2 and changed(cyc.gen, g0) == 0
            assert np.isfinite(float(cyc.run_lesson("auto"))) and cyc.opt._stash == []
            assert changed(cyc.gen, g0) == 64 and changed(cyc.style, s0) >= 40
            d0 = snap(cyc.disc)
            assert np.isfinite(float(cyc.run_lesson("disc"))) and changed(cyc.disc, d0) >= 28
        assert {"hwg_dtw_align", "hwg_insert_spaces_fill", "hwg_balance", "hwg_l1_halves", "hwg_shift_expand",
                "hwg_ctc_backward", "hwg_spectral_norm_bwd"} <= set(calls)
    finally:
        pkg.set_retain_graph(False)
