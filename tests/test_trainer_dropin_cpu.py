"""Build-container only (needs /root/reference): the UNMODIFIED reference trainer — `HWWithStyleTrainer._train_iteration`,
'gen' lesson of the shipped IAM GAN curriculum with `balance_loss` — run with the DROP-IN modules swapped in by
`integrate.install(retain_graph=True)`, on CPU through the interpreter of the C-ABI (tests/abi_emu.py), and compared with
the golden of the same lesson run with the reference's own classes (tests/golden/trainer_gen.npz).

This is the drop-in claim at the level a user of the reference meets it: no call site changes, the trainer's own
`self.model(label, label_lengths, style)` / `self.model.hwr(gen_image)` / `CTCLoss` / `self.model.discriminator(fake)`
calls, its three `.backward(retain_graph=True)` passes over one graph and its gradient stashing (trainer :300-338)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from oracle import ref_shim, synth
from oracle.make_golden import digest

from . import abi_emu
from .test_trainer_gen_cpu import build_inputs

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="/root/reference is not present on this machine")


def _rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def _oracle_sets(gold):
    """The two per-loss gradient sets of the fp32 oracle chain (== the reference trainer's, tests/test_trainer_gen_cpu.py)."""
    from oracle import disc as odisc
    from oracle import gen as ogen
    from oracle import hwr as ohwr
    from .test_trainer_gen_cpu import W_GEN, W_RECOG
    gsd, hsd, dsd, content, style, noise, masks = build_inputs(gold)
    B = style.size(0)
    label = torch.from_numpy(gold["label"]).int()
    lengths = torch.from_numpy(gold["label_lengths"]).int()
    out = {}
    for which in ("recog", "adv"):
        gp = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in gsd.items()}
        img = ogen.generator_forward(gp, content, style, noise)
        if which == "recog":
            lp = ohwr.hwr_forward({k: v.clone() for k, v in hsd.items()}, img, True, {})
            loss = W_RECOG * torch.nn.functional.ctc_loss(lp, label.permute(1, 0), torch.IntTensor([lp.size(0)] * B), lengths)
        else:
            loss = W_GEN * odisc.gen_loss(odisc.disc_forward(dsd, img, masks, training=True))
        loss.backward()
        out[which] = {k: v.grad for k, v in gp.items() if v.requires_grad and v.grad is not None}
    return out


def test_reference_trainer_gen_lesson_runs_on_the_drop_ins(golden_dir, hwg_lib, monkeypatch):
    import importlib
    import sys

    import handwriting_line_generation_b200 as pkg
    from handwriting_line_generation_b200 import _lib, integrate
    from oracle import make_trainer_golden as harness
    gold = np.load(f"{golden_dir}/trainer_gen.npz")
    gsd, _, _, content, style, noise, masks = build_inputs(gold)
    saved_path, saved_ds = list(sys.path), sys.modules.get("datasets")
    state = {}

    def install():
        hws = importlib.import_module("model.hw_with_style")
        mloss = importlib.import_module("model.loss")
        mauto = importlib.import_module("model.autoencoder")
        mtr = importlib.import_module("trainer.hw_with_style_trainer")
        state["orig"] = (hws, mloss, hws.SpacedGenerator, hws.CNNOnlyHWR, hws.DiscriminatorAP, mloss.CTCLoss)
        state["enc"] = (mauto, mtr, mauto.Encoder2, mtr.Encoder2)
        integrate.install(retain_graph=True, encoder=True)

    def hook(tr, model, rec):
        assert isinstance(model.generator, pkg.SpacedGenerator) and isinstance(model.hwr, pkg.CNNOnlyHWR)
        assert isinstance(model.discriminator, pkg.DiscriminatorAP) and tr.balance_loss
        # the trainer built the perceptual encoder from the drop-in class and loaded the checkpoint into it (:136-158)
        assert isinstance(tr.encoder, pkg.Encoder2) and tr.encoder.out_dim == 32
        state["encoder_sum"] = float(sum(p.detach().double().abs().sum() for p in tr.encoder.parameters()))
        model.discriminator.dropout_masks = masks                      # the reference run patched F.dropout2d with these
        recorded = model.generator.forward                              # the harness's recorder around the module's forward
        model.generator.forward = lambda content, style, *a, **k: recorded(content, style, *a, noise=noise, **k)

    try:
        with abi_emu.installed(monkeypatch) as calls:
            tr, log, rec, model = harness.run_lesson("gen", install=install, hook=hook)
    finally:
        hws, mloss, g, h, d, c = state["orig"]
        hws.SpacedGenerator, hws.CNNOnlyHWR, hws.DiscriminatorAP, mloss.CTCLoss = g, h, d, c
        if "enc" in state:
            mauto, mtr, e1, e2 = state["enc"]
            mauto.Encoder2, mtr.Encoder2 = e1, e2
        _lib.RETAIN_SAVED = False
        sys.path[:] = saved_path
        if saved_ds is not None:
            sys.modules["datasets"] = saved_ds
        else:
            sys.modules.pop("datasets", None)
    assert state["encoder_sum"] > 0
    # the trainer fed the drop-in generator exactly what it fed the reference's (same RNG consumption up to that point)
    assert np.array_equal(rec["gen_in"][0].numpy(), gold["content"]) and np.array_equal(rec["gen_in"][1].numpy(), gold["style"])
    # bf16 path: as close to the reference's fp32 image as plain torch with bf16 storage between the layers is (DESIGN
    # section 5; this 2-line, 128-px case: 3.7e-2 for the drop-in)
    from oracle import gen as ogen
    with torch.no_grad():
        emu = ogen.generator_forward({k: v.clone() for k, v in gsd.items()}, content, style, noise, emulate_bf16=True)
    ref_img = torch.from_numpy(gold["image"])
    e, e_emu = _rel_l2(rec["gen_out"], ref_img), _rel_l2(emu, ref_img)
    assert e <= 1.3 * e_emu + 2e-2, (e, e_emu)
    assert abs(log["genRecogLoss"] - gold["losses"][0]) <= 5e-2 * abs(gold["losses"][0])
    assert abs(log["generatorLoss"] - gold["losses"][1]) <= 2e-2 * abs(gold["losses"][1]) + 2e-3
    assert np.abs(model.discriminator.state_dict()["convs1.0.module.weight_u"].numpy() - gold["disc_u_after"]).max() <= 1e-4
    # the two gradient sets the trainer stashed (recognition loss first, then the adversarial loss), by direction against the
    # fp32 oracle chain on the FULL tensors (that chain reproduces the reference trainer's sets to a cosine of 0.99999,
    # tests/test_trainer_gen_cpu.py; the golden itself stores 256-entry samples).  22 bf16 layers deep on a 2-line, 128-px
    # case with train-mode BatchNorm over two samples: plain torch with bf16 storage emulation of the FORWARD reaches 0.74 /
    # 0.94 here, the drop-ins (which also keep the gradients in bf16 between the layers) 0.51 / 0.94; a wrong sign, scale,
    # stale or swapped set gives ~0 or < 0 (checked across the sets below).  Larger cases sit higher
    # (tests/test_chain_emulated_cpu.py: 0.80 / 0.99).
    assert len(tr.saved_grads) == 2
    names = [n for n, _ in model.named_parameters()]
    ref_sets = _oracle_sets(gold)
    for si, setname in enumerate(("recog", "adv")):
        got = {n[len("generator."):]: g for n, g in zip(names, tr.saved_grads[si]) if g is not None and n.startswith("generator.")}
        keys = [k for k in ref_sets[setname] if k in got]
        assert len(keys) >= 60, len(keys)
        num = sum(float((got[k].double() * ref_sets[setname][k].double()).sum()) for k in keys)
        d1 = sum(float((got[k].double() ** 2).sum()) for k in keys)
        d2 = sum(float((ref_sets[setname][k].double() ** 2).sum()) for k in keys)
        cos = num / (d1 * d2) ** 0.5
        print(f"trainer drop-in: {setname} gradient set, cosine with the fp32 oracle chain {cos:.3f}")
        assert cos >= (0.4 if setname == "recog" else 0.8), (setname, cos)
        other = ref_sets["adv" if setname == "recog" else "recog"]
        cross = sum(float((got[k].double() * other[k].double()).sum()) for k in keys) / (
            d1 * sum(float((other[k].double() ** 2).sum()) for k in keys)) ** 0.5
        print(f"trainer drop-in: {setname} set against the OTHER loss's reference gradient {cross:.3f}")
        assert cos >= cross + 0.3, (setname, cos, cross)      # it is the gradient of ITS loss
    assert {"hwg_ctc_forward", "hwg_ctc_backward", "hwg_spectral_norm", "hwg_gen_output_bwd"} <= set(calls)


def test_reference_trainer_disc_lesson_runs_on_the_drop_ins(golden_dir, hwg_lib, monkeypatch):
    """Curriculum slot ["disc"] (trainer :785-806, :381-388): real || generated (detached) lines through the drop-in
    discriminator, the trainer's hinge loss, backward to all 28 discriminator tensors, `clip_grad_value_`, Adam step of
    the trainer's own optimizer_discriminator — against the golden of the same lesson on the reference's classes."""
    import importlib
    import sys

    import handwriting_line_generation_b200 as pkg
    from handwriting_line_generation_b200 import _lib, integrate
    from oracle import make_trainer_golden as harness
    from oracle import synth
    gold = np.load(f"{golden_dir}/trainer_disc.npz")
    _, _, _, _, style, noise, _ = build_inputs(gold)
    B = style.size(0)
    masks = {k: torch.from_numpy(v) for k, v in synth.disc_masks(2 * B, int(gold["seeds"][4])).items()}
    saved_path, saved_ds = list(sys.path), sys.modules.get("datasets")
    state = {}

    def install():
        hws = importlib.import_module("model.hw_with_style")
        mloss = importlib.import_module("model.loss")
        state["orig"] = (hws, mloss, hws.SpacedGenerator, hws.CNNOnlyHWR, hws.DiscriminatorAP, mloss.CTCLoss)
        integrate.install(retain_graph=True)

    def hook(tr, model, rec):
        assert isinstance(model.discriminator, pkg.DiscriminatorAP)
        model.discriminator.dropout_masks = masks
        state["before"] = {n: p.detach().clone() for n, p in model.discriminator.named_parameters()}
        recorded = model.generator.forward
        model.generator.forward = lambda content, style, *a, **k: recorded(content, style, *a, noise=noise, **k)

    try:
        with abi_emu.installed(monkeypatch):
            tr, log, rec, model = harness.run_lesson("disc", install=install, hook=hook)
    finally:
        hws, mloss, g, h, d, c = state["orig"]
        hws.SpacedGenerator, hws.CNNOnlyHWR, hws.DiscriminatorAP, mloss.CTCLoss = g, h, d, c
        _lib.RETAIN_SAVED = False
        sys.path[:] = saved_path
        if saved_ds is not None:
            sys.modules["datasets"] = saved_ds
        else:
            sys.modules.pop("datasets", None)
    assert np.array_equal(rec["gen_in"][0].numpy(), gold["content"]) and np.array_equal(rec["gen_in"][1].numpy(), gold["style"])
    assert abs(log["discriminatorLoss"] - gold["losses"][0]) <= 2e-2 * abs(gold["losses"][0])
    num = d1 = d2 = 0.0
    checked = 0
    for n, p in model.discriminator.named_parameters():
        key = f"grad/disc/discriminator.{n}"
        if key + "/sample" not in gold.files:
            continue
        assert p.grad is not None, n
        ref = gold[key + "/sample"].astype(np.float64)
        samp = digest(p.grad.numpy())[1][:256].astype(np.float64)          # already clipped by the trainer (:381)
        num, d1, d2 = num + float((samp * ref).sum()), d1 + float((samp * samp).sum()), d2 + float((ref * ref).sum())
        checked += 1
        if p.requires_grad and float(np.abs(ref).max()) > 0:
            assert not torch.equal(p.detach(), state["before"][n]), n     # the trainer's optimizer stepped it
    assert checked == 28, checked
    cos = num / (d1 * d2) ** 0.5
    print(f"trainer drop-in: clipped discriminator gradients, cosine with the reference trainer's {cos:.3f}")
    assert cos >= 0.95, cos
    assert all(p.grad is None or float(p.grad.abs().max()) == 0 for n, p in model.named_parameters()
               if n.startswith("generator."))                              # fake.detach(): nothing reaches the generator


def test_reference_trainer_curriculum_cycle_runs_on_the_drop_ins(golden_dir, hwg_lib, monkeypatch):
    """Curriculum slots 1 -> 2 (["no-step","gen"] then ["auto","auto-gen"]) of the unmodified trainer on the drop-ins, incl.
    `Encoder2` at its own call site (`self.encoder(both_i)`, trainer :742), the DTW `correct_pred` (`autoencode`), the CTC
    of `reconRecog`, three backward passes over one retained graph per lesson, the trainer's own per-tensor balancing of the
    four stashed sets (:340-377), `clip_grad_value_` and `optimizer.step()`.  The second lesson draws its noise / dropout
    from RNG streams the reference run does not share, so this is a "runs end to end, finite, every generator tensor
    stepped" check; the per-lesson parity is what the two tests above pin.  The remaining lessons of the 7-lesson cycle
    (disc, gen, auto, disc, and the next cycle's count lesson) follow in the same trainer: finite logs throughout."""
    import handwriting_line_generation_b200 as pkg
    from handwriting_line_generation_b200 import _lib, integrate
    from oracle import make_trainer_golden as harness
    gold = np.load(f"{golden_dir}/trainer_gen.npz")
    _, _, _, _, _, noise, masks = build_inputs(gold)
    saved_path, saved_ds = list(sys.path), sys.modules.get("datasets")
    state = {}
    def install():
        hws = importlib.import_module("model.hw_with_style"); mloss = importlib.import_module("model.loss")
        mauto = importlib.import_module("model.autoencoder"); mtr = importlib.import_module("trainer.hw_with_style_trainer")
        state["orig"] = (hws, mloss, hws.SpacedGenerator, hws.CNNOnlyHWR, hws.DiscriminatorAP, mloss.CTCLoss, hws.correct_pred)
        state["enc"] = (mauto, mtr, mauto.Encoder2, mtr.Encoder2)
        state["f34"] = (hws.CountCNN, hws.HWWithStyle.insert_spaces, hws.CharStyleEncoder)
        integrate.install(retain_graph=True, encoder=True, dtw=True, spacer=True, style=True)
    def hook(tr, model, rec):
        model.discriminator.dropout_masks = masks
        recorded = model.generator.forward

        def first_lesson_noise(content, style, *a, **k):    # by shape: with the spacer drop-in the width is the product's own
            nz = [torch.from_numpy(z) for z in synth.gen_noise(synth.gen_noise_shapes(content.size(0), style.size(0)), 76)]
            return recorded(content, style, *a, noise=nz, **k)
        model.generator.forward = first_lesson_noise
    cwd = os.getcwd()
    try:
        with abi_emu.installed(monkeypatch) as calls:
            tr, log, rec, model = harness.run_lesson("gen", install=install, hook=hook)
            os.chdir(ref_shim.REF)
            model.discriminator.dropout_masks = None
            fwd = recorded_fwd = type(model.generator).forward.__get__(model.generator)
            def with_noise(content, style, *a, **k):
                B = style.size(0)
                nz = [torch.from_numpy(z) for z in synth.gen_noise(synth.gen_noise_shapes(content.size(0), B), 77)]
                return fwd(content, style, *a, noise=nz, **k)
            model.generator.forward = with_noise
            assert isinstance(model.style_extractor, pkg.CharStyleEncoder) and isinstance(model.spacer, pkg.CountCNN)
            before = {n: p.detach().clone() for n, p in model.generator.named_parameters()}
            before_style = {n: p.detach().clone() for n, p in model.style_extractor.named_parameters()}
            before_spacer = {n: p.detach().clone() for n, p in model.spacer.named_parameters()}
            n0 = len(calls)
            tr.iteration = 2
            log2 = tr._train_iteration(2)
            lesson2 = set(calls[n0:])
            changed = sum(int(not torch.equal(p.detach(), before[n])) for n, p in model.generator.named_parameters())
            # the 'auto' lesson trains the style extractor (its image path and the heads of the characters that occur)
            changed_style = [n for n, p in model.style_extractor.named_parameters() if not torch.equal(p.detach(), before_style[n])]
            # ... and the rest of the 7-lesson cycle (config :85-95): disc, gen, auto, disc, then the next cycle's count
            cycle = {}
            per_lesson = {1: n0, 2: len(calls) - n0}
            for it in (3, 4, 5, 6, 7):
                tr.iteration = it
                c0 = len(calls)
                cycle[it] = tr._train_iteration(it)
                per_lesson[it] = len(calls) - c0
            print("library calls per lesson (slot: calls):", per_lesson)
    finally:
        os.chdir(cwd)
        hws, mloss, g, h, d, c, cp = state["orig"]
        hws.SpacedGenerator, hws.CNNOnlyHWR, hws.DiscriminatorAP, mloss.CTCLoss, hws.correct_pred = g, h, d, c, cp
        mauto, mtr, e1, e2 = state["enc"]; mauto.Encoder2, mtr.Encoder2 = e1, e2
        hws.CountCNN, hws.HWWithStyle.insert_spaces, hws.CharStyleEncoder = state["f34"]
        _lib.RETAIN_SAVED = False
        sys.path[:] = saved_path
        if saved_ds is not None: sys.modules["datasets"] = saved_ds
        else: sys.modules.pop("datasets", None)
    for k in ("autoLoss", "perceptualLoss", "reconRecogLoss", "generatorLoss"):
        assert k in log2 and np.isfinite(log2[k]), (k, log2)
    assert 0.0 < log2["perceptualLoss"] < 5.0 and 0.0 < log2["autoLoss"] < 2.0
    assert changed == len(before) == 64                       # the balanced gradient reached every generator tensor
    assert tr.saved_grads == []                               # consumed by the trainer's balancing
    assert {"hwg_dtw_align", "hwg_add_stats", "hwg_ctc_backward", "hwg_spectral_norm", "hwg_gen_output_bwd",
            "hwg_hwr_stem_bwd_image", "hwg_norm_bwd_apply"} <= lesson2
    assert "discriminatorLoss" in cycle[3] and "discriminatorLoss" in cycle[6] and "perceptualLoss" in cycle[5]
    assert "countLoss" in cycle[7], cycle[7]
    assert any(n.startswith("down.0.") for n in changed_style) and any(n.startswith("char_extractor.") for n in changed_style)
    assert any(n.startswith("prep.") for n in changed_style) and len(changed_style) >= 40, len(changed_style)
    changed_spacer = [n for n, p in model.spacer.named_parameters() if not torch.equal(p.detach(), before_spacer[n])]
    assert len(changed_spacer) >= 14, changed_spacer          # the 'count' lesson stepped the spacer
    assert {"hwg_insert_spaces_plan", "hwg_insert_spaces_fill", "hwg_shift_expand"} <= set(calls)
    for it, lg in cycle.items():
        assert all(np.isfinite(v) for k, v in lg.items() if isinstance(v, float)), (it, lg)
