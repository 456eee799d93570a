"""CPU: the HOST side of the DiscriminatorAP drop-in run through the CPU interpreter of the C-ABI (tests/abi_emu.py) against
the goldens of the unmodified reference and the oracle — the same assertions as tests/test_disc_gpu.py makes on the real
kernels (which are GPU-verified), here pinning the module's composition (spectral-norm bookkeeping, one-launch operand
packing with the device-side 1/sigma, shift expansion, GroupNorm / Dropout2d / AvgPool passes, the backward to the image and
to all 28 parameters, the wgrad arena and its unpack table) in the build container."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import disc as odisc
from oracle import synth
from oracle.make_golden import DISC_CASES, DISC_LESSON_CASES

from . import abi_emu
from .test_disc_cpu import lesson_oracle_grads


def _rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _module(seed, training, frozen):
    from handwriting_line_generation_b200 import DiscriminatorAP
    torch.manual_seed(seed)
    m = DiscriminatorAP(64, use_low=True, use_med=True)
    sd = synth.perturb_disc(m.state_dict(), seed + 1)
    sd_cpu = {k: v.clone() for k, v in sd.items()}
    m.train(training)
    if frozen:
        for p in m.parameters():
            p.requires_grad_(False)
    return m, sd_cpu


@pytest.mark.parametrize("name", ["train_w128", "eval_w128"])
def test_gen_lesson_predictions_and_input_gradient(name, golden_dir, hwg_lib, monkeypatch):
    gold = np.load(f"{golden_dir}/disc.npz")
    B, W, wseed, iseed, training = DISC_CASES[name]
    m, sd = _module(wseed, training, frozen=True)
    masks = {k: torch.from_numpy(v) for k, v in synth.disc_masks(B, iseed + 7).items()}
    m.dropout_masks = masks
    img = torch.from_numpy(synth.hwr_case(B, W, iseed)).requires_grad_()
    with abi_emu.installed(monkeypatch) as calls:
        preds = m(img)
        loss = odisc.gen_loss(preds)
        loss.backward()
    for i, p in enumerate(preds):
        ref = torch.from_numpy(gold[f"{name}/pred{i}"])
        assert list(p.shape) == list(ref.shape)
        assert _rel_l2(p.detach(), ref) <= 2e-2, (i, _rel_l2(p.detach(), ref))
    assert abs(loss.item() - float(gold[f"{name}/loss"])) <= 2e-2 * abs(float(gold[f"{name}/loss"]))
    for k in ("convs1.0.module.weight_u", "convs3.4.module.weight_v", "convs4.14.module.weight_u"):
        assert np.abs(m.state_dict()[k].numpy() - gold[f"{name}/{k}"]).max() <= 1e-4
    img2 = torch.from_numpy(synth.hwr_case(B, W, iseed)).requires_grad_()
    odisc.gen_loss(odisc.disc_forward(sd, img2, masks, training=training)).backward()
    img3 = torch.from_numpy(synth.hwr_case(B, W, iseed)).requires_grad_()
    odisc.gen_loss(odisc.disc_forward(sd, img3, masks, training=training, emulate_bf16=True)).backward()
    e, e_emul = _rel_l2(img.grad, img2.grad), _rel_l2(img3.grad, img2.grad)
    cos = F.cosine_similarity(img.grad.double().flatten(), img2.grad.double().flatten(), dim=0).item()
    assert e <= 1.3 * e_emul + 2e-2 and cos >= 0.95, (e, e_emul, cos)
    assert {"hwg_spectral_norm", "hwg_stem_conv", "hwg_shift_collapse", "hwg_conv_fprop", "hwg_act_bwd"} <= set(calls)


def test_disc_lesson_parameter_gradients(golden_dir, hwg_lib, monkeypatch):
    name = "hinge_w200"
    gold = np.load(f"{golden_dir}/disc.npz")
    B, W, wseed, iseed = DISC_LESSON_CASES[name]
    m, sd = _module(wseed, True, frozen=False)
    m.dropout_masks = {k: torch.from_numpy(v) for k, v in synth.disc_masks(B, iseed + 7).items()}
    with abi_emu.installed(monkeypatch) as calls:
        preds = m(torch.from_numpy(synth.hwr_case(B, W, iseed)))
        loss = odisc.hinge_loss(preds, B // 2)
        loss.backward()
    assert abs(loss.item() - float(gold[f"{name}/loss"])) <= 2e-2 * abs(float(gold[f"{name}/loss"]))
    _, ref = lesson_oracle_grads(sd, B, W, iseed)
    _, emu = lesson_oracle_grads(sd, B, W, iseed, emulate_bf16=True)
    got = {n: p.grad for n, p in m.named_parameters() if p.requires_grad}
    assert sorted(got) == sorted(ref)
    for n in ref:
        assert got[n] is not None and got[n].shape == ref[n].shape, n
        if ref[n].abs().max() < 1e-7:
            assert got[n].abs().max() <= 1e-4, n
            continue
        e, e_emu = _rel_l2(got[n], ref[n]), _rel_l2(emu[n], ref[n])
        cos = F.cosine_similarity(got[n].double().flatten(), ref[n].double().flatten(), dim=0).item()
        assert e <= 1.3 * e_emu + 2e-2 and cos >= 0.95, (n, e, e_emu, cos)
    assert {"hwg_conv_wgrad", "hwg_channel_sum", "hwg_spectral_norm_bwd", "hwg_norm_bwd_apply"} <= set(calls)
