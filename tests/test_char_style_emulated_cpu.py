"""CPU: the CharStyleEncoder drop-in (handwriting_line_generation_b200/char_style.py + nhwc.py: space-to-depth strided
convolutions, replicate-padding gathers, grouped per-character heads, the hand-derived GroupNorm coefficient algebra, weight
re-layouts carried by autograd) through the CPU interpreter of the C-ABI against the goldens of the UNMODIFIED reference
(tests/golden/style.npz) and, for the gradients of all trainable parameters, against torch autograd over the oracle
(oracle/style.py) — the assertions of tests/test_char_style_gpu.py made in the build container."""
import numpy as np
import pytest
import torch

from oracle import style as ostyle
from oracle.make_golden import STYLE_CASES, keys_fixture, style_inputs, weights_digest

from . import abi_emu

BF16_REL = 2e-2


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm())


def build(seed):
    from handwriting_line_generation_b200.char_style import CharStyleEncoder
    torch.manual_seed(seed)
    return CharStyleEncoder(1, 64, 128, 128, 0, 'group', 'relu', 'replicate', 80, global_pool=False,
                            average_found_char_style=1.0, num_final_g_spacing_style=1, num_char_fc=1, vae=False, window=2,
                            small=False)


def test_state_dict_contract(golden_dir):
    gold = np.load(f"{golden_dir}/style.npz")
    m = build(600)
    assert keys_fixture(m.state_dict()).tolist() == gold["state_dict_keys"].tolist()
    d = gold["b2_w256/weights_digest"]
    assert abs(weights_digest(m.state_dict()) - d) <= 1e-6 * abs(d)       # same seed -> the reference's initial weights


def check_case(m, name, gold, dev, grads=True):
    """Shared with the GPU test."""
    B, W, wseed, iseed = STYLE_CASES[name]
    image, recog = style_inputs(B, W, iseed)
    m.eval()
    with torch.no_grad():
        style = m(image.to(dev), recog.to(dev))
    ref = torch.from_numpy(gold[f"{name}/style"])
    assert tuple(style.shape) == tuple(ref.shape)
    e_fwd = rel_l2(style.cpu(), ref)
    assert e_fwd <= BF16_REL, e_fwd
    if not grads:
        return e_fwd, {}
    # gradients of a linear loss w.r.t. every parameter that takes part, against autograd over the fp32 oracle
    R = torch.randn(ref.shape, generator=torch.Generator().manual_seed(iseed + 9))
    m.zero_grad()
    (m(image.to(dev), recog.to(dev)) * R.to(dev)).sum().backward()
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    (ostyle.char_style_forward(sd, image, recog) * R).sum().backward()
    got = {n: p.grad for n, p in m.named_parameters()}
    used = [k for k, v in sd.items() if v.grad is not None and float(v.grad.abs().max()) > 0]
    assert len(used) > 60
    num = d1 = d2 = 0.0
    worst = {}
    for k in used:
        assert got[k] is not None, k
        a, b = got[k].detach().cpu().double(), sd[k].grad.double()
        num, d1, d2 = num + float((a * b).sum()), d1 + float((a * a).sum()), d2 + float((b * b).sum())
        if b.numel() >= 64:
            c = float((a * b).sum() / (a.norm() * b.norm() + 1e-300))
            worst[k] = c
    cos = num / (d1 * d2) ** 0.5
    # parameters of classes that never occur get no gradient in either implementation
    unused = [k for k, v in sd.items() if v.grad is None or float(v.grad.abs().max()) == 0]
    for k in unused:
        assert got[k] is None or float(got[k].abs().max()) == 0.0, k
    return e_fwd, {"cos_all": cos, "worst": min(worst.items(), key=lambda kv: kv[1])}


@pytest.mark.parametrize("name", ["b2_w256"])
def test_char_style_encoder_through_the_interpreter(name, golden_dir, hwg_lib, monkeypatch):
    gold = np.load(f"{golden_dir}/style.npz")
    m = build(STYLE_CASES[name][2])
    with abi_emu.installed(monkeypatch) as calls:
        e_fwd, rep = check_case(m, name, gold, "cpu")
    print(f"CharStyleEncoder through the interpreter: style rel-L2 {e_fwd:.2e}, gradients {rep}")
    # bf16 forward through 7 + 3 stacked layers: the whole-gradient direction, and every large tensor's own direction
    assert rep["cos_all"] >= 0.97, rep
    assert rep["worst"][1] >= 0.8, rep
    assert {"hwg_shift_expand", "hwg_conv_fprop", "hwg_conv_wgrad", "hwg_norm_bwd_apply", "hwg_act_bwd", "hwg_channel_sum",
            "hwg_scale_shift_act"} <= set(calls)
