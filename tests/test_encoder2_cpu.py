"""CPU: the HOST side of the Encoder2 drop-in (handwriting_line_generation_b200/encoder2.py; SURVEY §8 f1, second half).

* state_dict contract and same-seed initialisation against the fixture made from the unmodified reference;
* every packed operand of its ONE hwg_linear_map table against the plain torch re-layout (CPU job-table interpreter);
* the composition — tap lists, shift expansion, GroupNorm / Dropout2d / residual bookkeeping, the two-launch (6,3) head,
  the backward chain over the recon half — run through the CPU interpreter of the C-ABI (tests/abi_emu.py) and compared
  with the oracle (pinned to the reference goldens by tests/test_enc_cpu.py) on the golden cases.

The module has not run on a GPU yet (tools/pending_test_enc_gpu.py is the parity test for the real kernels); what this file
pins is that, GIVEN kernels that follow include/hwg_b200.h, the module computes the reference's function."""
import numpy as np
import pytest
import torch

from handwriting_line_generation_b200 import conv
from handwriting_line_generation_b200.encoder2 import DROPOUT_SITES, Encoder2
from oracle import enc as oenc
from oracle import synth
from oracle.make_golden import ENC_CASES, keys_fixture, weights_digest

from . import abi_emu, ref_map
from .test_enc_cpu import encoder2_state_dict


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def test_state_dict_and_seeded_init_equal_the_reference(golden_dir):
    gold = np.load(f"{golden_dir}/enc.npz")
    torch.manual_seed(500)
    m = Encoder2(32)
    sd = m.state_dict()
    assert keys_fixture(sd).tolist() == gold["state_dict_keys"].tolist()
    d = gold["eval_w128/weights_digest"]
    assert abs(weights_digest(sd) - d) <= 1e-6 * abs(d)
    assert oenc.DROPOUT_SITES == DROPOUT_SITES


def test_packed_operands_equal_the_torch_relayout(hwg_lib):
    torch.manual_seed(3)
    m = Encoder2(32)
    table, c = m.build_table(torch.device("cpu"))
    ref_map.run_jobs_cpu(table)
    bf = lambda t: t.to(torch.bfloat16).float()                                           # noqa: E731
    w = m.down_conv1[0].weight.detach()
    f, d, taps = c["down_conv1.0"]
    ref = torch.zeros(5, 32, 16)
    ref[:, :, :5] = w[:, 0].permute(1, 0, 2)                                             # [dy][co][dx]
    assert torch.equal(f.float(), bf(ref)) and torch.equal(d.float(), bf(ref.permute(0, 2, 1)))
    assert taps == [(2 - dy, 0) for dy in range(5)]
    for site, mod, t in m.conv_layers():
        f, d, td = c[site]
        ref = conv.pack_conv2d_weight(mod.weight.detach()).float()                       # [taps][co][ci]
        assert torch.equal(f.float(), ref), site
        assert torch.equal(d.float(), ref.permute(0, 2, 1)), site
        assert td == [(-dh, -dw) for dh, dw in t]
    w = m.down_conv3[7].weight.detach()
    for half, (f, d, td) in enumerate(c["down_conv3.7"]):
        ref = conv.pack_conv2d_weight(w[:, :, 3 * half:3 * half + 3]).float()
        assert torch.equal(f.float(), ref) and torch.equal(d.float(), ref.permute(0, 2, 1))


def _case(name):
    B, W, wseed, iseed, training = ENC_CASES[name]
    sd = encoder2_state_dict(wseed)
    r = np.random.RandomState(iseed + 7)
    masks = [torch.from_numpy((r.rand(2 * B, c) >= 3 * p).astype(np.float32)) for _, c, p in oenc.DROPOUT_SITES]
    image = torch.from_numpy(synth.hwr_case(B, W, iseed))
    recon = torch.from_numpy(synth.hwr_case(B, W, iseed + 1))
    m = Encoder2(32)
    m.load_state_dict(sd)
    m.train(training)
    m.dropout_masks = masks
    return m, sd, masks, image, recon, training


@pytest.mark.parametrize("name", sorted(ENC_CASES))
def test_composition_through_the_abi_interpreter_matches_the_oracle(name, hwg_lib, monkeypatch):
    m, sd, masks, image, recon, training = _case(name)
    ofeat, omid = oenc.encoder2_forward(sd, torch.cat((image, recon), 0), masks, training)
    orecon = recon.clone().requires_grad_()
    oloss = oenc.perceptual_loss(sd, image, orecon, masks, training)
    oloss.backward()
    with abi_emu.installed(monkeypatch) as calls:
        with torch.no_grad():
            feat, mid = m(torch.cat((image, recon), 0))
        r = recon.clone().requires_grad_()
        loss = m.perceptual_loss(image, r)
        (3.0 * loss).backward()
    assert tuple(feat.shape) == tuple(ofeat.shape) and tuple(mid.shape) == tuple(omid.shape)
    # bf16 storage between the layers, fp32 arithmetic: the same 2e-2 per-tensor bound as the CUDA parity tests
    assert rel_l2(feat, ofeat) <= 1.5e-2, rel_l2(feat, ofeat)   # measured 5.3e-3 / 5.5e-3
    assert rel_l2(mid, omid) <= 1.5e-2, rel_l2(mid, omid)       # measured 6.9e-3 / 6.3e-3
    assert abs(loss.item() - oloss.item()) <= 5e-3 * abs(oloss.item())   # measured 7e-4 / 5e-6
    # the loss is an L1: its gradient is a sum of sign patterns, and a feature difference inside the bf16 rounding flips
    # a whole sign — direction and size are what can be compared
    g, og = r.grad / 3.0, orecon.grad
    cos = float((g.double() * og.double()).sum() / (g.double().norm() * og.double().norm()))
    assert cos >= 0.95, cos                                              # measured 0.984 / 0.982
    assert abs(float(g.norm() / og.norm()) - 1.0) <= 5e-2
    assert {"hwg_conv_fprop", "hwg_stem_conv", "hwg_shift_collapse", "hwg_gn_coeffs", "hwg_scale_shift_act",
            "hwg_avgpool_nhwc", "hwg_add_stats", "hwg_l1_halves", "hwg_norm_bwd_reduce", "hwg_gn_bwd_coeffs",
            "hwg_norm_bwd_apply", "hwg_act_bwd"} == set(calls)


def test_module_surface_backward_equals_the_fused_loss(hwg_lib, monkeypatch):
    """`Encoder2.forward` (the reference surface: NCHW fp32 features, autograd to the input) followed by the trainer's own
    chunk + l1_loss must give the gradient `perceptual_loss` computes with the fused L1 kernel on the recon half."""
    m, sd, masks, image, recon, training = _case("train_w200")
    with abi_emu.installed(monkeypatch):
        r1 = recon.clone().requires_grad_()
        l1 = m.perceptual_loss(image, r1)
        l1.backward()
        r2 = recon.clone().requires_grad_()
        feats = m(torch.cat((image, r2), 0))
        l2 = 0
        for f in feats:
            o_f, r_f = torch.chunk(f, 2, dim=0)
            l2 = l2 + torch.nn.functional.l1_loss(r_f, o_f)
        l2.backward()
    assert abs(l1.item() - l2.item()) <= 1e-5 * abs(l2.item())
    assert rel_l2(r1.grad, r2.grad) <= 1e-2, rel_l2(r1.grad, r2.grad)


def test_cpu_tensors_are_rejected_without_the_interpreter():
    m = Encoder2(32)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 1, 64, 128))


def test_repeated_backward_needs_and_honours_set_retain_graph(hwg_lib, monkeypatch):
    """The reference trainer's gradient balancing backpropagates several losses through ONE graph with retain_graph=True
    (trainer :300-338).  Default: the saved state goes with the first backward and a second one fails with torch's own
    wording; with set_retain_graph(True) the second backward reproduces the first."""
    import handwriting_line_generation_b200 as pkg
    from handwriting_line_generation_b200 import _lib
    m, sd, masks, image, recon, training = _case("eval_w128")
    with abi_emu.installed(monkeypatch):
        r = recon.clone().requires_grad_()
        loss = m.perceptual_loss(image, r)
        loss.backward(retain_graph=True)
        with pytest.raises(RuntimeError, match="second time"):
            loss.backward()
        monkeypatch.setattr(_lib, "RETAIN_SAVED", False)
        pkg.set_retain_graph(True)
        assert _lib.RETAIN_SAVED
        r = recon.clone().requires_grad_()
        loss = m.perceptual_loss(image, r)
        loss.backward(retain_graph=True)
        g1 = r.grad.clone()
        loss.backward()
        assert torch.equal(r.grad, 2 * g1)
        pkg.set_retain_graph(False)
