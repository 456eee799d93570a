"""CPU: the torch restatement of the discriminator (oracle/disc.py) against the golden outputs, input gradients and
spectral-norm updates of the unmodified reference (tests/golden/disc.npz, written by oracle/make_golden.py), and the
drop-in module's state_dict contract (names, shapes, seeded init) — no GPU needed."""
import numpy as np
import pytest
import torch

from oracle import disc as odisc
from oracle import synth
from oracle.make_golden import DISC_CASES, digest, keys_fixture, weights_digest

FP32_REL = 1e-4


def disc_module(seed):
    from handwriting_line_generation_b200 import DiscriminatorAP
    torch.manual_seed(seed)
    m = DiscriminatorAP(64, use_low=True, use_med=True)
    sd = synth.perturb_disc(m.state_dict(), seed + 1)
    return m, sd


def test_disc_state_dict_contract(golden_dir):
    gold = np.load(f"{golden_dir}/disc.npz")
    m, sd = disc_module(DISC_CASES["train_w128"][2])
    assert keys_fixture(sd).tolist() == gold["state_dict_keys"].tolist()
    # same seed -> same random init as the reference (construction order mirrors discriminator_ap.py:70-130,
    # including the u / v draws of SpectralNorm._make_params :45-61)
    assert abs(weights_digest(sd) - gold["train_w128/weights_digest"]) <= 1e-6 * abs(gold["train_w128/weights_digest"])
    assert sum(p.numel() for p in m.parameters()) == 1506308


@pytest.mark.parametrize("name", sorted(DISC_CASES))
def test_disc_oracle_matches_reference_golden(name, golden_dir):
    gold = np.load(f"{golden_dir}/disc.npz")
    B, W, wseed, iseed, training = DISC_CASES[name]
    _, sd = disc_module(wseed)
    masks = {k: torch.from_numpy(v) for k, v in synth.disc_masks(B, iseed + 7).items()}
    img = torch.from_numpy(synth.hwr_case(B, W, iseed)).requires_grad_()
    update = {}
    preds = odisc.disc_forward(sd, img, masks, training=training, update=update)
    loss = odisc.gen_loss(preds)
    loss.backward()
    for i, p in enumerate(preds):
        ref = gold[f"{name}/pred{i}"]
        assert list(p.shape) == list(ref.shape)
        assert np.abs(p.detach().numpy() - ref).max() <= FP32_REL * np.abs(ref).max()
    assert abs(loss.item() - float(gold[f"{name}/loss"])) <= FP32_REL * abs(float(gold[f"{name}/loss"]))
    _, samp = digest(img.grad.numpy())
    assert np.abs(samp - gold[f"{name}/grad_sample"]).max() <= FP32_REL * gold[f"{name}/grad_digest"][3]
    for k in ("convs1.0.module.weight_u", "convs3.4.module.weight_v", "convs4.14.module.weight_u"):
        assert np.abs(update[k].numpy() - gold[f"{name}/{k}"]).max() <= 1e-5


def test_disc_module_refuses_cpu_tensors():
    m, _ = disc_module(1)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 1, 64, 64))


def lesson_oracle_grads(sd, B, W, iseed, emulate_bf16=False):
    """Hinge-loss ('disc' lesson) gradients of every trainable parameter through the oracle."""
    masks = {k: torch.from_numpy(v) for k, v in synth.disc_masks(B, iseed + 7).items()}
    leaf = {k: v.clone().requires_grad_(not k.endswith(("weight_u", "weight_v"))) for k, v in sd.items()}
    preds = odisc.disc_forward(leaf, torch.from_numpy(synth.hwr_case(B, W, iseed)), masks, training=True,
                               emulate_bf16=emulate_bf16)
    loss = odisc.hinge_loss(preds, B // 2)
    loss.backward()
    return loss.item(), {k: v.grad for k, v in leaf.items() if v.requires_grad}


@pytest.mark.parametrize("name", ["hinge_w128", "hinge_w200"])
def test_disc_oracle_parameter_gradients_match_reference_golden(name, golden_dir):
    from oracle.make_golden import DISC_LESSON_CASES
    gold = np.load(f"{golden_dir}/disc.npz")
    B, W, wseed, iseed = DISC_LESSON_CASES[name]
    _, sd = disc_module(wseed)
    loss, grads = lesson_oracle_grads(sd, B, W, iseed)
    assert abs(loss - float(gold[f"{name}/loss"])) <= FP32_REL * abs(float(gold[f"{name}/loss"]))
    assert sorted(grads) == sorted(gold[f"{name}/param_names"].tolist())
    for n, g in grads.items():
        dig = gold[f"{name}/grad/{n}/digest"]
        _, samp = digest(g.numpy())
        ref = gold[f"{name}/grad/{n}/sample"]
        assert np.abs(samp[:ref.size] - ref).max() <= FP32_REL * dig[3] + 1e-9, n


def test_disc_job_tables_match_torch_relayouts():
    """The discriminator's two hwg_linear_map tables, interpreted on the CPU (tests/ref_map.py): (1) every forward /
    dgrad operand equals the torch re-layout of weight_bar * (1/sigma) (hwgMapJob.scale_dev), incl. the 7-tap x
    16-channel operand of in_conv and the 16-row padding of the one-channel heads; (2) the wgrad unpack is the adjoint
    of the forward pack scaled by 1/sigma, biases and GroupNorm sums land in their parameters' slots."""
    from handwriting_line_generation_b200 import conv
    from tests import ref_map
    m, sd = disc_module(5)
    plan = m._build_plan()
    g = torch.Generator().manual_seed(1)
    plan["inv_sigma"].copy_(torch.rand(plan["n_sn"], generator=g) + 0.5)
    ref_map.run_jobs_cpu(plan["table"])
    c = plan["c"]
    k = 0
    for site, mod, taps, spectral in m.conv_layers():
        w = (mod.weight_bar if spectral else mod.weight).detach()
        scale = plan["inv_sigma"][k] if spectral else torch.tensor(1.0)
        k += int(spectral)
        co, ci = w.size(0), w.size(1)
        if site == "in_conv.0":
            ref_f = torch.zeros(7, co, 16)
            ref_f[:, :, :7] = w[:, 0].permute(1, 0, 2)                       # [dy][co][dx]
            ref_d = ref_f.permute(0, 2, 1)
            got_d = c["dgrad"][site][0].float()
        else:
            ref_f = conv.pack_conv2d_weight(w * scale).float()              # [taps][co][ci]
            ref_d = ref_f.permute(0, 2, 1)
            got_d = c["dgrad"][site][0].float()[:, :ci, :co]
            assert c["dgrad"][site][1] == [(-dh, -dw) for dh, dw in taps]
        got_f = c[site].float()
        assert (got_f[:, :co] - ref_f.to(torch.bfloat16).float()).abs().max() <= 8e-3 * ref_f.abs().max(), site
        assert got_f[:, co:].abs().max() == 0 if got_f.size(1) > co else True
        assert (got_d[:, :, :co] - ref_d.to(torch.bfloat16).float()).abs().max() <= 8e-3 * ref_f.abs().max(), site
        bp, b = c["bias"][site]
        assert torch.equal(bp[:b.numel()] if b is not None else bp, mod.bias.detach())
    # ---- unpack
    m._plan = plan
    wp = m._wgrad_plan(torch.device("cpu"))
    arena = torch.randn(wp["arena_floats"], generator=g)
    ref_map.run_jobs_cpu(wp["table"], src_base=arena)
    params = dict(m.named_parameters())
    k = 0
    for site, (wname, bname, co, cop, ci, taps, spectral) in wp["meta"].items():
        scale = float(plan["inv_sigma"][k]) if spectral else 1.0
        k += int(spectral)
        o, n = wp["slots"][("w", site)]
        dw = arena[o:o + n].view(len(taps), cop, ci)
        got = wp["gflat"][wp["goff"][wname]:wp["goff"][wname] + params[wname].numel()].view_as(params[wname])
        if site == "in_conv.0":
            ref = dw[:, :co, :7].permute(1, 0, 2).reshape(co, 1, 7, 7)
        else:
            kh, kw = params[wname].shape[2:]
            ref = (dw[:, :co] * scale).view(kh, kw, co, ci).permute(2, 3, 0, 1)
        assert torch.allclose(got, ref, rtol=1e-6, atol=1e-6), site
        ob, _ = wp["slots"][("b", site)]
        assert torch.equal(wp["gflat"][wp["goff"][bname]:wp["goff"][bname] + co], arena[ob:ob + co]), site
    for gname in ("in_conv.1", "convs3.1"):
        for kind, suffix in (("gamma", ".weight"), ("beta", ".bias")):
            o, n = wp["slots"][(kind, gname)]
            go = wp["goff"][gname + suffix]
            assert torch.equal(wp["gflat"][go:go + n], arena[o:o + n])
