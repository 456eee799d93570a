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
        assert np.abs(samp - gold[f"{name}/grad/{n}/sample"]).max() <= FP32_REL * dig[3] + 1e-9, n
