"""CPU: the oracle chain (oracle/gen.py -> oracle/hwr.py -> CTC, oracle/disc.py -> adversarial loss) against the
UNMODIFIED reference TRAINER: tests/golden/trainer_gen.npz holds what `HWWithStyleTrainer._train_iteration` fed the
generator in its 'gen' lesson (curriculum slot ["no-step","gen"]), the two weighted losses and the two gradient sets it
stashes for balancing (trainer/hw_with_style_trainer.py:300-338).  Written by `python -m oracle.make_trainer_golden`
(SURVEY.md Appendix A recipe).  This anchors the oracle — and through the GPU parity tests the CUDA path — on the
reference's own call sites: `self.model(label, label_lengths, style_gen)` (:577), `self.model.hwr(gen_image)` + the
`genRecog` CTC call (:760-762), `self.model.discriminator(fake)` (:810)."""
import numpy as np
import pytest
import torch

from oracle import disc as odisc
from oracle import gen as ogen
from oracle import hwr as ohwr
from oracle import synth
from oracle.make_golden import digest

W_RECOG, W_GEN = 1e-4, 1.0          # loss_weights genRecog / generator of the IAM GAN config


def build_inputs(gold):
    """Weights (seeded drop-in construction == seeded reference construction), inputs, noise and masks of the lesson."""
    from handwriting_line_generation_b200 import CNNOnlyHWR, DiscriminatorAP, SpacedGenerator
    s_gen, s_hwr, s_disc, s_noise, s_masks = (int(v) for v in gold["seeds"][:5])
    torch.manual_seed(s_gen)
    gsd = SpacedGenerator(80, 128, 256, n_style_trans=6, emb_dropout=False, append_style=True, small=False).state_dict()
    torch.manual_seed(s_hwr)
    hsd = CNNOnlyHWR(80, norm='batch').state_dict()
    torch.manual_seed(s_disc)
    dsd = synth.perturb_disc(DiscriminatorAP(64, use_low=True, use_med=True).state_dict(), s_disc + 1)
    content, style = torch.from_numpy(gold["content"]), torch.from_numpy(gold["style"])
    noise = [torch.from_numpy(np.random.RandomState(s_noise + i).standard_normal(tuple(sh)).astype(np.float32))
             for i, sh in enumerate(gold["noise_shapes"].tolist())]
    B = style.size(0)
    masks = {k: torch.from_numpy(v) for k, v in synth.disc_masks(B, s_masks).items()}
    assert gold["mask_sites"].tolist() == odisc.DROPOUT_ORDER and gold["modes"].tolist() == [1, 1, 1]
    return gsd, hsd, dsd, content, style, noise, masks


def test_oracle_chain_reproduces_the_reference_trainers_gen_lesson(golden_dir):
    """Forward quantities (generated image, both weighted losses, the spectral-norm update) hold the north_star fp32
    bound of 1e-4.  The two gradient sets cross 22 (recognition) / 22 (adversarial) fp32 layers and end in sums of
    10^4..10^6 signed products (noise weights, biases): two fp32 evaluations of the SAME chain differ there by up to
    7e-3 of a tensor's max (measured: oracle fp32 vs the trainer, and oracle fp32 vs oracle fp64, on the W=60 and W=128
    variants of this case), so each tensor is held to 1e-2 of its max and each whole set to a cosine of 0.99999 — a
    modelling difference (a missing term, a wrong scale or padding) moves both by orders of magnitude more."""
    gold = np.load(f"{golden_dir}/trainer_gen.npz")
    gsd, hsd, dsd, content, style, noise, masks = build_inputs(gold)
    gp = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in gsd.items()}
    img = ogen.generator_forward(gp, content, style, noise)
    assert np.abs(img.detach().numpy() - gold["image"]).max() <= 1e-4 * np.abs(gold["image"]).max()
    B = style.size(0)
    label = torch.from_numpy(gold["label"]).int()                     # [L,B]
    lengths = torch.from_numpy(gold["label_lengths"]).int()
    lp = ohwr.hwr_forward({k: v.clone() for k, v in hsd.items()}, img, True, {})
    pred_size = torch.IntTensor([lp.size(0)] * B)                     # trainer :761
    recog = W_RECOG * torch.nn.functional.ctc_loss(lp, label.permute(1, 0), pred_size, lengths)
    upd = {}
    adv = W_GEN * odisc.gen_loss(odisc.disc_forward(dsd, img, masks, training=True, update=upd))
    assert abs(recog.item() - gold["losses"][0]) <= 1e-4 * abs(gold["losses"][0])
    assert abs(adv.item() - gold["losses"][1]) <= 1e-4 * abs(gold["losses"][1])
    assert np.abs(upd["convs1.0.module.weight_u"].numpy() - gold["disc_u_after"]).max() <= 1e-5
    names = [k for k, p in gp.items() if p.requires_grad]
    params = [gp[k] for k in names]
    worst = {}
    for setname, loss in (("recog", recog), ("adv", adv)):
        grads = torch.autograd.grad(loss, params, retain_graph=True, allow_unused=True)
        checked, worst[setname] = 0, 0.0
        num = d1 = d2 = 0.0
        for n, g in zip(names, grads):
            key = f"grad/{setname}/generator.{n}"
            if g is None or key + "/digest" not in gold.files:
                continue
            dig, ref = gold[key + "/digest"], gold[key + "/sample"].astype(np.float64)
            samp = digest(g.numpy())[1][:256].astype(np.float64)
            err = float(np.abs(samp - ref).max() / max(dig[3], 1e-30))
            worst[setname] = max(worst[setname], err)
            assert err <= 1e-2, (setname, n, err)
            num, d1, d2 = num + float((samp * ref).sum()), d1 + float((samp * samp).sum()), d2 + float((ref * ref).sum())
            checked += 1
        assert checked >= 60, checked
        assert num / (d1 * d2) ** 0.5 >= 0.99999, (setname, num / (d1 * d2) ** 0.5)
    print("worst relative error vs the trainer's stashed gradient sets:", worst)


def test_oracle_reproduces_the_reference_trainers_disc_lesson(golden_dir):
    """Curriculum slot ["disc"] through the unmodified trainer (:785-806): real lines || generated lines (detached)
    through the discriminator, hinge loss, `clip_grad_value_(…, 2)` (:381).  The oracle's loss holds 1e-4; the gradients
    of all 28 trainable discriminator tensors are compared after the same clipping."""
    gold = np.load(f"{golden_dir}/trainer_disc.npz")
    gsd, hsd, dsd, content, style, noise, _ = build_inputs(gold)
    B = style.size(0)
    with torch.no_grad():
        fake = ogen.generator_forward(gsd, content, style, noise)
    assert np.abs(fake.numpy() - gold["image"]).max() <= 1e-4 * np.abs(gold["image"]).max()
    real = torch.from_numpy(synth.hwr_case(B, fake.size(3), int(gold["seeds"][5])))
    masks = {k: torch.from_numpy(v) for k, v in synth.disc_masks(2 * B, int(gold["seeds"][4])).items()}
    leaf = {k: v.clone().requires_grad_(not k.endswith(("weight_u", "weight_v"))) for k, v in dsd.items()}
    loss = odisc.hinge_loss(odisc.disc_forward(leaf, torch.cat((real, fake), 0), masks, training=True), B)
    assert abs(loss.item() - gold["losses"][0]) <= 1e-4 * abs(gold["losses"][0])
    loss.backward()
    checked, worst = 0, 0.0
    for n, p in leaf.items():
        key = f"grad/disc/discriminator.{n}"
        if not p.requires_grad or key + "/digest" not in gold.files:
            continue
        dig, ref = gold[key + "/digest"], gold[key + "/sample"].astype(np.float64)
        samp = digest(p.grad.clamp(-2, 2).numpy())[1][:256].astype(np.float64)
        if dig[3] < 1e-9:
            assert np.abs(samp).max() <= 1e-6, n
        else:
            err = float(np.abs(samp - ref).max() / dig[3])
            worst = max(worst, err)
            assert err <= 1e-3, (n, err)          # 12 fp32 layers, sums over up to 10^6 pixels
        checked += 1
    assert checked == 28, checked
    print("worst relative error vs the trainer's discriminator gradients:", worst)
