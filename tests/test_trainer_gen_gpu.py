"""GPU: the CUDA 'gen' lesson against the UNMODIFIED reference TRAINER (tests/golden/trainer_gen.npz, see
tests/test_trainer_gen_cpu.py): same weights (seeded construction), the content / style / labels the trainer fed,
the same noise and Dropout2d masks.  Forward values are held to the bf16 bounds of DESIGN §5; the two gradient sets the
trainer stashes (recognition loss, adversarial loss) are compared by direction, like tests/test_gen_train_gpu.py does
for the chain (22 bf16 layers: LeakyReLU / ReLU / max-pool decisions flip within bf16 rounding)."""
import numpy as np
import pytest
import torch

from oracle import disc as odisc
from oracle.make_golden import digest
from tests.test_trainer_gen_cpu import W_GEN, W_RECOG, build_inputs

pytestmark = pytest.mark.gpu


def _rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def _set_cosine(grads, gold, setname):
    num = d1 = d2 = 0.0
    n_used = 0
    for n, g in grads.items():
        key = f"grad/{setname}/generator.{n}"
        if g is None or key + "/sample" not in gold.files:
            continue
        ref = gold[key + "/sample"].astype(np.float64)
        samp = digest(g.cpu().numpy())[1][:256].astype(np.float64)
        num, d1, d2 = num + float((samp * ref).sum()), d1 + float((samp * samp).sum()), d2 + float((ref * ref).sum())
        n_used += 1
    assert n_used >= 60, n_used
    return num / (d1 * d2) ** 0.5


def test_cuda_gen_lesson_against_the_reference_trainer():
    import os
    import handwriting_line_generation_b200 as pkg
    golden_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    gold = np.load(f"{golden_dir}/trainer_gen.npz")
    gsd, hsd, dsd, content, style, noise, masks = build_inputs(gold)

    def module(cls, sd, *a, **k):
        m = cls(*a, **k)
        m.load_state_dict(sd)
        return m.cuda().train()

    gen = module(pkg.SpacedGenerator, gsd, 80, 128, 256, n_style_trans=6, emb_dropout=False, append_style=True, small=False)
    hwr = module(pkg.CNNOnlyHWR, hsd, 80, norm='batch')
    disc = module(pkg.DiscriminatorAP, dsd, 64, use_low=True, use_med=True)
    for p in list(hwr.parameters()) + list(disc.parameters()):
        p.requires_grad_(False)
    disc.dropout_masks = masks
    c, s = content.cuda(), style.cuda()
    nz = [z.cuda() for z in noise]
    B = style.size(0)
    label = torch.from_numpy(np.ascontiguousarray(gold["label"].T)).int().cuda()          # [B,L]
    lengths = torch.from_numpy(gold["label_lengths"]).int()

    # ---- recognition loss (trainer :760-762) and its gradient set
    img = gen(c, s, noise=nz)
    # bf16 path on a 2-line, 128-px case: as close to the reference's fp32 image as plain torch with bf16 storage between
    # the layers is; the CPU interpreter of the same modules measures 3.7e-2
    from oracle import gen as ogen
    with torch.no_grad():
        emu = ogen.generator_forward({k: v.clone() for k, v in gsd.items()}, content, style, noise, emulate_bf16=True)
    ref_img = torch.from_numpy(gold["image"])
    e, e_emu = _rel_l2(img.detach().cpu(), ref_img), _rel_l2(emu, ref_img)
    assert e <= 1.3 * e_emu + 2e-2, (e, e_emu)
    lp = hwr(img)
    recog = W_RECOG * pkg.CTCLoss(lp, label, torch.IntTensor([lp.size(0)] * B), lengths)
    assert abs(recog.item() - gold["losses"][0]) <= 5e-2 * abs(gold["losses"][0]), (recog.item(), gold["losses"][0])
    recog.backward()
    g_recog = {n: (None if p.grad is None else p.grad.detach().clone()) for n, p in gen.named_parameters()}
    for p in gen.parameters():
        p.grad = None
    # ---- adversarial loss (trainer :810-821) on a fresh forward with the same noise, and its gradient set
    img = gen(c, s, noise=nz)
    adv = W_GEN * odisc.gen_loss(disc(img))
    assert abs(adv.item() - gold["losses"][1]) <= 2e-2 * abs(gold["losses"][1]) + 2e-3, (adv.item(), gold["losses"][1])
    assert np.abs(disc.state_dict()["convs1.0.module.weight_u"].cpu().numpy() - gold["disc_u_after"]).max() <= 1e-4
    adv.backward()
    g_adv = {n: (None if p.grad is None else p.grad.detach().clone()) for n, p in gen.named_parameters()}
    cos_r, cos_a = _set_cosine(g_recog, gold, "recog"), _set_cosine(g_adv, gold, "adv")
    # measured through the CPU interpreter of the same modules (tests/test_trainer_dropin_cpu.py): 0.53 / 0.94 on these
    # samples (plain torch with bf16 forward emulation: 0.77 / 0.94); a wrong sign, scale or a swapped set gives <= 0, an
    # unrelated gradient 0 +- 1e-3 (3.3 M components).  The recognition set crosses a random-init recognizer whose train-mode
    # BatchNorm normalises over TWO 128-pixel lines: on the B200 the fp32 atomics' summation order of the statistics moves
    # its cosine from run to run — ten runs in round 2: 0.27 ... 0.53 (profiles/microbench_r02.txt) — hence the low bar
    assert cos_r >= 0.15 and cos_a >= 0.8, (cos_r, cos_a)
