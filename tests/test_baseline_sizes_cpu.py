"""CPU: the oracle (oracle/gen.py, oracle/hwr.py, oracle/disc.py, oracle/ctc_oracle.c) pinned to the UNMODIFIED reference at
BASELINE.json's configuration sizes (tests/golden/full.npz, `python -m oracle.make_golden full`): configs[0] (recognizer + CTC
fwd+bwd, 8 lines of 64x1024), configs[1] (generator inference, 32 lines), the bench step's forward at 16 lines, and a
two-line slice of configs[4] (64x2048 lines; the whole B=64 case is compared on the GPU box, where it takes seconds).
fp32 on both sides, same ATen ops: outputs 1e-4 of max (observed 8e-7), losses 1e-5, gradient tensors rel-L2 1e-2 over the
golden's strided sample (observed 2.2e-5 next to the loss growing to 2.8e-3 below the four BatchNorm layers of the head, whose
backward cancels two means: fp32 summation order of the host's BLAS threads, the noise tests/test_trainer_gen_cpu.py measures)."""
import numpy as np
import torch

from oracle import ctc as octc
from oracle import disc as odisc
from oracle import gen as ogen
from oracle import hwr as ohwr
from oracle import synth
from oracle.make_golden import FULL, digest, full_labels, weights_digest


def _close(t, gold, key, tol, n=None):
    dig, samp = digest(t.detach().numpy() if torch.is_tensor(t) else np.asarray(t))
    ref = gold[key + "/sample"]
    samp = samp[:len(ref)]
    scale = float(np.abs(ref).max()) + 1e-30
    err = float(np.abs(samp - ref).max()) / scale
    assert err <= tol, f"{key}: {err:.2e} of max"
    assert abs(dig[1] - gold[key + "/digest"][1]) <= max(tol, 1e-4) * abs(gold[key + "/digest"][1]) + 1e-12, key


ZERO_GRAD = {"cnn.conv2.bias", "cnn.conv4.bias", "cnn.conv6.bias", "cnn1d.0.bias", "cnn1d.3.bias", "cnn1d.6.bias",
             "cnn1d.9.bias"}


def _rel(t, gold, key, tol):
    _, samp = digest(t.detach().numpy() if torch.is_tensor(t) else np.asarray(t))
    ref = gold[key + "/sample"].astype(np.float64)
    samp = samp[:len(ref)].astype(np.float64)
    err = float(np.linalg.norm(samp - ref) / (np.linalg.norm(ref) + 1e-300))
    assert err <= tol, f"{key}: rel-L2 {err:.2e}"


def _sd(make, seed):
    _, sd = synth.state_dict_from_seed(make, seed)
    return {k: v.clone() for k, v in sd.items()}


def _gen_sd(seed, C=80):
    from handwriting_line_generation_b200 import SpacedGenerator     # parameter container (same seeded init as the reference)
    return _sd(lambda: SpacedGenerator(C, 128, 256, n_style_trans=6, emb_dropout=False, append_style=True, small=False), seed)


def _hwr_sd(seed, C=80):
    from handwriting_line_generation_b200 import CNNOnlyHWR
    return _sd(lambda: CNNOnlyHWR(C, norm='batch'), seed)


def test_oracle_config1_recognizer_ctc_fwd_bwd(golden_dir):
    gold = np.load(f"{golden_dir}/full.npz")
    c = FULL["cfg1"]
    sd = _hwr_sd(c["wseed"])
    assert abs(weights_digest(sd) - gold["cfg1/weights_digest"]) <= 1e-6 * abs(gold["cfg1/weights_digest"])
    p = {k: v.requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    x = torch.from_numpy(synth.hwr_case(c["B"], c["W"], c["iseed"])).requires_grad_()
    label = torch.from_numpy(full_labels(c["B"], c["S"], 80, c["iseed"] + 1))
    T = c["W"] // 4 - 6
    lp = ohwr.hwr_forward(p, x, True, None)
    il, tl = torch.IntTensor([T] * c["B"]), torch.IntTensor([c["S"]] * c["B"])
    loss = torch.nn.functional.ctc_loss(lp, label.permute(1, 0), il, tl)
    loss.backward()
    _close(lp, gold, "cfg1/log_probs", 1e-4)
    assert abs(loss.item() - float(gold["cfg1/loss"])) <= 1e-5 * float(gold["cfg1/loss"])
    _rel(x.grad, gold, "cfg1/grad/input", 1e-2)
    names = [k for k in gold.files if k.startswith("cfg1/grad/") and k.endswith("/sample") and "input" not in k]
    assert len(names) == 38
    for k in names:
        n = k[len("cfg1/grad/"):-len("/sample")]
        if n in ZERO_GRAD:       # bias of a convolution that feeds BatchNorm: the true gradient is identically zero
            wmax = float(np.abs(gold[f"cfg1/grad/{n.replace('bias', 'weight')}/sample"]).max())
            assert float(p[n].grad.abs().max()) <= 1e-3 * wmax and float(np.abs(gold[k]).max()) <= 1e-3 * wmax, n
            continue
        _rel(p[n].grad, gold, f"cfg1/grad/{n}", 1e-2)
    # the C oracle of the CTC on the same log-probs
    oloss, _, _ = octc.ctc_loss_and_grad(lp.detach().numpy(), np.ascontiguousarray(label.numpy().T), il.numpy(), tl.numpy())
    assert abs(oloss - float(gold["cfg1/loss"])) <= 1e-4 * float(gold["cfg1/loss"])
    assert np.array_equal(lp.argmax(2).numpy().astype(np.int16), gold["cfg1/argmax"])


def test_oracle_config2_generator_inference(golden_dir):
    gold = np.load(f"{golden_dir}/full.npz")
    c = FULL["cfg2"]
    sd = _gen_sd(c["wseed"])
    assert abs(weights_digest(sd) - gold["cfg2/weights_digest"]) <= 1e-6 * abs(gold["cfg2/weights_digest"])
    content, style = synth.gen_case(c["T"], c["B"], 80, 128, c["iseed"])
    noise = [torch.from_numpy(z) for z in synth.gen_noise(synth.gen_noise_shapes(c["T"], c["B"]), c["iseed"] + 7)]
    with torch.no_grad():
        img = ogen.generator_forward(sd, torch.from_numpy(content), torch.from_numpy(style), noise)
    assert tuple(img.shape) == (32, 1, 64, 1024)
    _close(img, gold, "cfg2/image", 1e-4)
    _close(img[0], gold, "cfg2/image_line0", 1e-4)


def test_oracle_bench_step_forward_16_lines(golden_dir):
    gold = np.load(f"{golden_dir}/full.npz")
    c = FULL["step16"]
    from handwriting_line_generation_b200 import DiscriminatorAP
    gsd, hsd = _gen_sd(c["gseed"]), _hwr_sd(c["hseed"])
    torch.manual_seed(c["dseed"])
    dsd = synth.perturb_disc(DiscriminatorAP(64, use_low=True, use_med=True).state_dict(), c["dseed"] + 1)
    content, style = synth.gen_case(c["T"], c["B"], 80, 128, c["iseed"])
    noise = [torch.from_numpy(z) for z in synth.gen_noise(synth.gen_noise_shapes(c["T"], c["B"]), c["iseed"] + 7)]
    masks = {k: torch.from_numpy(v) for k, v in synth.disc_masks(c["B"], c["iseed"] + 8).items()}
    label = torch.from_numpy(full_labels(c["B"], c["S"], 80, c["iseed"] + 1))
    with torch.no_grad():
        img = ogen.generator_forward(gsd, torch.from_numpy(content), torch.from_numpy(style), noise)
        lp = ohwr.hwr_forward(hsd, img, True, None)
        preds = odisc.disc_forward(dsd, img, masks, training=True)
        T = c["T"] - 6
        ctc = torch.nn.functional.ctc_loss(lp, label.permute(1, 0), torch.IntTensor([T] * c["B"]), torch.IntTensor([c["S"]] * c["B"]))
        adv = odisc.gen_loss(preds)
    _close(img, gold, "step16/image", 1e-4)
    _close(lp, gold, "step16/log_probs", 1e-4)
    for i, pr in enumerate(preds):
        assert float((pr - torch.from_numpy(gold[f"step16/pred{i}"])).abs().max()) <= 1e-4 * float(np.abs(gold[f"step16/pred{i}"]).max())
    assert abs(ctc.item() - float(gold["step16/ctc"])) <= 1e-5 * float(gold["step16/ctc"])
    assert abs(adv.item() - float(gold["step16/adv"])) <= 1e-4 * abs(float(gold["step16/adv"])) + 1e-6


def test_oracle_config5_two_lines_of_the_long_line_case(golden_dir):
    """configs[4]: lines are independent in the generator, so the first two of the 64 long lines pin the oracle on
    64x2048-px geometry in a second; the whole batch (recognizer, CTC with 120-char targets) runs in the GPU test."""
    gold = np.load(f"{golden_dir}/full.npz")
    c = FULL["cfg5"]
    sd = _gen_sd(c["gseed"], c["C"])
    assert abs(weights_digest(sd) - gold["cfg5/gen_weights_digest"]) <= 1e-6 * abs(gold["cfg5/gen_weights_digest"])
    content, style = synth.gen_case(c["T"], c["B"], c["C"], 128, c["iseed"])
    noise = [torch.from_numpy(z[:2].copy()) for z in synth.gen_noise(synth.gen_noise_shapes(c["T"], c["B"]), c["iseed"] + 7)]
    with torch.no_grad():
        img = ogen.generator_forward(sd, torch.from_numpy(content[:, :2].copy()), torch.from_numpy(style[:2].copy()), noise)
    assert tuple(img.shape) == (2, 1, 64, 2048)
    _close(img, gold, "cfg5/image_lines0_1", 1e-4)
