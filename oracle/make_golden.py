"""TEST INFRASTRUCTURE — writes tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, CPU, torch 2.11.0) on deterministic synthetic inputs.

Run in the build container:  python -m oracle.make_golden [ctc|hwr|gen|all]
The GPU box has no /root/reference: tests there read the committed fixtures only.
Inputs come from oracle/synth.py (numpy RandomState, reproducible everywhere), so the
full-size cases store only outputs (loss, nll, gradient digests), not the inputs."""
import os
import sys

import numpy as np
import torch

from . import ref_shim, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

# name -> (T, B, C, S, seed, ragged)
CTC_CASES = {
    "small": (20, 3, 7, 5, 11, True),
    "cfg1": (250, 8, 80, 60, 12, False),      # BASELINE.json configs[0]: IAM charset, S=60
    "cfg1_ragged": (250, 8, 80, 90, 13, True),
    "cfg5": (506, 64, 78, 120, 14, False),    # configs[4]: RIMES charset, S=120, W=2048
    "odd_c": (64, 5, 37, 9, 15, True),        # C not a multiple of 2: 4-byte cp.async path
    "empty_target": (30, 4, 10, 6, 16, True),  # one sequence with target length 0
    "repeats": (40, 4, 6, 12, 17, True),      # few classes -> many repeated labels
}


def grad_digest(g):
    """Position-weighted digests + a strided sample: enough to pin a [T,B,C] gradient."""
    flat = g.reshape(-1).astype(np.float64)
    w = np.cos(np.arange(flat.size) * 0.37) + 1.5
    idx = np.arange(0, flat.size, max(1, flat.size // 4096))
    return np.array([flat.sum(), np.abs(flat).sum(), (flat * w).sum(), np.abs(flat).max()]), flat[idx].astype(np.float32)


def make_ctc():
    ref_shim.install()
    from model.loss import CTCLoss  # the reference's wrapper, model/loss.py:28
    from utils.string_utils import naive_decode  # utils/string_utils.py:51
    out = {}
    for name, (T, B, C, S, seed, ragged) in CTC_CASES.items():
        lp, tg, il, tl = synth.ctc_case(T, B, C, S, seed, ragged)
        if name == "empty_target":
            tl[1] = 0
            tg[1, :] = 0
        lpt = torch.from_numpy(lp).requires_grad_()
        # the trainer's call pattern: label [S,B] int32 -> label.permute(1,0); CPU IntTensors
        label = torch.from_numpy(np.ascontiguousarray(tg.T))
        loss = CTCLoss(lpt, label.permute(1, 0), torch.from_numpy(il), torch.from_numpy(tl))
        (loss * 1.0).backward()
        nll = torch.nn.functional.ctc_loss(lpt.detach(), label.permute(1, 0), torch.from_numpy(il),
                                           torch.from_numpy(tl), reduction="none")
        g = lpt.grad.numpy()
        dig, samp = grad_digest(g)
        out[f"{name}/loss"] = np.float32(loss.item())
        out[f"{name}/nll"] = nll.numpy().astype(np.float32)
        out[f"{name}/grad_digest"] = dig
        out[f"{name}/grad_sample"] = samp
        # fp64 run of the same call: measures the fp32 rounding noise of the reference itself
        lp64 = torch.from_numpy(lp).double().requires_grad_()
        torch.nn.functional.ctc_loss(lp64, label.permute(1, 0), torch.from_numpy(il),
                                     torch.from_numpy(tl)).backward()
        _, samp64 = grad_digest(lp64.grad.numpy())
        out[f"{name}/grad_fp32_noise"] = np.float64(np.abs(samp.astype(np.float64) - samp64).max())
        if T * B * C <= 20000:
            out[f"{name}/grad"] = g
        dec = [naive_decode(lp[:, b, :])[0] for b in range(B)]
        out[f"{name}/decoded_len"] = np.array([len(d) for d in dec], np.int32)
        out[f"{name}/decoded"] = np.array([x for d in dec for x in d], np.int32)
        print(f"ctc/{name}: T={T} B={B} C={C} S={S} loss={loss.item():.6f}")
    np.savez_compressed(os.path.join(GOLD, "ctc.npz"), **out)


# name -> (T, B, dense content, weight seed, input seed)
GEN_CASES = {
    "tiny": (8, 2, False, 100, 101),
    "small": (32, 3, False, 100, 102),
    "dense": (24, 2, True, 100, 103),
    "odd_T": (37, 2, False, 104, 105),   # width not a multiple of any tile
}
GEN_ARGS = dict(n_class=80, style_size=128, dim=256, n_style_trans=6, emb_dropout=False, append_style=True,
                small=False)

# name -> (B, W, weight seed, input seed, training)
HWR_CASES = {
    "train_w128": (2, 128, 200, 201, True),
    "train_w260": (3, 260, 200, 202, True),   # W/4+1 etc. not multiples of the tile
    "eval_w128": (2, 128, 200, 203, False),
}


def digest(t):
    """Digests + a strided sample of a float tensor (full tensor when small)."""
    flat = t.reshape(-1).astype(np.float64)
    w = np.cos(np.arange(flat.size) * 0.37) + 1.5
    idx = np.arange(0, flat.size, max(1, flat.size // 8192))
    return np.array([flat.sum(), np.abs(flat).sum(), (flat * w).sum(), np.abs(flat).max()]), flat[idx].astype(np.float32)


def weights_digest(sd):
    """One number per state_dict: catches any difference in init order or key naming."""
    tot = 0.0
    for i, k in enumerate(sorted(sd)):
        v = sd[k].double().reshape(-1)
        tot += float((v * torch.cos(torch.arange(v.numel(), dtype=torch.float64) * 0.11 + i)).sum())
    return np.float64(tot)


def keys_fixture(sd):
    return np.array([f"{k}:{'x'.join(str(d) for d in v.shape)}" for k, v in sd.items()])


def make_gen():
    ref_shim.install()
    from model.pure_gen import SpacedGenerator
    out = {}
    for name, (T, B, dense, wseed, iseed) in GEN_CASES.items():
        m, sd = synth.state_dict_from_seed(lambda: SpacedGenerator(80, 128, 256, n_style_trans=6, emb_dropout=False,
                                                                   append_style=True, small=False), wseed)
        m.eval()
        out[f"{name}/weights_digest"] = weights_digest(sd)
        out["state_dict_keys"] = keys_fixture(sd)
        content, style = synth.gen_case(T, B, 80, 128, iseed, dense)
        noise = [torch.from_numpy(z) for z in synth.gen_noise(synth.gen_noise_shapes(T, B), iseed + 7)]
        it = iter(noise)
        orig = torch.randn_like

        def fake_randn_like(x, **kw):
            z = next(it)
            assert z.shape == x.shape, (z.shape, x.shape)
            return z

        torch.randn_like = fake_randn_like   # the reference draws its noise here (pure_gen.py:206,212)
        try:
            with torch.no_grad():
                img = m(torch.from_numpy(content), torch.from_numpy(style))
        finally:
            torch.randn_like = orig
        dig, samp = digest(img.numpy())
        out[f"{name}/digest"], out[f"{name}/sample"] = dig, samp
        out[f"{name}/shape"] = np.array(img.shape)
        if img.numel() <= 40000:
            out[f"{name}/image"] = img.numpy()
        print(f"gen/{name}: T={T} B={B} -> {tuple(img.shape)} absmax {dig[3]:.4f}")
    np.savez_compressed(os.path.join(GOLD, "gen.npz"), **out)


def make_hwr():
    ref_shim.install()
    from model.cnn_only_hwr import CNNOnlyHWR
    out = {}
    for name, (B, W, wseed, iseed, training) in HWR_CASES.items():
        m, sd = synth.state_dict_from_seed(lambda: CNNOnlyHWR(80, norm='batch'), wseed)
        m.train(training)
        out[f"{name}/weights_digest"] = weights_digest(sd)
        out["state_dict_keys"] = keys_fixture(sd)
        img = synth.hwr_case(B, W, iseed)
        with torch.no_grad():
            lp = m(torch.from_numpy(img))
        dig, samp = digest(lp.numpy())
        out[f"{name}/digest"], out[f"{name}/sample"] = dig, samp
        out[f"{name}/shape"] = np.array(lp.shape)
        if lp.numel() <= 40000:
            out[f"{name}/log_probs"] = lp.numpy()
        out[f"{name}/argmax"] = lp.argmax(2).numpy().astype(np.int32)
        sd2 = m.state_dict()
        for k in ("cnn.batchnorm2.running_mean", "cnn.batchnorm6.running_var", "cnn1d.10.running_mean",
                  "cnn1d.1.running_var"):
            out[f"{name}/{k}"] = sd2[k].numpy()
        print(f"hwr/{name}: B={B} W={W} train={training} -> {tuple(lp.shape)}")
    np.savez_compressed(os.path.join(GOLD, "hwr.npz"), **out)


# name -> (B, W, weight seed, input seed, training)
DISC_CASES = {
    "train_w128": (2, 128, 300, 301, True),
    "train_w264": (3, 264, 300, 302, True),    # 264 -> 132 -> 66 -> 33 -> 16 -> 8: AvgPool floors, odd tile widths
    "eval_w128": (2, 128, 300, 303, False),
}


# 'disc' lesson: hinge loss on real || fake rows, gradients of every trainable parameter.  name -> (B, W, wseed, iseed)
DISC_LESSON_CASES = {"hinge_w128": (4, 128, 300, 304), "hinge_w200": (2, 200, 300, 305)}


def make_disc():
    """Outputs, generator-loss input gradient and updated spectral-norm vectors of the unmodified reference
    DiscriminatorAP (IAM GAN config: dim 64, 'use low', med on).  Dropout2d keep-masks are injected by patching
    torch.nn.functional.dropout2d (what nn.Dropout2d.forward calls)."""
    ref_shim.install()
    from model.discriminator_ap import DiscriminatorAP
    import torch.nn.functional as F
    from . import disc as odisc
    out = {}
    for name, (B, W, wseed, iseed, training) in DISC_CASES.items():
        torch.manual_seed(wseed)
        m = DiscriminatorAP(64, use_low=True, use_med=True)
        sd = synth.perturb_disc(m.state_dict(), wseed + 1)
        m.train(training)
        out[f"{name}/weights_digest"] = weights_digest(sd)
        out["state_dict_keys"] = keys_fixture(sd)
        masks = {k: torch.from_numpy(v) for k, v in synth.disc_masks(B, iseed + 7).items()}
        order = iter(odisc.DROPOUT_ORDER)
        orig = F.dropout2d

        def fake_dropout2d(x, p=0.5, training=True, inplace=False):
            if not training:
                return x
            site = next(order)
            assert abs(p - odisc.DROPOUT_P[site]) < 1e-9 and x.size(1) == masks[site].size(1), (site, p, x.shape)
            return x * (masks[site] / (1.0 - p))[:, :, None, None]

        F.dropout2d = fake_dropout2d
        try:
            img = torch.from_numpy(synth.hwr_case(B, W, iseed)).requires_grad_()
            preds = m(img)
            # the generator's adversarial loss, trainer/hw_with_style_trainer.py:810-821
            loss = 0
            for gp in preds:
                loss = loss - gp.mean()
            loss = loss / len(preds)
            loss.backward()
        finally:
            F.dropout2d = orig
        for i, pr in enumerate(preds):
            out[f"{name}/pred{i}"] = pr.detach().numpy()
        out[f"{name}/loss"] = np.float32(loss.item())
        dig, samp = digest(img.grad.numpy())
        out[f"{name}/grad_digest"], out[f"{name}/grad_sample"] = dig, samp
        sd2 = m.state_dict()
        for k in ("convs1.0.module.weight_u", "convs3.4.module.weight_v", "convs4.14.module.weight_u"):
            out[f"{name}/{k}"] = sd2[k].numpy()
        print(f"disc/{name}: B={B} W={W} train={training} -> {[tuple(p.shape) for p in preds]} loss {loss.item():.5f}")
    for name, (B, W, wseed, iseed) in DISC_LESSON_CASES.items():
        torch.manual_seed(wseed)
        m = DiscriminatorAP(64, use_low=True, use_med=True)
        synth.perturb_disc(m.state_dict(), wseed + 1)
        m.train()
        masks = {k: torch.from_numpy(v) for k, v in synth.disc_masks(B, iseed + 7).items()}
        order = iter(odisc.DROPOUT_ORDER)
        orig = F.dropout2d

        def fake_dropout2d(x, p=0.5, training=True, inplace=False):
            site = next(order)
            return x * (masks[site] / (1.0 - p))[:, :, None, None]

        F.dropout2d = fake_dropout2d
        try:
            preds = m(torch.from_numpy(synth.hwr_case(B, W, iseed)))
            # trainer/hw_with_style_trainer.py:797-804 (real rows first, then fake)
            loss = 0
            for pr in preds:
                loss = loss + F.relu(1.0 - pr[:B // 2]).mean() + F.relu(1.0 + pr[B // 2:]).mean()
            loss = loss / len(preds)
            loss.backward()
        finally:
            F.dropout2d = orig
        out[f"{name}/loss"] = np.float32(loss.item())
        names = [n for n, p in m.named_parameters() if p.requires_grad]
        out[f"{name}/param_names"] = np.array(names)
        for n, p in m.named_parameters():
            if p.requires_grad:
                dig, samp = digest(p.grad.numpy())
                out[f"{name}/grad/{n}/digest"], out[f"{name}/grad/{n}/sample"] = dig, samp[:512]
        print(f"disc/{name}: B={B} W={W} hinge loss {loss.item():.5f}, {len(names)} parameter gradients")
    np.savez_compressed(os.path.join(GOLD, "disc.npz"), **out)


# name -> (B, W, weight seed, input seed, training)
ENC_CASES = {"eval_w128": (2, 128, 500, 501, False), "train_w200": (2, 200, 500, 502, True)}


def make_enc():
    """Features and the perceptual-loss gradient (w.r.t. the reconstructed image) of the unmodified reference Encoder2(32)."""
    ref_shim.install()
    from model.autoencoder import Encoder2
    import torch.nn.functional as F
    from . import enc as oenc
    out = {}
    for name, (B, W, wseed, iseed, training) in ENC_CASES.items():
        torch.manual_seed(wseed)
        m = Encoder2(32)
        m.train(training)
        out["state_dict_keys"] = keys_fixture(m.state_dict())
        out[f"{name}/weights_digest"] = weights_digest(m.state_dict())
        r = np.random.RandomState(iseed + 7)
        masks = [torch.from_numpy((r.rand(2 * B, c) >= 3 * p).astype(np.float32)) for _, c, p in oenc.DROPOUT_SITES]
        it = iter(masks)
        orig = F.dropout2d

        def fake_dropout2d(x, p=0.5, training=True, inplace=False):
            if not training:
                return x
            k = next(it)
            assert k.shape == (x.size(0), x.size(1))
            return x * (k / (1.0 - p))[:, :, None, None]

        F.dropout2d = fake_dropout2d
        try:
            image = torch.from_numpy(synth.hwr_case(B, W, iseed))
            recon = torch.from_numpy(synth.hwr_case(B, W, iseed + 1)).requires_grad_()
            feats = m(torch.cat((image, recon), 0))                      # trainer :740-742
            loss = 0
            for f in feats:
                o_f, r_f = torch.chunk(f, 2, dim=0)
                loss = loss + F.l1_loss(r_f, o_f)
            loss.backward()
        finally:
            F.dropout2d = orig
        for i, f in enumerate(feats):
            dig, samp = digest(f.detach().numpy())
            out[f"{name}/feat{i}/digest"], out[f"{name}/feat{i}/sample"] = dig, samp[:2048]
            out[f"{name}/feat{i}/shape"] = np.array(f.shape)
        out[f"{name}/loss"] = np.float32(loss.item())
        dig, samp = digest(recon.grad.numpy())
        out[f"{name}/grad_digest"], out[f"{name}/grad_sample"] = dig, samp[:2048]
        print(f"enc/{name}: B={B} W={W} train={training} -> {[tuple(f.shape) for f in feats]} loss {loss.item():.5f}")
    np.savez_compressed(os.path.join(GOLD, "enc.npz"), **out)


# name -> (B, W, weight seed, input seed)
STYLE_CASES = {"b2_w256": (2, 256, 600, 601), "b3_w520": (3, 520, 600, 602)}
DTW_CASES = {"t58_l9": (58, 3, 9, 611), "t124_l30": (124, 2, 30, 612)}      # name -> (T, B, label length, seed)


def style_inputs(B, W, seed, n_class=80):
    """Image and a recognizer-like log-prob tensor [B, n_class, W/4-6]: spaced random text, sharpened, log-softmax."""
    image = torch.from_numpy(synth.hwr_case(B, W, seed))
    T = W // 4 - 6
    content, _ = synth.gen_case(T, B, n_class, 8, seed + 1)                               # [T,B,C] one-hot
    r = np.random.RandomState(seed + 2)
    logits = 5.0 * content + r.standard_normal(content.shape).astype(np.float32)
    recog = torch.log_softmax(torch.from_numpy(logits), 2).permute(1, 2, 0).contiguous()  # [B,C,T]
    return image, recog


def dtw_inputs(T, B, L, seed, n_class=80):
    r = np.random.RandomState(seed)
    pred = torch.log_softmax(torch.from_numpy(2.0 * r.standard_normal((T, B, n_class)).astype(np.float32)), 2)
    label = torch.from_numpy(r.randint(1, n_class, (L, B)).astype(np.int64))
    return pred, label


def make_style():
    """Style vectors of the unmodified reference CharStyleEncoder (IAM GAN configuration) and DTW alignments of the
    unmodified `correct_pred`."""
    ref_shim.install()
    from model.char_style import CharStyleEncoder
    from model.hw_with_style import correct_pred
    out = {}
    for name, (B, W, wseed, iseed) in STYLE_CASES.items():
        torch.manual_seed(wseed)
        m = CharStyleEncoder(1, 64, 128, 128, 0, 'group', 'relu', 'replicate', 80, global_pool=False,
                             average_found_char_style=1.0, num_final_g_spacing_style=1, num_char_fc=1, vae=False, window=2,
                             small=False)
        m.eval()
        out["state_dict_keys"] = keys_fixture(m.state_dict())
        out[f"{name}/weights_digest"] = weights_digest(m.state_dict())
        image, recog = style_inputs(B, W, iseed)
        with torch.no_grad():
            style = m(image, recog)
        out[f"{name}/style"] = style.numpy()
        print(f"style/{name}: B={B} W={W} -> {tuple(style.shape)} |style|max {style.abs().max():.4f}")
    for name, (T, B, L, seed) in DTW_CASES.items():
        pred, label = dtw_inputs(T, B, L, seed)
        out[f"dtw/{name}"] = correct_pred(pred, label).numpy()
        print(f"dtw/{name}: {out[f'dtw/{name}'].shape}")
    np.savez_compressed(os.path.join(GOLD, "style.npz"), **out)


# name -> (L, B, weight seed, input seed)
SPACER_CASES = {"l12_b3": (12, 3, 700, 701), "l40_b2": (40, 2, 700, 702)}


def spacer_inputs(L, B, seed, n_class=80):
    r = np.random.RandomState(seed)
    label = torch.from_numpy(r.randint(1, n_class, (L, B)).astype(np.int64))
    lengths = torch.IntTensor([L - (b % 3) for b in range(B)])
    style = torch.from_numpy(r.standard_normal((B, 128)).astype(np.float32))
    return label, lengths, style


def spacer_train_extras(L, B, seed):
    """Dropout2d keep-masks of the two sites ([B,128], [B,64]; drop rate inflated to 0.3 so small batches really drop
    channels) and the weights of the linear test loss."""
    r = np.random.RandomState(seed + 5)
    masks = [torch.from_numpy((r.rand(B, c) >= 0.3).astype(np.float32)) for c in (128, 64)]
    R = torch.from_numpy(r.standard_normal((L, B, 2)).astype(np.float32))
    return masks, R


def make_spacer():
    """Counts of the unmodified reference CountCNN and the spaced text of the unmodified `HWWithStyle.insert_spaces`
    (called unbound on a stand-in that carries the attributes it reads: no model weights are involved in it)."""
    ref_shim.install()
    import types
    from model.count_cnn import CountCNN
    from model.hw_with_style import HWWithStyle
    out = {}
    for name, (L, B, wseed, iseed) in SPACER_CASES.items():
        torch.manual_seed(wseed)
        m = CountCNN(80, 128, 128, 2).eval()
        out["state_dict_keys"] = keys_fixture(m.state_dict())
        out[f"{name}/weights_digest"] = weights_digest(m.state_dict())
        label, lengths, style = spacer_inputs(L, B, iseed)
        onehot = torch.zeros(L, B, 80).scatter_(2, label[..., None], 1.0)
        with torch.no_grad():
            counts = m(onehot, style)
        for std, tag in ((1e-8, "cfg"), (0.4, "noisy")):           # the config's count_std / dup_std, and a case where the draws matter
            host = types.SimpleNamespace(count_std=std, dup_std=std / 10, count_duplicates=True, num_class=80)
            np.random.seed(iseed)
            spaced, padded = HWWithStyle.insert_spaces(host, label, lengths, counts)
            out[f"{name}/{tag}/spaced"] = spaced.argmax(2).numpy().astype(np.int16)
            assert float(spaced.sum(2).min()) == 1.0 and float(spaced.max()) == 1.0
            out[f"{name}/{tag}/padded"] = np.array(padded, np.float64)
        out[f"{name}/counts"] = counts.numpy()
        print(f"spacer/{name}: counts {tuple(counts.shape)} spaced T={spaced.size(0)}")
    # train mode (the 'count' lesson trains the spacer): Dropout2d keep-masks injected, a linear loss, every gradient
    import torch.nn.functional as F
    for name, (L, B, wseed, iseed) in SPACER_CASES.items():
        torch.manual_seed(wseed)
        m = CountCNN(80, 128, 128, 2).train()
        label, lengths, style = spacer_inputs(L, B, iseed)
        style = style.clone().requires_grad_()
        onehot = torch.zeros(L, B, 80).scatter_(2, label[..., None], 1.0).requires_grad_()
        masks, R = spacer_train_extras(L, B, iseed)
        it = iter(masks)
        orig = F.dropout2d

        def fake_dropout2d(x, p=0.5, training=True, inplace=False):
            mk = next(it)
            assert training and abs(p - 0.1) < 1e-9 and tuple(mk.shape) == tuple(x.shape[:2]), (p, x.shape)
            return x * (mk / (1.0 - p))[:, :, None]

        F.dropout2d = fake_dropout2d
        try:
            counts = m(onehot, style)
        finally:
            F.dropout2d = orig
        (counts * R).sum().backward()
        out[f"{name}/train/counts"] = counts.detach().numpy()
        out[f"{name}/train/grad/style"] = style.grad.numpy()
        out[f"{name}/train/grad/input"] = onehot.grad.numpy()
        for n, p_ in m.named_parameters():
            out[f"{name}/train/grad/{n}"] = p_.grad.numpy()
        print(f"spacer/{name}: train-mode counts + {len(list(m.parameters()))} parameter gradients")
    np.savez_compressed(os.path.join(GOLD, "spacer.npz"), **out)


# ---- BASELINE-size cases (parity only at toy sizes) ---------------------------------------------------------
# cfg1: BASELINE configs[0] — CNNOnlyHWR + CTC fwd+bwd, B=8, 64x1024, 60-char targets
# cfg2: configs[1] — SpacedGenerator inference, B=32, T_s=256 -> [32,1,64,1024]
# step16: the bench step's forward at 16 lines: generator -> {recognizer + CTC, discriminator}
# cfg5: configs[4] — RIMES charset (78 classes), B=64, T_s=512 -> 64x2048 lines, recognizer, 120-char targets, CTC fwd+bwd
FULL = dict(cfg1=dict(B=8, W=1024, S=60, wseed=900, iseed=901),
            cfg2=dict(B=32, T=256, wseed=910, iseed=911),
            step16=dict(B=16, T=256, S=40, gseed=920, hseed=921, dseed=922, iseed=923),
            cfg5=dict(B=64, T=512, S=120, C=78, gseed=930, hseed=931, iseed=932))


def full_labels(B, S, C, seed):
    return np.random.RandomState(seed).randint(1, C, (S, B)).astype(np.int32)          # [S,B] like the trainer's `label`


def _ref_generate(m, content, style, noise):
    it = iter(noise)
    orig = torch.randn_like

    def fake_randn_like(x, **kw):
        z = next(it)
        assert z.shape == x.shape, (z.shape, x.shape)
        return z

    torch.randn_like = fake_randn_like
    try:
        return m(torch.from_numpy(content), torch.from_numpy(style))
    finally:
        torch.randn_like = orig


def _put(out, key, t, nsamp=8192):
    dig, samp = digest(t.detach().numpy() if torch.is_tensor(t) else t)
    out[key + "/digest"], out[key + "/sample"] = dig, samp[:nsamp]


def make_full():
    """Goldens of the UNMODIFIED reference modules at BASELINE.json's configuration sizes: digests + strided samples (the
    tensors themselves are tens of MB).  Written to tests/golden/full.npz."""
    ref_shim.install()
    from model.cnn_only_hwr import CNNOnlyHWR
    from model.discriminator_ap import DiscriminatorAP
    from model.loss import CTCLoss
    from model.pure_gen import SpacedGenerator
    import torch.nn.functional as F
    from . import disc as odisc
    out = {}
    torch.set_num_threads(os.cpu_count() or 1)

    def gen_module(seed, C=80):
        return synth.state_dict_from_seed(lambda: SpacedGenerator(C, 128, 256, n_style_trans=6, emb_dropout=False,
                                                                  append_style=True, small=False), seed)

    # ---- cfg1
    c = FULL["cfg1"]
    m, sd = synth.state_dict_from_seed(lambda: CNNOnlyHWR(80, norm='batch'), c["wseed"])
    m.train()
    out["cfg1/weights_digest"] = weights_digest(sd)
    x = torch.from_numpy(synth.hwr_case(c["B"], c["W"], c["iseed"])).requires_grad_()
    label = torch.from_numpy(full_labels(c["B"], c["S"], 80, c["iseed"] + 1))
    T = c["W"] // 4 - 6
    lp = m(x)
    loss = CTCLoss(lp, label.permute(1, 0), torch.IntTensor([T] * c["B"]), torch.IntTensor([c["S"]] * c["B"]))
    loss.backward()
    _put(out, "cfg1/log_probs", lp)
    out["cfg1/loss"] = np.float64(loss.item())
    _put(out, "cfg1/grad/input", x.grad)
    for n, p in m.named_parameters():
        _put(out, f"cfg1/grad/{n}", p.grad, 1024)
    out["cfg1/argmax"] = lp.argmax(2).numpy().astype(np.int16)
    print(f"full/cfg1: lp {tuple(lp.shape)} loss {loss.item():.5f}")

    # ---- cfg2
    c = FULL["cfg2"]
    m, sd = gen_module(c["wseed"])
    m.eval()
    out["cfg2/weights_digest"] = weights_digest(sd)
    content, style = synth.gen_case(c["T"], c["B"], 80, 128, c["iseed"])
    noise = [torch.from_numpy(z) for z in synth.gen_noise(synth.gen_noise_shapes(c["T"], c["B"]), c["iseed"] + 7)]
    with torch.no_grad():
        img = _ref_generate(m, content, style, noise)
    _put(out, "cfg2/image", img)
    _put(out, "cfg2/image_line0", img[0], 4096)
    del noise
    print(f"full/cfg2: image {tuple(img.shape)} absmax {float(img.abs().max()):.4f}")

    # ---- step16
    c = FULL["step16"]
    g, gsd = gen_module(c["gseed"])
    g.train()
    h, hsd = synth.state_dict_from_seed(lambda: CNNOnlyHWR(80, norm='batch'), c["hseed"])
    h.train()
    torch.manual_seed(c["dseed"])
    d = DiscriminatorAP(64, use_low=True, use_med=True)
    synth.perturb_disc(d.state_dict(), c["dseed"] + 1)
    d.train()
    for k, sdx in (("gen", gsd), ("hwr", hsd), ("disc", d.state_dict())):
        out[f"step16/{k}_weights_digest"] = weights_digest(sdx)
    content, style = synth.gen_case(c["T"], c["B"], 80, 128, c["iseed"])
    noise = [torch.from_numpy(z) for z in synth.gen_noise(synth.gen_noise_shapes(c["T"], c["B"]), c["iseed"] + 7)]
    masks = {k: torch.from_numpy(v) for k, v in synth.disc_masks(c["B"], c["iseed"] + 8).items()}
    label = torch.from_numpy(full_labels(c["B"], c["S"], 80, c["iseed"] + 1))
    order = iter(odisc.DROPOUT_ORDER)
    orig = F.dropout2d

    def fake_dropout2d(x, p=0.5, training=True, inplace=False):
        site = next(order)
        return x * (masks[site] / (1.0 - p))[:, :, None, None]

    F.dropout2d = fake_dropout2d
    try:
        with torch.no_grad():
            img = _ref_generate(g, content, style, noise)
            lp = h(img)
            preds = d(img)
            T = c["T"] - 6
            ctc = CTCLoss(lp, label.permute(1, 0), torch.IntTensor([T] * c["B"]), torch.IntTensor([c["S"]] * c["B"]))
            adv = -sum(p.mean() for p in preds) / len(preds)
    finally:
        F.dropout2d = orig
    _put(out, "step16/image", img)
    _put(out, "step16/log_probs", lp)
    for i, pr in enumerate(preds):
        out[f"step16/pred{i}"] = pr.numpy()
    out["step16/ctc"], out["step16/adv"] = np.float64(ctc.item()), np.float64(adv.item())
    del noise
    print(f"full/step16: ctc {ctc.item():.4f} adv {adv.item():.5f}")

    # ---- cfg5
    c = FULL["cfg5"]
    g, gsd = gen_module(c["gseed"], c["C"])
    g.eval()
    h, hsd = synth.state_dict_from_seed(lambda: CNNOnlyHWR(c["C"], norm='batch'), c["hseed"])
    h.train()
    out["cfg5/gen_weights_digest"], out["cfg5/hwr_weights_digest"] = weights_digest(gsd), weights_digest(hsd)
    content, style = synth.gen_case(c["T"], c["B"], c["C"], 128, c["iseed"])
    noise = [torch.from_numpy(z) for z in synth.gen_noise(synth.gen_noise_shapes(c["T"], c["B"]), c["iseed"] + 7)]
    label = torch.from_numpy(full_labels(c["B"], c["S"], c["C"], c["iseed"] + 1))
    with torch.no_grad():
        img = _ref_generate(g, content, style, noise)
    del noise
    lp = h(img).detach().requires_grad_()
    T = 4 * c["T"] // 4 - 6
    loss = CTCLoss(lp, label.permute(1, 0), torch.IntTensor([T] * c["B"]), torch.IntTensor([c["S"]] * c["B"]))
    loss.backward()
    _put(out, "cfg5/image", img)
    _put(out, "cfg5/image_lines0_1", img[:2], 4096)
    _put(out, "cfg5/log_probs", lp)
    out["cfg5/loss"] = np.float64(loss.item())
    _put(out, "cfg5/ctc_grad", lp.grad)
    out["cfg5/argmax"] = lp.argmax(2).numpy().astype(np.int16)
    print(f"full/cfg5: image {tuple(img.shape)} lp {tuple(lp.shape)} loss {loss.item():.4f}")
    np.savez_compressed(os.path.join(GOLD, "full.npz"), **out)


def main(argv):
    what = argv[1] if len(argv) > 1 else "all"
    os.makedirs(GOLD, exist_ok=True)
    if what in ("ctc", "all"):
        make_ctc()
    if what in ("hwr", "all") and "make_hwr" in globals():
        globals()["make_hwr"]()
    if what in ("gen", "all") and "make_gen" in globals():
        globals()["make_gen"]()
    if what in ("disc", "all"):
        make_disc()
    if what in ("enc", "all"):
        make_enc()
    if what in ("style", "all"):
        make_style()
    if what in ("spacer", "all"):
        make_spacer()
    if what in ("full", "all"):
        make_full()


if __name__ == "__main__":
    main(sys.argv)
