"""TEST INFRASTRUCTURE — runs the UNMODIFIED reference trainer's 'gen' lesson (HWWithStyleTrainer._train_iteration,
trainer/hw_with_style_trainer.py:210-400, lesson ["no-step","gen"] of the shipped IAM GAN curriculum) on CPU and writes
tests/golden/trainer_gen.npz: what the generator was fed (spaced content, style, noise, Dropout2d masks, labels), the
two weighted losses, and the two gradient sets the trainer stashes for its balancing step (recognition loss and
adversarial loss, trainer :300-338), as digests over the generator's parameters.

    python -m oracle.make_trainer_golden          (build container only: needs /root/reference)

The recipe is SURVEY.md Appendix A: stubs for the modules that are not installed, config edits that only touch IO, a
text-only instance.  Nothing on the hot path is modified; torch.randn_like and F.dropout2d are wrapped to RECORD what
they drew."""
import json
import os
import random
import sys
import tempfile

import numpy as np
import torch

from . import ref_shim, synth
from .make_golden import GOLD, digest

CFG = "configs/cf_IAMslant_noMask_charSpecSingleAppend_GANMedMT_autoAEMoPrcp2tightNewCTCUseGen_balB_hCF0.75_sMG.json"
B = 2              # lines (tiny: the CPU trainer step takes seconds)
L = int(os.environ.get("HWG_TRAINER_L", 15))          # characters per line
LABEL_SEED = int(os.environ.get("HWG_TRAINER_LABEL_SEED", 5))
SEEDS = dict(generator=400, hwr=401, discriminator=402, noise=403, masks=404, real=405)   # what the tests rebuild the inputs from


class _Log:
    def add_entry(self, *a, **k):
        pass

    def __getattr__(self, n):
        return lambda *a, **k: None


def run_lesson(kind, install=None, hook=None):
    """kind: 'gen' (curriculum slot 1, ["no-step","gen"]) or 'disc' (slot 3, ["disc"]).
    install / hook (tests only: the same unmodified trainer with the drop-in modules): `install()` runs before the model
    is built (the import swap of INTEGRATION.md), `hook(trainer, model, rec)` right before `_train_iteration`; with a hook
    nothing is written and (trainer, log, rec, model) is returned."""
    ref_shim.install()
    cwd = os.getcwd()
    os.chdir(ref_shim.REF)
    try:
        if install is not None:
            install()
        import torch.nn.functional as F
        from model import HWWithStyle
        import model.loss as mloss
        from model.autoencoder import Encoder2
        from trainer import HWWithStyleTrainer
        cfg = json.load(open(CFG))
        tmp = tempfile.mkdtemp(prefix="hwg_trainer_")
        torch.manual_seed(0)
        np.random.seed(0)
        random.seed(0)
        enc = Encoder2(32)
        torch.save({"state_dict": {"encoder." + k: v for k, v in enc.state_dict().items()}}, os.path.join(tmp, "enc.pth"))
        cfg["cuda"] = False
        cfg["model"]["pretrained_hwr"] = None
        cfg["trainer"].update(save_dir=tmp, print_dir=None, encoder_weights=os.path.join(tmp, "enc.pth"))
        model = HWWithStyle(cfg["model"])
        loss = {n: getattr(mloss, l) if hasattr(mloss, l) else eval(l, vars(mloss)) for n, l in cfg["loss"].items()}
        # The three hot-path modules get weights a test can rebuild WITHOUT the reference: each is re-initialised from
        # its own seed by constructing the reference class alone (the drop-ins mirror that construction, tested).
        from model.pure_gen import SpacedGenerator
        from model.cnn_only_hwr import CNNOnlyHWR
        from model.discriminator_ap import DiscriminatorAP
        torch.manual_seed(SEEDS["generator"])
        model.generator.load_state_dict(SpacedGenerator(80, 128, 256, n_style_trans=6, emb_dropout=False,
                                                        append_style=True, small=False).state_dict())
        torch.manual_seed(SEEDS["hwr"])
        model.hwr.load_state_dict(CNNOnlyHWR(80, norm='batch').state_dict())
        torch.manual_seed(SEEDS["discriminator"])
        dref = DiscriminatorAP(64, use_low=True, use_med=True)
        model.discriminator.load_state_dict(synth.perturb_disc(dref.state_dict(), SEEDS["discriminator"] + 1))
        tr = HWWithStyleTrainer(model, loss, [], None, cfg, None, None, _Log())

        r = np.random.RandomState(LABEL_SEED)
        label = torch.from_numpy(r.randint(1, 80, (L, B)).astype(np.int64)).int()
        lengths = torch.IntTensor([L] * B)

        class Text:
            max_len = 20

            def getInstance(self):
                return {"image": None, "label": label.clone(), "label_lengths": lengths.clone(), "gt": ["x" * L] * B,
                        "spaced_label": None, "author": ["a"] * B, "name": ["n"] * B}

        tr.text_data = Text()
        real = torch.from_numpy(synth.hwr_case(B, 128, SEEDS["real"]))      # 'disc' lesson: the real lines

        class Loader:                      # trainer :230 uses the py2 iterator protocol
            def next(self):
                inst = Text().getInstance()
                inst.update(image=real.clone(), fg_mask=torch.ones_like(real), a_batch_size=2)
                return inst

        tr.data_loader_iter = Loader()
        # ---- recorders
        from . import disc as odisc
        dmasks = {k: torch.from_numpy(v) for k, v in synth.disc_masks(B, SEEDS["masks"]).items()}
        rec = {"noise": [], "masks": [], "gen_in": None}
        orig_randn_like, orig_drop = torch.randn_like, F.dropout2d

        def randn_like(x, **kw):           # NoiseInjection (pure_gen.py:206,212): seeded numpy noise, one tensor per call
            z = torch.from_numpy(np.random.RandomState(SEEDS["noise"] + len(rec["noise"])).standard_normal(tuple(x.shape))
                                 .astype(np.float32))
            rec["noise"].append(tuple(x.shape))
            return z

        def dropout2d(x, p=0.5, training=True, inplace=False):
            if not training:
                return x
            if x.dim() == 4:                 # the discriminator's six sites, in forward order: injected synth masks
                site = odisc.DROPOUT_ORDER[len(rec["masks"])]
                keep = dmasks[site]
                assert abs(p - odisc.DROPOUT_P[site]) < 1e-9 and keep.shape == (x.size(0), x.size(1))
                rec["masks"].append(site)
            else:                            # the spacer's 1-D Dropout2d calls (only shape the recorded content)
                keep = (torch.rand(x.size(0), x.size(1)) >= p).float()
            return x * (keep / (1.0 - p)).view(x.size(0), x.size(1), *([1] * (x.dim() - 2)))

        gen_fwd = model.generator.forward

        def gen_forward(content, style, *a, **k):
            rec["gen_in"] = (content.detach().clone(), style.detach().clone())
            out = gen_fwd(content, style, *a, **k)
            rec["gen_out"] = out.detach().clone()
            return out

        model.generator.forward = gen_forward
        torch.randn_like, F.dropout2d = randn_like, dropout2d
        if kind == "balance":
            return _balance_golden(tr, model, (orig_randn_like, orig_drop), F, tmp)
        slot = {"gen": 1, "disc": 3}[kind]
        if kind == "disc":                  # 2B rows go through the discriminator: masks for 2B samples
            dmasks.update({k: torch.from_numpy(v) for k, v in synth.disc_masks(2 * B, SEEDS["masks"]).items()})
        if hook is not None:
            hook(tr, model, rec)
        try:
            tr.iteration = slot
            log = tr._train_iteration(slot)
        finally:
            torch.randn_like, F.dropout2d = orig_randn_like, orig_drop
            model.generator.forward = gen_fwd
        if hook is not None:
            return tr, log, rec, model
        out = {"content": rec["gen_in"][0].numpy(), "style": rec["gen_in"][1].numpy(), "image": rec["gen_out"].numpy(),
               "label": label.numpy(), "label_lengths": lengths.numpy(),
               "noise_shapes": np.array(rec["noise"], np.int64), "mask_sites": np.array(rec["masks"]),
               "modes": np.array([int(model.generator.training), int(model.hwr.training), int(model.discriminator.training)]),
               "seeds": np.array([SEEDS[k] for k in ("generator", "hwr", "discriminator", "noise", "masks", "real")], np.int64),
               "loss_keys": np.array(sorted(k for k in log if "Loss" in k))}
        names = [n for n, _ in model.named_parameters()]
        if kind == "gen":
            assert len(tr.saved_grads) == 2 and len(names) == len(tr.parameters)
            out["losses"] = np.array([log.get("genRecogLoss", np.nan), log.get("generatorLoss", np.nan)], np.float64)
            out["disc_u_after"] = model.discriminator.state_dict()["convs1.0.module.weight_u"].numpy().copy()
            for si, setname in enumerate(("recog", "adv")):
                for n, g in zip(names, tr.saved_grads[si]):
                    if g is None or not n.startswith("generator."):
                        continue
                    dig, samp = digest(g.numpy())
                    out[f"grad/{setname}/{n}/digest"] = dig
                    out[f"grad/{setname}/{n}/sample"] = samp[:256]
            fname = "trainer_gen.npz"
        else:
            # the trainer has clipped (clip_grad_value_ 2, :381) and stepped optimizer_discriminator; .grad still holds
            # the clipped gradients of this iteration
            out["losses"] = np.array([log["discriminatorLoss"]], np.float64)
            for n, p in model.named_parameters():
                if n.startswith("discriminator.") and p.grad is not None:
                    dig, samp = digest(p.grad.numpy())
                    out[f"grad/disc/{n}/digest"] = dig
                    out[f"grad/disc/{n}/sample"] = samp[:256]
            assert all(p.grad is None or float(p.grad.abs().max()) == 0 for n, p in model.named_parameters()
                       if n.startswith("generator."))             # fake.detach(): nothing reaches the generator
            fname = "trainer_disc.npz"
        print(kind, "losses", out["losses"], "noise tensors", len(rec["noise"]), "dropout sites", len(rec["masks"]),
              "content", out["content"].shape, "image", out["image"].shape)
        np.savez_compressed(os.path.join(GOLD, fname), **out)
        return tmp
    finally:
        os.chdir(cwd)


def _balance_golden(tr, model, originals, F, tmp):
    """Lessons 1 -> 2 of the curriculum (["no-step","gen"] then ["auto","auto-gen"]): the second one balances the four
    stashed gradient sets into p.grad, clips and steps (trainer :340-391).  Recorded for every parameter tensor with at
    most 4096 elements: the four sets, p.grad before balancing (after the last backward) and after balancing (on entry
    to clip_grad_value_), plus the statistics that involve all tensors."""
    torch.randn_like, F.dropout2d = originals            # the auto lesson draws in modules outside the hot path
    tr.iteration = 1
    tr._train_iteration(1)
    assert len(tr.saved_grads) == 2
    snaps = []
    orig_backward = torch.Tensor.backward
    orig_clip = torch.nn.utils.clip_grad_value_

    def backward(self, *a, **k):
        r = orig_backward(self, *a, **k)
        snaps.append([None if p.grad is None else p.grad.detach().clone() for p in tr.parameters])
        return r

    post = {}

    def clip(params, v):
        post["grads"] = [None if p.grad is None else p.grad.detach().clone() for p in tr.parameters]
        post["sets"] = len(tr.saved_grads)
        return orig_clip(params, v)

    saved_before = None
    torch.Tensor.backward = backward
    torch.nn.utils.clip_grad_value_ = clip
    try:
        tr.iteration = 2
        # the stash is emptied inside the call: keep a reference to the list objects
        stash = tr.saved_grads
        tr._train_iteration(2)
    finally:
        torch.Tensor.backward = orig_backward
        torch.nn.utils.clip_grad_value_ = orig_clip
    # inside the call: backward #1 = autoGenLoss (stashed, zeroed), #2 = recogLoss (stashed, zeroed), #3 = the rest -> D
    assert len(snaps) == 3, len(snaps)
    sets = list(stash[:2]) + [snaps[0], snaps[1]]          # order of self.saved_grads at balancing time
    D, after = snaps[2], post["grads"]
    names = [n for n, _ in model.named_parameters()]
    mult = None
    for it, m in tr.balance_var_x.items():
        if int(it) <= 2:
            mult = m
    from .balance import abs_means
    means, fill = abs_means(D)
    out = {"multipliers": np.array(mult, np.float64), "fill": np.float64(fill), "names": []}
    keep = [i for i, p in enumerate(tr.parameters) if p.numel() <= 4096 and D[i] is not None]
    # a spread over the model's parts, and every tensor whose own mean|D| is exactly 0 (the :354-359 rule)
    pick = [i for i in keep if float(means[i]) == 0][:8]
    for prefix in ("generator.", "style_extractor.", "spacer.", "hwr.", "discriminator."):
        pick += [i for i in keep if names[i].startswith(prefix)][:10]
    pick = sorted(set(pick))
    for i in pick:
        n = names[i]
        out["names"].append(n)
        out[f"D/{n}"] = D[i].numpy()
        out[f"after/{n}"] = after[i].numpy()
        out[f"meanD/{n}"] = np.float64(means[i])
        for k, st in enumerate(sets):
            if st[i] is not None:
                out[f"R{k}/{n}"] = st[i].numpy()
    out["names"] = np.array(out["names"])
    out["n_zero_mean"] = np.int64(sum(1 for i in pick if float(means[i]) == 0))
    print("balance: multipliers", mult, "tensors stored", len(pick), "with zero mean|D|", int(out["n_zero_mean"]), "fill", fill.item())
    np.savez_compressed(os.path.join(GOLD, "trainer_balance.npz"), **out)
    return tmp


if __name__ == "__main__":
    for k in (sys.argv[1:] or ["gen", "disc", "balance"]):
        run_lesson(k)
