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
SEEDS = dict(generator=400, hwr=401, discriminator=402, noise=403, masks=404)   # what the tests rebuild the inputs from


class _Log:
    def add_entry(self, *a, **k):
        pass

    def __getattr__(self, n):
        return lambda *a, **k: None


def main():
    ref_shim.install()
    cwd = os.getcwd()
    os.chdir(ref_shim.REF)
    try:
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
        try:
            tr.iteration = 1
            log = tr._train_iteration(1)            # curriculum slot 1: ["no-step", "gen"]
        finally:
            torch.randn_like, F.dropout2d = orig_randn_like, orig_drop
            model.generator.forward = gen_fwd
        assert len(tr.saved_grads) == 2, len(tr.saved_grads)
        names = [n for n, _ in model.named_parameters()]
        assert len(names) == len(tr.parameters)
        out = {"content": rec["gen_in"][0].numpy(), "style": rec["gen_in"][1].numpy(), "image": rec["gen_out"].numpy(),
               "label": label.numpy(), "label_lengths": lengths.numpy(),
               "losses": np.array([log.get("genRecogLoss", np.nan), log.get("generatorLoss", np.nan)], np.float64),
               "loss_keys": np.array(sorted(k for k in log if "Loss" in k))}
        out["noise_shapes"] = np.array(rec["noise"], np.int64)
        out["mask_sites"] = np.array(rec["masks"])
        out["modes"] = np.array([int(model.generator.training), int(model.hwr.training), int(model.discriminator.training)])
        out["seeds"] = np.array([SEEDS[k] for k in ("generator", "hwr", "discriminator", "noise", "masks")], np.int64)
        u = model.discriminator.state_dict()["convs1.0.module.weight_u"]
        out["disc_u_after"] = u.numpy().copy()
        for si, setname in enumerate(("recog", "adv")):
            for n, g in zip(names, tr.saved_grads[si]):
                if g is None or not n.startswith("generator."):
                    continue
                dig, samp = digest(g.numpy())
                out[f"grad/{setname}/{n}/digest"] = dig
                out[f"grad/{setname}/{n}/sample"] = samp[:256]
        print("losses", out["losses"], "noise tensors", len(rec["noise"]), "dropout sites", len(rec["masks"]),
              "content", out["content"].shape, "image", out["image"].shape)
        np.savez_compressed(os.path.join(GOLD, "trainer_gen.npz"), **out)
        return tmp
    finally:
        os.chdir(cwd)


if __name__ == "__main__":
    main()
