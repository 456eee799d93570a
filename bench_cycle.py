"""The reference curriculum's 7-lesson cycle (configs/cf_IAMslant_*sMG.json:85-96: count, [no-step gen], [auto, auto-gen], disc,
[no-step gen], [auto, auto-gen], disc) with EVERY module of the HWWithStyle model on the library's drop-ins — generator,
recognizer, CTC, discriminator, perceptual encoder, DTW alignment, spacer + insert_spaces, style extractor — written the way
`HWWithStyleTrainer.run_gen` / `_train_iteration` drive them (trainer/hw_with_style_trainer.py:514-830, :274-391), on
synthetic IAM-shaped data.  SURVEY.md §8d: "lines/sec ... reported per lesson type and for the 7-lesson cycle".

Unlike bench.py's headline step this is driven EAGERLY, call by call: the 'auto' and 'count' lessons have data-dependent
shapes (DTW path lengths, windows per character class), so they are not one fixed CUDA graph; the numbers include the host
time of ~330 ('gen') to ~1500 ('auto') library calls per lesson.

  python bench_cycle.py [--B 16] [--cycles 5]        (one GPU; bench.py reports it under extra_workloads.cycle)
"""
import argparse
import json
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

W_AUTO, W_PERC, W_COUNT, W_RECON_RECOG, W_GEN_RECOG, W_DISC, W_GEN = 0.5, 0.5, 0.5, 1e-6, 1e-4, 1.0, 1.0   # config :54-62
BALANCE_VAR_X = [0.6, 0.5, 0.4, 0.75]                                                                      # config :100
PADDING_CONSTANT = -1.0
CURRICULUM = ["count", "gen", "auto", "disc", "gen", "auto", "disc"]


class GanCycle:
    def __init__(self, dev, B=16, a_batch=2, W=1024, L=40, C=80, seed=0):
        import handwriting_line_generation_b200 as pkg
        from handwriting_line_generation_b200 import dtw, spacing
        import bench_inputs as synth                   # input builders (numpy)
        self.pkg, self.dev, self.B, self.a, self.W, self.L, self.C = pkg, dev, B, a_batch, W, L, C
        self.dtw, self.spacing = dtw, spacing
        torch.manual_seed(seed)
        self.gen = pkg.SpacedGenerator(C, 128, 256, n_style_trans=6, emb_dropout=False, append_style=True, small=False).to(dev).train()
        self.hwr = pkg.CNNOnlyHWR(C, norm='batch').to(dev).train()
        self.disc = pkg.DiscriminatorAP(64, use_low=True, use_med=True).to(dev).train()
        self.enc = pkg.Encoder2(32).to(dev).train()
        self.spacer = pkg.CountCNN(C, 128, 128, 2).to(dev).train()
        self.style = pkg.CharStyleEncoder(1, 64, 128, 128, 0, 'group', 'relu', 'replicate', C, global_pool=True,
                                          average_found_char_style=1.0, num_final_g_spacing_style=1, num_char_fc=1, vae=False,
                                          window=2, small=False).to(dev).train()
        for m in (self.hwr, self.enc):                 # hwr_frozen; the encoder is in no optimizer
            for p in m.parameters():
                p.requires_grad_(False)
        pkg.set_retain_graph(True)                     # several backward passes per graph (trainer :303-325)
        # the trainer's two optimizers (base_trainer.py:61-102): generator side = generator + style extractor + spacer
        gparams = list(self.gen.parameters()) + list(self.style.parameters()) + list(self.spacer.parameters())
        self.opt = pkg.FlatAdam(gparams, lr=2e-4, betas=(0.5, 0.999), clip_value=2.0)
        self.gen._grad_sink = self.opt
        self.opt_d = pkg.FlatAdam([p for p in self.disc.parameters() if p.requires_grad], lr=2e-4, betas=(0.5, 0.999),
                                  clip_value=2.0)
        self.host = types_ns(count_std=1e-8, dup_std=1e-9, count_duplicates=True, num_class=C)
        r = np.random.RandomState(seed + 1)
        self.batches = []
        for i in range(4):
            image = torch.from_numpy(synth.hwr_case(B, W, 100 + i)).to(dev)
            label = torch.from_numpy(r.randint(1, C, (L, B)).astype(np.int64)).to(dev)
            fg = (torch.from_numpy(r.rand(B, 1, 64, W).astype(np.float32)) > 0.5).float().to(dev)
            self.batches.append(dict(image=image, label=label, lengths=[L] * B, fg_mask=fg))
        self.prev_styles = [torch.randn(128) for _ in range(100)]
        self.i = 0

    # -- pieces of HWWithStyle --------------------------------------------------------------------------------------
    def onehot(self, label):
        return torch.zeros(label.size(0), label.size(1), self.C, device=self.dev).scatter_(2, label[..., None].long(), 1.0)

    def set_disc_trainable(self, flag):
        for n, p in self.disc.named_parameters():
            if not n.endswith(("weight_u", "weight_v")):
                p.requires_grad_(flag)

    def extract_style(self, image, pred):
        """hw_with_style.py:281-301 (`use_hwr_pred_for_style`): the lines of one author side by side."""
        B, a = image.size(0), self.a
        spaced = pred.permute(1, 2, 0)                                                   # [B,C,T]
        ci = image.permute(1, 2, 0, 3).contiguous().view(1, 64, B // a, image.size(3) * a).permute(2, 0, 1, 3).contiguous()
        cl = spaced.permute(1, 0, 2).contiguous().view(self.C, B // a, spaced.size(2) * a).permute(1, 0, 2).contiguous()
        style = self.style(ci, cl)
        return style.repeat_interleave(a, 0)

    def generate(self, label, lengths, style):
        """HWWithStyle.forward (:232-268): spacer -> insert_spaces -> generator."""
        with torch.no_grad():
            counts = self.spacer(self.onehot(label), style.detach())
        spaced, _ = self.spacing.insert_spaces(self.host, label, lengths, counts)
        return self.gen(spaced, style)

    def style_gen(self, B):
        """trainer :974-982: interpolated / extrapolated pairs from the bank of earlier styles."""
        idx = np.random.randint(0, len(self.prev_styles), (B, 2))
        mix = np.random.uniform(-0.5, 1.5, B)
        return torch.stack([self.prev_styles[i] * m + self.prev_styles[j] * (1 - m) for (i, j), m in zip(idx, mix)], 0).to(self.dev)

    def adversarial(self, fake):
        preds = self.disc(fake)
        return -(W_GEN / len(preds)) * sum(p.mean() for p in preds)

    # -- lessons ---------------------------------------------------------------------------------------------------
    def lesson_count(self, bt):
        image, label = bt["image"], bt["label"]
        with torch.no_grad():
            pred = self.hwr(image)
        style = self.extract_style(image, pred)
        spaced = self.dtw.correct_pred(pred, label)                                      # [T',B]
        counts = self.spacer(self.onehot(label), style)
        # gt_counts from the aligned label (trainer :670-697), on the host as the trainer computes them
        idx = spaced.cpu().numpy()
        lab = label.cpu().numpy()
        gt = np.zeros((label.size(0), label.size(1), 2), np.float32)
        for b in range(idx.shape[1]):
            c = d = pos = last = 0
            for i in range(idx.shape[0]):
                v = int(idx[i, b])
                if v == 0 and last == 0:
                    c += 1
                elif last == 0 or last == v:
                    d += 1
                    last = v
                else:
                    gt[pos, b] = (c, d)
                    c, d = (1, 0) if v == 0 else (0, 1)
                    pos += 1
                    last = v
        loss = W_COUNT * F.mse_loss(counts, torch.from_numpy(gt).to(self.dev))
        loss.backward()
        self.opt.step()
        return loss

    def lesson_gen(self, bt):
        """["no-step", "gen"]: two stashed gradient sets (trainer :312-338)."""
        label, lengths = bt["label"], bt["lengths"]
        B = label.size(1)
        self.set_disc_trainable(False)
        img = self.generate(label, lengths, self.style_gen(B))
        T = img.size(3) // 4 - 6
        il = torch.full((B,), T, dtype=torch.int32, device=self.dev)
        tl = torch.tensor(lengths, dtype=torch.int32, device=self.dev)
        recog = W_GEN_RECOG * self.pkg.CTCLoss(self.hwr(img), label.permute(1, 0), il, tl)
        adv = self.adversarial(img)
        recog.backward(retain_graph=True)
        self.opt.stash()
        adv.backward()
        self.opt.stash()
        return recog.detach() + adv.detach()

    def lesson_auto(self, bt):
        """["auto", "auto-gen"]: reconstruction through the style extractor; generator loss and reconRecog stashed, auto +
        perceptual as the main gradient, balancing of the four sets, clip + Adam (trainer :300-391)."""
        image, label, lengths, fg = bt["image"], bt["label"], bt["lengths"], bt["fg_mask"]
        B = image.size(0)
        self.set_disc_trainable(False)
        with torch.no_grad():
            pred = self.hwr(image)
        style = self.extract_style(image, pred)
        spaced = self.onehot(self.dtw.correct_pred(pred, label))
        recon = self.gen(spaced, style)                                                  # autoencode (:270-279)
        for k in range(0, B, self.a):                                                    # the style bank (:548-551)
            self.prev_styles.append(style[k].detach().cpu())
        self.prev_styles = self.prev_styles[-100:]
        img_a, rec_a = image, recon
        if recon.size(3) > image.size(3):                                                # :587-598
            img_a = F.pad(image, (0, recon.size(3) - image.size(3)), value=PADDING_CONSTANT)
            fg = F.pad(fg, (0, recon.size(3) - image.size(3)), value=0)
        elif recon.size(3) < image.size(3):
            rec_a = F.pad(recon, (0, image.size(3) - recon.size(3)), value=PADDING_CONSTANT)
        auto = W_AUTO * F.l1_loss(rec_a * fg, img_a * fg)                                # no_bg_loss (:596-605)
        perc = W_PERC * self.enc.perceptual_loss(image, recon)                           # :724-748
        T = recon.size(3) // 4 - 6
        il = torch.full((B,), T, dtype=torch.int32, device=self.dev)
        tl = torch.tensor(lengths, dtype=torch.int32, device=self.dev)
        recog = W_RECON_RECOG * self.pkg.CTCLoss(self.hwr(recon), label.permute(1, 0), il, tl)     # :752-757
        adv = self.adversarial(recon)                                                    # 'auto-gen': fake = recon (:783-784)
        adv.backward(retain_graph=True)                                                  # :303-311
        self.opt.stash()
        recog.backward(retain_graph=True)                                                # :312-323
        self.opt.stash()
        (auto + perc).backward()                                                         # :329
        self.opt.balance(BALANCE_VAR_X)                                                  # :340-377
        self.opt.step()                                                                  # :381-391
        return auto.detach() + perc.detach()

    def lesson_disc(self, bt):
        image, label, lengths = bt["image"], bt["label"], bt["lengths"]
        B = image.size(0)
        self.set_disc_trainable(True)
        with torch.no_grad():
            fake = self.generate(label, lengths, self.style_gen(B))
        if fake.size(3) > image.size(3):                                                 # :789-795
            image = F.pad(image, (0, fake.size(3) - image.size(3), 0, 0), 'replicate')
        elif fake.size(3) < image.size(3):
            fake = F.pad(fake, (0, image.size(3) - fake.size(3), 0, 0), 'replicate')
        preds = self.disc(torch.cat((image, fake), 0))
        loss = W_DISC * sum(F.relu(1.0 - p[:B]).mean() + F.relu(1.0 + p[B:]).mean() for p in preds) / len(preds)
        loss.backward()
        self.opt_d.step()
        return loss

    def run_lesson(self, name):
        bt = self.batches[self.i % len(self.batches)]
        self.i += 1
        return getattr(self, "lesson_" + name)(bt)


def types_ns(**kw):
    import types
    return types.SimpleNamespace(**kw)


def measure(dev, B=16, cycles=5, warmup=2):
    cyc = GanCycle(dev, B)
    pkg = cyc.pkg
    per = {n: [] for n in set(CURRICULUM)}
    calls = {}
    losses = {}
    for c in range(warmup + cycles):
        for name in CURRICULUM:
            torch.cuda.synchronize()
            n0 = pkg._lib.launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            loss = cyc.run_lesson(name)
            e1.record()
            torch.cuda.synchronize()
            ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)     # host-bound lessons: wall clock counts
            if c >= warmup:
                per[name].append(ms)
            calls[name] = pkg._lib.launch_count() - n0
            losses[name] = float(loss.detach())
    med = {n: float(np.median(v)) for n, v in per.items()}
    cycle_ms = sum(med[n] for n in CURRICULUM)
    assert all(np.isfinite(v) for v in losses.values()), losses
    return {"cycle_B%d" % B: {
        "what": "the reference curriculum's 7-lesson cycle (count, gen, auto+auto-gen, disc, gen, auto+auto-gen, disc) with every "
                "module on the library's drop-ins incl. spacer, DTW and style extractor, driven eagerly (data-dependent shapes), "
                f"{B} lines of 64x1024 px per lesson, synthetic data",
        "ms_per_lesson": med, "library_launches_per_lesson": calls, "ms_per_cycle": cycle_ms,
        "lines_per_s": 7 * B / cycle_ms * 1e3, "last_losses": losses}}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=16)
    ap.add_argument("--cycles", type=int, default=5)
    a = ap.parse_args()
    sys.path.insert(0, ".")
    torch.cuda.set_device(0)
    print(json.dumps(measure(torch.device("cuda", 0), a.B, a.cycles)))
