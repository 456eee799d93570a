"""TEST / BASELINE INFRASTRUCTURE — the balanced optimizer step of the HWWithStyle GAN curriculum written with stock PyTorch
(torch autograd over the fp32 restatements oracle/gen.py, oracle/hwr.py, oracle/disc.py, oracle/enc.py, `F.ctc_loss`,
oracle/balance.py, `clip_grad_value_`, `torch.optim.Adam`) — the reference's own arithmetic for the step bench.py times:

  lesson 1, 'gen' (no-step; trainer/hw_with_style_trainer.py:577, :760-764, :810-821)
      image = generator(spaced text, style); genRecog = 1e-4 * CTC(hwr(image)); generator loss = -mean D(image)
      genRecog.backward(retain_graph=True) -> stash (:312-323);  generator loss .backward() -> stash (:326-338)
  lesson 2, the perceptual part of 'auto' (:724-748)
      recon = generator(...); 0.5 * L1 between Encoder2 features of [real lines ; recon]; backward
  balance the two stashed sets into that gradient per tensor (:340-377, balance_var_x[:2]), clip_grad_value_(2) (:381),
  Adam lr 2e-4 betas (0.5, 0.999) (configs/cf_IAM*.json:35-46).

Runs on any torch device: on the host cores it is bench.py's `cpu_baseline` / `--impl reference` arm, on `cuda` (cuDNN, TF32
allowed — torch defaults for convolutions, as the reference would run) it is the `gpu_baseline`.  Weights are random-init
from the reference's state_dict key/shape fixtures (oracle/synth.random_state_dict): neither /root/reference nor the
product package is imported.  Never part of the product path."""
import numpy as np
import torch
import torch.nn.functional as F

from . import balance as obal
from . import disc as odisc
from . import enc as oenc
from . import gen as ogen
from . import hwr as ohwr
from . import synth

W_CTC, W_GEN, W_PERC = 1e-4, 1.0, 0.5       # loss_weights genRecog / generator / perceptual (config json :54-62)
BALANCE_VAR_X = [0.6, 0.5]                  # config :100, the entries of the two sets stashed here


def balance_nosync(main, saved_sets, multipliers):
    """oracle/balance.py without its data-dependent Python branches (same arithmetic through torch.where), so that the
    step can be captured in a CUDA graph for the graphed variant of the gpu baseline."""
    means = torch.stack([g.abs().mean() for g in main])
    nz = means != 0
    fill = (means * nz).sum() / nz.sum().clamp_min(1)
    means = torch.where(nz, means, fill)
    for x, saved in zip(multipliers, saved_sets):
        r = torch.stack([R.abs().mean() for R in saved])
        coef = torch.where(r != 0, x * means / torch.where(r != 0, r, torch.ones_like(r)), torch.zeros_like(r))
        torch._foreach_add_(main, torch._foreach_mul(list(saved), list(coef.unbind())))
    return main


class PortStep:
    def __init__(self, device, B, Ts=256, C=80, S=40, style_dim=128, dim=256, seed=0, capturable=False, step="balanced"):
        self.dev, self.B, self.Ts, self.kind = torch.device(device), B, Ts, step
        dev = self.dev
        self.gsd = {k: v.to(dev) for k, v in synth.random_state_dict("gen", seed, C).items()}
        # trainable = the 64 parameters; the 8 blur kernels (`conv.{1,2}.conv1.2`, `conv.{3,4}.conv1.1` weight / weight_flip,
        # [C,1,3,3]) are buffers and `gen.*` aliases `conv.*`
        blur = {f"conv.{i}.conv1.{j}.{n}" for i, j in ((1, 2), (2, 2), (3, 1), (4, 1)) for n in ("weight", "weight_flip")}
        self.trainable = [k for k, v in self.gsd.items() if v.is_floating_point() and not k.startswith("gen.") and k not in blur]
        assert len(self.trainable) == 64, len(self.trainable)
        for k in self.trainable:
            self.gsd[k].requires_grad_(True)
        for k in list(self.gsd):                     # `self.gen = self.conv` aliases (pure_gen.py:40)
            if k.startswith("gen."):
                self.gsd[k] = self.gsd["conv." + k[4:]]
        self.params = [self.gsd[k] for k in self.trainable]
        self.hsd = {k: v.to(dev) for k, v in synth.random_state_dict("hwr", seed + 1, C).items()}
        self.dsd = {k: v.to(dev) for k, v in synth.random_state_dict("disc", seed + 2).items()}
        self.esd = {k: v.to(dev) for k, v in synth.random_state_dict("enc", seed + 3).items()}
        self.opt = torch.optim.Adam(self.params, lr=2e-4, betas=(0.5, 0.999), capturable=capturable)
        content, style = synth.gen_case(Ts, B, C, style_dim, 3)
        self.c, self.s = torch.from_numpy(content).to(dev), torch.from_numpy(style).to(dev)
        self.real = torch.from_numpy(synth.hwr_case(B, 4 * Ts, 9)).to(dev)
        self.shapes = synth.gen_noise_shapes(Ts, B, dim)
        T = Ts - 6
        self.tg = torch.from_numpy(np.random.RandomState(7).randint(1, C, (B, S)).astype(np.int32)).to(dev)
        # tuples, not tensors: ATen's CUDA CTC then needs no device->host copy of the lengths
        self.il, self.tl = (T,) * B, (S,) * B
        self.nosync = capturable

    def _gen(self):
        noise = [torch.randn(sh, device=self.dev) for sh in self.shapes]     # the reference draws its noise in forward
        return ogen.generator_forward(self.gsd, self.c, self.s, noise)

    def _grads(self):
        return [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]

    def _stash(self):
        saved = [g.clone() for g in self._grads()]
        for p in self.params:
            if p.grad is not None:
                p.grad.zero_()
        return saved

    def __call__(self):
        B = self.B
        img = self._gen()
        upd = {}
        lp = ohwr.hwr_forward(self.hsd, img, True, upd)
        recog = W_CTC * F.ctc_loss(lp, self.tg, self.il, self.tl)
        masks = {site: (torch.rand(B, cm * 64, device=self.dev) >= p).float() for site, p, cm in synth.DISC_SITES}
        dupd = {}
        adv = W_GEN * odisc.gen_loss(odisc.disc_forward(self.dsd, img, masks, training=True, update=dupd))
        if self.kind != "balanced":                  # round-1 'gen'-lesson step: one backward over the summed losses
            (recog + adv).backward()
            self._finish(upd, dupd)
            return recog.detach() + adv.detach()
        recog.backward(retain_graph=True)
        set1 = self._stash()
        adv.backward()
        set2 = self._stash()
        emasks = [(torch.rand(2 * B, ch, device=self.dev) >= p).float() for _, ch, p in oenc.DROPOUT_SITES]
        perc = W_PERC * oenc.perceptual_loss(self.esd, self.real, self._gen(), emasks, training=True)
        perc.backward()
        main = self._grads()
        (balance_nosync if self.nosync else obal.balance)(main, [set1, set2], BALANCE_VAR_X)
        self._finish(upd, dupd)
        return recog.detach() + adv.detach() + perc.detach()

    def _finish(self, upd, dupd):
        with torch.no_grad():
            for k, v in upd.items():                 # BatchNorm running statistics
                if k in self.hsd and torch.is_tensor(v):
                    self.hsd[k].copy_(v)
            for k, v in dupd.items():                # the spectral-norm vectors advance on every forward
                self.dsd[k].copy_(v)
        torch.nn.utils.clip_grad_value_(self.params, 2)
        self.opt.step()
        self.opt.zero_grad(set_to_none=False)
