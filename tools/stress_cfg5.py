"""BASELINE configs[4] (long-line stress): width 2048, 120-char transcripts, RIMES charset (78 classes), batch 64 —
generation + recognition + CTC forward/backward through the generator (development aid; prints timings and checks
finiteness and the CTC loss against torch's CUDA implementation on the same log-probs)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import handwriting_line_generation_b200 as pkg
import bench_inputs as synth

dev = torch.device("cuda", 0)
B, Ts, C, S = 64, 512, 78, 120
torch.manual_seed(0)
gen = pkg.SpacedGenerator(C, 128, 256, n_style_trans=6, emb_dropout=False, append_style=True).to(dev).train()
hwr = pkg.CNNOnlyHWR(C, norm='batch').to(dev).train()
for p in hwr.parameters():
    p.requires_grad_(False)
opt = pkg.FlatAdam(gen.parameters(), lr=2e-4, betas=(0.5, 0.999), clip_value=2.0)
gen._grad_sink = opt
content, style = synth.gen_case(Ts, B, C, 128, 11)
c, s = torch.from_numpy(content).to(dev), torch.from_numpy(style).to(dev)
T = Ts - 6
tg = torch.from_numpy(np.random.RandomState(3).randint(1, C, (B, S)).astype(np.int32)).to(dev)
il = torch.full((B,), T, dtype=torch.int32, device=dev)
tl = torch.full((B,), S, dtype=torch.int32, device=dev)


def step():
    img = gen(c, s)
    lp = hwr(img)
    loss = pkg.CTCLoss(lp, tg, il, tl)
    loss.backward()
    return img, lp, loss


img, lp, loss = step()
torch.cuda.synchronize()
assert img.shape == (B, 1, 64, 4 * Ts) and lp.shape == (T, B, C), (img.shape, lp.shape)
ref = torch.nn.functional.ctc_loss(lp.detach(), tg.long(), il.long(), tl.long(), blank=0, reduction='mean')
assert torch.isfinite(loss) and abs(loss.item() - ref.item()) <= 1e-4 * abs(ref.item()), (loss.item(), ref.item())
assert all(torch.isfinite(p.grad).all() for p in gen.parameters())
gnorm = float(opt.flat_g.norm())
opt.step()
for _ in range(2):
    step(); opt.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    step(); opt.step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"cfg5 stress: B={B} width {4 * Ts} S={S} C={C}: loss {loss.item():.4f} (torch {ref.item():.4f}), |grad| {gnorm:.3e}, "
      f"{ms:.2f} ms/step eager = {B / ms * 1e3:.0f} lines/s, peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
