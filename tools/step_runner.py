"""Runs a few steps of one workload (development aid; run it under `ncu --metrics gpu__time_duration.sum` for a launch
list, or bare for an eager/graphed device time).

  python tools/step_runner.py {gen_infer|hwr_train|gen_train|gan_step} [--B n] [--steps k] [--graph]
gan_step = bench.py's headline step (bench_gan_train.GanStep; HWG_BENCH_STEP=balanced|gen_only)
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import handwriting_line_generation_b200 as pkg
from handwriting_line_generation_b200 import graphs
import bench_inputs as synth  # input builders (numpy)

ap = argparse.ArgumentParser()
ap.add_argument("workload")
ap.add_argument("--B", type=int, default=16)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--graph", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
B, Ts, W = a.B, 256, 1024
torch.manual_seed(0)

if a.workload == "gen_infer":
    gen = pkg.SpacedGenerator(80, 128, 256, n_style_trans=6, emb_dropout=False, append_style=True).to(dev).eval()
    content, style = synth.gen_case(Ts, B, 80, 128, 3)
    ins = [torch.from_numpy(content).to(dev), torch.from_numpy(style).to(dev)]
    mods = [gen]

    def step(c, s):
        with torch.no_grad():
            return gen(c, s)
elif a.workload == "hwr_train":
    hwr = pkg.CNNOnlyHWR(80, norm='batch').to(dev).train()
    opt = pkg.FlatAdam(hwr.parameters(), lr=1e-4, betas=(0.9, 0.999))
    hwr._grad_sink = opt
    T, S = W // 4 - 6, 60
    ins = [torch.from_numpy(synth.hwr_case(B, W, 1)).to(dev),
           torch.randint(1, 80, (B, S), dtype=torch.int32, device=dev)]
    il = torch.full((B,), T, dtype=torch.int32, device=dev)
    tl = torch.full((B,), S, dtype=torch.int32, device=dev)
    mods = [hwr]

    def step(img, tg):
        loss = pkg.CTCLoss(hwr(img), tg, il, tl)
        loss.backward()
        opt.step()
        return loss
elif a.workload == "gen_train":
    gen = pkg.SpacedGenerator(80, 128, 256, n_style_trans=6, emb_dropout=False, append_style=True).to(dev).train()
    hwr = pkg.CNNOnlyHWR(80, norm='batch').to(dev).train()
    for p in hwr.parameters():
        p.requires_grad_(False)
    disc = None
    if not os.environ.get("HWG_BENCH_NO_DISC"):
        disc = pkg.DiscriminatorAP(64, use_low=True, use_med=True).to(dev).train()
        for p in disc.parameters():
            p.requires_grad_(False)
    opt = pkg.FlatAdam(gen.parameters(), lr=2e-4, betas=(0.5, 0.999), clip_value=2.0)
    gen._grad_sink = opt
    content, style = synth.gen_case(Ts, B, 80, 128, 3)
    T, S = Ts - 6, 40
    ins = [torch.from_numpy(content).to(dev), torch.from_numpy(style).to(dev),
           torch.randint(1, 80, (B, S), dtype=torch.int32, device=dev)]
    il = torch.full((B,), T, dtype=torch.int32, device=dev)
    tl = torch.full((B,), S, dtype=torch.int32, device=dev)
    mods = [gen, hwr]

    def step(c, s, tg):
        img = gen(c, s)
        loss = 1e-4 * pkg.CTCLoss(hwr(img), tg, il, tl)
        if disc is not None:
            preds = disc(img)
            loss = loss - (1.0 / len(preds)) * sum(p.mean() for p in preds)
        loss.backward()
        opt.step()
        return loss
elif a.workload == "gan_step":
    import bench_gan_train as bg
    st = bg.GanStep(dev, B)
    content, style = synth.gen_case(Ts, B, 80, 128, 3)
    ins = [torch.from_numpy(content).to(dev), torch.from_numpy(style).to(dev),
           torch.randint(1, 80, (B, bg.GAN["S"]), dtype=torch.int32, device=dev),
           torch.from_numpy(synth.hwr_case(B, 4 * Ts, 9)).to(dev)]
    mods = [st.gen, st.hwr]
    step = st.train
else:
    raise SystemExit("unknown workload")

if a.graph:
    fn = graphs.GraphedStep(step, ins, modules=mods, warmup=a.warmup)
else:
    fn = step
    for _ in range(a.warmup):
        fn(*ins)
torch.cuda.synchronize()
n0 = pkg._lib.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    out = fn(*ins)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
print(f"{a.workload} B={B} graph={a.graph}: {ms:.3f} ms/step, {B / ms * 1e3:.1f} lines/s, "
      f"{(pkg._lib.launch_count() - n0) // a.steps} hwg launches/step (host path)", flush=True)
