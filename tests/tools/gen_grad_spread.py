"""Run-to-run spread of the generator's gradient error against the fp32 oracle (development aid: the InstanceNorm statistics
are summed with fp32 atomics, so every forward differs in the last bits and LeakyReLU masks tip).  Prints, for a few runs of
tests/test_gen_train_gpu.py's 'small' case, rel-L2(cuda, fp32) next to rel-L2(bf16-emulated torch, fp32) for the tensors
with the smallest margin.  HWG_LIB_PATH selects the build."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import test_gen_train_gpu as T

name = "small"
Tt, B, _, wseed, iseed = T.GEN_CASES[name]
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 6
ref = None
for r in range(runs):
    m, sd = T._gen_module(wseed)
    sd = {k: v.clone() for k, v in sd.items()}
    m = m.cuda().train()
    content, style = T.synth.gen_case(Tt, B, 80, 128, iseed, True)
    noise = T.synth.gen_noise(T.synth.gen_noise_shapes(Tt, B), iseed + 7)
    R = torch.randn(B, 1, 64, 4 * Tt, generator=torch.Generator().manual_seed(2))
    c = torch.from_numpy(content).cuda().requires_grad_()
    s = torch.from_numpy(style).cuda().requires_grad_()
    img = m(c, s, noise=[torch.from_numpy(z).cuda() for z in noise])
    (img * R.cuda()).sum().backward()
    torch.cuda.synchronize()
    if ref is None:
        _, g32 = T._oracle(sd, content, style, noise, R, False)
        _, gemu = T._oracle(sd, content, style, noise, R, True)
        ref = (g32, gemu)
    g32, gemu = ref
    got = {n: p.grad.cpu() for n, p in m.named_parameters() if not n.startswith("gen.")}
    rows = []
    for n, g in g32.items():
        if g.numel() == 1 or n not in got:
            continue
        ours, emu = T.rel_l2(got[n].numpy(), g.numpy()), T.rel_l2(gemu[n].numpy(), g.numpy())
        k = 1.3 if g.numel() >= 256 else 2.5
        rows.append((k * emu + T.BF16_REL - ours, n, ours, emu))
    rows.sort()
    print(f"run {r}: " + "; ".join(f"{n} {o:.3f}/{e:.3f} (margin {mg:+.3f})" for mg, n, o, e in rows[:3]), flush=True)
