import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import gen as ogen, synth
from oracle.make_golden import GEN_CASES
from tests.test_modules_cpu import _gen_module
from tests.test_modules_gpu import rel_l2
name = sys.argv[1] if len(sys.argv) > 1 else "small"
T, B, dense, wseed, iseed = GEN_CASES[name]
m, sd = _gen_module(wseed)
sd = {k: v.clone() for k, v in sd.items()}
m = m.cuda().train()
content, style = synth.gen_case(T, B, 80, 128, iseed, True)
noise = synth.gen_noise(synth.gen_noise_shapes(T, B), iseed + 7)
R = torch.randn(B, 1, 64, 4 * T, generator=torch.Generator().manual_seed(2))
c = torch.from_numpy(content).cuda().requires_grad_()
s = torch.from_numpy(style).cuda().requires_grad_()
img = m(c, s, noise=[torch.from_numpy(z).cuda() for z in noise])
(img * R.cuda()).sum().backward()
p = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
oc = torch.from_numpy(content).requires_grad_(); os_ = torch.from_numpy(style).requires_grad_()
oimg = ogen.generator_forward(p, oc, os_, [torch.from_numpy(z) for z in noise])
(oimg * R).sum().backward()
print("img", rel_l2(img.detach().cpu().numpy(), oimg.detach().numpy()))
def show(n, g, o):
    g, o = g.cpu().numpy(), o.numpy()
    print(f"{n:34s} rel {rel_l2(g, o):.4f} |got| {np.abs(g).max():.3e} |ref| {np.abs(o).max():.3e} cos {float((g*o).sum()/np.sqrt((g*g).sum()*(o*o).sum()+1e-30)):.4f}")
show("content", c.grad, oc.grad); show("style", s.grad, os_.grad)
for n, q in m.named_parameters():
    if n.startswith("gen."): continue
    if q.grad is None: print(n, "NO GRAD"); continue
    show(n, q.grad, p[n].grad)
