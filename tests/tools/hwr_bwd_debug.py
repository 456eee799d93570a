import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import hwr as ohwr, synth
from tests.test_modules_cpu import _hwr_module
from tests.test_modules_gpu import rel_l2
B, W = 2, 128
m, sd = _hwr_module(200)
sd = {k: v.clone() for k, v in sd.items()}
m = m.cuda().train()
img = synth.hwr_case(B, W, 31)
T = W // 4 - 6
R = torch.randn(T, B, 80, generator=torch.Generator().manual_seed(1))
lp = m(torch.from_numpy(img).cuda())
(lp * R.cuda()).sum().backward()
p = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
olp = ohwr.hwr_forward(p, torch.from_numpy(img), True, None, emulate_bf16=('emu' in sys.argv))
(olp * R).sum().backward()
print("lp err", rel_l2(lp.detach().cpu().numpy(), olp.detach().numpy()))
for n, q in m.named_parameters():
    g, o = q.grad.cpu().numpy(), p[n].grad.numpy()
    print(f"{n:28s} rel {rel_l2(g, o):.4f}  |got| {np.abs(g).max():.3e} |ref| {np.abs(o).max():.3e}  cos {float((g*o).sum()/np.sqrt((g*g).sum()*(o*o).sum()+1e-30)):.4f}")
