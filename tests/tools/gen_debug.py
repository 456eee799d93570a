"""Per-layer comparison of the CUDA generator with the CPU oracle (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import gen as ogen, synth
from oracle.make_golden import GEN_CASES
from tests.test_modules_cpu import _gen_module

name = sys.argv[1] if len(sys.argv) > 1 else "dense"
T, B, dense, wseed, iseed = GEN_CASES[name]
m, sd = _gen_module(wseed)
m = m.cuda().eval()
content, style = synth.gen_case(T, B, 80, 128, iseed, dense)
noise = synth.gen_noise(synth.gen_noise_shapes(T, B), iseed + 7)
trace = []
with torch.no_grad():
    ref = ogen.generator_forward(sd, torch.from_numpy(content), torch.from_numpy(style), [torch.from_numpy(z) for z in noise], trace=trace)
    out, saved = m._forward_impl(torch.from_numpy(content).cuda(), torch.from_numpy(style).cuda(),
                                 [torch.from_numpy(z).cuda() for z in noise], keep=True)
pre = [t for k, t in trace if k == "pre_adain"]
post = [t for k, t in trace if k == "post_adain"]
def rel(a, b): return ((a - b).abs().max() / b.abs().max()).item(), ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()
for i, (x, y, st, coef) in enumerate(saved):
    yy = y.float().permute(0, 3, 1, 2).cpu()
    # NOTE: y was normalised in place by scale_shift_act unless it is the last layer -> compare to post
    tgt = post[i] if i < len(saved) - 1 else pre[i]
    n = st.shape[0]; C = st.shape[1]; HW = yy.shape[2] * yy.shape[3]
    mean_ref = pre[i].mean((2, 3)); 
    mean_got = (st[:, :, 0] / HW).cpu()
    var_ref = pre[i].var((2, 3), unbiased=False); var_got = (st[:, :, 1] / HW).cpu() - mean_got ** 2
    print(i, tuple(yy.shape), "act max/rms rel", rel(yy, tgt), "mean", rel(mean_got, mean_ref), "var", rel(var_got, var_ref))
print("out", rel(out.cpu(), ref))
