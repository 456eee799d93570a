"""Development aid (CPU, build container): what limits the fidelity of the generator gradients on the reference trainer s own
'gen' lesson (tests/golden/trainer_gen.npz, 2 lines of 64x128 px) — bf16 rounding of the FORWARD activations (cosine 0.74 of
the recognition set with the fp32 chain) and not of the gradients between the layers (0.735 with, 0.737 without gradient
rounding); fp32 with inputs perturbed by 1e-3: 0.997.  python tests/tools/grad_sensitivity.py"""
import numpy as np, torch, sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from oracle import gen as ogen, hwr as ohwr, disc as odisc
from tests.test_trainer_gen_cpu import build_inputs
gold=np.load(__import__('os').path.join(sys.path[0],'tests','golden','trainer_gen.npz'))
gsd,hsd,dsd,content,style,noise,masks=build_inputs(gold)
label=torch.from_numpy(gold["label"]).int(); lengths=torch.from_numpy(gold["label_lengths"]).int()
def grads(emu, content=content, style=style, noise=noise):
    gp={k:v.clone().requires_grad_(v.is_floating_point()) for k,v in gsd.items()}
    img=ogen.generator_forward(gp,content,style,noise,emulate_bf16=emu)
    lp=ohwr.hwr_forward({k:v.clone() for k,v in hsd.items()},img,True,{},emulate_bf16=emu)
    B=style.size(0)
    recog=1e-4*torch.nn.functional.ctc_loss(lp,label.permute(1,0),torch.IntTensor([lp.size(0)]*B),lengths)
    adv=odisc.gen_loss(odisc.disc_forward(dsd,img,masks,training=True,update={},emulate_bf16=emu))
    names=[k for k,p in gp.items() if p.requires_grad]
    out={}
    for nm,loss in (("recog",recog),("adv",adv)):
        g=torch.autograd.grad(loss,[gp[k] for k in names],retain_graph=True,allow_unused=True)
        out[nm]=torch.cat([x.reshape(-1).double() for x in g if x is not None])
    return out
def cos(a,b): return float((a*b).sum()/(a.norm()*b.norm()))
ref=grads(False)
e1=grads(True)
print("emu fwd+grad rounding:", {k:round(cos(e1[k],ref[k]),3) for k in ref})
# gradient rounding off
for mod in (ogen,ohwr,odisc):
    mod._Q.backward=staticmethod(lambda ctx,g: g)
ohwr._QGradOnly.backward=staticmethod(lambda ctx,g: g)
e2=grads(True)
print("emu fwd rounding only:", {k:round(cos(e2[k],ref[k]),3) for k in ref})
for eps in (1e-3,1e-4,1e-5):
    torch.manual_seed(0)
    p=grads(False, style=style*(1+eps*torch.randn_like(style)))
    print(f"fp32, style perturbed by {eps:g} relative:", {k:round(cos(p[k],ref[k]),3) for k in ref})
for eps in (1e-3,):
    p=grads(False, noise=[z*(1+eps*torch.randn_like(z)) for z in noise])
    print(f"fp32, noise perturbed by {eps:g} relative:", {k:round(cos(p[k],ref[k]),3) for k in ref})
