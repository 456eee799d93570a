"""Device timing of the recognizer train step (config 1 shapes) — development aid."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import handwriting_line_generation_b200 as pkg
import bench_inputs as synth
B, W, S = int(os.environ.get("B", 8)), 1024, 60
torch.manual_seed(0)
m = pkg.CNNOnlyHWR(80, norm='batch').cuda().train()
opt = torch.optim.Adam(m.parameters(), lr=1e-4)
img = torch.from_numpy(synth.hwr_case(B, W, 1)).cuda()
T = W // 4 - 6
tg = torch.randint(1, 80, (B, S), dtype=torch.int32).cuda()
il = torch.full((B,), T, dtype=torch.int32); tl = torch.full((B,), S, dtype=torch.int32)
def step():
    opt.zero_grad(set_to_none=True)
    lp = m(img)
    loss = pkg.CTCLoss(lp, tg, il, tl)
    loss.backward()
    opt.step()
    return loss
def fwd():
    with torch.no_grad():
        return m(img)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
n0 = pkg._lib.launch_count()
l = step(); torch.cuda.synchronize()
print("launches per step", pkg._lib.launch_count() - n0, "loss", l.item())
tf = 1.0 if os.environ.get('ONLY_STEP') else timeit(fwd)
ts = timeit(step)
gf_fwd = 24.661 * B
print(f"B={B}: forward {tf:.3f} ms ({gf_fwd / tf:.1f} TFLOP/s eff), train step {ts:.3f} ms = {B / ts * 1e3:.0f} lines/s ({3 * gf_fwd / ts:.1f} TFLOP/s eff on 3x fwd FLOPs)")
