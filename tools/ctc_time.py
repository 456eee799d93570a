"""Quick device timing of the CTC kernels (development aid, not the bench)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import handwriting_line_generation_b200 as pkg
import bench_inputs as synth

def timeit(fn, n=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for a, b in evs:
        flush.zero_()
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2] * 1e3  # us

for (T, B, C, S) in [(250, 8, 80, 60), (506, 64, 78, 120), (506, 512, 78, 120), (250, 2048, 80, 60)]:
    lp, tg, il, tl = synth.ctc_case(T, B, C, S, 1, ragged=False)
    x = torch.from_numpy(lp).cuda().requires_grad_()
    tgt = torch.from_numpy(tg).cuda(); ilc = torch.from_numpy(il).cuda(); tlc = torch.from_numpy(tl).cuda()
    def ours():
        x.grad = None
        pkg.CTCLoss(x, tgt, ilc, tlc).backward()
    def aten():
        x.grad = None
        torch.nn.functional.ctc_loss(x, tgt, ilc, tlc).backward()
    def ours_fwd():
        with torch.no_grad(): pkg.CTCLoss(x, tgt, ilc, tlc)
    a, b, c = timeit(ours), timeit(aten), timeit(ours_fwd)
    ideal = 2 * T * B * C * 4
    print(f"T={T} B={B} C={C} S={S}: hwg fwd+bwd {a:.1f} us ({ideal / a / 1e3:.1f} GB/s ideal-bytes), "
          f"fwd-only {c:.1f} us, ATen CUDA fwd+bwd {b:.1f} us")
