"""Device timing of the HBM-bound passes on train-step shapes (development aid).  Buffers rotate through a pool larger
than L2 so that every launch streams from HBM.   python tools/stream_bench.py [names...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from handwriting_line_generation_b200 import ops, _lib

dev = "cuda"
which = sys.argv[1:]
REPS = 12


def bf(*shape):
    return torch.randn(*shape, device=dev).to(torch.bfloat16)


def timed(name, fn, nsets, bytes_per_launch):
    for i in range(3):
        fn(i % nsets)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(REPS):
        fn(i % nsets)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / REPS
    print(f"{name:28s} {us:8.1f} us  {bytes_per_launch / us / 1e3:8.1f} GB/s (algorithmic bytes {bytes_per_launch / 1e6:.1f} MB)", flush=True)


def want(n):
    return not which or n in which


NS = 6   # 6 x (2-3 tensors of 33 MB) > 126 MB L2
if want("bn_bwd"):
    N, H, W, C = 16, 8, 257, 512
    z = [bf(N, H, W, C) for _ in range(NS)]
    g = [bf(N, H, W, C) for _ in range(NS)]
    coef = torch.randn(C, 2, device=dev)
    save = torch.rand(C, 2, device=dev) + 0.5
    wgt = torch.randn(C, device=dev)
    nb = z[0].numel() * 2
    sums = torch.zeros(C, 2, device=dev)
    gz = torch.empty_like(z[0])
    dcb = torch.zeros(C, device=dev)
    timed("bn_bwd_reduce 512ch", lambda i: _lib.call("hwg_bn_bwd_reduce", g[i].data_ptr(), z[i].data_ptr(), coef.data_ptr(),
          save.data_ptr(), N * H * W, C, 1, sums.data_ptr(), _lib.stream()), NS, 2 * nb)
    timed("bn_bwd_apply 512ch", lambda i: _lib.call("hwg_bn_bwd_apply", g[i].data_ptr(), z[i].data_ptr(), coef.data_ptr(),
          save.data_ptr(), wgt.data_ptr(), sums.data_ptr(), N * H * W, 0, C, 1, gz.data_ptr(), dcb.data_ptr(), _lib.stream()), NS, 3 * nb)
if want("adain_bwd"):
    for (N, H, W, C) in ((16, 64, 1024, 16), (16, 8, 256, 128)):
        a = [bf(N, H, W, C) for _ in range(NS)]
        g = [bf(N, H, W, C) for _ in range(NS)]
        save = torch.rand(N, C, 2, device=dev) + 0.5
        coef = torch.randn(N, C, 2, device=dev)
        nb = a[0].numel() * 2
        timed(f"adain_lrelu_bwd {C}ch (2 passes)", lambda i: ops.adain_lrelu_bwd(g[i], a[i], save, coef, 0.2, None, 1, 0), NS, 5 * nb)
if want("scale_shift"):
    N, H, W, C = 16, 16, 256, 256
    x = [bf(N, H, W, C) for _ in range(NS)]
    out = torch.empty_like(x[0])
    coef = torch.randn(C, 2, device=dev)
    timed("scale_shift_act 256ch", lambda i: ops.scale_shift_act(x[i], coef, False, _lib.ACT_RELU, out=out), NS, 2 * x[0].numel() * 2)
if want("maxpool_bwd"):
    for name, (N, H, W, C), geom in (("2x2", (16, 32, 512, 128), ((2, 2), (2, 2), (0, 0))), ("2x2/2x1", (16, 16, 256, 256), ((2, 2), (2, 1), (0, 1)))):
        c = [torch.relu(bf(N, H, W, C)) for _ in range(NS)]
        k, s, p = geom
        Ho, Wo = (H + 2 * p[0] - k[0]) // s[0] + 1, (W + 2 * p[1] - k[1]) // s[1] + 1
        ga = bf(N, Ho, Wo, C)
        timed(f"relu_maxpool_bwd {name}", lambda i: ops.relu_maxpool_bwd(ga, c[i], k, s, p), NS, 2 * c[0].numel() * 2 + ga.numel() * 2)
if want("blur"):
    N, H, W, C = 16, 64, 1024, 16
    x = [bf(N, H, W, C) for _ in range(NS)]
    nw = torch.ones(C, device=dev)
    st = torch.zeros(N, C, 2, device=dev)
    timed("blur_noise_act_stats 16ch", lambda i: ops.blur_noise_act_stats(x[i], None, nw, st, _lib.ACT_LRELU, 0.2, 1, 0), NS, 2 * x[0].numel() * 2)
if want("copy"):
    x = [bf(16, 8, 257, 512) for _ in range(NS)]
    out = torch.empty_like(x[0])
    timed("torch copy_ (reference)", lambda i: out.copy_(x[i]), NS, 2 * x[0].numel() * 2)
