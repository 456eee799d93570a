"""Device timing of hwg_conv_fprop on representative layer shapes (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from handwriting_line_generation_b200 import conv, _lib

SHAPES = {
    # name: (N, H, W, Cin, Cout, taps(kh,kw,ph,pw), epilogue)
    "hwr_conv3": (8, 16, 256, 256, 256, (3, 3, 1, 1), "relu"),
    "hwr_conv5": (8, 8, 257, 512, 512, (3, 3, 0, 0), "relu"),
    "hwr_conv1": (8, 32, 512, 64, 128, (3, 3, 1, 1), "relu"),
    "gen_b0c2": (32, 4, 256, 256, 256, (3, 3, 1, 1), "noise_stats"),
    "gen_b1c2": (32, 8, 256, 128, 128, (3, 3, 1, 1), "noise_stats"),
    "gen_b3c2": (32, 32, 512, 32, 32, (3, 3, 1, 1), "noise_stats"),
    "gen_b4c2": (32, 64, 1024, 16, 16, (3, 3, 1, 1), "noise_stats"),
    "gen_b4c2_plain": (32, 64, 1024, 16, 16, (3, 3, 1, 1), "none"),
    # train-step shapes at 16 lines per GPU
    "t_hwr_1d": (16, 1, 256, 512, 512, (1, 3, 0, 0), "stats"),
    "t_hwr_conv6": (16, 3, 256, 512, 512, (3, 3, 0, 0), "stats"),
    "t_hwr_conv5": (16, 8, 257, 512, 512, (3, 3, 0, 0), "relu"),
    "t_hwr_conv4": (16, 8, 257, 256, 512, (3, 3, 1, 1), "stats"),
    "t_hwr_conv2": (16, 16, 256, 128, 256, (3, 3, 1, 1), "stats"),
    "t_hwr_conv1": (16, 32, 512, 64, 128, (3, 3, 1, 1), "relu"),
    "t_gen_b0c2": (16, 4, 256, 256, 256, (3, 3, 1, 1), "noise_stats"),
    "t_gen_b0c2_dgrad": (16, 4, 256, 256, 256, (3, 3, 1, 1), "none"),
    "t_gen_b1c2": (16, 8, 256, 128, 128, (3, 3, 1, 1), "noise_stats"),
    "t_gen_b2c2": (16, 16, 256, 64, 64, (3, 3, 1, 1), "noise_stats"),
    # discriminator, train-step shapes at 16 lines per GPU (HWG_CONV_TILE_W=32 reproduces the module's tile choice)
    "t_disc_convs1_0": (16, 58, 1024, 64, 64, (3, 3, 0, 1), "lrelu"),
    "t_disc_convs1_3": (16, 28, 512, 64, 128, (3, 3, 0, 1), "none"),
    "t_disc_convs2_0": (16, 26, 512, 128, 128, (3, 3, 0, 1), "lrelu"),
    "t_disc_convs3_0": (16, 12, 256, 128, 128, (3, 3, 0, 1), "stats"),
    "t_disc_convs3_4": (16, 5, 128, 128, 256, (3, 3, 0, 1), "none"),
    # the discriminator's in_conv as the module runs it: 7 vertical taps over the 16 shifted copies of the image
    "t_disc_in_conv": (16, 64, 1024, 16, 64, (7, 1, 0, 0), "stats"),
    "t_disc_in_conv_plain": (16, 64, 1024, 16, 64, (7, 1, 0, 0), "none"),
    "t_disc_in_conv_dgrad": (16, 64, 1024, 64, 16, (7, 1, 3, 0), "none"),
}
TILE_W = int(os.environ.get("HWG_CONV_TILE_W", "0"))
which = [a for a in sys.argv[1:] if a in SHAPES] or ([] if sys.argv[1:] else list(SHAPES))
reps = 10
for name in which:
    N, H, W, Cin, Cout, (kh, kw, ph, pw), epi = SHAPES[name]
    x = torch.randn(N, H, W, Cin, device="cuda").to(torch.bfloat16)
    w = (torch.randn(kh * kw, Cout, Cin, device="cuda") / (Cin * kh * kw) ** 0.5).to(torch.bfloat16)
    b = torch.randn(Cout, device="cuda")
    taps = conv.conv_taps(kh, kw, ph, pw)
    Ho, Wo = H + 2 * ph - kh + 1, W + 2 * pw - kw + 1
    kw_ = dict(bias=b)
    if TILE_W:
        kw_.update(tile_w=TILE_W)
    if epi == "relu":
        kw_.update(act=_lib.ACT_RELU)
    elif epi == "lrelu":
        kw_.update(act=_lib.ACT_LRELU, slope=0.1)
    elif epi == "stats":
        kw_.update(stats=torch.zeros(N, Cout, 2, device="cuda"))
    elif epi == "noise_stats":
        kw_.update(act=_lib.ACT_LRELU, slope=0.2, noise_w=torch.ones(Cout, device="cuda"), noise_seed=1,
                   stats=torch.zeros(N, Cout, 2, device="cuda"))
    out = torch.empty(N, Ho, Wo, Cout, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        conv.conv_fprop(x, w, taps, Ho, Wo, out=out, **kw_)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        conv.conv_fprop(x, w, taps, Ho, Wo, out=out, **kw_)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    fl = 2.0 * N * Ho * Wo * Cout * Cin * kh * kw
    by = (x.numel() + out.numel()) * 2
    kern = {1: "tcgen05", 2: "small", 3: "tcgen05+halo2d", 4: "tcgen05+halo-rows"}.get(_lib.load().hwg_last_conv_kernel(), "?")
    print(f"{name:16s} {us:9.1f} us  {fl / us / 1e6:8.1f} TFLOP/s  act-bytes {by / us / 1e3:8.1f} GB/s  [{kern}]", flush=True)

if "blur_b4" in sys.argv[1:]:
    from handwriting_line_generation_b200 import ops
    N, H, W, C = 32, 64, 1024, 16
    x = torch.randn(N, H, W, C, device="cuda").to(torch.bfloat16)
    nw = torch.ones(C, device="cuda")
    st = torch.zeros(N, C, 2, device="cuda")
    for _ in range(3):
        ops.blur_noise_act_stats(x, None, nw, st, _lib.ACT_LRELU, 0.2, 1, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.blur_noise_act_stats(x, None, nw, st, _lib.ACT_LRELU, 0.2, 1, 0)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"blur_b4 {us:9.1f} us  {2 * x.numel() * 2 / us / 1e3:8.1f} GB/s", flush=True)
