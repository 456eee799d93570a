"""Device timing of hwg_conv_wgrad on the generator's weight-gradient launches (development aid).

  python tools/wgrad_bench.py [--B 128] [names...]
Environment switches of the library (HWG_WGRAD_HALO, HWG_WGS_VARIANT, HWG_WGRAD_CTAS) apply.
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from handwriting_line_generation_b200 import conv, _lib, weightmap

B = 128
args = sys.argv[1:]
if "--B" in args:
    i = args.index("--B"); B = int(args[i + 1]); del args[i:i + 2]
T3 = conv.conv_taps(3, 3, 1, 1)
# name: (x shape HWC, gy shape HWC, [ (taps, kwargs) launches ])
F = weightmap.fused_taps()
CASES = {
    "b4c2 16->16": ((64, 1024, 16), (64, 1024, 16), [(T3, {})]),
    "b4c1 fused 32->16": ((32, 512, 32), (64, 1024, 16), [(F, dict(grid=(32, 512), gy_stride=(2, 2), tap_phase=weightmap.fused_phases()))]),
    "b3c2 32->32": ((32, 512, 32), (32, 512, 32), [(T3, {})]),
    "b3c1 fused 64->32": ((16, 256, 64), (32, 512, 32),
                          [(F[4 * q:4 * q + 4], dict(grid=(16, 256), gy_stride=(2, 2), gy_offset=(q // 2, q % 2))) for q in range(4)]),
    "b2c2 64->64": ((16, 256, 64), (16, 256, 64), [(T3, {})]),
    "b2c1 vert 128->64": ((8, 256, 128), (16, 256, 64),
                          [(weightmap.vert_taps(par), dict(grid=(8, 256), gy_stride=(2, 1), gy_offset=(par, 0))) for par in (0, 1)]),
    "b1c2 128->128": ((8, 256, 128), (8, 256, 128), [(T3, {})]),
    "b1c1 vert 256->128": ((4, 256, 256), (8, 256, 128),
                           [(weightmap.vert_taps(par), dict(grid=(4, 256), gy_stride=(2, 1), gy_offset=(par, 0))) for par in (0, 1)]),
    "b0c2 256->256": ((4, 256, 256), (4, 256, 256), [(T3, {})]),
    "hwr conv1 64->128": ((32, 512, 64), (32, 512, 128), [(T3, {})]),
    "hwr conv3 256->256": ((16, 256, 256), (16, 256, 256), [(T3, {})]),
    "hwr conv5 512->512": ((8, 257, 512), (6, 255, 512), [(conv.conv_taps(3, 3, 0, 0), {})]),
    "disc convs1 64->64": ((58, 1024, 64), (56, 1024, 64), [(conv.conv_taps(3, 3, 0, 1), {})]),
}
which = [n for n in CASES if not args or any(a in n for a in args)]
reps, nbuf = 5, 2
names = {0: "staged", 1: "tcgen05", 2: "tcgen05+halo"}
for name in which:
    (H, W, Ci), (Ho, Wo, Co), launches = CASES[name]
    Bc = B if H * W * Ci * B * 2 < (1 << 31) else B // 2
    xs = [torch.randn(Bc, H, W, Ci, device="cuda").to(torch.bfloat16) for _ in range(nbuf)]
    gs = [torch.randn(Bc, Ho, Wo, Co, device="cuda").to(torch.bfloat16) for _ in range(nbuf)]
    ntap = sum(len(t) for t, _ in launches)
    out = torch.zeros(ntap, Co, Ci, device="cuda")

    def run(i):
        o = 0
        for taps, kw in launches:
            conv.conv_wgrad(xs[i % nbuf], gs[i % nbuf], taps, Ci, Co, out=out[o:o + len(taps)], **kw)
            o += len(taps)
    run(0); run(1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    by = (xs[0].numel() + gs[0].numel()) * 2
    fl = 2.0 * Bc * Ho * Wo * Co * Ci * ntap / (4 if "fused" in name else (2 if "vert" in name else 1))
    print(f"{name:22s} B={Bc:4d} {len(launches)} launch(es) {us:9.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  operands once {by / us / 1e3:7.1f} GB/s  "
          f"[{names.get(_lib.load().hwg_last_wgrad_kernel(), '?')}]", flush=True)
