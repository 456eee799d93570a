#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_call30.log) 2>&1
timeout 900 python -m pytest tests/test_disc_gpu.py tests/test_enc_gpu.py tests/test_optim_gpu.py tests/test_gen_train_gpu.py tests/test_trainer_gen_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4
python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
from handwriting_line_generation_b200 import _lib, ops
for (N, H, W, kh, kw, ph, pw, C) in [(128, 64, 1024, 7, 7, 0, 3, 64), (256, 64, 1024, 5, 5, 2, 2, 32), (16, 64, 1024, 7, 7, 0, 3, 64)]:
    img = torch.randn(N, 1, H, W, device='cuda'); w = torch.randn(kh, C, 16, device='cuda').to(torch.bfloat16); b = torch.randn(C, device='cuda')
    Ho, Wo = H + 2 * ph - kh + 1, W + 2 * pw - kw + 1
    y = torch.empty(N, Ho, Wo, C, device='cuda', dtype=torch.bfloat16); st = torch.zeros(N, C, 2, device='cuda')
    f = lambda: _lib.call("hwg_stem_conv", img.data_ptr(), w.data_ptr(), b.data_ptr(), N, H, W, kh, kw, ph, pw, C, y.data_ptr(), st.data_ptr(), _lib.stream())
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(f"stem_conv N={N} {kh}x{kw} -> {C}: {us:.1f} us, output {y.numel() * 2 / us / 1e3:.0f} GB/s")
for B in (128, 16):
    x = torch.randn(B, 128, device='cuda'); yv = torch.randn(B, 128, device='cuda'); gy = torch.randn(B, 128, device='cuda'); Wm = torch.randn(128, 128, device='cuda')
    f = lambda: ops.linear_bwd(x, yv, gy, Wm, 2, 0.2)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    print(f"linear_bwd B={B}: {e0.elapsed_time(e1) * 50:.1f} us per call (incl. host launch)")
PY
for i in 1 2; do
  timeout 300 python tools/step_runner.py gan_step --B 128 --steps 10 --graph 2>&1 | tail -1
  timeout 300 python tools/step_runner.py gan_step --B 16 --steps 20 --graph 2>&1 | tail -1
done
