"""torchrun --nproc-per-node 2 tools/dp_syncbn_check.py — SyncBN parity on two GPUs (development aid / evidence):
a batch sharded over two ranks with CNNOnlyHWR.sync_bn_group set must give the log-probs, the input gradient and the
running statistics of ONE process running the whole batch (the reference is single-process)."""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import handwriting_line_generation_b200 as pkg
from oracle import synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


B, W, C = 8, 256, 80
torch.manual_seed(0)
hwr = pkg.CNNOnlyHWR(C, norm='batch').to(dev).train()
for p in hwr.parameters():
    p.requires_grad_(False)
single = copy.deepcopy(hwr)
x_all = torch.from_numpy(synth.hwr_case(B, W, 9)).to(dev)
T = W // 4 - 6
g_all = torch.randn(T, B, C, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
per = B // world
sl = slice(rank * per, (rank + 1) * per)
# sharded, synchronised
hwr.sync_bn_group = dist.group.WORLD
xs = x_all[sl].clone().requires_grad_()
lp = hwr(xs)
lp.backward(g_all[:, sl].contiguous())
torch.cuda.synchronize()
dist.barrier()
ok = True
if rank == 0:
    xf = x_all.clone().requires_grad_()
    lpf = single(xf)
    lpf.backward(g_all)
    torch.cuda.synchronize()
    e_lp = rel_l2(lp.detach(), lpf.detach()[:, sl])
    e_g = rel_l2(xs.grad, xf.grad[sl])
    e_rm = rel_l2(hwr.cnn.batchnorm4.running_mean, single.cnn.batchnorm4.running_mean)
    e_rv = rel_l2(hwr.cnn1d[10].running_var, single.cnn1d[10].running_var)
    # and what per-rank statistics would have given (the default without sync_bn_group)
    local_only = copy.deepcopy(single)
    lpl = local_only(x_all[sl].clone())
    e_local = rel_l2(lpl.detach(), lpf.detach()[:, sl])
    print(f"SyncBN 2-rank vs single process: log-probs {e_lp:.2e}, input gradient {e_g:.2e}, running_mean {e_rm:.2e}, "
          f"running_var {e_rv:.2e}; per-rank statistics instead: log-probs {e_local:.2e}", flush=True)
    ok = e_lp <= 2e-2 and e_g <= 2.5e-1 and e_rm <= 1e-3 and e_rv <= 1e-3 and (world == 1 or e_local > 2 * e_lp)
    print("SYNCBN_CHECK", "PASS" if ok else "FAIL", flush=True)
dist.barrier()
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0 if ok else 1)
