"""GPU tests of the peer-memory exchange (include/hwg_b200.h "Peer-memory exchange", dp.PeerExchange).

On the one-GPU box the process group has world size 1: the kernels still run their full protocol against the rank's
own mailbox (flagged stores, epoch counters, parity buffers, CUDA-graph replay), and hwg_bn_coeffs_peer must equal
hwg_bn_coeffs.  With >= 2 GPUs visible the two-rank SyncBN parity check (tests/tools/dp_syncbn_check.py: exchange vs NCCL,
sharded recognizer vs one process on the whole batch) is launched under torchrun."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.fixture(scope="module")
def px():
    import torch.distributed as dist
    from handwriting_line_generation_b200 import dp
    assert not dist.is_initialized()
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{_free_port()}", rank=0, world_size=1,
                            device_id=torch.device("cuda", 0))
    try:
        yield dp.PeerExchange(dist.group.WORLD)
    finally:
        dist.destroy_process_group()


def test_allreduce_world1_is_identity_over_epochs(px):
    g = torch.Generator(device="cuda").manual_seed(0)
    for it in range(40):                                   # both parities of a slot, several slots, odd sizes
        n = (1024, 1, 2, 777, 1023)[it % 5]
        v = torch.randn(n, device="cuda", generator=g)
        out = px.allreduce_(v.clone(), ("t", it % 3))
        assert torch.equal(out, v)
    px.check()
    assert int(px.epochs[px.slot(("t", 0))].item()) == 14  # 14 launches on slot ("t", 0): the epoch lives on the device


def test_allreduce_rejects_oversize(px):
    with pytest.raises(RuntimeError, match="hwg_peer_allreduce_f32"):
        px.allreduce_(torch.zeros(1025, device="cuda"), ("big", 0))


def test_exchange_replays_in_cuda_graph(px):
    src = torch.randn(1024, device="cuda")
    buf = torch.zeros_like(src)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        buf.copy_(src)
        px.allreduce_(buf, ("graph", 0))
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        buf.copy_(src)
        px.allreduce_(buf, ("graph", 0))
    before = int(px.epochs[px.slot(("graph", 0))].item())
    for _ in range(9):
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(buf, src)
    assert int(px.epochs[px.slot(("graph", 0))].item()) == before + 9
    px.check()


@pytest.mark.parametrize("N,C,HW", [(4, 512, 250), (3, 256, 16 * 65), (1, 8, 7)])
def test_bn_coeffs_peer_equals_bn_coeffs(px, N, C, HW):
    from handwriting_line_generation_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(C)
    x = torch.randn(N, HW, C, device="cuda", generator=g) * 1.5 + 0.3
    stats = torch.stack([x.sum(1), (x * x).sum(1)], -1).contiguous()          # [N,C,2]
    w = torch.rand(C, device="cuda", generator=g) + 0.5
    b = torch.randn(C, device="cuda", generator=g)
    rm0, rv0 = torch.randn(C, device="cuda", generator=g), torch.rand(C, device="cuda", generator=g) + 0.5
    rm1, rv1, rm2, rv2 = rm0.clone(), rv0.clone(), rm0.clone(), rv0.clone()
    coef1, save1 = ops.bn_coeffs(stats, N, C, HW, w, b, rm1, rv1, 0.1, 1e-5, True)
    coef2, save2 = ops.bn_coeffs_synced(stats, N, C, HW, w, b, rm2, rv2, 0.1, 1e-5, px, key=("bn", C))
    for a, r in ((coef2, coef1), (save2, save1), (rm2, rm1), (rv2, rv1)):
        torch.testing.assert_close(a, r, rtol=1e-6, atol=1e-6)                # same arithmetic, same fold order
    # and against the definition (fp64)
    xd = x.double().reshape(-1, C)
    mean, var = xd.mean(0), xd.var(0, unbiased=False)
    torch.testing.assert_close(save2[:, 0].double(), mean, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(save2[:, 1].double(), (var + 1e-5).rsqrt(), rtol=1e-4, atol=1e-4)
    px.check()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_syncbn_two_ranks_torchrun():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "tools", "dp_syncbn_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert "SYNCBN_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
