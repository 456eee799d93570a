"""Batch data parallelism helpers: one process per GPU (torchrun), lines sharded by rank.

The forward path has no cross-line coupling (InstanceNorm is per sample), so inference shards with
no data-path collective.  Training adds one exchange — the gradient all-reduce — implemented here
as flat fp32 buckets reduced with torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous, balanced [begin, end) of `total` lines for `rank` of `world` (sizes differ by <= 1)."""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_batch(t, dim, rank, world):
    b, e = shard_range(t.size(dim), rank, world)
    return t.narrow(dim, b, e - b)


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def max_over_ranks(value, device="cpu"):
    """Timing rule: a multi-GPU duration is the max over ranks."""
    rank, w = world()
    if w == 1:
        return float(value)
    t = torch.tensor([float(value)], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


class GradBuckets:
    """Flat fp32 gradient buckets for a fixed parameter list.

    `reduce()` copies every .grad into its slot of a flat buffer, all-reduces the buffers (sum) and
    scatters grad/world back.  Buckets are sized for launch latency/overlap (NVSwitch gives every peer
    full bandwidth, so there is no per-link tuning): `bucket_bytes` default 32 MiB.
    The reference's gradient balancing (trainer/hw_with_style_trainer.py:340-376) is nonlinear in the
    gradients, so each stashed gradient set must go through reduce() BEFORE it is balanced."""

    def __init__(self, params, bucket_bytes=32 << 20):
        self.params = [p for p in params if p.requires_grad]
        self.buckets, cur, size = [], [], 0
        for p in self.params:
            n = p.numel() * 4
            if cur and size + n > bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += n
        if cur:
            self.buckets.append(cur)
        self.flat = [torch.zeros(sum(p.numel() for p in b), dtype=torch.float32, device=b[0].device)
                     for b in self.buckets]

    def reduce(self, async_op=False):
        rank, w = world()
        handles = []
        for bucket, flat in zip(self.buckets, self.flat):
            off = 0
            for p in bucket:
                n = p.numel()
                if p.grad is None:
                    flat[off:off + n].zero_()
                else:
                    flat[off:off + n].copy_(p.grad.reshape(-1))
                off += n
            if w > 1:
                handles.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True))
        for h in handles:
            h.wait()
        for bucket, flat in zip(self.buckets, self.flat):
            off = 0
            for p in bucket:
                n = p.numel()
                g = flat[off:off + n].view_as(p) / w
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
                off += n
