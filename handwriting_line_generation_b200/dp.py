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


class GradReducer:
    """Gradient all-reduce overlapped with the backward pass (the reference itself is single-GPU,
    trainer/hw_with_style_trainer.py:300-391 being the step this slots into before the optimizer).

    Parameters are grouped, in reverse registration order (≈ the order the backward produces their gradients), into
    flat fp32 buckets; every `.grad` is a view into its bucket, so there is no gather/scatter copy.  A
    post-accumulate-grad hook counts the gradients of a bucket; when the last one lands, the bucket is pre-scaled by
    1/world and all-reduced (sum) on a side stream ordered after the producing stream by an event, while the
    backward keeps running on the main stream.  `finish()` makes the main stream wait for the outstanding buckets —
    call it between `loss.backward()` and `optimizer.step()`.  All of this is stream work only (no host
    synchronisation), so a whole step including the collectives can be captured in one CUDA graph.
    With world_size 1 the hooks are not installed and finish() is a no-op."""

    def __init__(self, params, bucket_bytes=4 << 20, group=None, flat=None):
        """flat: an optim.FlatAdam over the same parameters (same order) — its gradient buffer is bucketed in place
        instead of allocating new buckets."""
        self.group, self.flat = group, flat
        self.rank, self.world = world()
        self.params = [p for p in params if p.requires_grad]
        self.buckets = []            # dicts: params, flat, views, pending, work
        cur, size = [], 0
        for p in reversed(self.params):
            n = p.numel() * 4
            if cur and size + n > bucket_bytes:
                self._close(cur)
                cur, size = [], 0
            cur.append(p)
            size += n
        if cur:
            self._close(cur)
        self.cuda = bool(self.params) and self.params[0].is_cuda
        self.comm = torch.cuda.Stream() if (self.cuda and self.world > 1) else None
        self._handles, self._bucket_of = [], {}
        if self.world > 1:
            for bi, b in enumerate(self.buckets):
                for p in b["params"]:
                    self._handles.append(p.register_post_accumulate_grad_hook(self._make_hook(bi)))

    def _close(self, plist):
        dev = plist[0].device
        size = sum(-(-p.numel() // 4) * 4 for p in plist)
        if self.flat is not None:                 # contiguous range of the optimizer's flat gradient buffer
            lo = min(self.flat.offsets[id(p)][0] for p in plist)
            flat = self.flat.flat_g[lo:lo + size]
            views = [self.flat.grad_view(p) for p in plist]
            assert all(lo <= self.flat.offsets[id(p)][0] < lo + size for p in plist), "parameter order differs from the optimizer's"
        else:
            flat = torch.zeros(size, dtype=torch.float32, device=dev)
            views, off = [], 0
            for p in plist:
                v = flat[off:off + p.numel()].view_as(p)
                p.grad = v
                views.append(v)
                off += -(-p.numel() // 4) * 4          # 16-byte aligned slots
        self.buckets.append({"params": plist, "flat": flat, "views": views, "pending": len(plist), "work": None})

    def _make_hook(self, bi):
        def hook(p):
            b = self.buckets[bi]
            b["pending"] -= 1
            if b["pending"] == 0:
                self._launch(b)
        self._bucket_of.update({id(p): bi for p in self.buckets[bi]["params"]})
        return hook

    def mark_ready(self, p):
        """For gradients written straight into the flat buffer by a backward kernel (no AccumulateGrad, so no hook):
        the producer reports the parameter itself."""
        if self.world > 1 and id(p) in self._bucket_of:
            b = self.buckets[self._bucket_of[id(p)]]
            b["pending"] -= 1
            if b["pending"] == 0:
                self._launch(b)

    def _launch(self, b):
        for p, v in zip(b["params"], b["views"]):
            if p.grad is None:
                v.zero_()
                p.grad = v
            elif p.grad.data_ptr() != v.data_ptr():   # autograd installed its own tensor: move it into the bucket
                v.copy_(p.grad)
                p.grad = v
        if self.comm is not None:
            ev = torch.cuda.Event()
            ev.record()
            self.comm.wait_event(ev)
            with torch.cuda.stream(self.comm):
                b["flat"].mul_(1.0 / self.world)
                b["work"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            b["flat"].mul_(1.0 / self.world)
            b["work"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self):
        """Launches the buckets that never filled (parameters without a gradient this step), then waits."""
        if self.world == 1:
            return
        for b in self.buckets:
            if b["work"] is None:
                self._launch(b)
        for b in self.buckets:
            b["work"].wait()
            b["work"] = None
            b["pending"] = len(b["params"])
        if self.comm is not None:
            torch.cuda.current_stream().wait_stream(self.comm)

    def zero_grad(self):
        """One fill per bucket instead of one per parameter (gradients stay views into the buckets)."""
        for b in self.buckets:
            b["flat"].zero_()

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []
