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
        self._handles, self._bucket_of, self._producer_streams = [], {}, {}
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
        self.buckets.append({"params": plist, "flat": flat, "views": views, "ready": set(), "work": None})

    def _arrived(self, p):
        """Post-accumulate-grad hook: one parameter of a bucket has its gradient for this backward pass.  The autograd
        engine runs the hook ONCE per pass, after every node that feeds the parameter has run — also when the producing
        Functions returned None because their kernels wrote the gradient straight into the flat buffer, and also when the
        module was called several times in the graph (the reference trainer concatenates a reconstruction and a generated
        batch, trainer :538/:577) — so it is the only readiness signal; `mark_ready` does not count."""
        b = self.buckets[self._bucket_of[id(p)]]
        if b["work"] is not None:          # already on its way this step
            return
        b["ready"].add(id(p))
        if len(b["ready"]) == len(b["params"]):
            self._launch(b)

    def _make_hook(self, bi):
        self._bucket_of.update({id(p): bi for p in self.buckets[bi]["params"]})
        return self._arrived

    def mark_ready(self, p):
        """For gradients written straight into the flat buffer by a backward kernel: the producer reports, after its
        launches, the STREAM it wrote on — the all-reduce is ordered after that stream as well as after the stream the
        hook runs on.  Readiness itself comes from the hook (`_arrived`): a producer that runs twice in one pass must not
        release the bucket after its first run."""
        if self.world > 1 and self.comm is not None and id(p) in self._bucket_of:
            self._producer_streams[torch.cuda.current_stream().cuda_stream] = torch.cuda.current_stream()

    def _launch(self, b):
        for p, v in zip(b["params"], b["views"]):
            if p.grad is None:
                if self.flat is None:                 # (a flat slot is zeroed by step / zero_grad and may already hold what a
                    v.zero_()                         # backward kernel added)
                p.grad = v
            elif p.grad.data_ptr() != v.data_ptr():   # autograd installed its own tensor: move it into the bucket
                if self.flat is not None:             # a backward kernel may already have added into the flat slot
                    v.add_(p.grad)                    # (FlatAdam._rebind_grads does the same; the slot starts zeroed)
                else:
                    v.copy_(p.grad)
                p.grad = v
        if self.comm is not None:
            ev = torch.cuda.Event()
            ev.record()
            self.comm.wait_event(ev)
            for st in self._producer_streams.values():        # kernels that wrote the flat buffer directly
                self.comm.wait_stream(st)
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
            b["ready"].clear()
        self._producer_streams.clear()       # per step: a stream of an earlier (eager) step must not enter a later capture
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


class PeerExchange:
    """Peer-mapped mailboxes for the in-kernel exchanges of include/hwg_b200.h ("Peer-memory exchange"): SyncBN
    statistics of the data-parallel recognizer summed over the ranks inside `hwg_bn_coeffs_peer` /
    `hwg_peer_allreduce_f32` (NVLink stores with the flag in the payload) instead of one NCCL all-reduce + stream
    fork/join per BatchNorm layer.  Set `CNNOnlyHWR.sync_bn_group = PeerExchange(group)`.

    Every rank allocates a zero-filled mailbox and maps the mailboxes of all ranks of `group` (one node) into its own
    address space: torch symmetric memory (CUDA VMM handles) when available, else CUDA IPC handles exchanged with
    `all_gather_object`.  PyTorch owns the memory; the library only receives the device table of base addresses.
    A slot is handed out per call-site key in first-use order — every rank must run the same exchanges in the same
    order (they do: same model, same step)."""
    SLOTS = 64

    def __init__(self, group=None, device=None):
        from . import _lib
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        nbytes = _lib.load().hwg_peer_mailbox_bytes(self.world, self.SLOTS)
        if nbytes <= 0:
            raise RuntimeError(f"PeerExchange: world size {self.world} not supported")
        self._keep = []
        try:
            ptrs, self.mapping = self._map_symmetric(nbytes), "symmetric_memory"
        except Exception as e:                       # noqa: BLE001 - any failure of the VMM path: fall back to CUDA IPC
            self._symm_error = repr(e)
            ptrs, self.mapping = self._map_ipc(nbytes, _lib), "cuda_ipc"
        self.table = torch.tensor([int(p) for p in ptrs], dtype=torch.int64, device=self.device)
        self.epochs = torch.zeros(self.SLOTS, dtype=torch.int32, device=self.device)
        self.fault = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._slots = {}
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)               # every mailbox is zeroed and mapped before anyone writes

    def _map_symmetric(self, nbytes):
        import torch.distributed._symmetric_memory as symm_mem
        enable = getattr(symm_mem, "enable_symm_mem_for_group", None)
        if enable is not None:
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                enable(self.group.group_name)
        buf = symm_mem.empty(nbytes // 4, dtype=torch.int32, device=self.device)
        buf.zero_()
        torch.cuda.synchronize(self.device)
        hdl = symm_mem.rendezvous(buf, group=self.group)
        ptrs = list(hdl.buffer_ptrs)
        if len(ptrs) != self.world or int(ptrs[self.rank]) != buf.data_ptr():
            raise RuntimeError("symmetric memory rendezvous returned an unexpected pointer table")
        self._keep += [buf, hdl]
        return ptrs

    def _map_ipc(self, nbytes, _lib):
        from torch.multiprocessing.reductions import reduce_tensor
        buf = torch.zeros(nbytes // 4, dtype=torch.int32, device=self.device)
        torch.cuda.synchronize(self.device)
        fn, args = reduce_tensor(buf)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, (fn, args), group=self.group)
        views = [buf if r == self.rank else g[0](*g[1]) for r, g in enumerate(gathered)]
        with torch.cuda.device(self.device):
            for v in views:
                if v.device != self.device:
                    _lib.call("hwg_peer_enable_access", v.device.index)
        self._keep += views
        return [v.data_ptr() for v in views]

    def slot(self, key):
        s = self._slots.get(key)
        if s is None:
            s = self._slots[key] = len(self._slots)
            if s >= self.SLOTS:
                raise RuntimeError(f"PeerExchange: more than {self.SLOTS} exchange sites")
        return s

    def args(self, key):
        """(peer_mailboxes, world, rank, slot, slots, epochs, fault) as the C-ABI takes them."""
        return (self.table.data_ptr(), self.world, self.rank, self.slot(key), self.SLOTS, self.epochs.data_ptr(),
                self.fault.data_ptr())

    def allreduce_(self, t, key):
        """In-place sum over the ranks of a small contiguous fp32 tensor (<= 1024 values), one launch on the current
        stream."""
        from . import _lib
        assert t.dtype == torch.float32 and t.is_contiguous()
        _lib.call("hwg_peer_allreduce_f32", t.data_ptr(), t.data_ptr(), t.numel(), *self.args(key), _lib.stream())
        return t

    def check(self):
        """Host-side check (synchronises): raises if an exchange timed out waiting for a peer."""
        if int(self.fault.item()) != 0:
            raise RuntimeError("PeerExchange: a peer did not answer an in-kernel exchange within the time-out")
