"""Flat fused optimizer for one optimizer group of the reference trainer.

The reference builds `torch.optim.Adam(lr=2e-4, betas=(0.5, 0.999))` over the generator-side parameters
(base/base_trainer.py:61-102, configs/cf_IAM*.json:35-46) and runs `clip_grad_value_(params, 2)` before every
step (trainer/hw_with_style_trainer.py:381).  `FlatAdam` keeps the same mathematics but stores the group as ONE flat
fp32 buffer: parameters, gradients and both moments are slices of four flat tensors, so that
  * the optimizer step (clip + Adam + gradient zeroing) is one kernel launch (hwg_adam_flat) with its step counter on
    the device — replayable inside a CUDA graph;
  * the data-parallel all-reduce works on contiguous slices of the gradient buffer with no gather/scatter copies
    (dp.GradReducer(flat=optimizer));
  * backward kernels can accumulate straight into the gradient buffer (`grad_view`).
Parameters keep their identity, names and shapes (the state_dict contract is untouched): only `.data` and `.grad`
are re-pointed at views of the flat buffers.
"""
import torch

from . import _lib


class _Sink:
    """A gradient destination with the optimizer's slot layout: the main gradient buffer or one of its stash slots.  The
    backward kernels of a module whose `_grad_sink` is set ADD their parameter gradients straight into it."""

    def __init__(self, opt, buf):
        self.opt, self.buf = opt, buf

    def owns(self, p):
        return id(p) in self.opt.offsets

    def grad_view(self, p):
        o, k = self.opt.offsets[id(p)]
        return self.buf[o:o + k].view_as(p)


class FlatAdam:
    def __init__(self, params, lr=2e-4, betas=(0.5, 0.999), eps=1e-8, clip_value=None):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, "FlatAdam: no trainable parameters"
        dev = self.params[0].device
        _lib.require_cuda(*self.params)
        self.lr, self.betas, self.eps, self.clip_value = lr, betas, eps, clip_value
        self.offsets, n = {}, 0
        for p in self.params:
            assert p.dtype == torch.float32 and p.device == dev
            self.offsets[id(p)] = (n, p.numel())
            n += -(-p.numel() // 4) * 4          # 16-byte aligned slots
        self.numel = n
        kw = dict(device=dev, dtype=torch.float32)
        self.flat_p, self.flat_g = torch.zeros(n, **kw), torch.zeros(n, **kw)
        self.exp_avg, self.exp_avg_sq = torch.zeros(n, **kw), torch.zeros(n, **kw)
        self.step_dev = torch.zeros(1, **kw)
        with torch.no_grad():
            for p in self.params:
                o, k = self.offsets[id(p)]
                view = self.flat_p[o:o + k].view_as(p)
                view.copy_(p.data)
                p.data = view
                p.grad = self.flat_g[o:o + k].view_as(p)

    def grad_view(self, p):
        o, k = self.offsets[id(p)]
        return self.flat_g[o:o + k].view_as(p)

    # ---- stash slots: the trainer's `saved_grad` sets (:303-338) without the clone + zero round trip ----
    def sink(self, slot=None):
        """Gradient destination for `module._grad_sink`: None = the main gradient buffer (this optimizer itself), an int =
        stash slot `slot` (a persistent zero-initialised buffer with the same layout).  A backward pass whose module
        points at slot k leaves its gradient set there directly — the trainer's "backward, clone into saved_grad, zero"
        (:312-338) becomes "backward into the slot" — and the passes of one optimizer step no longer share a buffer, so
        they can run concurrently on different streams.  `balance()` consumes and re-zeroes the slots that were handed
        out."""
        if slot is None:
            return self
        slots = self.__dict__.setdefault("_slots", {})
        if slot not in slots:
            slots[slot] = _Sink(self, torch.zeros_like(self.flat_g))
        self.__dict__.setdefault("_slots_used", set()).add(slot)
        return slots[slot]

    def owns(self, p):
        return id(p) in self.offsets

    def _rebind_grads(self):
        """autograd replaces `.grad` when it was None (or a caller assigned one): fold such gradients back in."""
        for p in self.params:
            o, k = self.offsets[id(p)]
            if p.grad is None:
                p.grad = self.flat_g[o:o + k].view_as(p)
            elif p.grad.data_ptr() != self.flat_g.data_ptr() + 4 * o:
                self.flat_g[o:o + k].view_as(p).add_(p.grad)
                p.grad = self.flat_g[o:o + k].view_as(p)

    @torch.no_grad()
    def step(self, grad_scale=1.0):
        self._rebind_grads()
        _lib.call("hwg_adam_flat", self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.exp_avg.data_ptr(),
                  self.exp_avg_sq.data_ptr(), self.numel, self.lr, self.betas[0], self.betas[1], self.eps,
                  float(self.clip_value or 0.0), float(grad_scale), self.step_dev.data_ptr(), 1, _lib.stream())
        torch.autograd.graph.increment_version(self.params)     # derived-weight caches key on (data_ptr, _version)

    # ---- gradient balancing (trainer/hw_with_style_trainer.py:300-377): tests/test_balance_gpu.py ----
    def stash(self):
        """The trainer's `saved_grad` (:303-322, :330-338): keeps the gradient buffer as one more stashed set and zeroes
        it for the next backward."""
        self._rebind_grads()
        if not hasattr(self, "_stash"):
            self._stash = []
        self._stash.append(self.flat_g.clone())
        self.flat_g.zero_()

    @torch.no_grad()
    def balance(self, multipliers):
        """:340-377 on the flat buffers: adds every stashed set to the gradient buffer, rescaled per parameter to the
        gradient's mean magnitude and weighted by `multipliers` (`balance_var_x`); one hwg_balance call (three launches,
        no host synchronisation).  Clears the stash."""
        import ctypes
        used = sorted(getattr(self, "_slots_used", ()))
        sets = [self._slots[k].buf for k in used] + getattr(self, "_stash", [])     # slot order, then stash() order
        if not sets:
            return
        self._rebind_grads()
        assert len(multipliers) >= len(sets)
        dev = self.flat_g.device
        if getattr(self, "_bal", None) is None:
            chunk = _lib.load().hwg_balance_chunk()
            offs, lens, tab = [], [], []
            for si, p in enumerate(self.params):
                o, k = self.offsets[id(p)]
                offs.append(o)
                lens.append(k)
                tab += [(si, c) for c in range(-(-k // chunk))]
            self._bal = dict(off=torch.tensor(offs, dtype=torch.int64, device=dev),
                             len=torch.tensor(lens, dtype=torch.int64, device=dev),
                             tab=torch.tensor(tab, dtype=torch.int32, device=dev).reshape(-1, 2).contiguous(),
                             nseg=len(offs))
        b, K = self._bal, len(sets)
        key = tuple(float(v) for v in multipliers[:K])
        if b.get("x_key") != key:            # uploaded once per multiplier list: no host copy on the steady-state path,
            b["x_key"], b["x"] = key, torch.tensor(key, dtype=torch.float32, device=dev)   # so a captured step can replay it
        x = b["x"]
        sums = torch.empty(b["tab"].size(0) * (K + 1) + b["nseg"], dtype=torch.float32, device=dev)   # block partials + first blocks
        mult = torch.empty((K, b["nseg"]), dtype=torch.float32, device=dev)
        ptrs = (ctypes.c_void_p * K)(*[t.data_ptr() for t in sets])
        _lib.call("hwg_balance", self.flat_g.data_ptr(), ctypes.addressof(ptrs), K, x.data_ptr(), b["off"].data_ptr(),
                  b["len"].data_ptr(), b["nseg"], b["tab"].data_ptr(), b["tab"].size(0), sums.data_ptr(), mult.data_ptr(),
                  _lib.stream())
        self._stash = []
        for k in used:                       # ready for the next step's backward passes
            self._slots[k].buf.zero_()
        if used:
            self._slots_used = set()

    def zero_grad(self, set_to_none=False):
        """step() already leaves the gradient buffer zeroed; this is for steps that are skipped."""
        self.flat_g.zero_()

    def state_dict(self):
        return {"step": self.step_dev.clone(), "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "lr": self.lr, "betas": self.betas, "eps": self.eps, "clip_value": self.clip_value}

    def load_state_dict(self, sd):
        self.step_dev.copy_(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.lr, self.betas, self.eps, self.clip_value = sd["lr"], tuple(sd["betas"]), sd["eps"], sd["clip_value"]
