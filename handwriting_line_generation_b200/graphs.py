"""CUDA-graph capture of fixed-shape steps (north_star: "CUDA streams and graphs instead of a tracing compiler").

A step of this path is ~50-170 small-to-medium kernel launches; at the reference's batch sizes (4-16 lines)
the launch gaps, not the kernels, set the step time.  `GraphedStep` captures one call of a Python step function
(forward, or forward+loss+backward+optimizer) and replays it: inputs are copied into static buffers, outputs
live in static buffers.

Requirements the library meets for this: no host synchronisation inside a step (lengths are device tensors, the
output bias is read from device memory), no attribute/driver calls per launch after the first (one-time shared
memory opt-in), noise seeds that change per replay (device-side counter added to the captured seed), and
weight re-packing expressed as captured GPU work (a parameter update inside the graph bumps the version the
next capture-time forward saw, so the re-pack is part of the graph).
"""
import torch


def enable_device_noise_counter(*modules):
    for m in modules:
        for sub in m.modules():
            if hasattr(sub, "device_noise_counter"):
                sub.device_noise_counter = True


class GraphedStep:
    """graph = GraphedStep(fn, example_inputs, modules=[...]); out = graph(*inputs).

    fn(*tensors) -> tensor or tuple of tensors; every tensor it touches besides its arguments must be persistent
    (parameters, optimizer state, buffers).  For a training step use an optimizer built with capturable=True and
    call optimizer.zero_grad(set_to_none=True) yourself before constructing the GraphedStep (fn must not)."""

    def __init__(self, fn, example_inputs, modules=(), warmup=3):
        enable_device_noise_counter(*modules)
        self.static_in = [x.clone() for x in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = fn(*self.static_in)

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out


class GraphCache:
    """One captured graph per input-shape key, for steps whose shapes vary between calls (the reference trainer's batches:
    the generated width follows the text, the real lines are padded to the widest of the batch).

        step = GraphCache(fn, modules=[gen, hwr]);  out = step(*inputs)

    The first call with a new key captures (`GraphedStep`), later calls with that key replay; at most `max_graphs` graphs
    are kept (least recently used first out — a captured step holds its activations).  `key(*inputs)` defaults to the
    inputs' shapes; `bucket(*inputs)` may return padded inputs first (e.g. spaced text padded with blank rows up to a
    multiple of 32 columns: fewer distinct widths at the price of InstanceNorm statistics taken over the padded line — a
    change of the result, which is why no padding is applied by default)."""

    def __init__(self, fn, modules=(), warmup=3, max_graphs=8, key=None, bucket=None):
        self.fn, self.modules, self.warmup, self.max_graphs = fn, tuple(modules), warmup, max_graphs
        self.key = key or (lambda *ins: tuple((tuple(x.shape), x.dtype) for x in ins))
        self.bucket = bucket
        self._graphs = {}            # insertion order = recency
        self.captures = 0

    def __call__(self, *inputs):
        if self.bucket is not None:
            inputs = self.bucket(*inputs)
        k = self.key(*inputs)
        g = self._graphs.pop(k, None)
        if g is None:
            while len(self._graphs) >= self.max_graphs:
                self._graphs.pop(next(iter(self._graphs)))
            g = GraphedStep(self.fn, list(inputs), modules=self.modules, warmup=self.warmup)
            self.captures += 1
        self._graphs[k] = g
        return g(*inputs)


def pad_spaced_text(multiple=32, blank=0):
    """`bucket` for GraphCache over (content [T,B,C], ...): pads the spaced one-hot text with blank rows up to a multiple
    of `multiple` columns (the generated line gets 4 x as many extra blank pixel columns)."""
    def bucket(content, *rest):
        T = content.size(0)
        Tp = -(-T // multiple) * multiple
        if Tp != T:
            pad = torch.zeros((Tp - T,) + tuple(content.shape[1:]), device=content.device, dtype=content.dtype)
            pad[..., blank] = 1
            content = torch.cat((content, pad), 0)
        return (content,) + rest
    return bucket
