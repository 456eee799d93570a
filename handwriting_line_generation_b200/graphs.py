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
