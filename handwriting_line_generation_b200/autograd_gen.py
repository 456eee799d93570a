"""Training path of the generator (autograd.Function around the fused forward + backward kernels)."""


def generator_apply(module, content, style, noise):
    raise NotImplementedError(
        "SpacedGenerator backward on libhwg_b200 is not built yet (round 1 ships the forward / inference "
        "path); call under torch.no_grad() — there is deliberately no PyTorch fallback")
