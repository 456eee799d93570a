"""Training path of the recognizer (autograd.Function around the fused forward + backward kernels)."""


def hwr_apply(module, input):
    raise NotImplementedError(
        "CNNOnlyHWR backward on libhwg_b200 is not built yet (round 1 ships the forward path); call under "
        "torch.no_grad() — there is deliberately no PyTorch fallback")
