"""TEST INFRASTRUCTURE — CPU fp32 restatement of the reference recognizer CNNOnlyHWR
(model/cnn_only_hwr.py:21-56,78-107, norm='batch', no pad, not small) on a reference-format
state_dict.  Pinned by tests/golden/hwr_*.npz."""
import torch
import torch.nn.functional as F

PADS = [1, 1, 1, 1, 1, 0, 0]
BN2D = {2, 4, 6}
POOL_AFTER = {0: ((2, 2), (2, 2), (0, 0)), 1: ((2, 2), (2, 2), (0, 0)),
              3: ((2, 2), (2, 1), (0, 1)), 5: ((2, 2), (2, 1), (0, 1))}
CNN1D = [(0, 1, 2, 2), (3, 4, 4, 4), (6, 7, 0, 1), (9, 10, 8, 8)]  # (conv idx, bn idx, pad, dilation)


class _Q(torch.autograd.Function):
    """bf16 storage emulation: rounds the value in forward AND the gradient in backward — the CUDA path stores
    every activation and every activation gradient as bf16 between kernels."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def _q(x, on):
    return _Q.apply(x) if on else x


def _qw(w, on):
    """weights are consumed as bf16 by the tensor cores (gradient flows to the fp32 master copy unrounded)"""
    return w + (w.detach().to(torch.bfloat16).to(w.dtype) - w.detach()) if on else w


def batchnorm(x, sd, prefix, training, update, dims):
    """nn.BatchNorm: train = batch stats (biased var), running stats get the unbiased var, momentum .1."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    shape = [1, -1] + [1] * (x.dim() - 2)
    if training:
        mean = x.mean(dims)
        var = x.var(dims, unbiased=False)
        if update is not None:
            n = x.numel() / x.size(1)
            update[prefix + ".running_mean"] = 0.9 * sd[prefix + ".running_mean"] + 0.1 * mean
            update[prefix + ".running_var"] = 0.9 * sd[prefix + ".running_var"] + 0.1 * var * n / (n - 1)
    else:
        mean, var = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    return (x - mean.view(shape)) / torch.sqrt(var.view(shape) + 1e-5) * w.view(shape) + b.view(shape)


def hwr_forward(sd, img, training=True, update=None, emulate_bf16=False):
    """img [B,1,64,W] -> log-probs [W/4-6, B, C]; `update` (dict) receives the new running stats.
    emulate_bf16=True rounds weights and every stored activation (and its gradient) to bf16 at the points where
    the CUDA path does — same arithmetic otherwise — so that ReLU / max-pool decisions coincide and the
    backward kernels can be checked tightly."""
    e = emulate_bf16
    x = img
    for i in range(7):
        w = sd[f"cnn.conv{i}.weight"]
        x = F.conv2d(x, _qw(w, e and i > 0), sd[f"cnn.conv{i}.bias"], padding=PADS[i])          # :31
        if i in BN2D:
            x = _q(x, e)                                                                        # conv output stored
            x = batchnorm(x, sd, f"cnn.batchnorm{i}", training, update, (0, 2, 3))              # :36
        x = F.relu(x)
        if i != 0:
            x = _q(x, e)                                # relu(conv) / relu(bn) stored (the stem stores after its pool)
        if i in POOL_AFTER:
            k, s, p = POOL_AFTER[i]
            x = F.max_pool2d(x, k, s, p)                                                        # :46-56
        if i == 0:
            x = _q(x, e)
    b, c, h, w = x.shape
    x = x.view(b, -1, w)                                                                        # :100
    for ci, bi, pad, dil in CNN1D:
        x = F.conv1d(x, _qw(sd[f"cnn1d.{ci}.weight"], e), sd[f"cnn1d.{ci}.bias"], padding=pad, dilation=dil)  # :78-89
        x = _q(x, e)
        x = _q(F.relu(batchnorm(x, sd, f"cnn1d.{bi}", training, update, (0, 2))), e)
    x = F.conv1d(x, _qw(sd["cnn1d.12.weight"], e), sd["cnn1d.12.bias"])                        # :90
    x = _q(x, False)
    if e:
        x = _QGradOnly.apply(x)                          # the logit gradient is stored as bf16 for dgrad/wgrad
    return F.log_softmax(x, dim=1).permute(2, 0, 1)                                             # :91,105


class _QGradOnly(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)
