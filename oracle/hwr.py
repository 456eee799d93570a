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


def hwr_forward(sd, img, training=True, update=None):
    """img [B,1,64,W] -> log-probs [W/4-6, B, C]; `update` (dict) receives the new running stats."""
    x = img
    for i in range(7):
        x = F.conv2d(x, sd[f"cnn.conv{i}.weight"], sd[f"cnn.conv{i}.bias"], padding=PADS[i])   # :31
        if i in BN2D:
            x = batchnorm(x, sd, f"cnn.batchnorm{i}", training, update, (0, 2, 3))              # :36
        x = F.relu(x)
        if i in POOL_AFTER:
            k, s, p = POOL_AFTER[i]
            x = F.max_pool2d(x, k, s, p)                                                        # :46-56
    b, c, h, w = x.shape
    x = x.view(b, -1, w)                                                                        # :100
    for ci, bi, pad, dil in CNN1D:
        x = F.conv1d(x, sd[f"cnn1d.{ci}.weight"], sd[f"cnn1d.{ci}.bias"], padding=pad, dilation=dil)  # :78-89
        x = F.relu(batchnorm(x, sd, f"cnn1d.{bi}", training, update, (0, 2)))
    x = F.conv1d(x, sd["cnn1d.12.weight"], sd["cnn1d.12.bias"])                                # :90
    return F.log_softmax(x, dim=1).permute(2, 0, 1)                                             # :91,105
