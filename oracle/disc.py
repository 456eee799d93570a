"""TEST INFRASTRUCTURE — CPU fp32 restatement of the reference discriminator DiscriminatorAP
(model/discriminator_ap.py:68-161, dim=64, small=False) with its SpectralNorm wrapper (:11-65) on a
reference-format state_dict.  Pinned by tests/golden/disc.npz (outputs and input gradients of the unmodified
reference)."""
import torch
import torch.nn.functional as F

LEAK = 0.1
# (prefix, spectral-normalised) of the convolutions, in forward order; the Dropout2d sites (p) follow the conv they
# are listed with (discriminator_ap.py:81-130)
DROPOUT_P = {"convs1.3": 0.05, "convs3.4": 0.05, "convs4.0": 0.025, "convs4.4": 0.025, "convs4.7": 0.025,
             "convs4.11": 0.025}
DROPOUT_ORDER = ["convs1.3", "convs3.4", "convs4.0", "convs4.4", "convs4.7", "convs4.11"]
DROPOUT_C = {"convs1.3": 2, "convs3.4": 4, "convs4.0": 2, "convs4.4": 4, "convs4.7": 4, "convs4.11": 4}   # x dim


class _Q(torch.autograd.Function):
    """bf16 storage emulation (value in forward, gradient in backward), as oracle/hwr.py."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def _q(x, on):
    return _Q.apply(x) if on else x


def _qw(w, on):
    return w + (w.detach().to(torch.bfloat16).to(w.dtype) - w.detach()) if on else w


def l2normalize(v, eps=1e-12):
    return v / (v.norm() + eps)                                                   # :7-8


def spectral_weight(sd, prefix, update=None):
    """SpectralNorm._update_u_v (:19-32): ONE power iteration on (u, v) — run on every forward, in train and eval mode
    alike — then weight = w_bar / sigma with sigma = u . (W v); u, v are constants for autograd (.data)."""
    w = sd[prefix + ".module.weight_bar"]
    u, v = sd[prefix + ".module.weight_u"].detach(), sd[prefix + ".module.weight_v"].detach()
    wm = w.reshape(w.size(0), -1)
    v = l2normalize(torch.mv(wm.detach().t(), u))
    u = l2normalize(torch.mv(wm.detach(), v))
    sigma = u.dot(wm.mv(v))
    if update is not None:
        update[prefix + ".module.weight_u"], update[prefix + ".module.weight_v"] = u, v
    return w / sigma


def group_norm(x, sd, prefix, groups=8):
    return F.group_norm(x, groups, sd[prefix + ".weight"], sd[prefix + ".bias"], 1e-5)


def disc_forward(sd, x, masks=None, training=True, update=None, use_low=True, use_med=True, emulate_bf16=False):
    """x [B,1,64,W] -> list of [B,-1] predictions (medium scale, then low scale), discriminator_ap.py:140-161.
    masks: dict site -> [B,C] 0/1 float keep-masks of the Dropout2d layers (training only); `update` receives the new
    spectral-norm u / v vectors."""
    e = emulate_bf16

    def sn_conv(x, prefix, pad):
        w = spectral_weight(sd, prefix, update)
        return F.conv2d(x, _qw(w, e), sd[prefix + ".module.bias"], padding=pad)

    def drop(x, site):
        if not training:
            return x
        p = DROPOUT_P[site]
        return x * (masks[site] / (1.0 - p))[:, :, None, None]

    def lrelu(x):
        return _q(F.leaky_relu(x, LEAK), e)

    B = x.size(0)
    x = F.conv2d(x, _qw(sd["in_conv.0.weight"], e), sd["in_conv.0.bias"], padding=(0, 3))       # :75
    x = lrelu(group_norm(_q(x, e), sd, "in_conv.1"))                                        # :76-77
    x = lrelu(sn_conv(x, "convs1.0", (0, 1)))                                               # :82-83
    x = _q(F.avg_pool2d(x, 2), e)                                                           # :86
    x = lrelu(drop(sn_conv(x, "convs1.3", (0, 1)), "convs1.3"))                             # :88-90
    x = lrelu(sn_conv(x, "convs2.0", (0, 1)))                                               # :96-97
    x = _q(F.avg_pool2d(x, 2), e)                                                           # :98
    x = F.conv2d(x, _qw(sd["convs3.0.weight"], e), sd["convs3.0.bias"], padding=(0, 1))     # :100
    x = lrelu(group_norm(_q(x, e), sd, "convs3.1"))                                         # :101-102
    x = _q(F.avg_pool2d(x, 2), e)                                                           # :103
    mL = lrelu(drop(sn_conv(x, "convs3.4", (0, 1)), "convs3.4"))                            # :104-106
    out = []
    if use_med:
        out.append(sn_conv(mL, "finalMed.0", (0, 1)).view(B, -1))                            # :110-112
    if use_low:
        x = lrelu(drop(sn_conv(mL, "convs4.0", (0, 1)), "convs4.0"))                        # :115-117
        x = _q(F.avg_pool2d(x, (1, 2)), e)                                                  # :118
        x = lrelu(drop(sn_conv(x, "convs4.4", (0, 1)), "convs4.4"))                         # :119-121
        x = lrelu(drop(sn_conv(x, "convs4.7", (0, 1)), "convs4.7"))                         # :122-124
        x = _q(F.avg_pool2d(x, (1, 2)), e)                                                  # :125
        x = lrelu(drop(sn_conv(x, "convs4.11", (0, 1)), "convs4.11"))                       # :126-128
        out.append(sn_conv(x, "convs4.14", (0, 0)).view(B, -1))                              # :129
    return out


def gen_loss(preds):
    """Generator's adversarial loss, trainer/hw_with_style_trainer.py:810-821."""
    loss = 0
    for p in preds:
        loss = loss - p.mean()
    return loss / len(preds)


def hinge_loss(preds, n_real):
    """Discriminator hinge loss, trainer/hw_with_style_trainer.py:797-804 (real rows first, then fake)."""
    loss = 0
    for p in preds:
        loss = loss + F.relu(1.0 - p[:n_real]).mean() + F.relu(1.0 + p[n_real:]).mean()
    return loss / len(preds)
