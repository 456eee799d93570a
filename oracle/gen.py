"""TEST INFRASTRUCTURE — CPU fp32 restatement of the reference generator, written from the
semantics in SURVEY.md Appendix B (each step cites reference model/pure_gen.py).  It takes a
reference-format state_dict, so the same weights drive the reference, this oracle and the CUDA
path.  Pinned by tests/golden/gen_*.npz (outputs of the unmodified reference)."""
from math import sqrt

import torch
import torch.nn.functional as F

BLUR = torch.tensor([[1., 2., 1.], [2., 4., 2.], [1., 2., 1.]]) / 16.0


def adain(x, style, w, b):
    """pure_gen.py:62-69: gamma*InstanceNorm(x)+beta, biased variance, eps 1e-5."""
    gb = F.linear(style, w, b)
    C = x.size(1)
    gamma, beta = gb[:, :C, None, None], gb[:, C:, None, None]
    mean = x.mean((2, 3), keepdim=True)
    var = x.var((2, 3), keepdim=True, unbiased=False)
    return gamma * (x - mean) / torch.sqrt(var + 1e-5) + beta


def blur(x):
    """pure_gen.py:80-137: depthwise [1,2,1]x[1,2,1]/16, zero padding 1."""
    C = x.size(1)
    return F.conv2d(x, BLUR.to(x).view(1, 1, 3, 3).repeat(C, 1, 1, 1), padding=1, groups=C)


def noise_weight(sd, key):
    w = sd[key]
    return w * sqrt(2.0 / w.size(1))  # EqualLR, pure_gen.py:222-226


class _Q(torch.autograd.Function):
    """bf16 storage emulation (value in forward, gradient in backward) — see oracle/hwr.py."""

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


def generator_forward(sd, content, style, noise, n_blocks=5, trace=None, emulate_bf16=False):
    """sd: SpacedGenerator.state_dict(); content [T,B,C]; style [B,S]; noise: list of 2*n_blocks
    [B,C,H,W] tensors in the order the reference draws them (pure_gen.py:206,212).
    emulate_bf16=True rounds the conv weights and every image-sized tensor the CUDA path stores (and its
    gradient) to bf16 at the same points; the arithmetic is otherwise unchanged."""
    e = emulate_bf16
    x = content.permute(1, 2, 0).unsqueeze(2)                        # :43-44 -> [B,C,1,T]
    s = style / torch.sqrt((style * style).mean(1, keepdim=True) + 1e-8)  # PixelNorm :311
    i = 1
    while f"style_emb.{i}.weight" in sd:                                # :31-39 Linear + LeakyReLU(0.2)
        s = F.leaky_relu(F.linear(s, sd[f"style_emb.{i}.weight"], sd[f"style_emb.{i}.bias"]), 0.2)
        i += 2
    x = torch.cat((x, s[:, :, None, None].expand(-1, -1, 1, x.size(3))), 1)  # :47-48
    x = _q(x, e)
    k = 0
    for b in range(n_blocks):
        p = f"conv.{b}."
        if p + "conv1.weight" in sd and sd[p + "conv1.weight"].dim() == 4 and sd[p + "conv1.weight"].size(2) == 4:
            x = F.conv_transpose2d(x, _qw(sd[p + "conv1.weight"], e), sd[p + "conv1.bias"], padding=(0, 1))   # :161-163
        elif p + "conv1.1.weight" in sd and sd[p + "conv1.1.weight"].size(1) != 1:
            x = F.interpolate(x, scale_factor=(2, 1), mode="nearest")                                  # :181
            x = blur(_q(F.conv2d(x, _qw(sd[p + "conv1.1.weight"], e), sd[p + "conv1.1.bias"], padding=1), e))  # :182-185
        elif p + "conv1.0.weight" in sd:
            w = sd[p + "conv1.0.weight"]
            w = F.pad(w * sqrt(2.0 / (w.size(0) * 9)), [1, 1, 1, 1])                                     # :259-271
            w = (w[:, :, 1:, 1:] + w[:, :, :-1, 1:] + w[:, :, 1:, :-1] + w[:, :, :-1, :-1]) / 4
            x = blur(_q(F.conv_transpose2d(x, _qw(w, e), sd[p + "conv1.0.bias"], stride=2, padding=1), e))  # :277
        else:
            x = F.conv2d(x, _qw(sd[p + "conv1.weight"], e), sd[p + "conv1.bias"], padding=1)
        for j in (1, 2):
            if j == 2:
                x = F.conv2d(x, _qw(sd[p + "conv2.weight"], e), sd[p + "conv2.bias"], padding=1)       # :211
            x = x + noise_weight(sd, p + f"noise{j}.weight_orig") * noise[k]                           # :206,212
            x = _q(F.leaky_relu(x, 0.2), e)
            if trace is not None:
                trace.append(("pre_adain", x))
            x = adain(x, s, sd[p + f"adain{j}.style.weight"], sd[p + f"adain{j}.style.bias"])
            if not (b == n_blocks - 1 and j == 2):
                x = _q(x, e)          # the last AdaIN is fused into the output kernel, never stored
            if trace is not None:
                trace.append(("post_adain", x))
            k += 1
    w = sd["out.0.conv.weight_orig"]
    x = F.conv2d(x, w * sqrt(2.0 / w.size(1)), sd["out.0.conv.bias"])                                 # :29,285-288
    return torch.tanh(x)
