"""TEST INFRASTRUCTURE — CPU fp32 restatement of the reference perceptual encoder Encoder2 (model/autoencoder.py:341-410;
`encoder_type: "2tight"` -> Encoder2(32), trainer/hw_with_style_trainer.py:148-149) and of the perceptual loss the 'auto'
lessons build on it (trainer :724-748).  Groundwork for SURVEY.md §8 f1 (second half): there is no CUDA counterpart yet.
Pinned by tests/golden/enc.npz."""
import torch
import torch.nn.functional as F

# Dropout2d sites in forward order: (state_dict prefix of the module they follow, channels, p)
DROPOUT_SITES = (("conv1.2", 32, 0.1), ("conv2.0", 64, 0.1), ("conv2.4", 64, 0.1), ("down_conv3.4", 128, 0.1))


def _gn(x, sd, prefix):
    return F.group_norm(x, 8, sd[prefix + ".weight"], sd[prefix + ".bias"], 1e-5)     # getGroupSize(c) == 8 for c >= 32


def _conv(x, sd, prefix, padding=0):
    return F.conv2d(x, sd[prefix + ".weight"], sd[prefix + ".bias"], padding=padding)


def encoder2_forward(sd, x, masks=None, training=False):
    """x [B,1,64,W] -> (features [B,out_dim,1,W/8-4], mid_features [B,64,16,W/4]).  masks: list of four [B,C] keep-masks
    (training only).  The in-place ReLUs that open `conv1` / `down_conv2` … act on the tensor the residual aliases
    (autoencoder.py:399-406: `res = x; x = self.conv1(x); x += res`), so the residual is the ReLU'd tensor."""
    def drop(t, i):
        if not training:
            return t
        p = DROPOUT_SITES[i][2]
        return t * (masks[i] / (1.0 - p))[:, :, None, None]

    x = _conv(x, sd, "down_conv1.0", 2)                                   # :346
    x = F.avg_pool2d(F.relu(_gn(x, sd, "down_conv1.1")), 2)               # :347-349
    x = _conv(x, sd, "down_conv1.4")                                      # :350
    res = F.relu(x)                                                       # conv1[0] is ReLU(inplace) on the aliased tensor
    x = _conv(res, sd, "conv1.1", 1)                                      # :357
    x = F.relu(drop(_gn(x, sd, "conv1.2"), 0))                            # :358-360
    x = _conv(x, sd, "conv1.5", 1) + res                                  # :361, :401
    x = F.avg_pool2d(F.relu(_gn(x, sd, "down_conv2.0")), 2)               # :365-367
    x = _conv(x, sd, "down_conv2.3")                                      # :368
    res = x
    x = F.relu(drop(_gn(x, sd, "conv2.0"), 1))                            # :372-374 (GroupNorm first: no aliasing here)
    x = _conv(x, sd, "conv2.3", 1)
    x = F.relu(drop(_gn(x, sd, "conv2.4"), 2))
    x = _conv(x, sd, "conv2.7", 1) + res                                  # :379, :405
    mid = x
    x = F.avg_pool2d(F.relu(_gn(x, sd, "down_conv3.0")), 2)               # :383-385
    x = _conv(x, sd, "down_conv3.3")                                      # :386 (3x3, no padding)
    x = F.relu(drop(_gn(x, sd, "down_conv3.4"), 3))                       # :387-389
    x = _conv(x, sd, "down_conv3.7")                                      # :393 (6x3, no padding)
    return x, mid


def perceptual_loss(sd, image, recon, masks=None, training=False):
    """trainer :740-748 for equal widths >= 40: both images through the encoder as one batch, L1 between the halves of each
    returned feature tensor, summed."""
    both = encoder2_forward(sd, torch.cat((image, recon), 0), masks, training)
    loss = 0
    for f in both:
        o_f, r_f = torch.chunk(f, 2, dim=0)
        loss = loss + F.l1_loss(r_f, o_f)
    return loss
