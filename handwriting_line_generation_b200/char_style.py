"""Drop-in for the reference style extractor `CharStyleEncoder` (model/char_style.py:126-310) in the configuration the IAM /
RIMES GAN configs build (model/hw_with_style.py:108-131: `style: "char"`, `char_style_dim: 0` -> single style vector,
`char_style_window: 2` -> the small CharExtractor, `style_norm: group`, `style_activ: relu`, replicate padding) — SURVEY.md §8
row f3: the style vector of the 'auto' / 'count' lessons, 33.4 M parameters, ~20 GFLOP forward per 64x1024 line.

Same constructor signature, module names, construction order (same seed -> same initial weights) and `state_dict` keys as
the reference; the torch sub-modules are parameter containers.  Every convolution, GroupNorm and ReLU runs on libhwg_b200
through `nhwc.conv_block` (tcgen05 implicit GEMM forward / input gradient / weight gradient, statistics in the convolution
epilogue, normalisation passes), forward AND backward (the style extractor is trained):

* `down` (seven Conv2dBlocks, :154-166): the 5x5 stem from one input channel as a 5-tap launch over the 16-channel shift
  expansion of the replicate-padded image (`hwg_shift_expand`, the discriminator's in_conv construction); the stride-2 and
  stride-(2,1) 4x4 convolutions as stride-1 convolutions over a space-to-depth view (2x2 / 2x4 taps over 4x / 2x the channels,
  the kernel re-arranged by a differentiable view, so autograd returns the gradient in the parameter's own layout);
  replicate padding is an index gather;
* the per-character heads (`char_extractor[c]`, :82-124, 79 classes): the reference walks classes, samples and positions in
  Python (:205-232, `.nonzero()` / `.item()` per window); here ALL windows are gathered at once, sorted by class, and each
  layer of the heads is one GROUPED `conv_block` (one launch per present class for the three convolution kernels, every other
  pass over all windows at once) — ONE host read (the windows per class) instead of thousands;
* `prep` (:168-177) on the concatenation [ReLU(features) ; recognizer log-probs] as 1-D tap launches, the two small `fc` /
  `final_g_spacing_style` stacks as plain GEMMs (`F.linear`).

Only the shipped configuration family is implemented (single style, no VAE, window < 3, group norm, ReLU, replicate padding);
anything else raises.  There is no PyTorch fallback for the convolution / normalisation path."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, nhwc
from .discriminator_ap import get_group_size


class Conv2dBlock(nn.Module):
    """Parameter container with the reference's attribute names (char_style.py:9-80): `norm` (GroupNorm) before `conv`."""

    def __init__(self, input_dim, output_dim, kernel_size, stride, padding, norm=True):
        super().__init__()
        self.norm = nn.GroupNorm(get_group_size(output_dim), output_dim) if norm else None
        self.conv = nn.Conv2d(input_dim, output_dim, kernel_size, stride, bias=True)
        self.stride = stride if isinstance(stride, tuple) else (stride, stride)
        self.padding = padding if isinstance(padding, tuple) else (padding,) * 4       # (left, right, top, bottom)


class CharExtractor(nn.Module):
    """Parameter container of the small CharExtractor (char_style.py:82-124, `small=True`)."""

    def __init__(self, input_dim, dim, style_dim, num_fc=1):
        super().__init__()
        gs = get_group_size
        self.conv1 = nn.Sequential(nn.ReLU(), nn.Conv1d(input_dim, dim, 3, padding=1), nn.GroupNorm(gs(dim), dim), nn.ReLU(),
                                   nn.Conv1d(dim, input_dim, 3, padding=1))
        self.conv2 = nn.Sequential(nn.ReLU(), nn.Conv1d(input_dim, 2 * dim, 1), nn.GroupNorm(gs(2 * dim), 2 * dim), nn.ReLU())
        fc = [nn.Linear(2 * dim, 2 * dim), nn.ReLU(True)]
        for _ in range(style_dim, num_fc - 1):          # as the reference writes it (:113): empty for the shipped configs
            fc += [nn.Linear(2 * dim, 2 * dim), nn.Dropout(0.25, True), nn.ReLU(True)]
        fc.append(nn.Linear(2 * dim, style_dim))
        self.fc = nn.Sequential(*fc)


class CharStyleEncoder(nn.Module):
    def __init__(self, input_dim, dim, style_dim, char_dim, char_style_dim, norm, activ, pad_type, n_class, global_pool=False,
                 average_found_char_style=0, num_final_g_spacing_style=1, num_char_fc=1, vae=False, window=6, small=False):
        super().__init__()
        if (vae or char_style_dim > 0 or norm != 'group' or activ != 'relu' or pad_type != 'replicate' or small or window >= 3
                or num_final_g_spacing_style != 1 or num_char_fc != 1 or input_dim != 1):
            raise NotImplementedError("CharStyleEncoder: only the shipped GAN configuration family (single style, group norm, "
                                      "ReLU, replicate padding, window < 3) is implemented")
        self.vae, self.single_style = False, True
        self.n_class, self.window, self.char_style_dim = n_class, window, style_dim
        down = [Conv2dBlock(input_dim, dim, 5, 1, 2)]
        for _ in range(2):
            down.append(Conv2dBlock(dim, 2 * dim, 4, 2, 1))
            dim *= 2
            down.append(Conv2dBlock(dim, dim, 3, 1, (1, 1, 0, 0)))
        down.append(Conv2dBlock(dim, dim, 4, (2, 1), (1, 1, 0, 0)))
        down.append(Conv2dBlock(dim, dim, 4, (2, 1), (1, 1, 0, 0), norm=False))
        self.down = nn.Sequential(*down)
        p = dim
        self.prep = nn.Sequential(nn.Conv1d(dim + n_class, p, 5, 1, 2), nn.ReLU(True), nn.MaxPool1d(2, 2), nn.Conv1d(p, p, 3, 1, 1),
                                  nn.GroupNorm(get_group_size(p), p), nn.ReLU(True), nn.Conv1d(p, p, 3, 1, 1), nn.ReLU(True))
        self.final_g_spacing_style = nn.Sequential(nn.Linear(p + style_dim, p), nn.ReLU(True), nn.Linear(p, style_dim))
        self.char_extractor = nn.ModuleList([CharExtractor(dim, char_dim, style_dim, num_char_fc) for _ in range(n_class)])
        self.feat_dim = dim

    # -- the image path --------------------------------------------------------------------------------------------
    def _down(self, image):
        """[B,1,64,W] fp32 -> [B,1,W/4-2,feat_dim] bf16 NHWC (the height collapses to 1)."""
        blocks = list(self.down)
        b0 = blocks[0]
        B, one, H, W = image.shape
        l, r, t, b = b0.padding
        xp = F.pad(image.float(), (l, r, t, b), mode="replicate").contiguous()
        Hp, Wp = H + t + b, W + l + r
        x5 = torch.empty((B, Hp, Wp, 16), device=image.device, dtype=torch.bfloat16)
        _lib.call("hwg_shift_expand", xp.data_ptr(), x5.data_ptr(), B, Hp, Wp, 5, 0, _lib.stream())
        # kernel row dy = tap (dy, 0); kernel column dx = channel dx of the shift expansion
        w0 = b0.conv.weight[:, 0].permute(0, 2, 1).unsqueeze(3)                          # [Co, 5 (dx), 5 (dy), 1]
        x = nhwc.conv_block(x5, w0, b0.conv.bias, [(dy, 0) for dy in range(5)], H, W, b0.norm.weight, b0.norm.bias,
                            b0.norm.num_groups, b0.norm.eps, relu=True)
        for blk in blocks[1:]:
            l, r, t, b = blk.padding
            x = nhwc.replicate_pad(x, l, r, t, b)
            sh, sw = blk.stride
            kh, kw = blk.conv.kernel_size
            Hp, Wp = x.size(1), x.size(2)
            Ho, Wo = (Hp - kh) // sh + 1, (Wp - kw) // sw + 1
            w = blk.conv.weight
            if (sh, sw) != (1, 1):
                x = nhwc.space_to_depth(x, sh, sw)
                w = nhwc.strided_weight(w, sh, sw)
            gn = blk.norm
            x = nhwc.conv_block(x, w, blk.conv.bias, nhwc.valid_taps(w.size(2), w.size(3)), Ho, Wo,
                                None if gn is None else gn.weight, None if gn is None else gn.bias,
                                8 if gn is None else gn.num_groups, 1e-5 if gn is None else gn.eps, relu=gn is not None)
        if x.size(1) != 1:
            raise RuntimeError(f"CharStyleEncoder expects 64-px-high images (feature height {x.size(1)} != 1)")
        return x

    # -- the per-character heads -------------------------------------------------------------------------------------
    def _heads(self, patches, present, counts):
        """patches [P,1,2*window+1,feat_dim] bf16 sorted by class; present: the classes that occur (ascending), counts: windows
        per present class -> char styles [P, style_dim] fp32 (CharExtractor.forward, :116-124)."""
        ex = [self.char_extractor[c] for c in present]
        K = patches.size(2)
        t3 = [(0, -1), (0, 0), (0, 1)]

        def stack(get):
            return torch.stack([get(e) for e in ex], 0)

        gn1, gn2 = ex[0].conv1[2], ex[0].conv2[2]
        x = nhwc.conv_block(F.relu(patches), stack(lambda e: e.conv1[1].weight).unsqueeze(3), stack(lambda e: e.conv1[1].bias),
                            t3, 1, K, stack(lambda e: e.conv1[2].weight), stack(lambda e: e.conv1[2].bias), gn1.num_groups,
                            gn1.eps, relu=True, counts=counts)
        x = nhwc.conv_block(x, stack(lambda e: e.conv1[4].weight).unsqueeze(3), stack(lambda e: e.conv1[4].bias), t3, 1, K,
                            counts=counts)
        x = nhwc.conv_block(F.relu(x + patches), stack(lambda e: e.conv2[1].weight).unsqueeze(3),
                            stack(lambda e: e.conv2[1].bias), [(0, 0)], 1, K, stack(lambda e: e.conv2[2].weight),
                            stack(lambda e: e.conv2[2].bias), gn2.num_groups, gn2.eps, relu=True, counts=counts)
        v = x[:, 0].float().mean(1)                                                      # adaptive_avg_pool1d(x, 1)
        outs, s = [], 0
        for e, n in zip(ex, counts):
            h = v[s:s + n]
            for m in e.fc:
                h = m(h) if not isinstance(m, nn.ReLU) else F.relu(h)
            outs.append(h)
            s += n
        return torch.cat(outs, 0)

    # -- forward ---------------------------------------------------------------------------------------------------
    def forward(self, x, recog):
        """x [B,1,64,W] image, recog [B,n_class,T] recognizer log-probs (hw_with_style.py:284) -> style [B,style_dim]."""
        _lib.require_cuda(x, recog)
        B = x.size(0)
        dev = x.device
        feat = self._down(x)                                                             # [B,1,Wx,D] bf16
        recog = recog.float()
        diff = feat.size(2) - recog.size(2)
        if diff > 0:
            recog = F.pad(recog, (diff // 2, diff // 2 + diff % 2), mode="replicate")    # :196
        elif diff < 0:
            feat = nhwc.replicate_pad(feat, -diff // 2, (-diff // 2) + (-diff) % 2, 0, 0)    # :198
        Wx, D, w = feat.size(2), feat.size(3), self.window
        # every (sample, position) whose arg-max class is a character, in the reference's visiting order (class, b, pos)
        pred = recog.argmax(1)                                                           # [B,Wx]
        b_idx, pos = torch.nonzero(pred > 0, as_tuple=True)
        cls = pred[b_idx, pos]
        order = torch.argsort(cls * (pred.numel() + 1) + b_idx * Wx + pos)
        b_idx, pos, cls = b_idx[order], pos[order], cls[order]
        per_class = torch.bincount(cls, minlength=self.n_class).tolist()                 # the ONE host read
        total = torch.zeros((B, self.char_style_dim), device=dev, dtype=torch.float32)
        b_sum = torch.zeros(B, device=dev, dtype=torch.float32)
        if cls.numel():
            present = [c for c in range(1, self.n_class) if per_class[c]]
            counts = [per_class[c] for c in present]
            fp = F.pad(feat[:, 0], (0, 0, w, w))                                         # zero padding at the line ends (:218-219)
            idx = pos[:, None] + torch.arange(2 * w + 1, device=dev)[None, :]
            patches = fp[b_idx[:, None], idx].unsqueeze(1).contiguous()                  # [P,1,2w+1,D]
            styles = self._heads(patches, present, counts)
            score = torch.exp(recog[b_idx, cls, pos])                                    # :221
            total = total.index_add(0, b_idx, score[:, None] * styles)                   # :226-228
            b_sum = b_sum.index_add(0, b_idx, score)
        avg = torch.where(b_sum[:, None] != 0, total / b_sum[:, None].clamp_min(1e-30), total)     # :287
        # prep (:289-292) on [ReLU(features) ; log-probs], channels padded to a multiple of 64
        C = D + self.n_class
        Cp = -(-C // 64) * 64
        xr = torch.zeros((B, 1, Wx, Cp), device=dev, dtype=torch.bfloat16)
        xr = torch.cat((F.relu(feat), recog.permute(0, 2, 1).unsqueeze(1).to(torch.bfloat16),
                        xr[..., :Cp - C]), 3)
        p = self.prep
        t5, t3 = [(0, k - 2) for k in range(5)], [(0, -1), (0, 0), (0, 1)]
        y = nhwc.conv_block(xr, p[0].weight.unsqueeze(2), p[0].bias, t5, 1, Wx, relu=True)
        Wh = Wx // 2
        y = y[:, :, :2 * Wh].reshape(B, 1, Wh, 2, y.size(3)).amax(3)                     # MaxPool1d(2, 2)
        y = nhwc.conv_block(y, p[3].weight.unsqueeze(2), p[3].bias, t3, 1, Wh, p[4].weight, p[4].bias, p[4].num_groups,
                            p[4].eps, relu=True)
        y = nhwc.conv_block(y, p[6].weight.unsqueeze(2), p[6].bias, t3, 1, Wh, relu=True)
        comb = torch.cat((y[:, 0].float().mean(1), avg), 1)                              # :294-296
        f = self.final_g_spacing_style
        return f[2](F.relu(f[0](comb)))
