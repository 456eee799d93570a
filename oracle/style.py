"""TEST INFRASTRUCTURE — CPU fp32 restatement of the reference style extractor `CharStyleEncoder` (model/char_style.py:126-310)
in the shape the IAM / RIMES GAN configs build it (hw_with_style.py:108-131: `style: "char"`, `char_style_dim: 0` ->
single_style, `char_style_window: 2` -> the small CharExtractor, `style_norm: group`, `style_activ: relu`, replicate
padding) and of the DTW alignment `correct_pred` (model/hw_with_style.py:18-74).  Groundwork for SURVEY.md §8 f3: there is
no CUDA counterpart yet.  Pinned by tests/golden/style.npz (tests/test_style_cpu.py).

The reference walks the 79 character classes and, per class, the batch and every position in Python (`.nonzero()`,
`.item()`, one `F.pad` + one `CharExtractor` call per class, char_style.py:205-232).  This restatement is the form a device
version would take: ONE gather of all (sample, position) windows whose arg-max class is a character, the per-class
extractors applied to the windows grouped by class, and the score-weighted average per sample as an index_add."""
import numpy as np
import torch
import torch.nn.functional as F

WINDOW = 2


def _gn(x, sd, prefix, groups=8):
    return F.group_norm(x, groups, sd[prefix + ".weight"], sd[prefix + ".bias"], 1e-5)


def _block(x, sd, prefix, stride, pad, norm=True, act=True):
    """Conv2dBlock (char_style.py:9-80) with replicate padding (left, right, top, bottom), GroupNorm(8), ReLU."""
    x = F.pad(x, pad, mode="replicate")
    x = F.conv2d(x, sd[prefix + ".conv.weight"], sd[prefix + ".conv.bias"], stride=stride)
    if norm:
        x = _gn(x, sd, prefix + ".norm")
    return F.relu(x) if act else x


def down(sd, image):
    """`self.down` (:154-166): [B,1,64,W] -> [B,256,W/4-2] (height collapses to 1)."""
    x = _block(image, sd, "down.0", 1, (2, 2, 2, 2))
    x = _block(x, sd, "down.1", 2, (1, 1, 1, 1))
    x = _block(x, sd, "down.2", 1, (1, 1, 0, 0))
    x = _block(x, sd, "down.3", 2, (1, 1, 1, 1))
    x = _block(x, sd, "down.4", 1, (1, 1, 0, 0))
    x = _block(x, sd, "down.5", (2, 1), (1, 1, 0, 0))
    x = _block(x, sd, "down.6", (2, 1), (1, 1, 0, 0), norm=False, act=False)
    assert x.size(2) == 1
    return x[:, :, 0]


def char_extractor(sd, prefix, patches):
    """CharExtractor with small=True (:82-124): patches [P,256,5] -> [P,128].  `conv1` opens with a NON-inplace ReLU, so the
    residual is the raw patch."""
    x = F.conv1d(F.relu(patches), sd[prefix + ".conv1.1.weight"], sd[prefix + ".conv1.1.bias"], padding=1)
    x = F.relu(_gn(x, sd, prefix + ".conv1.2"))
    x = F.conv1d(x, sd[prefix + ".conv1.4.weight"], sd[prefix + ".conv1.4.bias"], padding=1)
    x = F.conv1d(F.relu(x + patches), sd[prefix + ".conv2.1.weight"], sd[prefix + ".conv2.1.bias"])
    x = F.relu(_gn(x, sd, prefix + ".conv2.2"))
    x = x.mean(2)                                                                        # adaptive_avg_pool1d(x, 1)
    x = F.relu(F.linear(x, sd[prefix + ".fc.0.weight"], sd[prefix + ".fc.0.bias"]))
    return F.linear(x, sd[prefix + ".fc.2.weight"], sd[prefix + ".fc.2.bias"])


def gather_windows(x, recog, window=WINDOW):
    """All (sample, position) pairs whose arg-max class is a character (> 0), in the reference's visiting order
    (class, sample, position): returns (cls [P], b [P], pos [P], score [P] = exp(recog[b, cls, pos]),
    patches [P, C, 2*window+1] zero-padded at the line ends)."""
    pred = recog.argmax(1)                                                               # [B, Wx]
    b, pos = torch.nonzero(pred > 0, as_tuple=True)
    cls = pred[b, pos]
    order = torch.argsort(cls * (pred.numel() + 1) + b * pred.size(1) + pos)
    b, pos, cls = b[order], pos[order], cls[order]
    xp = F.pad(x, (window, window))
    idx = pos[:, None] + torch.arange(2 * window + 1)[None, :]                           # [P, 5] into the padded line
    patches = xp[b[:, None], :, idx].permute(0, 2, 1)                                    # [P, C, 5]
    score = torch.exp(recog[b, cls, pos])
    return cls, b, pos, score, patches


def char_style_forward(sd, image, recog, n_class=80, window=WINDOW):
    """image [B,1,64,W], recog [B,n_class,T] (the recognizer's log-probs, hw_with_style.py:284) -> style [B,128]."""
    B = image.size(0)
    x = down(sd, image)
    diff = x.size(2) - recog.size(2)
    if diff > 0:
        recog = F.pad(recog, (diff // 2, diff // 2 + diff % 2), mode="replicate")       # :196
    elif diff < 0:
        x = F.pad(x, (-diff // 2, (-diff // 2) + (-diff) % 2), mode="replicate")        # :198
    cls, b, pos, score, patches = gather_windows(x, recog, window)
    style_dim = sd["char_extractor.1.fc.2.weight"].size(0)
    total = torch.zeros(B, style_dim)
    b_sum = torch.zeros(B)
    if cls.numel():
        styles = torch.empty(cls.numel(), style_dim)
        for c in torch.unique(cls).tolist():                        # grouped by class: one extractor call per class
            sel = (cls == c).nonzero()[:, 0]
            styles[sel] = char_extractor(sd, f"char_extractor.{c}", patches[sel])
        total = total.index_add(0, b, score[:, None] * styles)                           # :226-228
        b_sum = b_sum.index_add(0, b, score)
    avg = torch.where(b_sum[:, None] != 0, total / b_sum[:, None], total)                # :287
    xr = torch.cat((F.relu(x), recog), 1)                                                # :289
    xr = F.relu(F.conv1d(xr, sd["prep.0.weight"], sd["prep.0.bias"], padding=2))
    xr = F.max_pool1d(xr, 2, 2)
    xr = F.relu(_gn(F.conv1d(xr, sd["prep.3.weight"], sd["prep.3.bias"], padding=1), sd, "prep.4"))
    xr = F.relu(F.conv1d(xr, sd["prep.6.weight"], sd["prep.6.bias"], padding=1))
    comb = torch.cat((xr.mean(2), avg), 1)
    comb = F.relu(F.linear(comb, sd["final_g_spacing_style.0.weight"], sd["final_g_spacing_style.0.bias"]))
    return F.linear(comb, sd["final_g_spacing_style.2.weight"], sd["final_g_spacing_style.2.bias"])


def correct_pred(pred, label):
    """hw_with_style.py:18-74: DTW alignment of the label (blanks in front, behind and between the characters) to the
    recognizer output pred [T,B,C]; returns the aligned label [T',B] (int64, zero-padded).  Vectorised over the batch;
    the (i, j) recurrence is inherent.  Ties take the first of (up, diagonal, left), as torch.min does on the CPU."""
    pred = np.asarray(pred, np.float32)
    label = np.asarray(label, np.int64)
    T, B, _ = pred.shape
    L = label.shape[0] * 2 + 1
    lab = np.zeros((L, B), np.int64)
    lab[1::2] = label
    dtw = np.full((T + 1, L + 1, B), np.inf, np.float32)
    dtw[0, 0] = 0
    w = max(T // 2, abs(T - L))
    for i in range(1, T + 1):
        dtw[i, max(1, i - w):min(L, i + w) + 1] = 0
    hist = np.zeros((T, L, B), np.int32)
    ar = np.arange(B)
    for i in range(1, T + 1):
        for j in range(max(1, i - w), min(L, i + w) + 1):
            cost = np.float32(1) - pred[i - 1, ar, lab[j - 1]]
            cand = np.stack((dtw[i - 1, j], dtw[i - 1, j - 1], dtw[i, j - 1]))
            k = cand.argmin(0)
            hist[i - 1, j - 1] = k
            dtw[i, j] = cost + cand[k, ar]
    lines = []
    for bb in range(B):
        i, j = T - 1, L - 1
        out = [lab[j, bb]]
        while i > 0 or j > 0:
            h = hist[i, j, bb]
            if h == 0:
                i -= 1
            elif h == 1:
                i -= 1
                j -= 1
            else:
                j -= 1
            out.append(lab[j, bb])
        lines.append(out[::-1])
    maxlen = max(len(ln) for ln in lines)
    res = np.zeros((maxlen, B), np.int64)
    for bb, ln in enumerate(lines):
        res[:len(ln), bb] = ln
    return res
