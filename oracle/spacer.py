"""TEST INFRASTRUCTURE — CPU restatement of the text-spacing front end of `HWWithStyle.forward`:
the spacer `CountCNN` (model/count_cnn.py:7-45, built by hw_with_style.py:200-204 as CountCNN(num_class, style_dim, 128, 2)
for `spacer: "CNN duplicates"`) and `insert_spaces` (hw_with_style.py:302-328) — the checker of the product's
`count_cnn.CountCNN` and `spacing.insert_spaces`.  Pinned by tests/golden/spacer.npz (tests/test_spacer_cpu.py).

The reference's insert_spaces walks batch x characters in Python with two `np.random.normal(...)` + `.item()` per character
(2*L*B host synchronisations on a GPU).  The restatement draws the same normals in ONE vectorised call (same legacy
RandomState stream, same order: sample-major, then character, count before duplicates), rounds half-to-even as Python's
round() does, and builds the spaced one-hot text from cumulative offsets — the form a device version would take."""
import math

import numpy as np
import torch
import torch.nn.functional as F


def count_cnn_forward(sd, label_onehot, style, masks=None):
    """label_onehot [L,B,C], style [B,S] -> counts [L,B,2] = (blanks before the character, repetitions of it).  masks: None
    = eval mode (Dropout2d off), else the two [B,C] keep-masks of the train-mode Dropout2d(0.1) sites (count_cnn.py:14,18;
    on a [B,C,L] tensor Dropout2d drops whole (sample, channel) rows)."""
    def drop(t, i):
        return t if masks is None else t * (masks[i] / 0.9)[:, :, None]

    x = torch.cat((label_onehot.permute(1, 2, 0), style[..., None].expand(-1, -1, label_onehot.size(0))), 1)
    x = F.relu(drop(F.group_norm(F.conv1d(x, sd["cnn.0.weight"], sd["cnn.0.bias"], padding=1), 8, sd["cnn.1.weight"], sd["cnn.1.bias"]), 0))
    x = F.relu(drop(F.group_norm(F.conv1d(x, sd["cnn.4.weight"], sd["cnn.4.bias"], padding=1), 8, sd["cnn.5.weight"], sd["cnn.5.bias"]), 1))
    x = F.relu(F.group_norm(F.conv1d(x, sd["cnn.8.weight"], sd["cnn.8.bias"], padding=1), 8, sd["cnn.9.weight"], sd["cnn.9.bias"]))
    x = F.conv1d(x, sd["cnn.11.weight"], sd["cnn.11.bias"])
    return x.permute(2, 0, 1) * sd["std"] + sd["mean"]


def insert_spaces(label, label_lengths, counts, num_class, count_std, dup_std, rng=np.random):
    """label [L,B] int, counts [L,B,2] -> (spaced one-hot [T,B,num_class] fp32, padded fractions), consuming `rng` exactly as
    the reference does."""
    label = np.asarray(label)
    lengths = [int(v) for v in label_lengths]
    c = np.asarray(counts, np.float32)
    B = label.shape[1]
    max_count = max(math.ceil(float(c.max())), 3)
    loc = np.concatenate([c[:n, b, :].reshape(-1) for b, n in enumerate(lengths)]).astype(np.float64)
    scale = np.tile(np.array([count_std, dup_std], np.float64), loc.size // 2)
    draws = np.rint(rng.normal(loc, scale)).astype(np.int64).reshape(-1, 2)            # python round(): half to even
    draws = np.maximum(draws, 0)                                                       # [0]*negative == []
    lines, off = [], 0
    for b, n in enumerate(lengths):
        d = draws[off:off + n]
        off += n
        reps = d.reshape(-1)                                                           # blanks, chars, blanks, chars, ...
        vals = np.stack((np.zeros(n, np.int64), label[:n, b].astype(np.int64)), 1).reshape(-1)
        lines.append(np.repeat(vals, reps))
    T = max(len(ln) for ln in lines) + max_count
    idx = np.zeros((T, B), np.int64)
    for b, ln in enumerate(lines):
        idx[:len(ln), b] = ln
    spaced = torch.zeros(T, B, num_class)
    spaced.scatter_(2, torch.from_numpy(idx)[..., None], 1.0)
    return spaced, [(T - len(ln)) / T for ln in lines]
