"""`HWWithStyle.insert_spaces` on the device (model/hw_with_style.py:302-328).

The reference builds the spaced text in Python, with two `np.random.normal(counts[i,b,k].item(), std)` calls per character —
2*L*B device->host synchronisations per generated batch.  Here the SAME standard normals are drawn from the same numpy stream
in one vectorised call (`np.random.standard_normal`: the legacy generator hands out the same values whether asked one by one
or at once, and `normal(loc, scale)` is `loc + scale * gauss` in doubles), uploaded, and two launches do the rest
(`hwg_insert_spaces_plan` / `hwg_insert_spaces_fill`).  One read of B+1 integers remains: the length of the result is data
dependent.  The spaced text is bit-identical to the reference's for the same RNG state (tests/test_spacing_gpu.py against the
goldens of the unmodified `insert_spaces`).

`insert_spaces(self, label, label_lengths, counts)` has the reference's signature and is bound as a method by
`integrate.install(spacer=True)`; `self` supplies `count_std`, `dup_std`, `count_duplicates`, `num_class`."""
import math

import numpy as np
import torch

from . import _lib


def insert_spaces(self, label, label_lengths, counts, rng=None):
    """label [L,B] integer class indices, label_lengths: B lengths (list / CPU tensor, as the data loader supplies them),
    counts [L,B,1 or 2] fp32 CUDA (the spacer's output) -> (spaced one-hot [T,B,num_class] fp32 on counts' device,
    padded: list of B floats).  rng: a numpy RandomState (default: the global stream the reference consumes)."""
    rng = np.random if rng is None else rng
    _lib.require_cuda(counts)
    dev = counts.device
    L, B = int(label.size(0)), int(label.size(1))
    n_out = int(counts.size(2))
    dup = bool(self.count_duplicates)
    if dup and n_out < 2:
        raise RuntimeError("insert_spaces: count_duplicates needs counts [L,B,2]")
    lens = [int(v) for v in label_lengths]
    if len(lens) != B or max(lens) > L or min(lens) < 0:
        raise RuntimeError("insert_spaces: label_lengths must hold one length in [0, L] per line")
    per_char = 2 if dup else 1
    # the reference's draws, in its order: line, character, count before duplicates — one call on the same stream
    z_host = torch.from_numpy(np.asarray(rng.standard_normal(per_char * sum(lens)), np.float64))
    z_off = np.concatenate(([0], np.cumsum([per_char * n for n in lens])[:-1])).astype(np.int64)
    z = torch.empty(max(1, z_host.numel()), dtype=torch.float64, device=dev)
    z[:z_host.numel()].copy_(z_host, non_blocking=True)
    z_off_d = torch.from_numpy(z_off).to(dev, non_blocking=True)
    lens_d = torch.tensor(lens, dtype=torch.int32).to(dev, non_blocking=True)
    c = counts.detach().float().contiguous()
    if not dup and n_out > 1:
        c = c[:, :, :1].contiguous()                    # the duplicates channel is not read (hw_with_style.py:311-314) ...
    reps = torch.empty((B, L, 2), dtype=torch.int32, device=dev)
    offsets = torch.empty((B, L + 1), dtype=torch.int32, device=dev)
    info = torch.empty(B + 1, dtype=torch.int32, device=dev)
    st = _lib.stream()
    _lib.call("hwg_insert_spaces_plan", lens_d.data_ptr(), c.data_ptr(), c.size(2), z.data_ptr(), z_off_d.data_ptr(), L, B,
              float(self.count_std), float(self.dup_std) if dup else 0.0, reps.data_ptr(), offsets.data_ptr(),
              info.data_ptr(), st)
    if not dup and n_out > 1:                           # ... but max_count looks at the whole tensor (:303)
        info[B] = torch.maximum(info[B], torch.ceil(counts.detach().float().max()).to(torch.int32))
    host = info.cpu().tolist()                          # the ONE synchronisation: line lengths + ceil(max counts)
    line_len, max_count = host[:B], max(host[B], 3)
    T = max(line_len) + max_count
    lab = label if label.is_cuda else label.to(dev, non_blocking=True)
    if lab.dtype not in (torch.int32, torch.int64):
        lab = lab.long()
    spaced = torch.empty((T, B, int(self.num_class)), dtype=torch.float32, device=dev)
    _lib.call("hwg_insert_spaces_fill", lab.data_ptr(), int(lab.dtype == torch.int64), lab.stride(0), lab.stride(1),
              lens_d.data_ptr(), reps.data_ptr(), offsets.data_ptr(), L, B, T, int(self.num_class), spaced.data_ptr(), st)
    return spaced, [(T - n) / T for n in line_len]
