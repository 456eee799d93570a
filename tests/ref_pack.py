"""TEST INFRASTRUCTURE — the generator's weight re-layouts written with plain torch ops, straight from the
reference's definitions (model/pure_gen.py:161-163,176-186,250-279).  The hwg_linear_map job tables
(handwriting_line_generation_b200/weightmap.py) are checked against these."""
import torch
import torch.nn.functional as F


def pack_taps(mats, cin_pad=None):
    w = torch.stack(list(mats), 0)
    cin = w.size(2)
    cin_pad = cin_pad or ((cin + 15) // 16) * 16
    if cin_pad != cin:
        w = F.pad(w, (0, cin_pad - cin))
    return w.to(torch.bfloat16).contiguous()


def conv3x3_fwd(w):
    return pack_taps([w[:, :, i, j] for i in range(3) for j in range(3)])


def conv3x3_dgrad(w):
    return pack_taps([w[:, :, i, j].t() for i in range(3) for j in range(3)])


def initial_fwd(w, cin_pad):
    """[3][4*Co][cin_pad]: the four output rows as channel folds (row r at channels r*Co..)."""
    packs = [pack_taps([w[:, :, r, kx].t() for kx in range(3)], cin_pad) for r in range(4)]
    return torch.cat(packs, 1).contiguous()


def initial_dgrad(w):
    cin = w.size(0)
    cin16 = ((cin + 15) // 16) * 16
    return pack_taps([F.pad(w[:, :, r, kx], (0, 0, 0, cin16 - cin)) for r in range(4) for kx in range(3)])


def vert_up_fwd(w):
    out = []
    for par in (0, 1):
        rows = {}
        for kh in range(3):
            src = (par + kh - 1) // 2
            rows[src] = rows[src] + w[:, :, kh, :] if src in rows else w[:, :, kh, :]
        mats = [rows[dh][:, :, kw] for dh in sorted(rows) for kw in range(3)]
        out.append(pack_taps(mats))
    return out


def vert_up_dgrad(w):
    comb = {-1: [2], 0: [1, 2], 1: [0, 1], 2: [0]}
    mats = []
    for dh in (-1, 0, 1, 2):
        wk = sum(w[:, :, kh, :] for kh in comb[dh])
        for kw in range(3):
            mats.append(wk[:, :, kw].t())
    return pack_taps(mats)


def fused_w4(w, multiplier):
    wp = F.pad(w * multiplier, [1, 1, 1, 1])
    return (wp[:, :, 1:, 1:] + wp[:, :, :-1, 1:] + wp[:, :, 1:, :-1] + wp[:, :, :-1, :-1]) / 4


_SEL = {0: [(0, 1), (-1, 3)], 1: [(1, 0), (0, 2)]}


def fused_up_fwd(w, multiplier):
    w4 = fused_w4(w, multiplier)
    mats, taps = [], []
    for py in (0, 1):
        for px in (0, 1):
            for dh, ky in _SEL[py]:
                for dw, kx in _SEL[px]:
                    taps.append((dh, dw))
                    mats.append(w4[:, :, ky, kx].t())
    return pack_taps(mats), taps


def fused_up_dgrad(w, multiplier):
    w4 = fused_w4(w, multiplier)
    return pack_taps([w4[:, :, ky, kx] for ky in range(4) for kx in range(4)])


TAPS_UNION = [(dh, dw) for dh in (-1, 0, 1) for dw in (-1, 0, 1)]


def fused_up_folded(w, multiplier):
    """The four output parities of FusedUpsample as ONE launch over the 9 union taps, Cout = 4 folds x C (fold =
    2*py+px); (tap, parity) pairs a parity does not use carry zero weights."""
    w4 = fused_w4(w, multiplier)
    cin, cout = w4.shape[:2]
    sel = {0: {0: 1, -1: 3}, 1: {1: 0, 0: 2}}
    mats = []
    for dh, dw in TAPS_UNION:
        rows = []
        for py in (0, 1):
            for px in (0, 1):
                if dh in sel[py] and dw in sel[px]:
                    rows.append(w4[:, :, sel[py][dh], sel[px][dw]].t())
                else:
                    rows.append(torch.zeros((cout, cin), device=w4.device, dtype=w4.dtype))
        mats.append(torch.cat(rows, 0))
    return pack_taps(mats)
