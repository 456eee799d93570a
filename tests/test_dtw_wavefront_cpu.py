"""CPU: the anti-diagonal wavefront formulation of the DTW that `hwg_dtw_align` (csrc/hwg_dtw.cu) uses — thread j owns column
j, cell (i, j) is computed at step i + j from the two previous diagonals, first-minimum ties, byte history, backtrack —
mirrored lane for lane in numpy and compared with the oracle (bit-exact against the reference goldens,
tests/test_style_cpu.py).  The kernel itself has not run on a GPU yet (tools/pending_test_dtw_gpu.py); this pins its
index arithmetic."""
import numpy as np
import pytest

from oracle import style as ostyle
from oracle.make_golden import dtw_inputs


def wavefront_align(pred, label):
    """Same variables as dtw_align_kernel: d1 / d2 = diagonals s-1 / s-2 indexed by column, hist [T, L] bytes."""
    T, B, C = pred.shape
    S = label.shape[0]
    L = 2 * S + 1
    w = max(T // 2, abs(T - L))
    res = []
    for b in range(B):
        lab = np.zeros(L, np.int64)
        lab[1::2] = label[:, b]
        j = np.arange(L + 1)
        cls = np.where(j >= 1, lab[np.maximum(j - 1, 0)], 0)
        d1 = np.full(L + 1, np.inf, np.float32)
        d2 = np.full(L + 1, np.inf, np.float32)
        hist = np.zeros((T, L), np.uint8)
        for s in range(T + L + 1):
            i = s - j
            v = np.full(L + 1, np.inf, np.float32)
            v[(i == 0) & (j == 0)] = 0
            act = (i >= 1) & (i <= T) & (j >= 1) & (j >= i - w) & (j <= i + w)
            up = d1
            left = np.concatenate(([np.inf], d1[:-1])).astype(np.float32)
            diag = np.concatenate(([np.inf], d2[:-1])).astype(np.float32)
            cost = np.float32(1) - pred[np.clip(i - 1, 0, T - 1), b, cls]
            k = np.zeros(L + 1, np.int64)
            m = up.copy()
            t = diag < m
            m[t], k[t] = diag[t], 1
            t = left < m
            m[t], k[t] = left[t], 2
            v[act] = (cost + m)[act]
            hist[(i - 1)[act], (j - 1)[act]] = k[act]
            d2, d1 = d1, v
        i, jj = T - 1, L - 1
        rev = [lab[jj]]
        while i > 0 or jj > 0:
            h = hist[i, jj]
            if h == 0:
                i -= 1
            elif h == 1:
                i -= 1
                jj -= 1
            else:
                jj -= 1
            rev.append(lab[jj])
        res.append(rev[::-1])
    out = np.zeros((max(len(r) for r in res), B), np.int64)
    for b, r in enumerate(res):
        out[:len(r), b] = r
    return out


@pytest.mark.parametrize("T,B,L,seed", [(58, 3, 9, 611), (26, 3, 15, 2), (40, 2, 40, 3), (124, 2, 30, 612)])
def test_wavefront_equals_the_row_major_recurrence(T, B, L, seed):
    pred, label = dtw_inputs(T, B, L, seed)
    assert np.array_equal(wavefront_align(pred.numpy(), label.numpy()), ostyle.correct_pred(pred.numpy(), label.numpy()))
