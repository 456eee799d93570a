"""TEST INFRASTRUCTURE — deterministic synthetic inputs shared by the golden
generator, the tests and bench.py (numpy RandomState: identical on every box)."""
import numpy as np


def ctc_case(T, B, C, S, seed, ragged=True, min_frac=0.5, sharp=3.0):
    """log-probs [T,B,C] (log-softmax of sharp*N(0,1)), targets [B,S] in 1..C-1 (0-padded past
    the length), input_lengths [B] (= T, like the trainer passes), target_lengths [B]."""
    r = np.random.RandomState(seed)
    x = (r.standard_normal((T, B, C)) * sharp).astype(np.float32)
    m = x.max(axis=2, keepdims=True)
    lp = (x - m - np.log(np.exp(x - m).sum(axis=2, keepdims=True))).astype(np.float32)
    tg = r.randint(1, C, size=(B, S)).astype(np.int32)
    if ragged and S > 0:
        tl = r.randint(max(1, int(S * min_frac)), S + 1, size=B).astype(np.int32)
    else:
        tl = np.full(B, S, np.int32)
    for b in range(B):
        tg[b, tl[b]:] = 0
    il = np.full(B, T, np.int32)
    return lp, tg, il, tl
