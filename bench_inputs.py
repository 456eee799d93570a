"""Deterministic synthetic inputs of the benchmarks, tools and tests (numpy RandomState: identical on every box).

Input builders only — nothing here computes a reference result, and the product benchmarks import this module instead
of anything under oracle/ (which is test infrastructure: the checker).  oracle/synth.py re-exports these for the golden
generator and the tests."""
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


def gen_case(T, B, n_class=80, style_dim=128, seed=0, dense=False):
    """Generator inputs: spaced one-hot content [T,B,C] built from random text —
    each character index ~U{1..C-1} emitted twice, separated by two blanks (class 0), truncated to T —
    or, with dense=True, softmax(randn) content as generate.py:834 feeds; style ~ N(0,1) [B,S]."""
    r = np.random.RandomState(seed)
    content = np.zeros((T, B, n_class), np.float32)
    if dense:
        x = r.standard_normal((T, B, n_class)).astype(np.float32)
        e = np.exp(x - x.max(2, keepdims=True))
        content = (e / e.sum(2, keepdims=True)).astype(np.float32)
    else:
        for b in range(B):
            seq = []
            while len(seq) < T:
                ch = int(r.randint(1, n_class))
                seq += [ch, ch, 0, 0]
            content[np.arange(T), b, np.array(seq[:T])] = 1.0
    style = r.standard_normal((B, style_dim)).astype(np.float32)
    return content, style


def gen_noise(shapes, seed):
    """The ten N(0,1) tensors NoiseInjection consumes, [B,C,H,W] each, in call order."""
    r = np.random.RandomState(seed)
    return [r.standard_normal(s).astype(np.float32) for s in shapes]


def gen_noise_shapes(T, B, dim=256):
    shapes = []
    H, W = 4, T
    for i, c in enumerate([dim, dim // 2, dim // 4, dim // 8, dim // 16]):
        if i in (1, 2):
            H *= 2
        elif i in (3, 4):
            H, W = H * 2, W * 2
        shapes += [(B, c, H, W)] * 2
    return shapes


def hwr_case(B, W, seed, H=64):
    r = np.random.RandomState(seed)
    return (r.rand(B, 1, H, W).astype(np.float32) * 2 - 1)
