"""TEST INFRASTRUCTURE — ctypes front end of oracle/ctc_oracle.c.

Restates F.ctc_loss as called by CTCLoss (reference model/loss.py:28-30) and
naive_decode (reference utils/string_utils.py:51-57) on numpy arrays.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "ctc_oracle.c")
SO = os.path.join(HERE, "_build", "ctc_oracle.so")
_lib = None


def build(force=False):
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=c99", "-fno-fast-math",
                               "-o", SO, SRC, "-lm"])
    return SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.ctc_oracle_reduce_mean.restype = ctypes.c_float
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _prep(log_probs, targets, input_lengths, target_lengths):
    lp = np.ascontiguousarray(log_probs, dtype=np.float32)
    tg = np.ascontiguousarray(targets, dtype=np.int32)
    il = np.ascontiguousarray(input_lengths, dtype=np.int32)
    tl = np.ascontiguousarray(target_lengths, dtype=np.int32)
    T, B, C = lp.shape
    assert tg.shape[0] == B
    S = tg.shape[1] if tg.ndim == 2 else 0
    return lp, tg, il, tl, T, B, C, S


def ctc_forward(log_probs, targets, input_lengths, target_lengths, blank=0):
    """-> (nll[B], log_alpha[B,T,2S+1])"""
    lp, tg, il, tl, T, B, C, S = _prep(log_probs, targets, input_lengths, target_lengths)
    nll = np.empty(B, np.float32)
    la = np.empty((B, T, 2 * S + 1), np.float32)
    lib().ctc_oracle_forward(_p(lp), T, B, C, _p(tg), ctypes.c_int64(S), ctypes.c_int64(1), S,
                             _p(il), _p(tl), blank, _p(nll), _p(la))
    return nll, la


def ctc_loss_mean(nll, target_lengths):
    """reduction='mean' + the reference wrapper's inf->0 (model/loss.py:30)."""
    nll = np.ascontiguousarray(nll, np.float32)
    tl = np.ascontiguousarray(target_lengths, np.int32)
    return float(lib().ctc_oracle_reduce_mean(_p(nll), _p(tl), len(nll)))


def ctc_backward(grad_nll, log_probs, targets, input_lengths, target_lengths, nll, log_alpha, blank=0):
    """-> (grad[T,B,C], log_beta[B,T,2S+1])"""
    lp, tg, il, tl, T, B, C, S = _prep(log_probs, targets, input_lengths, target_lengths)
    gn = np.ascontiguousarray(grad_nll, np.float32)
    nll = np.ascontiguousarray(nll, np.float32)
    la = np.ascontiguousarray(log_alpha, np.float32)
    lb = np.empty_like(la)
    grad = np.empty_like(lp)
    lib().ctc_oracle_backward(_p(gn), _p(lp), T, B, C, _p(tg), ctypes.c_int64(S), ctypes.c_int64(1),
                              S, _p(il), _p(tl), blank, _p(nll), _p(la), _p(lb), _p(grad))
    return grad, lb


def ctc_loss_and_grad(log_probs, targets, input_lengths, target_lengths, blank=0, grad_out=1.0):
    """CTCLoss(...) value and d loss / d log_probs, as the reference computes them."""
    nll, la = ctc_forward(log_probs, targets, input_lengths, target_lengths, blank)
    loss = ctc_loss_mean(nll, target_lengths)
    B = len(nll)
    tl = np.maximum(np.asarray(target_lengths, np.float32), 1.0)
    gn = (grad_out / (B * tl)).astype(np.float32)
    grad, _ = ctc_backward(gn, log_probs, targets, input_lengths, target_lengths, nll, la, blank)
    return loss, grad, nll


def greedy_decode(log_probs, input_lengths=None, blank=0):
    """-> (raw[T,B] int32, list of decoded int lists) — naive_decode per line."""
    lp = np.ascontiguousarray(log_probs, dtype=np.float32)
    T, B, C = lp.shape
    il = np.full(B, T, np.int32) if input_lengths is None else np.ascontiguousarray(input_lengths, np.int32)
    raw = np.empty((T, B), np.int32)
    dec = np.zeros((B, T), np.int32)
    dl = np.empty(B, np.int32)
    lib().ctc_oracle_decode(_p(lp), T, B, C, _p(il), blank, _p(raw), _p(dec), _p(dl))
    return raw, [dec[b, :dl[b]].tolist() for b in range(B)]
