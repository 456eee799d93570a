"""CTC loss and best-path decode on libhwg_b200's kernels.

Drop-in for the reference's `CTCLoss` (model/loss.py:28-30; call sites
trainer/hw_with_style_trainer.py:503,756,762) and `naive_decode`
(utils/string_utils.py:51-57).
"""
import torch

from . import _lib


def _lengths_to_device(x, B, device, what):
    if not torch.is_tensor(x):
        x = torch.tensor(list(x), dtype=torch.int32)
    if x.numel() != B:
        raise RuntimeError(f"{what} must have one entry per batch element ({B}), got {x.numel()}")
    return x.to(device=device, dtype=torch.int32, non_blocking=True).contiguous()


class _CTCLossFn(torch.autograd.Function):
    """loss = where(isinf(l), 0, l), l = F.ctc_loss(lp, tgt, il, tl)  (blank 0, mean)."""

    @staticmethod
    def forward(ctx, log_probs, targets, input_lengths, target_lengths, blank):
        _lib.require_cuda(log_probs)
        if log_probs.dim() != 3:
            raise RuntimeError("log_probs must be [T,B,C]")
        if log_probs.dtype != torch.float32:
            raise RuntimeError("log_probs must be float32 (the reference computes CTC in fp32)")
        T, B, C = log_probs.shape
        lp = log_probs.contiguous()  # reference HWR hands over a permuted view; ours is contiguous
        dev = lp.device
        if targets.dim() != 2 or targets.size(0) != B:
            raise RuntimeError("targets must be [B,S] (the trainer passes label.permute(1,0))")
        if not targets.is_cuda and targets.numel() and (int(targets.max()) >= C or int(targets.min()) < 0):
            raise RuntimeError(f"targets must be class indices in [0, {C})")   # CUDA targets: clamped in-kernel (no sync)
        tg = targets.to(device=dev, dtype=torch.int32, non_blocking=True)  # keeps strides
        S = tg.size(1)
        # the reference hands over CPU IntTensors: validate there like ATen does, without a sync
        for name, x, hi in (("input_lengths", input_lengths, T), ("target_lengths", target_lengths, S)):
            if torch.is_tensor(x) and not x.is_cuda and x.numel() and (int(x.max()) > hi or int(x.min()) < 0):
                raise RuntimeError(f"{name} must be in [0, {hi}]")
        il = _lengths_to_device(input_lengths, B, dev, "input_lengths")
        tl = _lengths_to_device(target_lengths, B, dev, "target_lengths")
        L = 2 * S + 1
        need_grad = ctx.needs_input_grad[0]
        nll = torch.empty(B, device=dev, dtype=torch.float32)
        log_alpha = torch.empty((B, T, L), device=dev, dtype=torch.float32)
        log_beta = torch.empty((B, T, L), device=dev, dtype=torch.float32) if need_grad else None
        loss = torch.empty((), device=dev, dtype=torch.float32)
        unit = torch.empty(B, device=dev, dtype=torch.float32)
        st = _lib.stream()
        _lib.call("hwg_ctc_forward", lp.data_ptr(), T, B, C, tg.data_ptr() if S else None,
                  tg.stride(0), tg.stride(1) if S else 1, S, il.data_ptr(), tl.data_ptr(), blank,
                  nll.data_ptr(), log_alpha.data_ptr(), _lib.ptr(log_beta), st)
        _lib.call("hwg_ctc_reduce_mean", nll.data_ptr(), tl.data_ptr(), B, loss.data_ptr(),
                  unit.data_ptr(), st)
        if need_grad:
            ctx.save_for_backward(lp, tg, il, tl, nll, log_alpha, log_beta, unit)
            ctx.blank = blank
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        lp, tg, il, tl, nll, log_alpha, log_beta, unit = ctx.saved_tensors
        T, B, C = lp.shape
        S = tg.size(1)
        go = grad_out.to(torch.float32).contiguous()
        grad = torch.empty_like(lp)
        _lib.call("hwg_ctc_backward", go.data_ptr(), unit.data_ptr(), lp.data_ptr(), T, B, C,
                  tg.data_ptr() if S else None, tg.stride(0), tg.stride(1) if S else 1, S,
                  il.data_ptr(), tl.data_ptr(), ctx.blank, nll.data_ptr(), log_alpha.data_ptr(),
                  log_beta.data_ptr(), 1, grad.data_ptr(), _lib.stream())
        return grad, None, None, None, None


def CTCLoss(input, target, input_len, target_len):
    """Same signature and result as the reference's CTCLoss (model/loss.py:28-30):
    input [T,B,C] log-probs, target [B,S] int (may be a strided view), input_len /
    target_len IntTensor[B] (CPU or CUDA).  Returns a 0-dim tensor; inf -> 0."""
    return _CTCLossFn.apply(input, target, input_len, target_len, 0)


def ctc_greedy_decode(log_probs, input_lengths=None, blank=0):
    """Best-path decode of [T,B,C] scores on the device.

    Returns (raw [T,B] int32, decoded [B,T] int32, decoded_len [B] int32), all CUDA
    tensors; row b's first decoded_len[b] entries are what naive_decode returns for line b."""
    _lib.require_cuda(log_probs)
    lp = log_probs.detach()
    if lp.dtype != torch.float32:
        lp = lp.float()
    lp = lp.contiguous()
    T, B, C = lp.shape
    dev = lp.device
    il = None if input_lengths is None else _lengths_to_device(input_lengths, B, dev, "input_lengths")
    raw = torch.empty((T, B), device=dev, dtype=torch.int32)
    dec = torch.zeros((B, T), device=dev, dtype=torch.int32)
    dl = torch.empty(B, device=dev, dtype=torch.int32)
    _lib.call("hwg_ctc_greedy_decode", lp.data_ptr(), T, B, C, _lib.ptr(il), blank, raw.data_ptr(),
              dec.data_ptr(), dl.data_ptr(), _lib.stream())
    return raw, dec, dl


def naive_decode(output):
    """Reference signature (utils/string_utils.py:51): output is ONE line's [T,C] scores
    (CUDA tensor).  Returns (predData, rawPredData) as python int lists."""
    raw, dec, dl = ctc_greedy_decode(output.unsqueeze(1))
    n = int(dl[0])
    return dec[0, :n].tolist(), raw[:, 0].tolist()
