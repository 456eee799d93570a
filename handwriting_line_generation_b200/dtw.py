"""Drop-in for the reference's DTW alignment `correct_pred` (model/hw_with_style.py:18-74): the label,
with blanks in front, behind and between its characters, aligned to the recognizer output by dynamic time warping — what
`HWWithStyle.autoencode` / `extract_style` use to place the ground-truth characters along the line (:279-291).

The reference copies the prediction to the CPU and runs a Python double loop of ~T*(2S+1) torch calls; here it is one
launch (`hwg_dtw_align`: one CTA per line, anti-diagonal wavefront, byte history, in-kernel backtrack) and one small
device-to-host read of the path lengths (the result's height is data dependent, as in the reference).

Bit-exact on the B200 against the reference's goldens (tests/test_dtw_gpu.py).  There is no CPU fallback."""
import torch

from . import _lib


def correct_pred(pred, label):
    """pred [T,B,C] float CUDA tensor (the recognizer's output), label [S,B] integer tensor -> LongTensor [T',B] on
    label's device: the aligned label, zero-padded to the longest path (same signature and result as the reference)."""
    _lib.require_cuda(pred)
    if pred.dim() != 3 or label.dim() != 2 or label.size(1) != pred.size(1):
        raise RuntimeError("correct_pred: pred must be [T,B,C] and label [S,B]")
    p = pred.detach().float().contiguous()
    T, B, C = p.shape
    S = label.size(0)
    L = 2 * S + 1
    dev = p.device
    lab = label.to(device=dev, dtype=torch.int32)
    hist = torch.empty((B, T, L), device=dev, dtype=torch.uint8)
    out = torch.zeros((T + L, B), device=dev, dtype=torch.int32)
    out_len = torch.empty(B, device=dev, dtype=torch.int32)
    scratch = torch.empty((B, T + L), device=dev, dtype=torch.int32)
    _lib.call("hwg_dtw_align", p.data_ptr(), T, B, C, lab.data_ptr(), lab.stride(0), lab.stride(1), S, hist.data_ptr(),
              out.data_ptr(), out_len.data_ptr(), scratch.data_ptr(), _lib.stream())
    maxlen = int(out_len.max())                       # data-dependent height: the one host read
    return out[:maxlen].long().to(label.device)
