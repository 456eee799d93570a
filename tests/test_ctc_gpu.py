"""GPU parity: libhwg_b200's CTC kernels (through the C-ABI, via the autograd Function the
trainer would call) against the CPU oracle and the reference-made golden vectors."""
import numpy as np
import pytest
import torch

from oracle import ctc as octc
from oracle import synth
from oracle.make_golden import CTC_CASES, grad_digest
from tests.test_ctc_oracle import REL, _case, grad_tolerance

pytestmark = pytest.mark.gpu


def run_cuda(lp, tg, il, tl, strided_targets=True, grad_out=1.0):
    from handwriting_line_generation_b200 import CTCLoss, _lib
    n0 = _lib.launch_count()
    x = torch.from_numpy(lp).cuda().requires_grad_()
    if strided_targets:  # the trainer's pattern: label [S,B] -> permute(1,0), a strided view
        label = torch.from_numpy(np.ascontiguousarray(tg.T)).cuda()
        target = label.permute(1, 0)
    else:
        target = torch.from_numpy(tg).cuda()
    loss = CTCLoss(x, target, torch.from_numpy(il), torch.from_numpy(tl))
    (loss * grad_out).backward()
    torch.cuda.synchronize()
    assert _lib.launch_count() - n0 >= 3, "the CUDA extension did not launch"
    return loss.item(), x.grad.cpu().numpy()


@pytest.mark.parametrize("name", sorted(CTC_CASES))
def test_cuda_matches_golden_and_oracle(name, golden_dir):
    gold = np.load(f"{golden_dir}/ctc.npz")
    lp, tg, il, tl = _case(name)
    loss, grad = run_cuda(lp, tg, il, tl)
    assert abs(loss - gold[f"{name}/loss"]) <= REL * abs(gold[f"{name}/loss"])
    tol = grad_tolerance(gold, name)
    _, samp = grad_digest(grad)
    assert np.abs(samp - gold[f"{name}/grad_sample"]).max() <= tol
    oloss, ograd, _ = octc.ctc_loss_and_grad(lp, tg, il, tl)
    assert abs(loss - oloss) <= REL * abs(oloss)
    assert np.abs(grad - ograd).max() <= tol
    assert np.isfinite(grad).all()


@pytest.mark.parametrize("name", sorted(CTC_CASES))
def test_decode_bit_exact(name, golden_dir):
    from handwriting_line_generation_b200 import ctc_greedy_decode
    gold = np.load(f"{golden_dir}/ctc.npz")
    lp, tg, il, tl = _case(name)
    raw, dec, dl = ctc_greedy_decode(torch.from_numpy(lp).cuda())
    oraw, odec = octc.greedy_decode(lp)
    assert np.array_equal(raw.cpu().numpy(), oraw)
    dl = dl.cpu().numpy()
    assert dl.tolist() == gold[f"{name}/decoded_len"].tolist()
    got = [x for b in range(lp.shape[1]) for x in dec[b, :dl[b]].cpu().tolist()]
    assert got == gold[f"{name}/decoded"].tolist()


def test_decode_ties_and_reference_signature():
    from handwriting_line_generation_b200 import naive_decode
    lp = np.zeros((6, 4), np.float32)
    lp[:, 2] = 1.0
    lp[:, 3] = 1.0
    lp[3, 1] = 5.0
    pred, raw = naive_decode(torch.from_numpy(lp).cuda())
    assert raw == [2, 2, 2, 1, 2, 2] and pred == [2, 1, 2]


def test_random_shapes_against_oracle():
    r = np.random.RandomState(7)
    for _ in range(12):
        T, B, C, S = int(r.randint(3, 130)), int(r.randint(1, 9)), int(r.randint(2, 100)), int(r.randint(0, 40))
        T = max(T, 2 * S + 1)
        lp, tg, il, tl = synth.ctc_case(T, B, C, S, int(r.randint(1 << 30)))
        il = r.randint(max(1, 2 * S + 1), T + 1, size=B).astype(np.int32)  # ragged input lengths
        if S == 0:
            tg = np.zeros((B, 1), np.int32)
        loss, grad = run_cuda(lp, tg, il, tl, strided_targets=bool(r.randint(2)), grad_out=0.5)
        oloss, ograd, _ = octc.ctc_loss_and_grad(lp, tg, il, tl, grad_out=0.5)
        assert abs(loss - oloss) <= REL * max(abs(oloss), 1e-6)
        assert np.abs(grad - ograd).max() <= 5 * REL * np.abs(ograd).max() + 1e-7
        for b in range(B):  # frames past the input length get exactly zero gradient
            assert not grad[il[b]:, b].any()


def test_infeasible_batch_is_zero_loss():
    from handwriting_line_generation_b200 import CTCLoss
    lp, tg, il, tl = synth.ctc_case(4, 2, 6, 8, 3, ragged=False)
    loss = CTCLoss(torch.from_numpy(lp).cuda(), torch.from_numpy(tg).cuda(), torch.from_numpy(il), torch.from_numpy(tl))
    assert loss.item() == 0.0


def test_full_size_property_gradient_rows_sum_to_zero():
    # size-independent property at BASELINE config-5 size: sum_c grad[t,b,c] = 0 because both
    # softmax probabilities and state occupancies sum to one per frame.
    lp, tg, il, tl = synth.ctc_case(506, 64, 78, 120, 99, ragged=True)
    loss, grad = run_cuda(lp, tg, il, tl)
    scale = np.abs(grad).max()
    assert np.abs(grad.sum(axis=2)).max() <= 2e-3 * scale * 78 ** 0.5
    assert loss > 0


def test_matches_torch_cuda_ctc():
    # the kernel the reference would run on this GPU (ATen LossCTC.cu)
    lp, tg, il, tl = _case("cfg1_ragged")
    loss, grad = run_cuda(lp, tg, il, tl)
    x = torch.from_numpy(lp).cuda().requires_grad_()
    ref = torch.nn.functional.ctc_loss(x, torch.from_numpy(tg).cuda(), torch.from_numpy(il), torch.from_numpy(tl))
    ref.backward()
    assert abs(loss - ref.item()) <= REL * abs(ref.item())
    g = x.grad.cpu().numpy()
    assert np.abs(grad - g).max() <= 2e-3 * np.abs(g).max()
