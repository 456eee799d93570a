"""CPU: the CTC oracle (oracle/ctc_oracle.c) against the golden vectors produced by the
unmodified reference (tests/golden/ctc.npz, oracle/make_golden.py) and against torch's own
CPU F.ctc_loss, which is the arithmetic the reference calls (model/loss.py:29)."""
import numpy as np
import pytest
import torch

from oracle import ctc as octc
from oracle import synth
from oracle.make_golden import CTC_CASES, grad_digest

REL = 1e-4  # north_star: fp32 rel 1e-4


def grad_tolerance(gold, name):
    """Per-tensor tolerance of the CTC gradient: max|a-b| <= REL*max|ref| + 2*noise, where
    noise = max|ref_fp32 - ref_fp64| is the reference's own fp32 rounding error on this case
    (alpha+beta+nll-lp cancels at magnitude ~T*|lp|: 7e-4..4e-3 of max|grad| at T=250..506,
    far above 1e-4; loss and nll are held to 1e-4 strictly)."""
    return REL * gold[f"{name}/grad_digest"][3] + 2.0 * float(gold[f"{name}/grad_fp32_noise"])


def _case(name):
    T, B, C, S, seed, ragged = CTC_CASES[name]
    lp, tg, il, tl = synth.ctc_case(T, B, C, S, seed, ragged)
    if name == "empty_target":
        tl[1] = 0
        tg[1, :] = 0
    return lp, tg, il, tl


@pytest.mark.parametrize("name", sorted(CTC_CASES))
def test_oracle_matches_reference_golden(name, golden_dir):
    gold = np.load(f"{golden_dir}/ctc.npz")
    lp, tg, il, tl = _case(name)
    loss, grad, nll = octc.ctc_loss_and_grad(lp, tg, il, tl)
    assert abs(loss - gold[f"{name}/loss"]) <= REL * abs(gold[f"{name}/loss"])
    np.testing.assert_allclose(nll, gold[f"{name}/nll"], rtol=REL)
    dig, samp = grad_digest(grad)
    tol = grad_tolerance(gold, name)
    assert np.abs(samp - gold[f"{name}/grad_sample"]).max() <= tol
    if f"{name}/grad" in gold:
        assert np.abs(grad - gold[f"{name}/grad"]).max() <= tol
    raw, dec = octc.greedy_decode(lp)
    assert [len(d) for d in dec] == gold[f"{name}/decoded_len"].tolist()
    assert [x for d in dec for x in d] == gold[f"{name}/decoded"].tolist()


def test_oracle_matches_torch_cpu_random_shapes():
    r = np.random.RandomState(5)
    for _ in range(6):
        T, B, C, S = int(r.randint(5, 60)), int(r.randint(1, 6)), int(r.randint(3, 30)), int(r.randint(1, 8))
        if T < 2 * S + 1:
            T = 2 * S + 1
        lp, tg, il, tl = synth.ctc_case(T, B, C, S, int(r.randint(1 << 30)))
        lpt = torch.from_numpy(lp).requires_grad_()
        loss = torch.nn.functional.ctc_loss(lpt, torch.from_numpy(tg), torch.from_numpy(il), torch.from_numpy(tl))
        loss.backward()
        l, g, _ = octc.ctc_loss_and_grad(lp, tg, il, tl)
        assert abs(l - loss.item()) <= REL * abs(loss.item())
        assert np.abs(g - lpt.grad.numpy()).max() <= REL * np.abs(lpt.grad.numpy()).max()


def test_oracle_infeasible_gives_zero_loss():
    # T < S: no alignment exists -> F.ctc_loss is inf -> the reference wrapper returns 0
    lp, tg, il, tl = synth.ctc_case(4, 2, 6, 8, 3, ragged=False)
    nll, _ = octc.ctc_forward(lp, tg, il, tl)
    assert np.isinf(nll).all()
    assert octc.ctc_loss_mean(nll, tl) == 0.0


def test_decode_ties_first_max_wins():
    lp = np.zeros((6, 1, 4), np.float32)  # all equal -> argmax 0 = blank everywhere
    raw, dec = octc.greedy_decode(lp)
    assert raw[:, 0].tolist() == [0] * 6 and dec == [[]]
    lp[:, 0, 2] = 1.0
    lp[:, 0, 3] = 1.0  # tie between 2 and 3 -> 2
    lp[3, 0, 1] = 5.0
    raw, dec = octc.greedy_decode(lp)
    assert raw[:, 0].tolist() == [2, 2, 2, 1, 2, 2] and dec == [[2, 1, 2]]
