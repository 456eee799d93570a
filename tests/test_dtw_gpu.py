"""`handwriting_line_generation_b200.dtw.correct_pred` (hwg_dtw_align) against the alignments of the UNMODIFIED reference
`correct_pred` (tests/golden/style.npz, `dtw/*`) and against the oracle on more shapes — bit-exact (integer work)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from oracle import style as ostyle                                   # noqa: E402
from oracle.make_golden import DTW_CASES, dtw_inputs                 # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(DTW_CASES))
def test_dtw_matches_the_reference_golden(name):
    from handwriting_line_generation_b200.dtw import correct_pred
    gold = np.load(os.path.join(ROOT, "tests", "golden", "style.npz"))
    T, B, L, seed = DTW_CASES[name]
    pred, label = dtw_inputs(T, B, L, seed)
    got = correct_pred(pred.cuda(), label)
    assert got.dtype == torch.int64 and got.device == label.device
    assert np.array_equal(got.numpy(), gold[f"dtw/{name}"])


@pytest.mark.parametrize("T,B,L,seed", [(250, 8, 40, 1), (26, 3, 15, 2), (60, 4, 60, 3), (506, 2, 120, 4)])
def test_dtw_matches_the_oracle(T, B, L, seed):
    from handwriting_line_generation_b200.dtw import correct_pred
    pred, label = dtw_inputs(T, B, L, seed)
    got = correct_pred(pred.cuda(), label.cuda()).cpu().numpy()
    assert np.array_equal(got, ostyle.correct_pred(pred.numpy(), label.numpy()))
