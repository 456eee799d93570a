"""GPU: the CharStyleEncoder drop-in on the real kernels against the goldens of the UNMODIFIED reference
(tests/golden/style.npz: style vectors, 1e-4-pinned oracle) and, for the gradients of every trainable parameter that takes
part, against torch autograd over the fp32 oracle (same assertions as the CPU run through the C-ABI interpreter)."""
import numpy as np
import pytest
import torch

from oracle.make_golden import STYLE_CASES
from tests.test_char_style_emulated_cpu import build, check_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(STYLE_CASES))
def test_char_style_encoder_cuda(name, golden_dir):
    from handwriting_line_generation_b200 import _lib
    gold = np.load(f"{golden_dir}/style.npz")
    m = build(STYLE_CASES[name][2]).cuda()
    n0 = _lib.launch_count()
    e_fwd, rep = check_case(m, name, gold, "cuda")
    assert _lib.launch_count() - n0 >= 100, "the style extractor did not run on the CUDA extension"
    print(f"CharStyleEncoder on the B200 [{name}]: style rel-L2 {e_fwd:.2e}, gradients {rep}")
    assert rep["cos_all"] >= 0.97, rep
    assert rep["worst"][1] >= 0.8, rep
