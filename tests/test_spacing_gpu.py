"""GPU: `spacing.insert_spaces` (hwg_insert_spaces_plan / _fill) against the goldens of the UNMODIFIED reference
`HWWithStyle.insert_spaces` (tests/golden/spacer.npz) — same numpy RNG state, the reference's own counts: the spaced text
must be BIT-IDENTICAL (integer work), at the config's count_std / dup_std (1e-8: the draws only break exact .5 ties) and at a
noisy setting where the draws change the rounding; plus ragged / empty lines and a one-channel (no duplicates) spacer against
the oracle; and the CountCNN drop-in on the real kernels (counts, every gradient)."""
import types

import numpy as np
import pytest
import torch

from oracle import spacer as ospacer
from oracle.make_golden import SPACER_CASES, spacer_inputs

pytestmark = pytest.mark.gpu


def _host(std, dup=True):
    return types.SimpleNamespace(count_std=std, dup_std=std / 10, count_duplicates=dup, num_class=80)


@pytest.mark.parametrize("name", sorted(SPACER_CASES))
@pytest.mark.parametrize("std,tag", [(1e-8, "cfg"), (0.4, "noisy")])
def test_insert_spaces_bit_exact_against_the_reference(name, std, tag, golden_dir):
    from handwriting_line_generation_b200.spacing import insert_spaces
    gold = np.load(f"{golden_dir}/spacer.npz")
    L, B, wseed, iseed = SPACER_CASES[name]
    label, lengths, _ = spacer_inputs(L, B, iseed)
    counts = torch.from_numpy(gold[f"{name}/counts"]).cuda()
    spaced, padded = insert_spaces(_host(std), label, lengths, counts, rng=np.random.RandomState(iseed))
    assert spaced.is_cuda and spaced.dtype == torch.float32
    ref = gold[f"{name}/{tag}/spaced"]
    assert tuple(spaced.shape) == (ref.shape[0], B, 80)
    assert float(spaced.sum(2).min()) == 1.0 and float(spaced.max()) == 1.0 and float(spaced.min()) == 0.0
    assert np.array_equal(spaced.argmax(2).cpu().numpy(), ref)
    assert np.allclose(padded, gold[f"{name}/{tag}/padded"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("dup", [True, False])
def test_insert_spaces_ragged_lines_against_the_oracle(dup):
    """Lengths from 0 to L, int32 and strided int64 labels, counts with negative and large entries, 300 lines x 70 characters
    (more characters than one scan pass holds would be > 256: L = 300 in the second round)."""
    from handwriting_line_generation_b200.spacing import insert_spaces
    for L, B, seed in ((70, 300, 1), (300, 5, 2)):
        r = np.random.RandomState(seed)
        label = torch.from_numpy(r.randint(1, 80, (B, L)).astype(np.int64)).t()          # [L,B] strided view
        lengths = r.randint(0, L + 1, B).tolist()
        lengths[0], lengths[-1] = L, 0
        counts = torch.from_numpy((r.standard_normal((L, B, 2)) * 1.5 + 1.5).astype(np.float32))
        host = _host(0.3, dup)
        spaced, padded = insert_spaces(host, label.cuda(), lengths, counts.cuda(), rng=np.random.RandomState(seed + 10))
        c_or = counts if dup else counts[:, :, :1]
        # the oracle restates the reference for the duplicates spacer; without duplicates every character appears once
        if dup:
            ospaced, opadded = ospacer.insert_spaces(label.numpy(), lengths, counts.numpy(), 80, 0.3, 0.03,
                                                     np.random.RandomState(seed + 10))
            assert tuple(spaced.shape) == tuple(ospaced.shape)
            assert np.array_equal(spaced.argmax(2).cpu().numpy(), ospaced.argmax(2).numpy())
            assert np.allclose(padded, opadded, rtol=0, atol=1e-12)
        else:
            rng = np.random.RandomState(seed + 10)
            lines = []
            for b in range(B):
                line = []
                for i in range(lengths[b]):
                    cnt = round(rng.normal(float(counts[i, b, 0]), 0.3))
                    line += [0] * cnt + [int(label[i, b])]
                lines.append(line)
            T = max(len(ln) for ln in lines) + max(int(np.ceil(float(counts.max()))), 3)
            assert spaced.size(0) == T
            got = spaced.argmax(2).cpu().numpy()
            for b, ln in enumerate(lines):
                assert got[:len(ln), b].tolist() == ln and not got[len(ln):, b].any(), b
        assert float(spaced.sum(2).min()) == 1.0


@pytest.mark.parametrize("name", sorted(SPACER_CASES))
def test_count_cnn_cuda_matches_the_reference(name, golden_dir):
    """The CountCNN drop-in on the real kernels: eval counts, train-mode counts with injected Dropout2d masks and the gradients
    of all 16 parameters, the style vector and the text input against the unmodified reference (same assertions as the CPU
    run through the C-ABI interpreter)."""
    from handwriting_line_generation_b200 import _lib
    from tests.test_count_cnn_emulated_cpu import build, check_case
    gold = np.load(f"{golden_dir}/spacer.npz")
    m = build(SPACER_CASES[name][2]).cuda()
    n0 = _lib.launch_count()
    worst = check_case(m, name, gold, "cuda")
    assert _lib.launch_count() - n0 >= 30, "CountCNN did not run on the CUDA extension"
    print("CountCNN on the B200, gradient rel-L2 per tensor:", {k: round(v, 4) for k, v in worst.items()})
