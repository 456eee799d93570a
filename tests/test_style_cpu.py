"""CPU: oracle/style.py — the vectorised restatement of the style extractor `CharStyleEncoder` (IAM GAN configuration) and
of the DTW alignment `correct_pred` — against outputs of the unmodified reference (tests/golden/style.npz).  Groundwork for
SURVEY §8 f3: no CUDA counterpart yet, so there is no product module to compare here."""
import numpy as np
import pytest
import torch

from oracle import style as ostyle
from oracle.make_golden import DTW_CASES, STYLE_CASES, dtw_inputs, keys_fixture, style_inputs, weights_digest

FP32_REL = 1e-4


def char_style_state_dict(seed, n_class=80, dim=64, char_dim=128, style_dim=128):
    """Random-init CharStyleEncoder weights, reproducible from the seed WITHOUT the reference: the same parameterised layers
    under the same names in the same construction order (model/char_style.py:154-188, :84-112) consume the torch RNG
    identically."""
    nn = torch.nn
    torch.manual_seed(seed)

    def block(ci, co, k, norm=True):
        m = nn.Module()
        if norm:
            m.norm = nn.GroupNorm(8, co)
        m.conv = nn.Conv2d(ci, co, k)
        return m

    root = nn.Module()
    d = dim
    root.down = nn.Sequential(block(1, d, 5), block(d, 2 * d, 4), block(2 * d, 2 * d, 3), block(2 * d, 4 * d, 4),
                              block(4 * d, 4 * d, 3), block(4 * d, 4 * d, 4), block(4 * d, 4 * d, 4, norm=False))
    d *= 4
    root.prep = nn.Sequential(nn.Conv1d(d + n_class, d, 5, 1, 2), nn.ReLU(True), nn.MaxPool1d(2, 2), nn.Conv1d(d, d, 3, 1, 1),
                              nn.GroupNorm(8, d), nn.ReLU(True), nn.Conv1d(d, d, 3, 1, 1), nn.ReLU(True))
    root.final_g_spacing_style = nn.Sequential(nn.Linear(d + style_dim, d), nn.ReLU(True), nn.Linear(d, style_dim))
    root.char_extractor = nn.ModuleList()
    for _ in range(n_class):
        e = nn.Module()
        e.conv1 = nn.Sequential(nn.ReLU(), nn.Conv1d(d, char_dim, 3, padding=1), nn.GroupNorm(8, char_dim), nn.ReLU(),
                                nn.Conv1d(char_dim, d, 3, padding=1))
        e.conv2 = nn.Sequential(nn.ReLU(), nn.Conv1d(d, 2 * char_dim, 1), nn.GroupNorm(8, 2 * char_dim), nn.ReLU())
        e.fc = nn.Sequential(nn.Linear(2 * char_dim, 2 * char_dim), nn.ReLU(True), nn.Linear(2 * char_dim, style_dim))
        root.char_extractor.append(e)
    return root.state_dict()


@pytest.fixture(scope="module")
def weights():
    return char_style_state_dict(600)


@pytest.mark.parametrize("name", sorted(STYLE_CASES))
def test_style_extractor_oracle_matches_reference_golden(name, golden_dir, weights):
    gold = np.load(f"{golden_dir}/style.npz")
    B, W, wseed, iseed = STYLE_CASES[name]
    assert wseed == 600
    assert sorted(keys_fixture(weights).tolist()) == sorted(gold["state_dict_keys"].tolist())
    d = gold[f"{name}/weights_digest"]
    assert abs(weights_digest(weights) - d) <= 1e-6 * abs(d)
    image, recog = style_inputs(B, W, iseed)
    with torch.no_grad():
        style = ostyle.char_style_forward(weights, image, recog)
    ref = gold[f"{name}/style"]
    assert style.shape == ref.shape
    assert np.abs(style.numpy() - ref).max() <= FP32_REL * np.abs(ref).max()


def test_window_gather_visits_what_the_reference_loops_visit():
    """(class, sample, position) order, zero padding at the line ends, score = exp(log-prob)."""
    torch.manual_seed(0)
    x = torch.randn(2, 4, 12)
    recog = torch.log_softmax(torch.randn(2, 6, 12) * 3, 1)
    cls, b, pos, score, patches = ostyle.gather_windows(x, recog, 2)
    pred = recog.argmax(1)
    want = [(c, bb, p) for c in range(1, 6) for bb in range(2) for p in range(12) if pred[bb, p] == c]
    assert list(zip(cls.tolist(), b.tolist(), pos.tolist())) == want
    for i, (c, bb, p) in enumerate(want):
        for k in range(5):
            q = p - 2 + k
            ref = x[bb, :, q] if 0 <= q < 12 else torch.zeros(4)
            assert torch.equal(patches[i, :, k], ref)
        assert abs(score[i].item() - float(torch.exp(recog[bb, c, p]))) <= 1e-7


@pytest.mark.parametrize("name", sorted(DTW_CASES))
def test_dtw_alignment_is_bit_exact(name, golden_dir):
    gold = np.load(f"{golden_dir}/style.npz")
    T, B, L, seed = DTW_CASES[name]
    pred, label = dtw_inputs(T, B, L, seed)
    got = ostyle.correct_pred(pred.numpy(), label.numpy())
    assert got.shape == gold[f"dtw/{name}"].shape and np.array_equal(got, gold[f"dtw/{name}"])
