"""CPU: oracle/spacer.py — the spacer `CountCNN` and the vectorised restatement of `HWWithStyle.insert_spaces` — against the
unmodified reference (tests/golden/spacer.npz): counts to fp32 tolerance, the spaced text BIT-EXACT, including a case whose
per-character normal draws change the rounding (same numpy RNG stream consumed in one vectorised call)."""
import numpy as np
import pytest
import torch

from oracle import spacer as ospacer
from oracle.make_golden import SPACER_CASES, keys_fixture, spacer_inputs, weights_digest


def count_cnn_state_dict(seed):
    """Random-init CountCNN(80, 128, 128, 2) weights without the reference: same layers, names and construction order
    (model/count_cnn.py:11-32)."""
    nn = torch.nn
    torch.manual_seed(seed)
    m = nn.Module()
    m.cnn = nn.Sequential(nn.Conv1d(208, 128, 3, 1, 1), nn.GroupNorm(8, 128), nn.Dropout2d(0.1), nn.ReLU(True),
                          nn.Conv1d(128, 64, 3, 1, 1), nn.GroupNorm(8, 64), nn.Dropout2d(0.1), nn.ReLU(True),
                          nn.Conv1d(64, 32, 3, 1, 1), nn.GroupNorm(8, 32), nn.ReLU(True), nn.Conv1d(32, 2, 1, 1, 0))
    m.mean = nn.Parameter(torch.FloatTensor([2.0, 0.0]))
    m.std = nn.Parameter(torch.FloatTensor([1.5, 0.5]))
    return m.state_dict()


@pytest.mark.parametrize("name", sorted(SPACER_CASES))
def test_spacer_and_insert_spaces_match_the_reference(name, golden_dir):
    gold = np.load(f"{golden_dir}/spacer.npz")
    L, B, wseed, iseed = SPACER_CASES[name]
    sd = count_cnn_state_dict(wseed)
    assert sorted(keys_fixture(sd).tolist()) == sorted(gold["state_dict_keys"].tolist())
    d = gold[f"{name}/weights_digest"]
    assert abs(weights_digest(sd) - d) <= 1e-6 * abs(d)
    label, lengths, style = spacer_inputs(L, B, iseed)
    onehot = torch.zeros(L, B, 80).scatter_(2, label[..., None], 1.0)
    with torch.no_grad():
        counts = ospacer.count_cnn_forward(sd, onehot, style)
    ref = gold[f"{name}/counts"]
    assert np.abs(counts.numpy() - ref).max() <= 1e-4 * np.abs(ref).max()
    for std, tag in ((1e-8, "cfg"), (0.4, "noisy")):
        rng = np.random.RandomState(iseed)                    # np.random.seed(iseed) of the golden run
        spaced, padded = ospacer.insert_spaces(label.numpy(), lengths, torch.from_numpy(ref), 80, std, std / 10, rng)
        assert np.array_equal(spaced.argmax(2).numpy(), gold[f"{name}/{tag}/spaced"]), tag
        assert float(spaced.sum(2).min()) == 1.0
        assert np.allclose(padded, gold[f"{name}/{tag}/padded"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", sorted(SPACER_CASES))
def test_oracle_spacer_train_mode_gradients_match_the_reference(name, golden_dir):
    """Train mode (the 'count' lesson trains the spacer): injected Dropout2d keep-masks, a linear loss; counts and the
    gradients of all 16 parameters, of the style vector and of the text input against the unmodified reference."""
    from oracle.make_golden import spacer_train_extras
    gold = np.load(f"{golden_dir}/spacer.npz")
    L, B, wseed, iseed = SPACER_CASES[name]
    sd = {k: v.clone().requires_grad_(True) for k, v in count_cnn_state_dict(wseed).items()}
    label, lengths, style = spacer_inputs(L, B, iseed)
    style = style.clone().requires_grad_()
    onehot = torch.zeros(L, B, 80).scatter_(2, label[..., None], 1.0).requires_grad_()
    masks, R = spacer_train_extras(L, B, iseed)
    counts = ospacer.count_cnn_forward(sd, onehot, style, masks)
    (counts * R).sum().backward()
    ref = gold[f"{name}/train/counts"]
    assert np.abs(counts.detach().numpy() - ref).max() <= 1e-4 * np.abs(ref).max()
    for key, t in [("style", style), ("input", onehot)] + list(sd.items()):
        g = gold[f"{name}/train/grad/{key}"]
        assert np.abs(t.grad.numpy() - g).max() <= 1e-4 * np.abs(g).max() + 1e-7, key


@pytest.mark.parametrize("name", sorted(SPACER_CASES))
def test_spacing_host_wrapper_through_the_interpreter(name, golden_dir, hwg_lib, monkeypatch):
    """`spacing.insert_spaces` (the product's host side: one vectorised draw from the reference's numpy stream, draw offsets
    per line, the single read-back that sizes the result) with the two launches interpreted on CPU: the spaced text of the
    unmodified reference, bit for bit."""
    import types
    from handwriting_line_generation_b200.spacing import insert_spaces
    from . import abi_emu
    gold = np.load(f"{golden_dir}/spacer.npz")
    L, B, wseed, iseed = SPACER_CASES[name]
    label, lengths, _ = spacer_inputs(L, B, iseed)
    counts = torch.from_numpy(gold[f"{name}/counts"])
    with abi_emu.installed(monkeypatch) as calls:
        for std, tag in ((1e-8, "cfg"), (0.4, "noisy")):
            host = types.SimpleNamespace(count_std=std, dup_std=std / 10, count_duplicates=True, num_class=80)
            spaced, padded = insert_spaces(host, label, lengths, counts, rng=np.random.RandomState(iseed))
            assert np.array_equal(spaced.argmax(2).numpy(), gold[f"{name}/{tag}/spaced"]), tag
            assert np.allclose(padded, gold[f"{name}/{tag}/padded"], rtol=0, atol=1e-12)
    assert calls == ["hwg_insert_spaces_plan", "hwg_insert_spaces_fill"] * 2
