"""CPU: the torch restatements (oracle/gen.py, oracle/hwr.py) against the golden outputs of the unmodified
reference, and the drop-in modules' state_dict contract (names, shapes, seeded init) — no GPU needed."""
import numpy as np
import pytest
import torch

from oracle import gen as ogen
from oracle import hwr as ohwr
from oracle import synth
from oracle.make_golden import GEN_CASES, HWR_CASES, digest, keys_fixture, weights_digest

FP32_REL = 1e-4


def _gen_module(seed):
    from handwriting_line_generation_b200 import SpacedGenerator
    return synth.state_dict_from_seed(lambda: SpacedGenerator(80, 128, 256, n_style_trans=6, emb_dropout=False,
                                                              append_style=True, small=False), seed)


def _hwr_module(seed):
    from handwriting_line_generation_b200 import CNNOnlyHWR
    return synth.state_dict_from_seed(lambda: CNNOnlyHWR(80, norm='batch'), seed)


def test_generator_state_dict_contract(golden_dir):
    gold = np.load(f"{golden_dir}/gen.npz")
    m, sd = _gen_module(GEN_CASES["tiny"][3])
    assert keys_fixture(sd).tolist() == gold["state_dict_keys"].tolist()
    # same seed -> same random init as the reference (construction order mirrors pure_gen.py:13-40)
    assert abs(weights_digest(sd) - gold["tiny/weights_digest"]) <= 1e-6 * abs(gold["tiny/weights_digest"])


def test_hwr_state_dict_contract(golden_dir):
    gold = np.load(f"{golden_dir}/hwr.npz")
    m, sd = _hwr_module(HWR_CASES["train_w128"][2])
    assert keys_fixture(sd).tolist() == gold["state_dict_keys"].tolist()
    assert abs(weights_digest(sd) - gold["train_w128/weights_digest"]) <= 1e-6 * abs(gold["train_w128/weights_digest"])


@pytest.mark.parametrize("name", sorted(GEN_CASES))
def test_generator_oracle_matches_reference_golden(name, golden_dir):
    gold = np.load(f"{golden_dir}/gen.npz")
    T, B, dense, wseed, iseed = GEN_CASES[name]
    _, sd = _gen_module(wseed)
    content, style = synth.gen_case(T, B, 80, 128, iseed, dense)
    noise = [torch.from_numpy(z) for z in synth.gen_noise(synth.gen_noise_shapes(T, B), iseed + 7)]
    with torch.no_grad():
        img = ogen.generator_forward(sd, torch.from_numpy(content), torch.from_numpy(style), noise).numpy()
    assert list(img.shape) == gold[f"{name}/shape"].tolist()
    _, samp = digest(img)
    assert np.abs(samp - gold[f"{name}/sample"]).max() <= FP32_REL * gold[f"{name}/digest"][3]
    if f"{name}/image" in gold:
        assert np.abs(img - gold[f"{name}/image"]).max() <= FP32_REL * gold[f"{name}/digest"][3]


@pytest.mark.parametrize("name", sorted(HWR_CASES))
def test_hwr_oracle_matches_reference_golden(name, golden_dir):
    gold = np.load(f"{golden_dir}/hwr.npz")
    B, W, wseed, iseed, training = HWR_CASES[name]
    _, sd = _hwr_module(wseed)
    upd = {}
    with torch.no_grad():
        lp = ohwr.hwr_forward(sd, torch.from_numpy(synth.hwr_case(B, W, iseed)), training, upd).numpy()
    assert list(lp.shape) == gold[f"{name}/shape"].tolist()
    assert np.abs(lp - gold[f"{name}/log_probs"]).max() <= FP32_REL * gold[f"{name}/digest"][3]
    assert np.array_equal(lp.argmax(2), gold[f"{name}/argmax"])
    if training:
        for k in ("cnn.batchnorm2.running_mean", "cnn.batchnorm6.running_var", "cnn1d.10.running_mean",
                  "cnn1d.1.running_var"):
            np.testing.assert_allclose(upd[k].numpy(), gold[f"{name}/{k}"], rtol=1e-4, atol=1e-6)


def test_modules_refuse_cpu_tensors():
    m, _ = _gen_module(1)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(4, 1, 80), torch.zeros(1, 128))
    h, _ = _hwr_module(1)
    with pytest.raises(RuntimeError, match="CUDA"):
        h(torch.zeros(1, 1, 64, 64))
