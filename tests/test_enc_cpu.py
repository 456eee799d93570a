"""CPU: oracle/enc.py (Encoder2, the perceptual encoder of the 'auto' lessons, and the perceptual loss built on it) against
the features and the loss gradient of the unmodified reference (tests/golden/enc.npz).  Groundwork for SURVEY §8 f1's second
half: no CUDA counterpart yet, so there is no product module to compare here."""
import numpy as np
import pytest
import torch

from oracle import enc as oenc
from oracle import synth
from oracle.make_golden import ENC_CASES, digest

FP32_REL = 1e-4


def encoder2_state_dict(seed):
    """Random-init Encoder2(32) weights, reproducible from the seed WITHOUT the reference: the same layers in the same
    construction order (model/autoencoder.py:344-394) consume the torch RNG identically."""
    nn = torch.nn
    torch.manual_seed(seed)
    gn = lambda c: nn.GroupNorm(8, c)                                             # noqa: E731
    m = nn.ModuleDict()
    m["down_conv1"] = nn.Sequential(nn.Conv2d(1, 32, 5, padding=2), gn(32), nn.ReLU(True), nn.AvgPool2d(2), nn.Conv2d(32, 32, 1))
    m["conv1"] = nn.Sequential(nn.ReLU(True), nn.Conv2d(32, 32, 3, padding=1), gn(32), nn.Dropout2d(0.1, True), nn.ReLU(True),
                               nn.Conv2d(32, 32, 3, padding=1))
    m["down_conv2"] = nn.Sequential(gn(32), nn.ReLU(True), nn.AvgPool2d(2), nn.Conv2d(32, 64, 1))
    m["conv2"] = nn.Sequential(gn(64), nn.Dropout2d(0.1, True), nn.ReLU(True), nn.Conv2d(64, 64, 3, padding=1), gn(64),
                               nn.Dropout2d(0.1, True), nn.ReLU(True), nn.Conv2d(64, 64, 3, padding=1))
    m["down_conv3"] = nn.Sequential(gn(64), nn.ReLU(True), nn.AvgPool2d(2), nn.Conv2d(64, 128, 3), gn(128),
                                    nn.Dropout2d(0.1, True), nn.ReLU(True), nn.Conv2d(128, 32, (6, 3)))
    return m.state_dict()


@pytest.mark.parametrize("name", sorted(ENC_CASES))
def test_encoder2_oracle_matches_reference_golden(name, golden_dir):
    from oracle.make_golden import keys_fixture, weights_digest
    gold = np.load(f"{golden_dir}/enc.npz")
    B, W, wseed, iseed, training = ENC_CASES[name]
    sd = encoder2_state_dict(wseed)
    assert keys_fixture(sd).tolist() == gold["state_dict_keys"].tolist()
    assert abs(weights_digest(sd) - gold[f"{name}/weights_digest"]) <= 1e-6 * abs(gold[f"{name}/weights_digest"])
    r = np.random.RandomState(iseed + 7)
    masks = [torch.from_numpy((r.rand(2 * B, c) >= 3 * p).astype(np.float32)) for _, c, p in oenc.DROPOUT_SITES]
    image = torch.from_numpy(synth.hwr_case(B, W, iseed))
    recon = torch.from_numpy(synth.hwr_case(B, W, iseed + 1)).requires_grad_()
    feats = oenc.encoder2_forward(sd, torch.cat((image, recon), 0), masks, training)
    for i, f in enumerate(feats):
        assert list(f.shape) == gold[f"{name}/feat{i}/shape"].tolist()
        _, samp = digest(f.detach().numpy())
        assert np.abs(samp[:2048] - gold[f"{name}/feat{i}/sample"]).max() <= FP32_REL * gold[f"{name}/feat{i}/digest"][3]
    loss = oenc.perceptual_loss(sd, image, recon, masks, training)
    assert abs(loss.item() - float(gold[f"{name}/loss"])) <= FP32_REL * abs(float(gold[f"{name}/loss"]))
    loss.backward()
    _, samp = digest(recon.grad.numpy())
    assert np.abs(samp[:2048] - gold[f"{name}/grad_sample"]).max() <= FP32_REL * gold[f"{name}/grad_digest"][3]
