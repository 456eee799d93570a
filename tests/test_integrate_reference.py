"""Build-container only (needs /root/reference; skipped on the GPU box): the drop-ins slot into the UNMODIFIED
reference's own construction code — `HWWithStyle(config['model'])` from the shipped IAM GAN config — and give the same
parameter names, shapes and (same seed) values as the reference's classes."""
import json
import os

import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="/root/reference is not present on this machine")
CFG = "configs/cf_IAMslant_noMask_charSpecSingleAppend_GANMedMT_autoAEMoPrcp2tightNewCTCUseGen_balB_hCF0.75_sMG.json"


def _build():
    from model import HWWithStyle
    cfg = json.load(open(os.path.join(ref_shim.REF, CFG)))
    cfg["model"]["pretrained_hwr"] = None
    torch.manual_seed(7)
    return HWWithStyle(cfg["model"])


def test_hwwithstyle_builds_with_the_drop_ins_and_matches_the_reference_state_dict():
    import importlib
    import sys
    saved_path, saved_ds = list(sys.path), sys.modules.get("datasets")
    ref_shim.install()
    cwd = os.getcwd()
    os.chdir(ref_shim.REF)                         # the configs use ./data/IAM_char_set.json
    try:
        hws = importlib.import_module("model.hw_with_style")
        mloss = importlib.import_module("model.loss")
        orig = (hws.SpacedGenerator, hws.CNNOnlyHWR, mloss.CTCLoss, hws.DiscriminatorAP)
        ref_model = _build()
        from handwriting_line_generation_b200 import CNNOnlyHWR, CTCLoss, DiscriminatorAP, SpacedGenerator, integrate
        swapped = integrate.install()
        try:
            assert ("model.hw_with_style", "SpacedGenerator") in swapped and ("model.loss", "CTCLoss") in swapped
            ours = _build()
        finally:
            hws.SpacedGenerator, hws.CNNOnlyHWR, mloss.CTCLoss, hws.DiscriminatorAP = orig
    finally:
        os.chdir(cwd)
        sys.path[:] = saved_path                   # leave the interpreter as we found it for the other tests
        if saved_ds is not None:
            sys.modules["datasets"] = saved_ds
        else:
            sys.modules.pop("datasets", None)
    assert isinstance(ours.generator, SpacedGenerator) and isinstance(ours.hwr, CNNOnlyHWR)
    assert isinstance(ours.discriminator, DiscriminatorAP) and ours.discriminator.use_low and ours.discriminator.use_med
    assert mloss.CTCLoss is orig[2] and CTCLoss is not orig[2]
    a, b = ref_model.state_dict(), ours.state_dict()
    # our modules add no persistent state; the reference's blur buffers etc. keep their names
    assert set(a) == set(b), (sorted(set(a) ^ set(b))[:10])
    for k in a:
        assert a[k].shape == b[k].shape, k
        if k.startswith(("generator.", "hwr.", "discriminator.")) and a[k].is_floating_point():
            assert torch.equal(a[k], b[k]), k           # same seed, same construction order -> same init
