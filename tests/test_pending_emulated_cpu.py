"""CPU: four of the GPU parity tests (trainer-level 'gen' lesson, Encoder2, gradient balancing, DTW — green on the B200
since round 2) executed through the CPU interpreter of the C-ABI with `.cuda()` made a no-op: their own code and thresholds
stay checked in the build container, where the kernels cannot run (this is how a 3e-2 image bound and a 0.6 cosine bound
that the bf16 pipeline cannot meet on the trainer's 2-line case were found and replaced before the first GPU run)."""
import importlib.util
import os

import pytest
import torch

from . import abi_emu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tests", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.fixture
def emulated_gpu(monkeypatch, hwg_lib):
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.nn.Module, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    with abi_emu.installed(monkeypatch) as calls:
        yield calls


def test_parked_trainer_gen_lesson_test(emulated_gpu):
    _load("test_trainer_gen_gpu").test_cuda_gen_lesson_against_the_reference_trainer()
    assert "hwg_ctc_backward" in emulated_gpu


@pytest.mark.parametrize("name", ["eval_w128", "train_w200"])
def test_parked_encoder2_test(emulated_gpu, name):
    _load("test_enc_gpu").test_encoder2_cuda_matches_reference_golden_and_oracle(name)
    assert "hwg_l1_halves" in emulated_gpu


def test_parked_balance_test(emulated_gpu):
    _load("test_balance_gpu").test_flat_balance_matches_the_oracle()
    assert emulated_gpu.count("hwg_balance") == 1


@pytest.mark.parametrize("name", ["t58_l9", "t124_l30"])
def test_parked_dtw_test(emulated_gpu, name):
    """The host wrapper `dtw.correct_pred` (buffers, zero padding to the longest path, dtype / device of the result) around
    an interpreter of hwg_dtw_align that is the oracle itself; the kernel's own arithmetic is mirrored in
    tests/test_dtw_wavefront_cpu.py."""
    _load("test_dtw_gpu").test_dtw_matches_the_reference_golden(name)
    assert emulated_gpu.count("hwg_dtw_align") == 1
