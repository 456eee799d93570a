"""CPU: `__graft_entry__._smoke_more` — the toy-size 'gen' lesson chain smoke() runs on cuda:0 after its CTC check — executed
through the CPU interpreter of the C-ABI with `.cuda()` made a no-op: the same code, assertions and thresholds the driver
runs on the B200, so that an edit to a module cannot break smoke() unnoticed in the build container."""
import torch

from . import abi_emu


def test_smoke_chain_passes_through_the_interpreter(hwg_lib, monkeypatch):
    import __graft_entry__ as entry
    import handwriting_line_generation_b200 as pkg
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.nn.Module, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    with abi_emu.installed(monkeypatch) as calls:
        msg = entry._smoke_more(pkg)
    print(msg)
    assert "generator-gradient cosine" in msg and len(calls) > 200
