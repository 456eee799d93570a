"""CPU: libhwg_b200.so loads without a GPU and exports every symbol include/hwg_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "hwg_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(hwg_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(hwg_lib):
    names = declared_symbols()
    assert "hwg_ctc_forward" in names and "hwg_ctc_greedy_decode" in names
    for n in names:
        assert hasattr(hwg_lib, n), f"{n} declared in include/hwg_b200.h but not exported"


def test_binding_table_matches_header(hwg_lib):
    from handwriting_line_generation_b200 import _lib
    assert _lib.exported_symbols() == declared_symbols()


def test_version_and_error_string(hwg_lib):
    assert hwg_lib.hwg_version() >= 100
    assert isinstance(hwg_lib.hwg_last_error(), bytes)


def test_invalid_arguments_fail_loudly(hwg_lib):
    # argument validation happens before any CUDA call: testable without a GPU
    rc = hwg_lib.hwg_ctc_forward(None, 10, 2, 5, None, 0, 0, 3, None, None, 0, None, None, None, None)
    assert rc != 0 and b"null" in hwg_lib.hwg_last_error()
    rc = hwg_lib.hwg_ctc_greedy_decode(None, 1, 1, 1, None, 0, None, None, None, None)
    assert rc != 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "handwriting_line_generation_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("no CPU", ""), f"{f} mentions the oracle"


def test_cpu_tensor_is_rejected():
    import torch
    import pytest
    from handwriting_line_generation_b200 import CTCLoss
    lp = torch.zeros(4, 1, 3).log_softmax(2)
    with pytest.raises(RuntimeError, match="CUDA"):
        CTCLoss(lp, torch.ones(1, 1, dtype=torch.int32), torch.tensor([4]), torch.tensor([1]))


def test_header_is_plain_c(tmp_path):
    """include/hwg_b200.h is the whole boundary: it must compile as C99 (what a cgo / JNI / FFI binding would include)."""
    import shutil
    import subprocess
    import pytest
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    src = tmp_path / "hdr.c"
    src.write_text('#include "hwg_b200.h"\nint main(void) { return hwg_version() == 0; }\n')
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                        str(src)], capture_output=True, text=True)
    assert r.returncode == 0 and "warning" not in r.stderr, r.stderr
