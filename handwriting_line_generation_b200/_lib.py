"""ctypes binding of libhwg_b200.so (the C-ABI declared in include/hwg_b200.h).

There is no CPU fallback: if the library is missing or a call fails, the
caller gets a RuntimeError.  Tensors cross the boundary as raw device
pointers (`tensor.data_ptr()`), sizes and the current CUDA stream handle.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libhwg_b200.so")
_lib = None

c_int, c_i64, c_vp, c_f = ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_float

# name -> (restype, argtypes); mirrors include/hwg_b200.h one to one
_SIGNATURES = {
    "hwg_version": (c_int, []),
    "hwg_last_error": (ctypes.c_char_p, []),
    "hwg_launch_count": (ctypes.c_uint64, []),
    "hwg_ctc_forward": (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_i64, c_i64, c_int, c_vp, c_vp,
                                c_int, c_vp, c_vp, c_vp, c_vp]),
    "hwg_ctc_reduce_mean": (c_int, [c_vp, c_vp, c_int, c_vp, c_vp, c_vp]),
    "hwg_ctc_backward": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_i64, c_i64, c_int,
                                 c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_vp]),
    "hwg_ctc_greedy_decode": (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_vp]),
}


def exported_symbols():
    return sorted(_SIGNATURES)


def load():
    """Loads the library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m handwriting_line_generation_b200.build` "
                "(there is no CPU or PyTorch fallback for this path)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {lib.hwg_last_error().decode()}")


def launch_count():
    return int(load().hwg_launch_count())


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("handwriting_line_generation_b200 runs on CUDA tensors only "
                               "(sm_100a kernels; no CPU fallback)")
