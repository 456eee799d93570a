"""ctypes binding of libhwg_b200.so (the C-ABI declared in include/hwg_b200.h).

There is no CPU fallback: if the library is missing or a call fails, the
caller gets a RuntimeError.  Tensors cross the boundary as raw device
pointers (`tensor.data_ptr()`), sizes and the current CUDA stream handle.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# HWG_LIB_PATH: development switch for same-box A/B timing of two builds of THIS library (tools/gpu_*.sh)
LIB_PATH = os.environ.get("HWG_LIB_PATH") or os.path.join(_HERE, "lib", "libhwg_b200.so")
_lib = None

c_int, c_i64, c_vp, c_f = ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_float

# name -> (restype, argtypes); mirrors include/hwg_b200.h one to one
_SIGNATURES = {
    "hwg_version": (c_int, []),
    "hwg_last_error": (ctypes.c_char_p, []),
    "hwg_launch_count": (ctypes.c_uint64, []),
    "hwg_last_conv_kernel": (c_int, []),
    "hwg_last_wgrad_kernel": (c_int, []),
    "hwg_ctc_forward": (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_i64, c_i64, c_int, c_vp, c_vp,
                                c_int, c_vp, c_vp, c_vp, c_vp]),
    "hwg_ctc_reduce_mean": (c_int, [c_vp, c_vp, c_int, c_vp, c_vp, c_vp]),
    "hwg_ctc_backward": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_i64, c_i64, c_int,
                                 c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_vp]),
    "hwg_ctc_greedy_decode": (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_vp]),
    "hwg_conv_fprop": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "hwg_conv_wgrad": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp]),
    "hwg_logsoftmax_bwd": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    "hwg_balance_chunk": (c_int, []),
    "hwg_balance": (c_int, [c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_vp]),
    "hwg_shift_expand": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "hwg_stem_conv": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    "hwg_shift_collapse": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "hwg_gn_coeffs": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_i64, c_f, c_vp, c_vp, c_vp]),
    "hwg_avgpool_nhwc": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "hwg_act_bwd": (c_int, [c_vp, c_vp, c_vp, c_f, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "hwg_norm_bwd_reduce": (c_int, [c_vp, c_vp, c_vp, c_f, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "hwg_gn_bwd_coeffs": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "hwg_norm_bwd_apply": (c_int, [c_vp, c_vp, c_vp, c_vp, c_f, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "hwg_spectral_norm": (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    "hwg_channel_sum": (c_int, [c_vp, c_i64, c_int, c_vp, c_vp]),
    "hwg_dtw_align": (c_int, [c_vp, c_int, c_int, c_int, c_vp, c_i64, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "hwg_insert_spaces_plan": (c_int, [c_vp, c_vp, c_int, c_vp, c_vp, c_int, c_int, ctypes.c_double, ctypes.c_double, c_vp, c_vp,
                                       c_vp, c_vp]),
    "hwg_insert_spaces_fill": (c_int, [c_vp, c_int, c_i64, c_i64, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "hwg_add_stats": (c_int, [c_vp, c_vp, c_vp, c_int, c_i64, c_int, c_vp, c_vp]),
    "hwg_l1_halves": (c_int, [c_vp, c_int, c_i64, c_f, c_f, c_vp, c_vp, c_vp]),
    "hwg_spectral_norm_bwd": (c_int, [c_vp, c_int, c_i64, c_vp, c_vp, c_vp]),
    "hwg_peer_mailbox_bytes": (c_i64, [c_int, c_int]),
    "hwg_peer_enable_access": (c_int, [c_int]),
    "hwg_bn_coeffs_peer": (c_int, [c_vp, c_int, c_int, c_i64, c_vp, c_vp, c_vp, c_vp, c_f, c_f, c_vp, c_vp,
                                   c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    "hwg_peer_allreduce_f32": (c_int, [c_vp, c_vp, c_int, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    "hwg_bn_bwd_reduce": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_vp, c_vp]),
    "hwg_bn_bwd_apply": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_int, c_int, c_vp, c_vp, c_vp]),
    "hwg_relu_maxpool_bwd": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_int, c_int, c_vp, c_vp, c_vp]),
    "hwg_hwr_stem_bwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    "hwg_hwr_stem_bwd_image": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "hwg_hwr_stem_bwd_expand": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "hwg_linear_f32": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_f, c_vp]),
    "hwg_pixelnorm_f32": (c_int, [c_vp, c_vp, c_int, c_int, c_vp]),
    "hwg_gen_pack_input": (c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "hwg_adain_coeffs": (c_int, [c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_int, c_f, c_vp, c_vp, c_vp]),
    "hwg_adain_bwd_reduce": (c_int, [c_vp, c_vp, c_vp, c_int, c_i64, c_int, c_vp, c_vp]),
    "hwg_adain_bwd_apply": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_f, c_vp,
                                    ctypes.c_uint64, ctypes.c_uint64, c_vp, c_int, c_vp, c_vp, c_vp]),
    "hwg_gen_output_bwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_i64, c_int, c_vp, c_vp, c_vp]),
    "hwg_bn_coeffs": (c_int, [c_vp, c_int, c_int, c_i64, c_vp, c_vp, c_vp, c_vp, c_f, c_f, c_int, c_vp, c_vp, c_vp]),
    "hwg_scale_shift_act": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_i64, c_int, c_int, c_f, c_vp]),
    "hwg_blur_noise_act_stats": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, ctypes.c_uint64,
                                         ctypes.c_uint64, c_vp, c_int, c_f, c_vp, c_vp]),
    "hwg_gen_output": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_i64, c_int, c_vp, c_vp]),
    "hwg_hwr_stem": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "hwg_adam_flat": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_f, c_f, c_f, c_f, c_f, c_f, c_vp, c_int, c_vp]),
    "hwg_linear_bwd_f32": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_f, c_vp, c_vp, c_vp, c_int, c_vp]),
    "hwg_linear_map": (c_int, [c_vp, c_vp, c_int, c_vp, c_vp, c_vp]),
    "hwg_map_items_per_block": (c_int, []),
    "hwg_maxpool_nhwc": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                 c_int, c_int, c_vp]),
}


HWG_MAX_TAPS = 16
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_LOGSOFTMAX = 0, 1, 2, 3
DT_BF16, DT_F32 = 0, 1


class ConvDesc(ctypes.Structure):
    """struct hwgConvDesc (include/hwg_b200.h)."""
    _fields_ = [
        ("N", ctypes.c_int32), ("H", ctypes.c_int32), ("W", ctypes.c_int32),
        ("Cin", ctypes.c_int32), ("x_pitch", ctypes.c_int32), ("Cout", ctypes.c_int32),
        ("Ho", ctypes.c_int32), ("Wo", ctypes.c_int32), ("ntaps", ctypes.c_int32),
        ("tap_dh", ctypes.c_int32 * HWG_MAX_TAPS), ("tap_dw", ctypes.c_int32 * HWG_MAX_TAPS),
        ("y_stride_n", ctypes.c_int64), ("y_stride_h", ctypes.c_int64), ("y_stride_w", ctypes.c_int64),
        ("y_dtype", ctypes.c_int32), ("act", ctypes.c_int32), ("slope", ctypes.c_float),
        ("tile_w", ctypes.c_int32),
        ("nz_stride_n", ctypes.c_int64), ("nz_stride_h", ctypes.c_int64), ("nz_stride_w", ctypes.c_int64),
        ("in_stride_h", ctypes.c_int32), ("in_stride_w", ctypes.c_int32),
        ("noise_seed", ctypes.c_uint64), ("noise_subseq", ctypes.c_uint64), ("noise_seed_dev", ctypes.c_uint64),
        ("fold_c", ctypes.c_int32), ("fold_w", ctypes.c_int32),
        ("fold_stride_h", ctypes.c_int64), ("fold_stride_w", ctypes.c_int64),
        ("fold_taps", ctypes.c_int32), ("force_tcgen05", ctypes.c_int32),
    ]


class WgradDesc(ctypes.Structure):
    """struct hwgWgradDesc (include/hwg_b200.h)."""
    _fields_ = [
        ("N", ctypes.c_int32), ("H", ctypes.c_int32), ("W", ctypes.c_int32), ("Cin", ctypes.c_int32),
        ("x_pitch", ctypes.c_int32), ("Ho", ctypes.c_int32), ("Wo", ctypes.c_int32), ("Cout", ctypes.c_int32),
        ("gy_pitch", ctypes.c_int32), ("ntaps", ctypes.c_int32),
        ("tap_dh", ctypes.c_int32 * HWG_MAX_TAPS), ("tap_dw", ctypes.c_int32 * HWG_MAX_TAPS),
        ("Hi", ctypes.c_int32), ("Wi", ctypes.c_int32), ("gy_stride_h", ctypes.c_int32),
        ("gy_stride_w", ctypes.c_int32), ("gy_off_h", ctypes.c_int32), ("gy_off_w", ctypes.c_int32),
        ("tap_gy_h", ctypes.c_int32 * HWG_MAX_TAPS), ("tap_gy_w", ctypes.c_int32 * HWG_MAX_TAPS),
    ]


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
            fn = getattr(lib, name, None)
            if fn is None:
                if os.environ.get("HWG_LIB_PATH"):      # an older build under the A/B switch: the entry simply cannot be called
                    continue
                raise RuntimeError(f"{LIB_PATH} does not export {name}: rebuild it (python -m handwriting_line_generation_b200.build)")
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


# The autograd.Functions of the drop-in modules keep their saved-for-backward state in a python object on ctx and drop it
# at the end of backward (what autograd does with retain_graph=False).  The reference trainer's gradient balancing calls
# .backward(retain_graph=True) several times on ONE graph (trainer/hw_with_style_trainer.py:303,314,325): with
# RETAIN_SAVED the state is kept until the graph itself is released, so repeated backward passes work.
RETAIN_SAVED = bool(os.environ.get("HWG_RETAIN_GRAPH"))


def set_retain_graph(flag=True):
    global RETAIN_SAVED
    RETAIN_SAVED = bool(flag)


def saved_state(state):
    """The saved-for-backward state of a module Function, or torch's own complaint when it has been released."""
    if state is None:
        raise RuntimeError("Trying to backward through the graph a second time: the saved state of this "
                           "handwriting_line_generation_b200 module was released by the first backward.  Call "
                           "handwriting_line_generation_b200.set_retain_graph(True) (or integrate.install(retain_graph=True)) "
                           "before the forward when .backward(retain_graph=True) is used, as the reference trainer's "
                           "gradient balancing does.")
    return state


_DEBUG_SYNC = bool(os.environ.get("HWG_DEBUG_SYNC"))   # development aid: synchronise after every launch, name the faulting one


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {lib.hwg_last_error().decode()}")
    if _DEBUG_SYNC:
        try:
            torch.cuda.synchronize()
        except RuntimeError as e:
            raise RuntimeError(f"{name}: device fault after this launch: {e}") from None


def named_params(module):
    """[(name, parameter)] of a module, computed once: `nn.Module.named_parameters()` walks the whole module tree on every
    call (65 parameters x ten calls per train step of the generator alone = 40 % of the step's host time).  Parameter
    OBJECTS persist over `.to()`, `load_state_dict` and optimizer steps (those change `.data` / `_version`, which the
    derived-weight caches key on), so the list is cached on the module; registering a new parameter invalidates it."""
    cache = module.__dict__.get("_hwg_named_params")
    n = sum(len(m._parameters) for m in module.modules()) if cache is None else cache[0]
    if cache is None:
        cache = module.__dict__["_hwg_named_params"] = (n, list(module.named_parameters()))
    return cache[1]


def params(module):
    return [p for _, p in named_params(module)]


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
