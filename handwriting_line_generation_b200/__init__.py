"""handwriting_line_generation_b200 — sm_100a implementation of the HWWithStyle
training-step hot path (generator, recognizer, CTC) of herobd/handwriting_line_generation.

The CUDA lives in lib/libhwg_b200.so (C-ABI: include/hwg_b200.h); this package is the
host-side mirror of the reference's PyTorch surface."""
from . import _lib  # noqa: F401
from .ctc import CTCLoss, ctc_greedy_decode, naive_decode  # noqa: F401
from .pure_gen import SpacedGenerator  # noqa: F401
from .cnn_only_hwr import CNNOnlyHWR  # noqa: F401
from .optim import FlatAdam  # noqa: F401
from .discriminator_ap import DiscriminatorAP  # noqa: F401
from .encoder2 import Encoder2  # noqa: F401
from .count_cnn import CountCNN  # noqa: F401
from .char_style import CharStyleEncoder  # noqa: F401

__all__ = ["CTCLoss", "ctc_greedy_decode", "naive_decode", "SpacedGenerator", "CNNOnlyHWR", "FlatAdam",
           "DiscriminatorAP", "Encoder2", "CountCNN", "CharStyleEncoder"]

set_retain_graph = _lib.set_retain_graph   # keep saved state over repeated .backward(retain_graph=True) calls (see _lib.py)
