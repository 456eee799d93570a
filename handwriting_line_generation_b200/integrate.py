"""The import swap of INTEGRATION.md §1 as a function.

The reference looks its hot-path classes up by name inside `model/hw_with_style.py` (`SpacedGenerator` :188,
`CNNOnlyHWR` :160, `DiscriminatorAP` :198) and resolves the loss string "CTCLoss" by `eval()` after `from model.loss import *` (train.py:48).
`install()` rebinds those four names in an importable reference tree, so `HWWithStyle(config['model'])`, the trainer
and generate.py run the sm_100a path without a source change."""
import importlib


def install():
    """Call after the reference's root is on sys.path and before HWWithStyle / the trainer are built.
    Returns the list of (module name, attribute) pairs that were rebound."""
    from . import CNNOnlyHWR, CTCLoss, DiscriminatorAP, SpacedGenerator
    swapped = []
    hws = importlib.import_module("model.hw_with_style")
    for name, cls in (("SpacedGenerator", SpacedGenerator), ("CNNOnlyHWR", CNNOnlyHWR),
                      ("DiscriminatorAP", DiscriminatorAP)):
        setattr(hws, name, cls)
        swapped.append(("model.hw_with_style", name))
    for modname in ("model.loss", "model"):
        try:
            mod = importlib.import_module(modname)
        except ImportError:
            continue
        if hasattr(mod, "CTCLoss"):
            setattr(mod, "CTCLoss", CTCLoss)
            swapped.append((modname, "CTCLoss"))
    return swapped
