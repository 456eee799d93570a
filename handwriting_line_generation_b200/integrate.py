"""The import swap of INTEGRATION.md §1 as a function.

The reference looks its hot-path classes up by name inside `model/hw_with_style.py` (`SpacedGenerator` :188,
`CNNOnlyHWR` :160, `DiscriminatorAP` :198) and resolves the loss string "CTCLoss" by `eval()` after `from model.loss import *` (train.py:48).
`install()` rebinds those four names in an importable reference tree, so `HWWithStyle(config['model'])`, the trainer
and generate.py run the sm_100a path without a source change."""
import importlib


def install(encoder=False, retain_graph=False, dtw=False, spacer=False, style=False):
    """Call after the reference's root is on sys.path and before HWWithStyle / the trainer are built.
    Returns the list of (module name, attribute) pairs that were rebound.

    encoder=True also rebinds `Encoder2`, the perceptual encoder the trainer builds for `encoder_type: "2tight"`
    (`from model.autoencoder import Encoder2`, trainer/hw_with_style_trainer.py:15,148-149) — opt-in until the module has
    a green GPU parity run (encoder2.py: status).
    retain_graph=True keeps the modules' saved-for-backward state over repeated `.backward(retain_graph=True)` calls on one
    graph — what the trainer does when `balance_loss` is configured (trainer/hw_with_style_trainer.py:300-338).
    dtw=True rebinds `correct_pred` (model/hw_with_style.py:18, the DTW label alignment `autoencode` / `extract_style`
    call) to the one-launch version.
    spacer=True rebinds `CountCNN` (hw_with_style.py:204) and `HWWithStyle.insert_spaces` (:302-328).
    style=True rebinds `CharStyleEncoder` (hw_with_style.py:122)."""
    from . import CNNOnlyHWR, CTCLoss, DiscriminatorAP, SpacedGenerator, set_retain_graph
    swapped = []
    if spacer:
        # the text -> spacing front end of HWWithStyle.forward (hw_with_style.py:236-238): the CountCNN spacer and
        # insert_spaces (one host read instead of 2*L*B `.item()` calls; same numpy RNG stream, bit-identical text)
        from .count_cnn import CountCNN
        from .spacing import insert_spaces
        hws_ = importlib.import_module("model.hw_with_style")
        setattr(hws_, "CountCNN", CountCNN)
        setattr(hws_.HWWithStyle, "insert_spaces", insert_spaces)
        swapped += [("model.hw_with_style", "CountCNN"), ("model.hw_with_style", "HWWithStyle.insert_spaces")]
    if style:
        # the style extractor of the 'auto' / 'count' lessons (hw_with_style.py:122): model/char_style.py CharStyleEncoder
        from .char_style import CharStyleEncoder
        setattr(importlib.import_module("model.hw_with_style"), "CharStyleEncoder", CharStyleEncoder)
        swapped.append(("model.hw_with_style", "CharStyleEncoder"))
    if retain_graph:
        set_retain_graph(True)
    if dtw:
        from .dtw import correct_pred
        setattr(importlib.import_module("model.hw_with_style"), "correct_pred", correct_pred)
        swapped.append(("model.hw_with_style", "correct_pred"))
    if encoder:
        from .encoder2 import Encoder2
        for modname in ("model.autoencoder", "trainer.hw_with_style_trainer"):
            try:
                mod = importlib.import_module(modname)
            except ImportError:
                continue
            setattr(mod, "Encoder2", Encoder2)
            swapped.append((modname, "Encoder2"))
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
