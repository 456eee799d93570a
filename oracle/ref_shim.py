"""TEST INFRASTRUCTURE — imports the unmodified reference from /root/reference (present only in
the build container, never on the GPU box).  Used by oracle/make_golden.py alone.

Shims (none touch hot-path arithmetic): stub `skimage`/`editdistance`
(not installed), and register the reference's `datasets/` directory as a namespace
package so it is not shadowed by the installed HuggingFace `datasets`."""
import os
import sys
import types

REF = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF, "model"))


def install():
    if not available():
        raise RuntimeError("/root/reference is not present on this machine")
    if REF not in sys.path:
        sys.path.insert(0, REF)

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    if "skimage" not in sys.modules:
        sk = stub("skimage")
        sk.draw = stub("skimage.draw", line=lambda *a: None)
        sk.morphology = stub("skimage.morphology", skeletonize=lambda x: x)
    if "editdistance" not in sys.modules:
        stub("editdistance", eval=lambda a, b: 0)
    ds = types.ModuleType("datasets")
    ds.__path__ = [os.path.join(REF, "datasets")]
    sys.modules["datasets"] = ds
