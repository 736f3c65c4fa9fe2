"""Import the *reference* (read-only, /root/reference) in the build container.

Used only by ``oracle/gen_golden.py`` (golden-vector generation / oracle pinning).
/root/reference does not exist on the GPU box, so nothing under tests/, smoke() or
bench.py imports this module at run time.
"""
import os
import sys
import types

REF = os.environ.get("PCAA_REFERENCE", "/root/reference")


def load_reference():
    if not os.path.isdir(REF):
        raise RuntimeError(f"reference tree not found at {REF}")
    # matplotlib / umap are imported at module top by the reference's utils.py:4-10 but absent here
    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules.setdefault(name, m)
        return sys.modules[name]
    mpl = stub("matplotlib")
    plt = stub("matplotlib.pyplot", Axes=object)
    stub("matplotlib.lines", Line2D=object)
    stub("matplotlib.colors")
    stub("umap")
    mpl.pyplot = plt
    os.environ.setdefault("WANDB_MODE", "disabled")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import constants, models, utils  # noqa: E401
    constants.WANDB_MODE = "disabled"
    constants.DEVICE = "cpu"
    return constants, models, utils
