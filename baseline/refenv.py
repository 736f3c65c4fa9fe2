"""Import the UNMODIFIED reference from ``baseline/_ref`` (see install_ref.py) and, optionally, swap this repository's
drop-in modules under it -- the recipe of INTEGRATION.md section 1 / SURVEY.md section 10 as importable functions.

Used by ``bench.py --impl reference`` / ``gpu_eager_baseline`` (the reference's own modules, CPU or ``cuda:0`` eager), by
``scripts/run_reference_scripts.py`` and the tests built on it.  Never imported by the product package.
"""
from __future__ import annotations

import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
ROOT = os.path.dirname(HERE)


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "models.py")) and os.path.exists(os.path.join(REF_DIR, "PCAA_ablation.py"))


def _stub(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        try:
            __import__(name)
            return sys.modules[name]
        except Exception:                               # noqa: BLE001 -- absent (or broken) optional plotting dependency
            m = types.ModuleType(name)
            sys.modules[name] = m
    for k, v in attrs.items():
        if not hasattr(m, k):
            setattr(m, k, v)
    return m


def activate(device: str = "cpu"):
    """Make ``import constants, models, utils, datasets, PCAA_ablation, inference_PCAA`` resolve to the reference.
    matplotlib / umap are imported at the top of the reference's utils.py:4-10 but are plotting-only and absent from this
    image: empty stand-in modules are registered for them.  Returns the reference's ``constants`` module with
    DEVICE / WANDB_MODE set (both are read at call time everywhere: train_AAE.py:39, utils.py:124)."""
    if not available():
        raise RuntimeError(f"reference not installed under {REF_DIR}: run `python baseline/install_ref.py` in the build container")
    mpl = _stub("matplotlib")
    plt = _stub("matplotlib.pyplot", Axes=object)
    _stub("matplotlib.lines", Line2D=object)
    _stub("matplotlib.colors")
    _stub("umap")
    if not hasattr(mpl, "pyplot"):
        mpl.pyplot = plt
    os.environ.setdefault("WANDB_MODE", "disabled")
    os.environ.setdefault("WANDB_SILENT", "true")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import constants                                    # the reference's own constants.py
    if os.path.dirname(os.path.abspath(constants.__file__)) != REF_DIR:
        raise RuntimeError(f"`constants` resolved to {constants.__file__}, not the reference")
    constants.WANDB_MODE = "disabled"
    constants.DEVICE = device
    return constants


def swap_in_b200():
    """register this repository's ``models`` / hot-path ``utils`` pieces under the names the
    reference's scripts import -- must run after activate() and BEFORE importing PCAA_ablation / inference_PCAA (they bind
    names at import, PCAA_ablation.py:12-25).  No reference file is edited."""
    if ROOT not in sys.path:
        sys.path.insert(1, ROOT)
    for name in ("models", "PCAA_ablation", "inference_PCAA", "train_AAE"):
        if name in sys.modules and name != "models":
            raise RuntimeError(f"swap_in_b200: {name} is already imported (it has bound the reference's classes)")
    import opensetgaitrecognition_pcaa_b200.models as b200_models        # picks up the already-imported `constants`
    import opensetgaitrecognition_pcaa_b200.utils as b200_utils
    import utils as ref_utils                                            # plotting / filename helpers stay the reference's
    if b200_models.constants is not sys.modules["constants"]:
        raise RuntimeError("the B200 modules were imported before the reference's constants: import order")
    shim = types.ModuleType("utils")
    shim.__dict__.update({k: v for k, v in ref_utils.__dict__.items() if not k.startswith("__")})
    shim.SeqChamferLoss = b200_utils.SeqChamferLoss
    shim.sample_distant_points = b200_utils.sample_distant_points
    sys.modules["models"] = b200_models
    sys.modules["utils"] = shim
    return b200_models, shim
