"""Shape / hyper-parameter contract of the hot path (values of the reference's constants.py:29-71).

When the reference's own ``constants`` module is importable (drop-in use with its scripts) that module is used
instead, so edits made there (NMAX, BATCH_SIZE, DEVICE ...) are honoured; this file is the stand-alone default.
"""
NMAX = 150
NSTEPS = 30
CROP_STEP = 6
NFEATURES = 4
POINTNET_OUT_DIM = 1024
DTC_FILTERS = [16, 32, 64, 128, 256, 512]
SUP_LATENT_DIM = 32
DEC_MLP_SIZE = NSTEPS * NMAX * NFEATURES
LR = 1e-4
B1 = 0.9
B2 = 0.99
BATCH_SIZE = 16
EPOCHS = 50
CHECKPOINT_FREQUENCY = 5
GP_WEIGHT = 15
ADV_WEIGHT = 1
SUPERVISION_FREQUENCY = 1
DEVICE = "cuda"


def get():
    """The reference's ``constants`` module if it is importable, else this module."""
    import sys
    mod = sys.modules.get("constants")
    if mod is not None and hasattr(mod, "NSTEPS") and hasattr(mod, "DTC_FILTERS"):
        return mod
    return sys.modules[__name__]
