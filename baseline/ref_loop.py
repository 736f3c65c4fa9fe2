"""Time the reference's OWN training loop and its OWN eval-mode encoder, unmodified, from ``baseline/_ref``.

``time_train_variant4`` calls ``PCAA_ablation.train_variant4`` itself (PCAA_ablation.py:746-1122: the reference's module
construction, optimizers, DataLoader(num_workers=0), host RNG draws, ``.to(DEVICE)`` copies, ``.item()`` reads, prints)
on ``constants.DEVICE`` = "cpu" or "cuda" (stock torch eager, fp32; on CUDA cuDNN may use TF32 for the 1x1 convolutions
exactly as the stock reference would).  Only the data source is supplied from outside: the name ``MSRadarDataset`` inside
``PCAA_ablation`` is bound to an in-memory dataset of synthetic crops with the reference's ``__getitem__`` contract
(datasets.py:466-479) -- the raw mmGait10 recordings are not available offline, and reading .npy files would time the disk.
Iteration boundaries are observed with a forward pre-hook on the training-mode ``CGEncoder`` (first call of every
iteration, PCAA_ablation.py:889); the loop is left by an exception raised from that hook once enough iterations are timed.

Used by ``bench.py --impl reference`` (CPU arm, all host cores) and for ``gpu_eager_baseline`` (the same loop on cuda:0).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import tempfile
import time

import numpy as np
import torch

from . import refenv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Done(Exception):
    pass


def _synthetic(n_distinct: int, nmax: int, classes: int, seed: int):
    if ROOT not in sys.path:
        sys.path.insert(1, ROOT)
    from opensetgaitrecognition_pcaa_b200 import synth            # numpy-only crop synthesiser (datasets.py:98-161 shaped)
    return synth.synth_batch(n_distinct, nmax, classes, seed=seed)


def time_train_variant4(device: str, batch: int, nmax: int, classes: int, warmup: int, steps: int,
                        budget_s: float = 240.0, quiet: bool = True, seed: int = 0):
    """Returns dict(times=[seconds per timed iteration], warmup=w_done, device=...).  At least one warm-up and one timed
    iteration always run; fewer than `steps` are timed when `budget_s` of wall time is exceeded (the caller reports the
    count that was actually timed)."""
    constants = refenv.activate(device)
    import PCAA_ablation                                           # binds the reference's models / utils / datasets

    n_iter = warmup + steps + 1                                    # the hook of iteration n_iter closes iteration n_iter-1
    pcs, gt = _synthetic(min(3, n_iter) * batch, nmax, classes, seed=1234)
    distinct = pcs.shape[0]

    class InMemoryCrops(torch.utils.data.Dataset):
        """MSRadarDataset's item contract (datasets.py:466-479): ((4, 30, N) float32, LongTensor label)."""

        def __init__(self, split, *a, **k):
            self.split = split
            self.n = n_iter * batch if split == constants.SPLIT.TRAIN else batch

        def __len__(self):
            return self.n

        def __getitem__(self, idx):
            j = idx % distinct
            return pcs[j], gt[j].type(torch.LongTensor)

    stamps = []
    t_start = time.perf_counter()
    sync = torch.cuda.synchronize if device.startswith("cuda") else (lambda: None)

    def pre_hook(mod, args):
        if type(mod).__name__ == "CGEncoder" and mod.training:
            sync()
            now = time.perf_counter()
            stamps.append(now)
            done = len(stamps) - 1                                  # completed iterations
            timed = done - warmup
            if done >= warmup + steps or (timed >= 1 and now - t_start > budget_s):
                raise _Done()
        return None

    cfg = constants.CONFIG
    cfg.update(MODEL_NAME="ref_timing", TRAIN_CLASSES=list(range(classes)), EPOCHS=1, BATCH_SIZE=batch, NMAX=nmax)
    constants.BATCH_SIZE = batch                                   # the gradient-penalty alphas read the module constant
    saved_ds = PCAA_ablation.MSRadarDataset
    PCAA_ablation.MSRadarDataset = InMemoryCrops
    handle = torch.nn.modules.module.register_module_forward_pre_hook(pre_hook)
    cwd = os.getcwd()
    scratch = tempfile.mkdtemp(prefix="pcaa_ref_loop_")            # train_variant4 writes models/<name>/config.pkl etc.
    os.chdir(scratch)
    torch.manual_seed(seed)
    np.random.seed(seed)
    out = io.StringIO()
    try:
        with contextlib.redirect_stdout(out) if quiet else contextlib.nullcontext():
            try:
                PCAA_ablation.train_variant4(cfg, wandb_mode="disabled", proj_head_on_discriminator=False)
            except _Done:
                pass
    finally:
        handle.remove()
        PCAA_ablation.MSRadarDataset = saved_ds
        os.chdir(cwd)
    durations = [b - a for a, b in zip(stamps[:-1], stamps[1:])]
    return {"times": durations[warmup:], "warmup": min(warmup, len(durations)), "device": device,
            "last_print": out.getvalue().strip().splitlines()[-1:] if quiet else []}


def time_eval_encoder(device: str, nmax: int, classes: int, k: int, windows: int, seed: int = 0):
    """The reference's CGEncoder (models.py:232-292) in eval mode on windows of k crops, as phase 2 of
    inference_PCAA.naive_sequential_procedure feeds it (:251, :290).  Returns (embeddings (n,32), logits (n,C), seconds)."""
    constants = refenv.activate(device)
    import models
    torch.manual_seed(seed)
    enc = models.CGEncoder(n_out_labels=classes, use_projection_head=True, nmax_points=nmax).to(device).float().eval()
    pcs, _ = _synthetic(windows * k, nmax, classes, seed=99)
    sync = torch.cuda.synchronize if device.startswith("cuda") else (lambda: None)
    fvs, logits = [], []
    with torch.no_grad():
        enc(pcs[:k].to(device))                                     # warm-up
        sync()
        t0 = time.perf_counter()
        for w in range(windows):
            lg, fv = enc(pcs[w * k:(w + 1) * k].to(device))
            fvs.append(fv.cpu()), logits.append(lg.cpu())
        sync()
        dt = time.perf_counter() - t0
    return torch.cat(fvs), torch.cat(logits), dt
