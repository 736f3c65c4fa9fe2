"""Synthetic mmGait10-shaped radar point-cloud crops (the real dataset is not available offline).

Mimics what the reference's data layer produces (datasets.py:98-161, 290-295, 466-479): 30 frames per crop, a
per-frame cloud of c ~ U{8..220} points (x, y, z, doppler) around a slowly moving offset, padded to `nmax` by
repeating random real points or sub-sampled to `nmax`, per-frame per-feature mean removed; crops are stored as
float64 ``(30, nmax, 4)`` ``.npy`` files and served as float32 ``(4, 30, nmax)`` tensors.
"""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np
import torch

NSTEPS, NFEATURES = 30, 4
_SIGMA = np.array([0.35, 0.35, 0.55, 1.2])


def synth_crops(n: int, nmax: int, seed: int = 0) -> np.ndarray:
    """n crops as float64 (n, 30, nmax, 4) -- the on-disk layout of datasets.py:141-150."""
    rng = np.random.default_rng(seed)
    out = np.empty((n, NSTEPS, nmax, NFEATURES), dtype=np.float64)
    for i in range(n):
        off = rng.normal(0, 1.0, 4) * np.array([1.0, 1.0, 0.2, 0.5])
        vel = rng.normal(0, 0.05, 4)
        card = rng.integers(8, 221, NSTEPS)
        for t in range(NSTEPS):
            c = int(card[t])
            pts = rng.normal(0, 1, (c, 4)) * _SIGMA + off + vel * t
            if c < nmax:
                idx = np.concatenate([np.arange(c), rng.integers(0, c, nmax - c)])
            else:
                idx = rng.permutation(c)[:nmax]
            pts = pts[idx]
            out[i, t] = pts - pts.mean(axis=0, keepdims=True)
    return out


def synth_batch(batch: int, nmax: int, n_classes: int, seed: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """(pcs (B,4,30,nmax) float32, labels (B,) int64) as MSRadarDataset.__getitem__ + default collate give them."""
    crops = synth_crops(batch, nmax, seed)
    pcs = torch.from_numpy(crops.astype(np.float32)).permute(0, 3, 1, 2).contiguous()
    labels = torch.from_numpy(np.random.default_rng(seed + 7919).integers(0, n_classes, batch).astype(np.int64))
    return pcs, labels


def write_dataset(root: str, nmax: int, train_subjects, unseen_subjects, crops_per_track: int = 8,
                  tracks_per_subject: int = 2, seed: int = 0) -> None:
    """Write crops in the reference's file convention (datasets.py:62-76):
    ``<root>/{train,valid,test,unseen}/crop{i}_subj{s}_{scenario}_track{id}.npy``, float64 (30, nmax, 4)."""
    scenarios = ["free_walk", "hands_in_pockets", "smartphone"]
    rng = np.random.default_rng(seed)
    for split, subjects in (("train", train_subjects), ("valid", train_subjects), ("test", train_subjects),
                            ("unseen", unseen_subjects)):
        d = os.path.join(root, split)
        os.makedirs(d, exist_ok=True)
        for s in subjects:
            for tr in range(tracks_per_subject):
                crops = synth_crops(crops_per_track, nmax, int(rng.integers(1 << 30)))
                tid = f"{s:02d}{tr:02d}{ {'train': 0, 'valid': 1, 'test': 2, 'unseen': 3}[split] }"
                for c in range(crops_per_track):
                    np.save(os.path.join(d, f"crop{c}_subj{s}_{scenarios[tr % 3]}_track{tid}.npy"), crops[c])
