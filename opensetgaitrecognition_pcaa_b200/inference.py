"""Open-set inference of PCAA on the B200 path (reference: inference_PCAA.py:117-347, 382-469).

The reference encodes every crop twice (batch 1 for the threshold search, batch k for the vote) and evaluates the
Gaussian-mixture likelihood per sample on the host with scipy.  Here every crop is encoded ONCE in large eval-mode
batches (the eval-mode encoder gives the same embedding whatever the batch), embeddings and class predictions stay on
the device, the float64 log-likelihood and the k-window vote are fused kernels, and only the ROC / Youden threshold
(a function of all phase-1 scores) is computed on the host after one gather.  Skip rules, window order, label
conventions and metrics follow the reference line by line (cited below).

Batch-sharded over ranks with no data-path collective: each rank encodes / scores its slice of the crop stream
(``sharded_stream_inference``); ``dp.gather_scores`` collects the phase-1 scores once when the threshold is searched.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import engine, ops

# exp(ll) > 0 in float64  <=>  ll > log(2^-1075): the reference's linear-domain pdf underflows to exactly 0.0 below
# this, and `0.0 > threshold` is then false even for threshold == 0 (tie class (i))
LOG_MIN_POSITIVE = -1075 * math.log(2.0)


@torch.no_grad()
def encode(encoder, pcs: torch.Tensor, batch: int = 1024):
    """Eval-mode encoder forward (BatchNorm running statistics) over crops (M,4,30,N) in chunks of `batch`.
    Returns (sup_fv (M,32) fp32, pred (M,) int32 = argmax of the class logits, inference_PCAA.py:252-253)."""
    if not pcs.is_cuda:
        raise RuntimeError("PCAA B200 inference needs CUDA tensors (no CPU fallback)")
    P = {k: v for k, v in encoder.named_parameters()}
    P.update({k: v for k, v in encoder.named_buffers()})
    M = pcs.shape[0]
    fv = torch.empty((M, 32), device=pcs.device, dtype=torch.float32)
    pred = torch.empty(M, device=pcs.device, dtype=torch.int32)
    cache = encoder.__dict__.setdefault("_pcaa_eval_wb16", {})
    wb16, tcn_wb16 = _eval_weight_copies(encoder, P, cache)
    dummy_gt = torch.zeros(min(batch, max(M, 1)), device=pcs.device, dtype=torch.int64)
    for s in range(0, M, batch):
        x = pcs[s:s + batch].contiguous()
        logits, f, _ = engine.encoder_forward(x, P, False, encoder.use_projection_head, wb16, tcn_wb16)
        _, _, pr = ops.softmax_ce(logits, dummy_gt[: x.shape[0]], want_grad=False)        # softmax is monotone: same argmax
        fv[s:s + x.shape[0]] = f
        pred[s:s + x.shape[0]] = pr
    return fv, pred


def _eval_weight_copies(encoder, P, cache):
    """bf16 tensor-core operand copies of the PointNet / TCN weights, refreshed when a parameter changes."""
    key = tuple((P[n].data_ptr(), P[n]._version) for n in sorted(P) if n.endswith("module.0.weight") or n.endswith("conv1d.weight"))
    if cache.get("key") != key:
        wb = {}
        for l in (2, 3, 4):
            W = P[f"pc_block.pointnet{l}.module.0.weight"]
            wb[l] = ops.pack_bf16(W.view(W.shape[0], W.shape[1]))
        tw = {}
        for l in range(1, 7):
            W = P[f"tc_block.dtc{l}.conv1d.weight"]
            tw[l] = ops.pack_bf16(W.view(W.shape[0], W.shape[1] * 3))
        cache.update(key=key, wb=wb, tw=tw)
    return cache["wb"], cache["tw"]


def roc_youden_threshold(labels: np.ndarray, scores: np.ndarray) -> float:
    """``thresholds[argmax(tpr - fpr)]`` of ``sklearn.metrics.roc_curve(labels, scores)`` (drop_intermediate=True,
    thresholds[0] = +inf), the threshold rule of inference_PCAA.py:230-231, restated in numpy."""
    y = np.asarray(labels) == 1
    s = np.asarray(scores, dtype=np.float64)
    order = np.argsort(s, kind="mergesort")[::-1]
    s, y = s[order], y[order]
    idx = np.r_[np.where(np.diff(s))[0], y.size - 1]
    tps = np.cumsum(y)[idx].astype(np.float64)
    fps = 1.0 + idx - tps
    thr = s[idx]
    if len(fps) > 2:
        keep = np.where(np.r_[True, np.logical_or(np.diff(fps, 2), np.diff(tps, 2)), True])[0]
        fps, tps, thr = fps[keep], tps[keep], thr[keep]
    tps, fps, thr = np.r_[0, tps], np.r_[0, fps], np.r_[np.inf, thr]
    return float(thr[np.argmax(tps / tps[-1] - fps / fps[-1])])


def log_threshold(threshold: float) -> float:
    """Decision `pdf > threshold` of the reference (float64, linear domain) as a test on the log-likelihood."""
    if threshold <= 0.0:
        return LOG_MIN_POSITIVE if threshold == 0.0 else -math.inf
    return math.inf if math.isinf(threshold) else math.log(threshold)


def f1_scores(labels: np.ndarray, preds: np.ndarray) -> Dict[str, float]:
    """accuracy, f1 micro / macro / weighted as inference_PCAA.py:326-332 (sklearn.metrics.f1_score semantics for
    single-label multi-class input: classes = union of labels and predictions)."""
    labels, preds = np.asarray(labels).astype(np.int64), np.asarray(preds).astype(np.int64)
    classes = np.unique(np.concatenate([labels, preds]))
    f1, support = [], []
    for c in classes:
        tp = float(np.sum((preds == c) & (labels == c)))
        fp = float(np.sum((preds == c) & (labels != c)))
        fn = float(np.sum((preds != c) & (labels == c)))
        f1.append(0.0 if 2 * tp + fp + fn == 0 else 2 * tp / (2 * tp + fp + fn))
        support.append(float(np.sum(labels == c)))
    f1, support = np.array(f1), np.array(support)
    acc = float(np.mean(labels == preds))
    return {"accuracy": acc, "f1_micro": acc, "f1_macro": float(f1.mean()),
            "f1_weighted": float((f1 * support).sum() / support.sum())}


def _uniform_windows(labels: np.ndarray, k: int) -> np.ndarray:
    """Mask over the len//k windows of k consecutive crops (DataLoader(batch_size=k, drop_last=True, shuffle=False)):
    True where all k labels agree (inference_PCAA.py:243-244, 279-280 skip the others)."""
    nw = len(labels) // k
    w = np.asarray(labels[: nw * k]).reshape(nw, k)
    return (w == w[:, :1]).all(axis=1)


def score_and_vote(fv: torch.Tensor, pred: torch.Tensor, means: torch.Tensor, k: int, log_thr: float, n_labels: int):
    """Fused device path of inference_PCAA.py:255-271 for ALL len//k windows: float64 log-likelihood of every
    embedding, then per window `#(lik > thr) > k/2 ? lowest most-frequent class : n_labels`."""
    ll = ops.openset_score(fv, means)
    nw = fv.shape[0] // k
    return ll, ops.openset_vote(ll[: nw * k], pred[: nw * k].contiguous(), k, log_thr, n_labels)


@torch.no_grad()
def naive_sequential_procedure(k: int, encoder, discriminator_means: torch.Tensor, test_pcs: torch.Tensor,
                               test_labels: Sequence[int], unseen_pcs: torch.Tensor, unseen_labels: Sequence[int],
                               seed: int = 0, unseen_valid_ratio: float = 0.2, encode_batch: int = 1024,
                               embeddings: Optional[dict] = None):
    """inference_PCAA.py:117-347 on device tensors.  test_* is the TEST split, unseen_* the UNSEEN split, both in the
    dataset's `sequential=True` order; labels are the per-crop subject labels.  `embeddings` may carry pre-computed
    {"test": (fv, pred), "unseen": (fv, pred)} (e.g. from a previous k: the reference re-encodes for every k).
    Returns dict(threshold, preds, labels, metrics, embeddings)."""
    rng = np.random.default_rng(seed)                                              # :127
    test_labels, unseen_labels = np.asarray(test_labels), np.asarray(unseen_labels)
    means = discriminator_means.to(test_pcs.device if test_pcs is not None else "cuda").float().contiguous()
    # 1.2 unseen subjects used for the threshold search (:178-187)
    subj = np.unique(unseen_labels)
    val_subj = rng.choice(subj, size=np.ceil(unseen_valid_ratio * len(subj)).astype(int), replace=False)
    is_val = np.isin(unseen_labels, val_subj)
    # 1.3 every crop encoded once; float64 log-likelihoods on the device (:195-217)
    if embeddings is None:
        embeddings = {"test": encode(encoder, test_pcs, encode_batch), "unseen": encode(encoder, unseen_pcs, encode_batch)}
    (fv_t, pr_t), (fv_u, pr_u) = embeddings["test"], embeddings["unseen"]
    ll_t = ops.openset_score(fv_t, means)
    ll_u = ops.openset_score(fv_u, means)
    # 1.4 ROC / Youden threshold on the host, in the reference's linear float64 domain (underflow ties included)
    lik_t = np.exp(ll_t.cpu().numpy())
    lik_u = np.exp(ll_u.cpu().numpy())
    scores = np.concatenate([lik_u[is_val], lik_t])                                # :225-228
    det = np.concatenate([np.zeros(int(is_val.sum())), np.ones(len(lik_t))])
    thr = roc_youden_threshold(det, scores)                                        # :230-231
    lthr = log_threshold(thr)
    # 2. k-window vote (:239-314)
    n_labels = len(np.unique(test_labels))                                         # :237
    nw_t, nw_u = len(test_labels) // k, len(unseen_labels) // k
    votes_t = ops.openset_vote(ll_t[: nw_t * k], pr_t[: nw_t * k].contiguous(), k, lthr, n_labels).cpu().numpy()
    votes_u = ops.openset_vote(ll_u[: nw_u * k], pr_u[: nw_u * k].contiguous(), k, lthr, n_labels).cpu().numpy()
    keep_t = _uniform_windows(test_labels, k)                                      # :243-244
    keep_u = _uniform_windows(unseen_labels, k)                                    # :279-280
    first_u = unseen_labels[: nw_u * k].reshape(nw_u, k)[:, 0]
    keep_u &= ~np.isin(first_u, val_subj)                                          # :284
    first_t = test_labels[: nw_t * k].reshape(nw_t, k)[:, 0]
    preds = np.concatenate([votes_t[keep_t], votes_u[keep_u]]).astype(np.int64)
    labels = np.concatenate([first_t[keep_t], np.full(int(keep_u.sum()), n_labels)]).astype(np.int64)   # :246, :286
    return {"threshold": thr, "log_threshold": lthr, "preds": preds, "labels": labels, "val_subjects": np.sort(val_subj),
            "metrics": dict(n_steps=k, **f1_scores(labels, preds)), "embeddings": embeddings}


@torch.no_grad()
def sharded_stream_inference(encoder, discriminator_means: torch.Tensor, pcs_local: torch.Tensor, k: int,
                             log_thr: float, n_labels: int, encode_batch: int = 1024):
    """Config 5 of BASELINE.json: this rank's contiguous shard of a crop stream (a multiple of k crops, so no window
    straddles ranks) -> (log-likelihoods, window labels), no collective on the data path."""
    fv, pred = encode(encoder, pcs_local, encode_batch)
    ll, votes = score_and_vote(fv, pred, discriminator_means.to(fv.device).float().contiguous(), k, log_thr, n_labels)
    return ll, votes, pred
