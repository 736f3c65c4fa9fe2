"""CPU oracle for the PCAA train / open-set-inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the shipped package may import this file;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it, and only as the checker / the reported CPU arm.

It restates, in explicit fp32 (torch CPU) and float64 (numpy) arithmetic, the
algorithm of the reference (rmazzier/OpenSetGaitRecognition_PCAA):

* PointNet shared MLP, BatchNorm2d (train + eval), ELU        models.py:6-34, 82-105
* mean pooling over points / frames                            models.py:242-249, 282-284
* causal dilated Conv1d + BatchNorm1d + ELU                    models.py:37-79, 108-160
* encoder heads                                                models.py:252-292
* decoder                                                      models.py:340-385
* conditional discriminator                                    models.py:405-421
* sequence Chamfer loss                                        utils.py:98-132
* prototype sampler                                            utils.py:216-251
* variant-4 (paper PCAA) train step                            PCAA_ablation.py:882-1021
* variant-2 (base train_CGAAE) / variant-3 (no decoder) steps  train_AAE.py:126-290, PCAA_ablation.py:500-660
* open-set likelihood, ROC/Youden threshold, k-window vote     inference_PCAA.py:129-136, 225-231, 239-314

Parity pin: the reference has no tests or golden vectors, so
this oracle is pinned against outputs of the reference itself executed in the
build container (``oracle/gen_golden.py`` imports /root/reference, runs the
reference modules *and* the unmodified ``train_variant4`` trainer, checks this
file against them and writes ``tests/golden/*.npz``).  ``tests/test_oracle_golden.py``
re-checks this file against those committed vectors on every run.

Parameters are plain dicts ``name -> tensor`` keyed with the reference's
``state_dict`` keys, prefixed by ``E.`` (CGEncoder), ``G.`` (CGDecoder), ``D.``
(CGDiscriminator) and ``GPH.`` (decoder projection head, ``PCAA_ablation.py:778-781``).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

# shape contract, constants.py:29-63
NSTEPS = 30
NFEATURES = 4
POINTNET_DIMS = [4, 512, 512, 1024, 1024]       # constants.py:36, models.py:86-98
DTC_FILTERS = [16, 32, 64, 128, 256, 512]       # constants.py:37
DTC_DILATIONS = [1, 2, 4, 1, 2, 4]              # models.py:111-151
SUP_LATENT_DIM = 32
BN_EPS = 1e-5
BN_MOMENTUM = 0.1

Params = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------- #
# elementary pieces
# --------------------------------------------------------------------------- #
def elu(x: torch.Tensor) -> torch.Tensor:
    """ELU(alpha=1): x if x>0 else exp(x)-1 (models.py:18,54)."""
    return torch.where(x > 0, x, torch.expm1(x))


def batchnorm_rows(y: torch.Tensor, gamma, beta, rmean, rvar, training: bool,
                   update: Optional[dict] = None, key: str = "") -> torch.Tensor:
    """BatchNorm over the rows of a channels-last matrix ``y[R, C]``.

    train: biased batch variance for normalisation, running stats updated with
    momentum 0.1 and the unbiased variance; eval: running stats
    (torch.nn.BatchNorm{1,2}d semantics used at models.py:29 and models.py:72).
    """
    if training:
        mean = y.mean(dim=0)
        var = ((y - mean) ** 2).mean(dim=0)
        if update is not None:
            n = y.shape[0]
            update[key + "running_mean"] = (1 - BN_MOMENTUM) * rmean + BN_MOMENTUM * mean.detach()
            update[key + "running_var"] = (1 - BN_MOMENTUM) * rvar + BN_MOMENTUM * var.detach() * n / max(n - 1, 1)
    else:
        mean, var = rmean, rvar
    return (y - mean) / torch.sqrt(var + BN_EPS) * gamma + beta


def pointnet_block(p: Params, x: torch.Tensor, training: bool, update: Optional[dict] = None,
                   pre: str = "E.pc_block.") -> torch.Tensor:
    """models.py:82-105.  x (B,4,T,N) -> channels-last activations [B*T*N, 1024]."""
    B, C, T, N = x.shape
    a = x.permute(0, 2, 3, 1).reshape(B * T * N, C)
    for l in range(1, 5):
        k = f"{pre}pointnet{l}.module."
        w = p[k + "0.weight"].reshape(p[k + "0.weight"].shape[0], -1)   # (Cout,Cin,1,1)
        y = a @ w.t() + p[k + "0.bias"]
        z = batchnorm_rows(y, p[k + "1.weight"], p[k + "1.bias"], p[k + "1.running_mean"],
                           p[k + "1.running_var"], training, update, k + "1.")
        a = elu(z)
    return a


def causal_dilated_conv(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, dil: int) -> torch.Tensor:
    """models.py:59-76: Conv1d(k=3, dilation=d, padding=2d) then drop the last 2d outputs.

    x [B,T,Cin] channels-last, w (Cout,Cin,3).  y[b,t] = b + sum_k W[:,:,k] x[b, t-(2-k)d]
    with zero for negative time indices.
    """
    B, T, Cin = x.shape
    y = b.expand(B, T, -1).clone()
    for k in range(3):
        shift = (2 - k) * dil
        if shift >= T:
            continue
        xs = torch.zeros_like(x)
        if shift == 0:
            xs = x
        else:
            xs[:, shift:, :] = x[:, : T - shift, :]
        y = y + xs @ w[:, :, k].t()
    return y


def tcn_block(p: Params, x: torch.Tensor, training: bool, update: Optional[dict] = None,
              pre: str = "E.tc_block.") -> torch.Tensor:
    """models.py:108-160.  x [B,T,1024] -> [B,T,512]."""
    B, T, _ = x.shape
    for l in range(1, 7):
        k = f"{pre}dtc{l}."
        y = causal_dilated_conv(x, p[k + "conv1d.weight"], p[k + "conv1d.bias"], DTC_DILATIONS[l - 1])
        C = y.shape[-1]
        z = batchnorm_rows(y.reshape(B * T, C), p[k + "batch_norm.weight"], p[k + "batch_norm.bias"],
                           p[k + "batch_norm.running_mean"], p[k + "batch_norm.running_var"],
                           training, update, k + "batch_norm.")
        x = elu(z).reshape(B, T, C)
    return x


def encoder_forward(p: Params, x: torch.Tensor, training: bool, use_projection_head: bool,
                    update: Optional[dict] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """CGEncoder.forward, models.py:279-292.  Returns (out_classes (B,C), sup_fv (B,32))."""
    B, _, T, N = x.shape
    a4 = pointnet_block(p, x, training, update)                     # [B*T*N, 1024]
    pooled = a4.reshape(B, T, N, -1).mean(dim=2)                    # AvgPool2d((1,N)), models.py:242,282
    h = tcn_block(p, pooled, training, update)                      # [B,T,512]
    g = h.mean(dim=1)                                               # AvgPool1d(NSTEPS), models.py:249,284
    fv = elu(g @ p["E.MLP_sup1.0.weight"].t() + p["E.MLP_sup1.0.bias"])
    hh = fv
    if use_projection_head:
        hh = elu(fv @ p["E.MLP_head.0.weight"].t() + p["E.MLP_head.0.bias"])
    logits = elu(hh @ p["E.MLP_sup2.0.weight"].t() + p["E.MLP_sup2.0.bias"])   # ELU on logits, models.py:271-277
    return logits, fv


def decoder_forward(p: Params, h: torch.Tensor, nmax: int, pre: str = "G.") -> torch.Tensor:
    """CGDecoder.forward, models.py:373-385 (bn1-4 exist but are never applied)."""
    x = h
    for l in range(1, 6):
        x = x @ p[f"{pre}dense{l}.weight"].t() + p[f"{pre}dense{l}.bias"]
        if l < 5:
            x = elu(x)
    return x.view(-1, NFEATURES, NSTEPS, nmax)


def proj_head_forward(p: Params, fv: torch.Tensor, pre: str = "GPH.") -> torch.Tensor:
    """decoder_projection_head: Linear(32,64)+ELU, PCAA_ablation.py:778-781."""
    return elu(fv @ p[pre + "0.weight"].t() + p[pre + "0.bias"])


def mean_learner_forward(p: Params, onehot: torch.Tensor, training: bool, update: Optional[dict] = None,
                         pre: str = "ML.") -> torch.Tensor:
    """GaussianMeanLearner.forward, models.py:424-443: 3 x (Linear, BatchNorm1d, ELU) + Linear(64, 32)."""
    h = onehot
    for i in (0, 3, 6):
        k = f"{pre}model.{i}."
        kb = f"{pre}model.{i + 1}."
        y = h @ p[k + "weight"].t() + p[k + "bias"]
        h = elu(batchnorm_rows(y, p[kb + "weight"], p[kb + "bias"], p[kb + "running_mean"], p[kb + "running_var"],
                               training, update, kb))
    return h @ p[f"{pre}model.9.weight"].t() + p[f"{pre}model.9.bias"]


def disc_forward(p: Params, x: torch.Tensor, onehot: torch.Tensor, pre: str = "D.") -> torch.Tensor:
    """CGDiscriminator.forward, models.py:418-421."""
    h = torch.cat([x, onehot], dim=-1)
    h = elu(h @ p[pre + "model.0.weight"].t() + p[pre + "model.0.bias"])
    h = elu(h @ p[pre + "model.2.weight"].t() + p[pre + "model.2.bias"])
    return h @ p[pre + "model.4.weight"].t() + p[pre + "model.4.bias"]


# --------------------------------------------------------------------------- #
# Chamfer, utils.py:98-132
# --------------------------------------------------------------------------- #
def pairwise_dist(gts: torch.Tensor, preds: torch.Tensor) -> torch.Tensor:
    """batch_pairwise_dist(x=gts, y=preds), utils.py:109-132.

    Inputs (B,C,T,N).  P[b,t,i,j] = |gt_i|^2 + |pred_j|^2 - 2 gt_i.pred_j (expanded form, unclamped).
    """
    x = gts.permute(0, 2, 3, 1)
    y = preds.permute(0, 2, 3, 1)
    rx = (x * x).sum(-1).unsqueeze(3)
    ry = (y * y).sum(-1).unsqueeze(2)
    zz = x @ y.transpose(2, 3)
    return rx + ry - 2 * zz


def chamfer(preds: torch.Tensor, gts: torch.Tensor, avg_out: bool = True):
    """SeqChamferLoss.forward, utils.py:98-107.

    Returns (loss, idx_gt_for_pred (B,T,N) int64, idx_pred_for_gt (B,T,N) int64);
    ties resolve to the lowest index (CPU torch.min behaviour).
    """
    P = pairwise_dist(gts, preds)
    m1, i1 = torch.min(P, 2)          # over gt i, per pred j
    m2, i2 = torch.min(P, 3)          # over pred j, per gt i
    per_frame = m1.sum(2) + m2.sum(2)
    loss = per_frame.mean() if avg_out else per_frame.mean(dim=1)
    return loss, i1, i2


# --------------------------------------------------------------------------- #
# prototype sampler, utils.py:216-251
# --------------------------------------------------------------------------- #
def sample_distant_points(dimension: int, n: int, min_dist: float, sphere_radius: float, seed: int = 42):
    rng = np.random.default_rng(seed)
    npoints = 10000
    vec = rng.standard_normal(size=(dimension, npoints))
    vec /= np.linalg.norm(vec, axis=0)
    vec = vec * sphere_radius
    pts = vec.T
    best = 0.0
    while best < min_dist:
        distances = np.ones(npoints) * 1e10
        far = rng.integers(low=0, high=npoints)
        sel = [far]
        for _ in range(n - 1):
            d = np.sum((pts - pts[far]) ** 2, axis=1)
            distances = np.minimum(distances, d)
            far = int(np.argmax(distances))
            sel.append(far)
        s = pts[sel]
        dd = np.sqrt(((s[:, None, :] - s[None, :, :]) ** 2).sum(-1))
        best = dd[dd > 0].min()
    return torch.tensor(s)


# --------------------------------------------------------------------------- #
# losses of the variant-4 step, PCAA_ablation.py:900-1013
# --------------------------------------------------------------------------- #
def d_loss_fn(p: Params, fv_detached, z, onehot, alphas, gp_weight: float):
    """WGAN-GP critic loss.  ``z = z0 + mus`` ; ``alphas`` (B,1) (PCAA_ablation.py:939-973)."""
    real = disc_forward(p, z, onehot)
    fake = disc_forward(p, fv_detached, onehot)
    interp = (z + alphas * (fv_detached - z)).detach().requires_grad_(True)
    di = disc_forward(p, interp, onehot)
    (grad,) = torch.autograd.grad(di, interp, torch.ones_like(di), create_graph=True)
    slopes = torch.sqrt((grad ** 2).sum(dim=1) + 1e-12)
    gp = ((slopes - 1) ** 2).mean()
    return fake.mean() - real.mean() + gp_weight * gp, gp


def cross_entropy(logits, gt):
    """torch.nn.CrossEntropyLoss() (mean), PCAA_ablation.py:1009."""
    lse = torch.logsumexp(logits, dim=1)
    return (lse - logits.gather(1, gt[:, None]).squeeze(1)).mean()


def adam_update(params: List[torch.Tensor], grads: List[Optional[torch.Tensor]], state: List[dict],
                lr: float, b1: float, b2: float, eps: float = 1e-8) -> None:
    """torch.optim.Adam single-tensor update (no weight decay / amsgrad), in place.

    Parameters whose grad is None are skipped (decoder bn1-4).
    """
    for prm, g, st in zip(params, grads, state):
        if g is None:
            continue
        if "step" not in st:
            st["step"] = 0
            st["m"] = torch.zeros_like(prm)
            st["v"] = torch.zeros_like(prm)
        st["step"] += 1
        t = st["step"]
        st["m"].mul_(b1).add_(g, alpha=1 - b1)
        st["v"].mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1 = 1 - b1 ** t
        bc2 = 1 - b2 ** t
        denom = (st["v"].sqrt() / math.sqrt(bc2)).add_(eps)
        prm.addcdiv_(st["m"], denom, value=-(lr / bc1))


G_PREFIXES = ("E.", "GPH.", "G.")      # optimizer_G chain order, PCAA_ablation.py:821-826
D_PREFIXES = ("DPH.", "D.")            # optimizer_D chain order, PCAA_ablation.py:828-833
_BUFFER_SUFFIX = ("running_mean", "running_var", "num_batches_tracked")


def is_buffer(name: str) -> bool:
    return name.endswith(_BUFFER_SUFFIX)


def trainable_names(p: Params, prefixes) -> List[str]:
    return [k for pre in prefixes for k in p if k.startswith(pre) and not is_buffer(k)]


def train_step(p: Params, opt_state: dict, pcs, gt, z0, alphas, means, cfg: dict, variant: int = 4,
               clone_leaves: bool = True, supervised: bool = True) -> dict:
    """One training iteration of the reference's AAE loops (in place on ``p``):

    * variant 4 -- the paper's PCAA, ``PCAA_ablation.py:882-1021``: encoder with projection head, decoder fed by
      the 32->64 decoder projection head;
    * variant 2 -- the base ``train_CGAAE`` loop, ``train_AAE.py:126-290`` (= ``train_variant2``,
      ``PCAA_ablation.py:381-389``): encoder without projection head, ``CGDecoder()`` fed by ``sup_fv`` directly;
    * variant 3 -- no decoder, ``PCAA_ablation.py:500-660``: ``tot = loss_g + sup``, and optimizer_G uses
      ``betas=(B1, B1)`` (``PCAA_ablation.py:452-456``);
    * variant 1 -- variant 4's networks with the GaussianMeanLearner's prototypes, ``PCAA_ablation.py:130-330``: ``mus``
      is the learner's output (train-mode BatchNorm1d over the batch of one-hot labels, so it depends on the batch
      composition), ``means`` is only used for its class count.  Reference quirk (verified by running it): ``z =
      Variable(z0 + mus)`` (``:186``) DETACHES, so although optimizer_D is built over ``chain(mean_learner,
      discriminator)`` (``:108-112``) the learner's parameters never receive a gradient (``grad is None``, Adam skips
      them); only its BatchNorm running statistics move.

    ``supervised=False`` is an iteration with ``i % SUPERVISION_FREQUENCY != 0`` (``PCAA_ablation.py:1005-1018``): the
    cross-entropy term is left out of ``tot``, so the classifier layers (``MLP_head``, ``MLP_sup2``) have ``grad is None`` after
    ``zero_grad()`` and ``torch.optim.Adam`` skips them entirely (their moments and per-parameter step counts do not move).

    cfg: LR, B1, B2, GP_WEIGHT, ADV_WEIGHT, NMAX.  ``z0`` (B,32) and ``alphas`` (B,1) are the host RNG draws of
    PCAA_ablation.py:915-931 / 944-948.  Returns losses, the class predictions and every gradient (by name).
    """
    assert variant in (1, 2, 3, 4)
    C = means.shape[0]
    nmax = cfg["NMAX"]
    head = variant in (1, 4)
    g_prefixes = {4: G_PREFIXES, 1: G_PREFIXES, 2: ("E.", "G."), 3: ("E.",)}[variant]
    d_prefixes = ("ML.", "D.") if variant == 1 else D_PREFIXES
    b2_g = cfg["B1"] if variant == 3 else cfg["B2"]
    out: dict = {}
    # leaf copies that require grad (clone_leaves=False: differentiate the stored tensors in place -- what the
    # reference's own modules do; used by the timed CPU baseline to avoid an 861 MB copy per step)
    if clone_leaves:
        leaves = {k: v.detach().clone().requires_grad_(True) for k, v in p.items() if not is_buffer(k)}
    else:
        leaves = {k: v.requires_grad_(True) for k, v in p.items() if not is_buffer(k)}
    q = dict(p)
    q.update(leaves)
    upd: dict = {}
    logits, fv = encoder_forward(q, pcs, True, head, upd)
    out["logits"], out["fv"] = logits.detach(), fv.detach()
    out["pred"] = torch.argmax(torch.softmax(logits.detach(), dim=1), dim=1)
    onehot = torch.nn.functional.one_hot(gt, C).float()
    mus = mean_learner_forward(q, onehot, True, upd).detach() if variant == 1 else onehot @ means
    z = z0 + mus

    # ---- discriminator step
    dl, gp = d_loss_fn(q, fv.detach(), z, onehot, alphas, cfg["GP_WEIGHT"])
    d_names = trainable_names(p, d_prefixes)
    d_grads = torch.autograd.grad(dl, [q[n] for n in d_names], allow_unused=True)
    out["d_loss"], out["gp"] = dl.detach(), gp.detach()
    out["d_grads"] = {n: (None if g is None else g.detach()) for n, g in zip(d_names, d_grads)}
    with torch.no_grad():
        adam_update([p[n] for n in d_names], list(d_grads), opt_state.setdefault("D", [dict() for _ in d_names]),
                    cfg["LR"], cfg["B1"], cfg["B2"])
    # generator step sees the *updated* critic (PCAA_ablation.py:996)
    if clone_leaves:
        for n in d_names:
            q[n] = p[n].detach().clone().requires_grad_(True)

    # ---- generator step
    loss_g = -disc_forward(q, fv, onehot).mean() * cfg["ADV_WEIGHT"]
    sup = cross_entropy(logits, gt)
    if variant == 3:
        rec_loss = torch.zeros(())
        tot = loss_g + sup if supervised else loss_g
    else:
        rec = decoder_forward(q, proj_head_forward(q, fv) if head else fv, nmax)
        rec_loss, i1, i2 = chamfer(rec, pcs)
        tot = rec_loss + loss_g + sup if supervised else rec_loss + loss_g
        out.update(rec=rec.detach(), idx_gt_for_pred=i1, idx_pred_for_gt=i2)
    g_names = trainable_names(p, g_prefixes)
    g_grads = torch.autograd.grad(tot, [q[n] for n in g_names], allow_unused=True)
    out.update(rec_loss=rec_loss.detach(), loss_g=loss_g.detach(), sup_loss=sup.detach(), tot_loss=tot.detach())
    out["g_grads"] = {n: (None if g is None else g.detach()) for n, g in zip(g_names, g_grads)}
    with torch.no_grad():
        adam_update([p[n] for n in g_names], list(g_grads), opt_state.setdefault("G", [dict() for _ in g_names]),
                    cfg["LR"], cfg["B1"], b2_g)
        for k, v in upd.items():
            p[k] = v
        for k in p:
            if k.endswith("num_batches_tracked") and (k.startswith("E.") or (variant == 1 and k.startswith("ML."))):
                p[k] = p[k] + 1
    return out


def train_step_variant4(p: Params, opt_state: dict, pcs, gt, z0, alphas, means, cfg: dict,
                        clone_leaves: bool = True) -> dict:
    """The paper-PCAA iteration, PCAA_ablation.py:882-1021 (see train_step)."""
    return train_step(p, opt_state, pcs, gt, z0, alphas, means, cfg, 4, clone_leaves)


# --------------------------------------------------------------------------- #
# open-set scoring, inference_PCAA.py:129-136, 225-231, 255-271
# --------------------------------------------------------------------------- #
def subject_sigma_scale(subject: int) -> np.ndarray:
    """Deterministic per-subject spread factors for `synth_batch` (synthetic identities: body size / gait dynamics)."""
    s = int(subject)
    return np.array([0.6 + 0.13 * (s % 7), 0.6 + 0.11 * ((3 * s) % 7), 0.7 + 0.09 * ((5 * s) % 7), 0.5 + 0.17 * ((2 * s + 1) % 7)])


def synth_subject_stream(subjects, crops_per_subject: int, nmax: int, seed: int):
    """An identity-ordered crop stream as MSRadarDataset(sequential=True) serves it: for every subject,
    `crops_per_subject` consecutive crops drawn with that subject's spread factors.  Returns (pcs, subject ids)."""
    xs, ids = [], []
    for i, s in enumerate(subjects):
        x, _ = synth_batch(crops_per_subject, nmax, 1, seed=seed + 1009 * i, sigma_scale=subject_sigma_scale(s))
        xs.append(x)
        ids += [int(s)] * crops_per_subject
    return torch.cat(xs), np.array(ids, dtype=np.int64)


def calibrate_bn(p: Params, pcs: torch.Tensor, use_projection_head: bool = True) -> Dict[str, torch.Tensor]:
    """Running statistics := the batch statistics of one training-mode forward over `pcs` (momentum undone), so that
    the eval-mode encoder is well conditioned on this data.  Returns {buffer name: tensor} (also written into p)."""
    upd: dict = {}
    q = dict(p)
    for k in list(q):
        if k.startswith("E.") and k.endswith("running_mean"):
            q[k] = torch.zeros_like(q[k])
        elif k.startswith("E.") and k.endswith("running_var"):
            q[k] = torch.zeros_like(q[k])
    with torch.no_grad():
        encoder_forward(q, pcs, True, use_projection_head, upd)
    out = {}
    for k, v in upd.items():
        out[k] = (v / BN_MOMENTUM).float()
        p[k] = out[k]
    return out


def joint_log_likelihood(x: np.ndarray, means: np.ndarray) -> np.ndarray:
    """log of the equal-weight mixture pdf (1/C) sum_c N(x; mu_c, I_d), float64.  x (M,d)."""
    x = np.asarray(x, dtype=np.float64).reshape(-1, means.shape[1])
    mu = np.asarray(means, dtype=np.float64)
    d = mu.shape[1]
    e = -0.5 * ((x[:, None, :] - mu[None, :, :]) ** 2).sum(-1)          # (M,C)
    m = e.max(axis=1, keepdims=True)
    lse = m[:, 0] + np.log(np.exp(e - m).sum(axis=1))
    return lse - 0.5 * d * math.log(2 * math.pi) - math.log(mu.shape[0])


def joint_likelihood(x: np.ndarray, means: np.ndarray) -> np.ndarray:
    """Linear-domain value the reference thresholds (underflows to 0.0 for far samples)."""
    x = np.asarray(x, dtype=np.float64).reshape(-1, means.shape[1])
    mu = np.asarray(means, dtype=np.float64)
    d = mu.shape[1]
    e = -0.5 * ((x[:, None, :] - mu[None, :, :]) ** 2).sum(-1)
    return (np.exp(e) * (2 * math.pi) ** (-0.5 * d)).sum(axis=1) / mu.shape[0]


def roc_youden_threshold(labels: np.ndarray, scores: np.ndarray) -> float:
    """thresholds[argmax(tpr-fpr)] of sklearn.metrics.roc_curve (drop_intermediate=True),
    inference_PCAA.py:225-231.  thresholds[0] = +inf (sklearn >= 1.3)."""
    labels = np.asarray(labels) == 1
    scores = np.asarray(scores, dtype=np.float64)
    order = np.argsort(scores, kind="mergesort")[::-1]
    s = scores[order]
    y = labels[order]
    distinct = np.where(np.diff(s))[0]
    idx = np.r_[distinct, y.size - 1]
    tps = np.cumsum(y)[idx].astype(np.float64)
    fps = (1 + idx - tps).astype(np.float64)
    thr = s[idx]
    if len(fps) > 2:
        keep = np.where(np.r_[True, np.logical_or(np.diff(fps, 2), np.diff(tps, 2)), True])[0]
        fps, tps, thr = fps[keep], tps[keep], thr[keep]
    tps = np.r_[0, tps]
    fps = np.r_[0, fps]
    thr = np.r_[np.inf, thr]
    fpr = fps / fps[-1]
    tpr = tps / tps[-1]
    return float(thr[np.argmax(tpr - fpr)])


def openset_vote(likelihoods: np.ndarray, preds: np.ndarray, threshold: float, k: int, n_labels: int) -> np.ndarray:
    """Windows of k consecutive crops: strict majority of likelihood > thr -> argmax(bincount(preds))
    (lowest class on ties) else ``n_labels`` (unknown).  inference_PCAA.py:255-271."""
    lk = np.asarray(likelihoods).reshape(-1, k)
    pr = np.asarray(preds).reshape(-1, k)
    out = np.empty(lk.shape[0], dtype=np.int64)
    for w in range(lk.shape[0]):
        n_above = int(np.sum(lk[w] > threshold))
        if n_above > k / 2:
            out[w] = int(np.argmax(np.bincount(pr[w])))
        else:
            out[w] = n_labels
    return out


def f1_scores(labels: np.ndarray, preds: np.ndarray) -> Dict[str, float]:
    """accuracy and micro / macro / weighted F1 as sklearn.metrics.f1_score computes them for single-label
    multi-class input (classes = union of labels and preds; zero_division -> 0).  inference_PCAA.py:326-332."""
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


def naive_sequential_procedure(k: int, test_emb, test_pred, test_labels, unseen_emb, unseen_pred, unseen_labels, means,
                               seed: int = 0, unseen_valid_ratio: float = 0.2, vote_test_emb=None, vote_unseen_emb=None):
    """inference_PCAA.py:117-347 restated on per-crop embeddings / class predictions (the eval-mode encoder gives the
    same embedding for a crop whether it is encoded alone (phase 1) or inside a batch of k (phase 2)).

    test_* : the TEST split in `sequential=True` order; unseen_* : the UNSEEN split (labels = unseen subject ids).
    vote_*_emb (optional): the embeddings of the SECOND encoding pass (batches of k, :251, :290) when they are available
    separately -- the threshold is searched on the first pass (:195-231), the votes use the second (:255-263); the two can
    differ in the last bits (batch-1 vs batch-k arithmetic).
    Returns dict(threshold, preds, labels, val_subjects, metrics)."""
    rng = np.random.default_rng(seed)                                             # :127
    test_labels, unseen_labels = np.asarray(test_labels), np.asarray(unseen_labels)
    subj = np.unique(unseen_labels)                                               # :178
    val_subj = rng.choice(subj, size=np.ceil(unseen_valid_ratio * len(subj)).astype(int), replace=False)   # :181-182
    is_val = np.isin(unseen_labels, val_subj)                                     # :184-187
    lik_test = joint_likelihood(test_emb, means)                                  # :195-202
    lik_unseen = joint_likelihood(unseen_emb, means)                              # :204-217
    scores = np.concatenate([lik_unseen[is_val], lik_test])                       # :225-228
    det = np.concatenate([np.zeros(int(is_val.sum())), np.ones(len(lik_test))])
    thr = roc_youden_threshold(det, scores)                                       # :230-231
    n_labels = len(np.unique(test_labels))                                        # :237
    if vote_test_emb is not None:
        lik_test = joint_likelihood(vote_test_emb, means)                         # :255-257
    if vote_unseen_emb is not None:
        lik_unseen = joint_likelihood(vote_unseen_emb, means)                     # :293-295
    preds, labels = [], []
    test_pred, unseen_pred = np.asarray(test_pred), np.asarray(unseen_pred)
    for w in range(len(test_labels) // k):                                        # :241, DataLoader(batch_size=k, drop_last=True)
        sl = slice(w * k, (w + 1) * k)
        if len(np.unique(test_labels[sl])) != 1:                                  # :243-244
            continue
        labels.append(int(test_labels[sl][0]))
        if np.sum(lik_test[sl] > thr) > k / 2:                                    # :262-263
            preds.append(int(np.argmax(np.bincount(test_pred[sl]))))              # :265-266
        else:
            preds.append(n_labels)                                                # :270
    for w in range(len(unseen_labels) // k):                                      # :276
        sl = slice(w * k, (w + 1) * k)
        if len(np.unique(unseen_labels[sl])) != 1:                                # :279-280
            continue
        if unseen_labels[sl][0] in val_subj:                                      # :284
            continue
        labels.append(n_labels)
        if np.sum(lik_unseen[sl] > thr) > k / 2:
            preds.append(int(np.argmax(np.bincount(unseen_pred[sl]))))
        else:
            preds.append(n_labels)
    preds, labels = np.array(preds, dtype=np.int64), np.array(labels, dtype=np.int64)
    return {"threshold": thr, "preds": preds, "labels": labels, "val_subjects": np.sort(val_subj),
            "metrics": f1_scores(labels, preds)}


# --------------------------------------------------------------------------- #
# deterministic parameters / inputs shared by the golden generator and the tests
# --------------------------------------------------------------------------- #
def param_shapes(n_classes: int, nmax: int, use_projection_head: bool = True,
                 dec_in: int = 64, mean_learner: bool = False) -> Dict[str, Tuple[int, ...]]:
    """state_dict names and shapes of CGEncoder / proj-head / CGDecoder / CGDiscriminator (/ GaussianMeanLearner)."""
    s: Dict[str, Tuple[int, ...]] = {}
    if mean_learner:
        dims = [n_classes, 16, 32, 64, SUP_LATENT_DIM]
        for j, i in enumerate((0, 3, 6, 9)):
            s[f"ML.model.{i}.weight"] = (dims[j + 1], dims[j])
            s[f"ML.model.{i}.bias"] = (dims[j + 1],)
            if i < 9:
                for nm in ("weight", "bias", "running_mean", "running_var"):
                    s[f"ML.model.{i + 1}.{nm}"] = (dims[j + 1],)
                s[f"ML.model.{i + 1}.num_batches_tracked"] = ()
    for l in range(1, 5):
        ci, co = POINTNET_DIMS[l - 1], POINTNET_DIMS[l]
        k = f"E.pc_block.pointnet{l}.module."
        s[k + "0.weight"] = (co, ci, 1, 1)
        s[k + "0.bias"] = (co,)
        s[k + "1.weight"] = (co,)
        s[k + "1.bias"] = (co,)
        s[k + "1.running_mean"] = (co,)
        s[k + "1.running_var"] = (co,)
        s[k + "1.num_batches_tracked"] = ()
    chans = [1024] + DTC_FILTERS
    for l in range(1, 7):
        ci, co = chans[l - 1], chans[l]
        k = f"E.tc_block.dtc{l}."
        s[k + "conv1d.weight"] = (co, ci, 3)
        s[k + "conv1d.bias"] = (co,)
        s[k + "batch_norm.weight"] = (co,)
        s[k + "batch_norm.bias"] = (co,)
        s[k + "batch_norm.running_mean"] = (co,)
        s[k + "batch_norm.running_var"] = (co,)
        s[k + "batch_norm.num_batches_tracked"] = ()
    s["E.MLP_sup1.0.weight"] = (32, 512)
    s["E.MLP_sup1.0.bias"] = (32,)
    head = 16 if use_projection_head else 32
    if use_projection_head:
        s["E.MLP_head.0.weight"] = (16, 32)
        s["E.MLP_head.0.bias"] = (16,)
    s["E.MLP_sup2.0.weight"] = (n_classes, head)
    s["E.MLP_sup2.0.bias"] = (n_classes,)
    s["GPH.0.weight"] = (dec_in, 32)
    s["GPH.0.bias"] = (dec_in,)
    S = NSTEPS * NFEATURES * nmax
    dims = [dec_in, S // 16, S // 8, S // 4, S // 2, S]
    for l in range(1, 6):
        s[f"G.dense{l}.weight"] = (dims[l], dims[l - 1])
        s[f"G.dense{l}.bias"] = (dims[l],)
        if l < 5:
            for nm in ("weight", "bias", "running_mean", "running_var"):
                s[f"G.bn{l}.{nm}"] = (dims[l],)
            s[f"G.bn{l}.num_batches_tracked"] = ()
    s["DPH.0.weight"] = (32, dec_in)      # discriminator_projection_head, PCAA_ablation.py:783-786 (unused)
    s["DPH.0.bias"] = (32,)
    s["D.model.0.weight"] = (64, 32 + n_classes)
    s["D.model.0.bias"] = (64,)
    s["D.model.2.weight"] = (32, 64)
    s["D.model.2.bias"] = (32,)
    s["D.model.4.weight"] = (1, 32)
    s["D.model.4.bias"] = (1,)
    return s


def det_params(n_classes: int, nmax: int, seed: int = 0, use_projection_head: bool = True,
               dec_in: int = 64, mean_learner: bool = False) -> Params:
    """Deterministic (numpy PCG64) parameters with torch-default-like scales and
    non-trivial BatchNorm affine / running statistics."""
    rng = np.random.default_rng(seed)
    p: Params = {}
    shapes = param_shapes(n_classes, nmax, use_projection_head, dec_in, mean_learner)
    ml_bn = tuple(f"ML.model.{i}." for i in (1, 4, 7))
    for name, shp in shapes.items():
        if name.endswith("num_batches_tracked"):
            p[name] = torch.tensor(0, dtype=torch.int64)
            continue
        if name.endswith("running_mean"):
            a = 0.1 * rng.standard_normal(shp)
        elif name.endswith("running_var"):
            a = rng.uniform(0.5, 1.5, shp)
        elif name.startswith(ml_bn) and name.endswith("weight"):
            a = 1.0 + 0.1 * rng.standard_normal(shp)
        elif name.startswith(ml_bn) and name.endswith("bias"):
            a = 0.1 * rng.standard_normal(shp)
        elif (".1.weight" in name or "batch_norm.weight" in name or (".bn" in name and name.endswith("weight"))):
            a = 1.0 + 0.1 * rng.standard_normal(shp)
        elif (".1.bias" in name or "batch_norm.bias" in name or (".bn" in name and name.endswith("bias"))):
            a = 0.1 * rng.standard_normal(shp)
        else:
            wname = name.rsplit(".", 1)[0] + ".weight"
            wshape = shapes[wname]
            fan_in = int(np.prod(wshape[1:]))
            bound = 1.0 / math.sqrt(fan_in)
            a = rng.uniform(-bound, bound, shp)
        p[name] = torch.from_numpy(np.asarray(a, dtype=np.float32).reshape(shp).copy())
    return p


def synth_batch(B: int, nmax: int, n_classes: int, seed: int = 1234, sigma_scale=None):
    """Synthetic mmGait10-shaped crops, SURVEY section 8(d) (mimics datasets.py:98-161,290-295):
    per-frame cardinality c~U{8..220}, points ~N(0,diag(.35,.35,.55,1.2)^2) around a drifting
    offset, pad by repeating random real points / subsample to nmax, subtract per-frame mean.
    ``sigma_scale`` (4,) optionally scales the per-feature spread (a stand-in for subject-specific gait statistics).
    Returns (pcs (B,4,T,N) float32, labels (B,) int64)."""
    rng = np.random.default_rng(seed)
    sig = np.array([0.35, 0.35, 0.55, 1.2])
    if sigma_scale is not None:
        sig = sig * np.asarray(sigma_scale, dtype=np.float64)
    out = np.empty((B, NSTEPS, nmax, NFEATURES), dtype=np.float64)
    for b in range(B):
        off = rng.normal(0, 1.0, 4) * np.array([1.0, 1.0, 0.2, 0.5])
        vel = rng.normal(0, 0.05, 4)
        for t in range(NSTEPS):
            c = int(rng.integers(8, 221))
            pts = rng.normal(0, 1, (c, 4)) * sig + off + vel * t
            if c < nmax:
                extra = rng.choice(c, nmax - c)
                pts = np.concatenate([pts, pts[extra]], axis=0)
            else:
                pts = pts[rng.choice(c, nmax, replace=False)]
            pts = pts - pts.mean(axis=0, keepdims=True)
            out[b, t] = pts
    pcs = torch.from_numpy(out.astype(np.float32)).permute(0, 3, 1, 2).contiguous()   # datasets.py:472
    labels = torch.from_numpy(rng.integers(0, n_classes, B).astype(np.int64))
    return pcs, labels
