"""Drop-in replacements of the hot-path pieces of the reference's ``utils.py``.

Reference: utils.py:88-132 (SeqChamferLoss), 135-157 (gradient_penalty helper), 160-161 (save_model),
216-251 (sample_distant_points).  Plotting helpers of the reference are out of scope.
"""
from __future__ import annotations

import numpy as np
import torch

from . import constants as _c
from . import ops

constants = _c.get()


class _ChamferFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, preds, gts, avg_out):
        frame_loss, i1, i2 = ops.chamfer_fwd(preds, gts, want_idx=True)
        ctx.save_for_backward(preds, gts, i1, i2)
        ctx.avg_out = avg_out
        return ops.chamfer_reduce(frame_loss, avg_out)

    @staticmethod
    def backward(ctx, gout):
        preds, gts, i1, i2 = ctx.saved_tensors
        gout = gout.contiguous().float()
        g = ops.chamfer_bwd(preds, gts, i1, i2, gout, ctx.avg_out) if ctx.needs_input_grad[0] else None
        # the loss is symmetric in its two clouds: the gradient with respect to gts (only asked for when a caller makes the
        # "ground truth" differentiable; the reference's trainers never do) is the same kernel with the roles exchanged
        g_gts = ops.chamfer_bwd(gts, preds, i2, i1, gout, ctx.avg_out) if ctx.needs_input_grad[1] else None
        return g, g_gts, None


class SeqChamferLoss(torch.nn.Module):
    """Sequence Chamfer loss: sum over points of the squared nearest-neighbour distance in both directions,
    mean over frames and batch (utils.py:98-107); one fused shared-memory kernel per direction pair."""

    def __init__(self):
        super().__init__()
        self.use_cuda = torch.cuda.is_available()

    def forward(self, preds, gts, avg_out=True):
        if not preds.is_cuda:
            raise RuntimeError("SeqChamferLoss: the PCAA B200 implementation runs on CUDA tensors only")
        return _ChamferFn.apply(preds.float().contiguous(), gts.float().contiguous(), bool(avg_out))

    def batch_pairwise_dist(self, x, y):
        """P[b,t,i,j] = |x_i|^2 + |y_j|^2 - 2 x_i.y_j for x, y of shape (B, C, T, N) (utils.py:109-132)."""
        if x.shape != y.shape:
            raise ValueError("batch_pairwise_dist: this implementation needs clouds of equal size")
        return ops.pairwise_dist(x.float().contiguous(), y.float().contiguous())


def gradient_penalty(critic, z_noise, codes, latent_dim):
    """utils.py:135-157 (unused helper duplicate of the inline WGAN-GP term); kept for API completeness."""
    alphas = torch.rand(size=(constants.BATCH_SIZE, 1)).repeat(1, latent_dim).to(constants.DEVICE)
    interpolates = z_noise + alphas * (codes - z_noise)
    disc = critic(interpolates)
    grads = torch.autograd.grad(outputs=disc, inputs=interpolates, grad_outputs=torch.ones_like(disc),
                                create_graph=True, retain_graph=True, only_inputs=True)[0]
    slopes = torch.sqrt(torch.sum(grads ** 2, dim=1) + 1e-12)
    return ((slopes - 1) ** 2).mean()


def save_model(_model: torch.nn.Module, _path):
    """utils.save_model of the reference (`torch.save(model.state_dict(), path)`, utils.py:254-256) -- same keys, shapes
    and file format.  Tensors are cloned first: inside a PCAATrainer the parameters are views of one flat buffer, and
    torch.save serialises the WHOLE storage behind a view (871 MB per file instead of the module's own bytes)."""
    torch.save({k: v.detach().clone() for k, v in _model.state_dict().items()}, _path)


def openness(n_train, n_total):
    return 1 - np.sqrt(2 * n_train / (n_train + n_total))


def sample_distant_points(dimension, n, min_dist, sphere_radius, seed=42):
    """n prototypes on the radius-`sphere_radius` sphere of R^dimension by farthest-point sampling over 10 000
    candidates, repeated until their minimum pairwise distance reaches `min_dist` (utils.py:216-251).  Host-side,
    one-off, float64 -- it only defines the *input* `discriminator_means` of the kernels."""
    rng = np.random.default_rng(seed)
    n_candidates = 10000
    cand = rng.standard_normal(size=(dimension, n_candidates))
    cand /= np.linalg.norm(cand, axis=0)
    cand = (cand * sphere_radius).T                      # (n_candidates, dimension)
    closest = 0.0
    while closest < min_dist:
        dist_to_set = np.full(n_candidates, 1e10)
        current = rng.integers(low=0, high=n_candidates)
        chosen = [current]
        for _ in range(n - 1):
            dist_to_set = np.minimum(dist_to_set, np.sum((cand - cand[current]) ** 2, axis=1))
            current = int(np.argmax(dist_to_set))
            chosen.append(current)
        pts = cand[chosen]
        pair = torch.cdist(torch.tensor(pts), torch.tensor(pts))
        closest = torch.min(pair[pair > 0])
    return torch.tensor(pts)
