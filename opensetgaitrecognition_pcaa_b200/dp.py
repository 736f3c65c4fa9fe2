"""Data-parallel plumbing of the PCAA train step and of batch-sharded open-set inference (one process per GPU).

The reference has no distributed code at all; this is the new capability BASELINE.json asks for:
training shards the batch over ranks and exchanges parameter gradients once per step (NCCL all-reduce over
NVLink 5 / NVSwitch), inference shards the stream of crops with no collective.

Everything here is backend-agnostic host logic (``nccl`` on the B200 box, ``gloo`` in the CPU tests):

* ``shard_range``     which samples of a global batch / stream a rank owns;
* ``global_draws``    the host RNG draws of PCAA_ablation.py:915-931, 944-948 made for the GLOBAL batch with the
                      reference's generators and sliced per rank, so an N-rank run consumes exactly the random
                      numbers a single-process run of the same global batch would;
* ``GradExchange``    bucketed sum-all-reduce of contiguous spans of one flat gradient buffer, each span launched
                      as soon as its producer kernels are enqueued (decoder gradients, 99 % of the bytes, go first
                      and overlap the PointNet backward); the 1/world factor is folded into the fused Adam kernel;
* ``gather_scores``   the one gather inference needs (per-sample scores for the host-side ROC threshold).
"""
from __future__ import annotations

import os

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def world_info(group=None) -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, end) of rank's contiguous shard of n items; the first n % world ranks take one extra item."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def global_draws(global_batch: int, latent_dim: int, rank: int, world: int, np_rng=None,
                 torch_gen: Optional[torch.Generator] = None):
    """z0 (B_local, latent) float32 ~ N(0,1) and alphas (B_local, 1) float32 ~ U(0,1), drawn for the global batch.

    z0 follows PCAA_ablation.py:915-921 (``np.random.normal`` in float64, cast to float32), alphas follow
    :944-948 (CPU ``torch.rand``).  With np_rng / torch_gen = None the global generators are used, exactly as the
    reference does; every rank must hold identically seeded generators.
    """
    normal = np.random.normal if np_rng is None else np_rng.normal
    z0 = torch.from_numpy(normal(0, 1, (global_batch, latent_dim))).float()
    alphas = torch.rand(size=(global_batch, 1), generator=torch_gen)
    s, e = shard_range(global_batch, rank, world)
    return z0[s:e].contiguous(), alphas[s:e].contiguous()


_SKIP_EXCHANGE = os.environ.get("PCAA_DP_SKIP_EXCHANGE", "0") == "1"


class GradExchange:
    """Sum-all-reduce of spans of a flat gradient buffer, overlappable with the kernels that follow.

    ``start(lo, hi)`` launches the reduction of ``flat[lo:hi]``: on CUDA it runs on a side stream that first waits
    for everything already enqueued on the current stream (the kernels that produced the span), so later kernels
    on the current stream overlap it; on CPU (gloo) it is an async work item.  ``finish()`` makes the current
    stream (or the host) wait for all started reductions.  With world size 1 both are no-ops.
    """

    def __init__(self, flat: torch.Tensor, group=None, side_stream: bool = False):
        self.flat = flat
        self.group = group
        self.rank, self.world = world_info(group)
        self._pending: List = []
        # side_stream: keep the side stream even for one rank, for work the caller chains behind a span's reduction
        # (`then`: e.g. the optimizer update of that span) so that it too overlaps the kernels that follow
        self._stream = torch.cuda.Stream(device=flat.device) if ((self.world > 1 or side_stream) and flat.is_cuda) else None
        self.bytes_reduced = 0

    def start(self, lo: int, hi: int, then=None) -> None:
        """Reduce flat[lo:hi]; `then()` (optional) is enqueued right behind the reduction -- on the side stream when
        there is one, so it overlaps whatever the caller enqueues next on the current stream."""
        if hi <= lo:
            return
        if self.world == 1 and (then is None or self._stream is None):
            if then is not None:
                then()
            return
        buf = self.flat[lo:hi]
        if _SKIP_EXCHANGE and self.world > 1:          # timing probe only (wrong numerics): measures what the collective costs
            if then is not None:
                then()
            return
        if self._stream is not None:
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(self._stream):
                self._stream.wait_event(ev)
                if self.world > 1:
                    self.bytes_reduced += buf.numel() * buf.element_size()
                    w = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                    if then is not None:
                        w.wait()                      # side stream waits for the collective, not the host
                    else:
                        self._pending.append(w)
                if then is not None:
                    then()
        else:
            self.bytes_reduced += buf.numel() * buf.element_size()
            w = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            if then is not None:
                w.wait()
                then()
            else:
                self._pending.append(w)

    def finish(self) -> None:
        for w in self._pending:
            w.wait()
        self._pending.clear()
        if self._stream is not None:
            torch.cuda.current_stream().wait_stream(self._stream)

    @property
    def grad_scale(self) -> float:
        """Factor that turns the reduced sum into the global-batch mean gradient (applied inside the Adam kernel)."""
        return 1.0 / self.world


def split_spans(lo: int, hi: int, max_elems: int, align: int = 8) -> List[Tuple[int, int]]:
    """Cut [lo, hi) into buckets of at most max_elems elements (multiples of `align`), in order."""
    if max_elems <= 0:
        return [(lo, hi)] if hi > lo else []
    step = max(align, max_elems // align * align)
    return [(s, min(hi, s + step)) for s in range(lo, hi, step)]


def gather_scores(local: torch.Tensor, counts: Sequence[int], group=None) -> Optional[torch.Tensor]:
    """Concatenate the per-rank 1-D score tensors (rank r holds counts[r] entries) on every rank.

    Used once per inference run: the ROC / Youden threshold of inference_PCAA.py:225-231 is a host-side function
    of ALL phase-1 scores.  Single process: returns `local`."""
    rank, world = world_info(group)
    if world == 1:
        return local
    n = max(counts)
    pad = torch.zeros(n, dtype=local.dtype, device=local.device)
    pad[: local.numel()] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[:c] for o, c in zip(outs, counts)])
