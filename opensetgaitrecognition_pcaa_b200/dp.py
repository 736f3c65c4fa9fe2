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
* ``PeerExchange``    the same sum-all-reduce done by the COPY ENGINES over NVLink peer memory instead of NCCL kernels:
                      every rank pulls its chunk of the span from each peer's gradient buffer (symmetric memory, P2P
                      mapped), adds the pulled chunks locally (``pcaa_sum_into``) and pulls the other ranks' reduced
                      chunks back.  No SM is taken from the persistent tcgen05 GEMMs of the encoder backward that
                      runs beside it (NCCL's all-reduce CTAs cost the step ~1 ms on 2 GPUs);
                      ``GradExchange.start_sharded`` uses its two halves as a sharded optimizer (ZeRO-1): reduce-scatter
                      the gradients, update the owned 1/world of the parameters, all-gather the updated parameters;
* ``gather_scores``   the one gather inference needs (per-sample scores for the host-side ROC threshold).
"""
from __future__ import annotations

import os

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


SINGLE = "single"      # process_group sentinel: behave as one rank even inside an initialised multi-rank job (parity checks)


def world_info(group=None) -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised or group is dp.SINGLE."""
    if isinstance(group, str) and group == SINGLE:
        return 0, 1
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


def exchange_mode() -> str:
    """"peer" (copy engines over NVLink peer memory; default, falls back to NCCL when the symmetric-memory rendezvous is
    not possible) or "nccl"; PCAA_DP_EXCHANGE overrides."""
    return os.environ.get("PCAA_DP_EXCHANGE", "peer").lower()


def alloc_symmetric(n: int, dtype, device, group=None):
    """zeros(n) of `dtype` in symmetric (peer-mapped) memory plus its rendezvous handle, or (plain zeros, None) when the peer
    mode is off / not possible.  Collective when world > 1 and the peer mode is on: all ranks must call it in the same order."""
    rank, world = world_info(group)
    if world > 1 and exchange_mode() == "peer" and torch.device(device).type == "cuda":
        try:
            import torch.distributed._symmetric_memory as symm
            t = symm.empty(n, dtype=dtype, device=torch.device(device))
            t.zero_()
            return t, symm.rendezvous(t, group if group is not None else dist.group.WORLD)
        except Exception as e:                       # noqa: BLE001 -- any failure here means "no peer access"
            if rank == 0:
                print(f"[pcaa dp] symmetric memory unavailable ({type(e).__name__}: {e})", flush=True)
    return torch.zeros(n, device=device, dtype=dtype), None


def alloc_exchange_buffer(n: int, device, group=None):
    """fp32 zeros(n) for a gradient buffer that will be exchanged, plus a PeerExchange when the peer mode is on and the
    symmetric-memory rendezvous succeeds (one process per GPU of one NVLink domain); otherwise (tensor, None) and the
    exchange goes through NCCL.  Collective when world > 1 and the peer mode is on."""
    t, hdl = alloc_symmetric(n, torch.float32, device, group)
    return t, (PeerExchange(t, hdl) if hdl is not None else None)


class PeerExchange:
    """Sum-all-reduce of spans of a symmetric-memory fp32 buffer by peer copies (copy engines) + one local add kernel.

    For a span [lo, hi) split into `world` chunks (multiples of 8 elements), on the CURRENT stream:
      barrier                      every rank's span is final (each rank's stream already waited for its producers)
      pull   stage[k] <- peer_k.buf[my chunk]      one cudaMemcpyAsync per peer, each on its own stream (own copy engine)
      add    buf[my chunk] += sum_k stage[k]       pcaa_sum_into
      barrier                      every chunk is reduced
      pull   buf[chunk_k] <- peer_k.buf[chunk_k]   the all-gather half
      barrier                      nobody still reads this rank's buffer (the next backward overwrites it)
    Per-rank NVLink traffic equals a ring all-reduce's (2 (w-1)/w of the span), SM time is one short HBM-bound kernel."""

    def __init__(self, flat: torch.Tensor, handle):
        self.flat, self.hdl = flat, handle
        self.rank, self.world = handle.rank, handle.world_size
        self.peers = [(self.rank + k) % self.world for k in range(1, self.world)]
        self.streams = [torch.cuda.Stream(device=flat.device) for _ in self.peers]
        self.stage: Optional[torch.Tensor] = None
        self.stage_small: Optional[torch.Tensor] = None
        self.bytes_pulled = 0

    def all_reduce_small_(self, lo: int, hi: int) -> None:
        """One-shot sum-all-reduce of a SMALL span (the critic's 4.5 k gradients): barrier, every rank copies every
        rank's span (its own included) into a staging matrix whose rows are in RANK order, barrier (nobody may overwrite
        its span while a peer still reads it), span = sum of the rows in that order -- all ranks compute bit-identical
        sums.  Two barriers, w-1 small peer copies on their own streams, one tiny kernel; stream-ordered on the current
        stream and capturable in a CUDA graph (no NCCL call, no host synchronisation)."""
        from . import ops
        n = hi - lo
        if n <= 0:
            return
        if n % 4 or lo % 4:
            raise ValueError("PeerExchange: spans must be multiples of 4 floats (16-byte copies)")
        cur = torch.cuda.current_stream(self.flat.device)
        if self.stage_small is None or self.stage_small.shape[1] < n:
            self.stage_small = torch.empty((self.world, (n + 7) // 8 * 8), device=self.flat.device, dtype=torch.float32)
        self.hdl.barrier(channel=0)
        self.stage_small[self.rank, :n].copy_(self.flat[lo:hi])
        for peer, st in zip(self.peers, self.streams):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                self.stage_small[peer, :n].copy_(self.hdl.get_buffer(peer, (n,), torch.float32, lo))
            self.bytes_pulled += 4 * n
        for st in self.streams:
            cur.wait_stream(st)
        self.hdl.barrier(channel=1)
        ops.sum_rows(self.flat[lo:hi], self.stage_small[:, :n])

    def chunks(self, lo: int, hi: int) -> List[Tuple[int, int]]:
        n = hi - lo
        c = ((n + self.world - 1) // self.world + 7) // 8 * 8
        return [(min(hi, lo + r * c), min(hi, lo + (r + 1) * c)) for r in range(self.world)]

    def reduce_scatter_(self, lo: int, hi: int) -> List[Tuple[int, int]]:
        """First half of `all_reduce_`: afterwards flat[chunk of this rank] holds the sum over ranks (the other chunks keep this
        rank's own values).  Returns the chunk list (index = rank)."""
        from . import ops
        if (hi - lo) % 4 or lo % 4:
            raise ValueError("PeerExchange: spans must be multiples of 4 floats (16-byte copies)")
        cur = torch.cuda.current_stream(self.flat.device)
        ch = self.chunks(lo, hi)
        mlo, mhi = ch[self.rank]
        width = max(e - b for b, e in ch)
        if self.stage is None or self.stage.shape[1] < width:
            self.stage = torch.empty((len(self.peers), width), device=self.flat.device, dtype=torch.float32)
        self.hdl.barrier(channel=0)
        if mhi > mlo:
            for k, (peer, st) in enumerate(zip(self.peers, self.streams)):
                st.wait_stream(cur)
                with torch.cuda.stream(st):
                    self.stage[k, :mhi - mlo].copy_(self.hdl.get_buffer(peer, (mhi - mlo,), torch.float32, mlo))
                self.bytes_pulled += 4 * (mhi - mlo)
            for st in self.streams:
                cur.wait_stream(st)
            ops.sum_into(self.flat[mlo:mhi], self.stage[:, :mhi - mlo])
        return ch

    def all_gather_(self, ch: List[Tuple[int, int]], bufs=None) -> None:
        """Second half: every rank pulls chunk k of each buffer from rank k.  `bufs` = [(local tensor, symmetric-memory handle)],
        all indexed like the gradient buffer; default = the gradient buffer itself (all-reduce).  Barrier before (every chunk
        is final on its owner) and after (nobody still reads this rank's chunk when its owner next overwrites it)."""
        cur = torch.cuda.current_stream(self.flat.device)
        if bufs is None:
            bufs = [(self.flat, self.hdl)]
        self.hdl.barrier(channel=1)
        for peer, st in zip(self.peers, self.streams):
            b, e = ch[peer]
            if e <= b:
                continue
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                for t, h in bufs:
                    t[b:e].copy_(h.get_buffer(peer, (e - b,), t.dtype, b))
                    self.bytes_pulled += t.element_size() * (e - b)
        for st in self.streams:
            cur.wait_stream(st)
        self.hdl.barrier(channel=2)

    def all_reduce_(self, lo: int, hi: int) -> None:
        if hi <= lo:
            return
        self.all_gather_(self.reduce_scatter_(lo, hi))


PEER_MIN = 1 << 20      # spans below 4 MB take the one-shot path (every rank pulls every rank's whole span): latency bound


class GradExchange:
    """Sum-all-reduce of spans of a flat gradient buffer, overlappable with the kernels that follow.

    ``start(lo, hi)`` launches the reduction of ``flat[lo:hi]``: on CUDA it runs on a side stream that first waits
    for everything already enqueued on the current stream (the kernels that produced the span), so later kernels
    on the current stream overlap it; on CPU (gloo) it is an async work item.  ``finish()`` makes the current
    stream (or the host) wait for all started reductions.  With world size 1 both are no-ops.
    """

    def __init__(self, flat: torch.Tensor, group=None, side_stream: bool = False, peer: Optional["PeerExchange"] = None):
        self.flat = flat
        self.group = group
        self.peer = peer            # spans of >= PEER_MIN elements go through the copy engines instead of NCCL
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
                if self.world > 1 and self.peer is not None:
                    self.bytes_reduced += buf.numel() * buf.element_size()
                    # stream-ordered on the side stream: nothing to wait for
                    if hi - lo >= PEER_MIN:
                        self.peer.all_reduce_(lo, hi)
                    else:
                        self.peer.all_reduce_small_(lo, hi)
                elif self.world > 1:
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

    def start_sharded(self, lo: int, hi: int, update, bufs) -> None:
        """Sharded-optimizer variant of ``start(lo, hi, then=update)`` for the peer exchange (ZeRO-1 over NVLink peer memory):
        reduce-scatter flat[lo:hi], call ``update(b, e)`` for THIS rank's chunk only (its Adam step: parameters, moments and
        the bf16 operand copy of [b, e)), then all-gather the UPDATED buffers `bufs` = [(tensor, symmetric handle)] instead
        of the reduced gradient.  Every rank ends with identical parameters, and an optimizer pass over 1/world of the
        span instead of all of it -- on the side stream like `start`.  The reduced gradient exists only chunk-wise on its
        owners afterwards (PCAATrainer.reduced_gradient assembles it for diagnostics)."""
        if hi <= lo:
            return
        if self.world == 1 or self.peer is None or self._stream is None:
            raise RuntimeError("start_sharded needs the peer exchange on more than one rank")
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self._stream):
            self._stream.wait_event(ev)
            self.bytes_reduced += 4 * (hi - lo)
            if _SKIP_EXCHANGE:                          # timing probe only (wrong numerics)
                b, e = self.peer.chunks(lo, hi)[self.rank]
                update(b, e)
                return
            ch = self.peer.reduce_scatter_(lo, hi)
            b, e = ch[self.rank]
            if e > b:
                update(b, e)
            self.peer.all_gather_(ch, bufs)

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


def _checksum(t: torch.Tensor) -> torch.Tensor:
    """Two int64 checksums of the bit pattern of an fp32 tensor (plain and position-weighted sums of its int32 view)."""
    v = t.view(torch.int32).to(torch.int64)
    w = (torch.arange(v.numel(), device=t.device, dtype=torch.int64) % 65521) + 1
    return torch.stack([v.sum(), (v * w).sum()])


def per_tensor_relnorm(flat, a: torch.Tensor, b: torch.Tensor, floor: float = 1e-12):
    """max over the tensors of a flat layout (train._Flat) of ||a_t - b_t|| / ||b_t||, and the name where it is reached;
    tensors whose reference norm is below `floor` of the whole buffer's norm (identically-zero gradients) are skipped."""
    worst, where = 0.0, ""
    total = float(b.norm())
    for name in flat.names:
        ta, tb = flat.view(a, name), flat.view(b, name)
        nb = float(tb.norm())
        if nb <= floor * max(total, 1e-30):
            continue
        e = float((ta - tb).norm()) / nb
        if e > worst:
            worst, where = e, name
    return worst, where


@torch.no_grad()
def graphed_step_parity(trainer, inputs, make_single, group=None) -> dict:
    """Data-parallel parity of ONE iteration run the way the benchmark runs it (`trainer.step_graphed`: split CUDA graphs
    + the gradient exchange actually selected), checked on every rank against a single-device emulation of the same
    global iteration.  Collective: call on all ranks with this rank's `inputs` = (pcs, gt, z0, alphas).

    Emulation (trainer built by `make_single()` with process_group=dp.SINGLE, stepped phase by phase from a snapshot of
    this trainer's state): pass A runs encoder forward + critic gradients on every rank's shard and sums the critic
    gradients; pass B repeats the forward per shard, installs the MEAN critic gradient (what the exchange + Adam give every
    rank), runs the generator forward / backward and sums the generator gradients; pass B is run TWICE, the deviation between
    its two results is the run-to-run noise of the kernels themselves (atomically accumulated statistics and split-K sums
    change the summation order, and a last-bit difference can flip the bf16 rounding of a stored activation).  Reported:
      rel_g, rel_d   ||g_dp - sum_shards g|| / ||sum_shards g||  (generator / critic flat gradients)
      worst_tensor   the same per parameter tensor, maximum (a wrong factor on a small tensor cannot hide in the flat norm)
      noise_g, noise_tensor   the two emulation runs against each other: the floor the figures above are judged against
      frac_p_off, max_dp     generator weights vs Adam(snapshot, mean gradient): fraction off by > 2e-6, largest deviation
      replicas_identical     every rank holds bit-identical generator and critic weights after the iteration
    Verdict `ok`: replicas identical, rel_g <= max(2e-3, 4 noise_g), worst_tensor <= max(5e-2, 8 noise_tensor) (a maximum over
    ~100 tensors of a noise-limited quantity), rel_d <= 1e-3, max_dp <= 2.02 lr (an Adam step is at most ~lr; entries whose
    gradient is noise may step the other way).  A wrong exchange (a missing rank, a wrong scale, a stale chunk) moves rel_g
    and worst_tensor to O(1).
    Returns the dict (same on all ranks)."""
    rank, world = world_info(group)
    dev = trainer.dev
    if hasattr(trainer, "sync_optimizer_state"):
        trainer.sync_optimizer_state()                                 # sharded optimizer: moments live on their chunk's owner
    snap = trainer.snapshot()
    trainer.step_graphed(*inputs)
    torch.cuda.synchronize(dev)
    g_dp = trainer.reduced_gradient() if hasattr(trainer, "reduced_gradient") else trainer.G.g.clone()
    d_dp = trainer.D.g.clone()
    gathered = []
    for t in inputs:
        outs = [torch.empty_like(t) for _ in range(world)]
        if world > 1:
            dist.all_gather(outs, t.contiguous(), group=group)
        else:
            outs = [t]
        gathered.append(outs)
    ref = make_single()
    if ref.world != 1:
        raise RuntimeError("graphed_step_parity: make_single() must build the trainer with process_group=dp.SINGLE")
    dsum = torch.zeros_like(d_dp)
    for r in range(world):                                             # pass A: critic gradients of every shard
        ref.restore(snap)
        phases, _ = ref._phases(*[g[r] for g in gathered])
        phases[0][1]()
        dsum += ref.D.g
    gsums = []
    for _ in range(2):                                                 # pass B (twice): generator gradients under the mean critic step
        gsum = torch.zeros_like(g_dp)
        for r in range(world):
            ref.restore(snap)
            phases, _ = ref._phases(*[g[r] for g in gathered])
            phases[0][1]()
            ref.D.g.copy_(dsum / world)
            for _, fn in phases[1:]:                                   # (the exchange phases are no-ops on one rank)
                fn()
            gsum += ref.G.g
        gsums.append(gsum)
    torch.cuda.synchronize(dev)
    gsum = gsums[0]
    rel_g = float((g_dp - gsum).norm() / gsum.norm())
    rel_d = float((d_dp - dsum).norm() / dsum.norm())
    noise_g = float((gsums[1] - gsum).norm() / gsum.norm())
    worst_t, worst_name = per_tensor_relnorm(trainer.G, g_dp, gsum)
    noise_t, _ = per_tensor_relnorm(trainer.G, gsums[1], gsum)
    # weights: Adam of the mean gradient from the snapshot, through the same fused kernel
    ref.restore(snap)
    ref.G.g.copy_(gsum / world)
    from . import ops
    cfg = ref.cfg
    b2_g = cfg.get("B2_G", cfg["B2"])
    ops.adam_advance(ref.G.step_dev, ref.G.coef_dev, cfg["LR"], cfg["B1"], b2_g)
    ops.adam_flat_dev(ref.G.p, ref.G.g, ref.G.m, ref.G.v, cfg["B1"], b2_g, 1e-8, ref.G.coef_dev, 1.0, ref.G.shadow)
    dp_abs = (trainer.G.p - ref.G.p).abs()
    frac_p_off, max_dp = float((dp_abs > 2e-6).float().mean()), float(dp_abs.max())
    cs = torch.cat([_checksum(trainer.G.p), _checksum(trainer.D.p)])
    allcs = [torch.empty_like(cs) for _ in range(world)]
    if world > 1:
        dist.all_gather(allcs, cs, group=group)
    else:
        allcs = [cs]
    same = all(torch.equal(allcs[0], c) for c in allcs)
    res = torch.tensor([rel_g, rel_d, frac_p_off, max_dp, noise_g, worst_t, noise_t], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(res, op=dist.ReduceOp.MAX, group=group)
    rel_g, rel_d, frac_p_off, max_dp, noise_g, worst_t, noise_t = (float(x) for x in res)
    peer = getattr(trainer.G, "peer", None)
    lr = float(cfg["LR"])
    out = {"rel_g": rel_g, "rel_d": rel_d, "worst_tensor": worst_t, "worst_tensor_name": worst_name, "noise_g": noise_g,
           "noise_tensor": noise_t, "frac_p_off": frac_p_off, "max_dp_over_lr": max_dp / lr, "replicas_identical": bool(same),
           "exchange": ("peer, sharded decoder update" if getattr(trainer, "shard_adam", False) else "peer") if peer is not None
                       else ("nccl" if world > 1 else "none"), "world": world,
           "path": "step_graphed (split graphs)" if trainer.split_graphs else "step_graphed (one graph)",
           "ok": bool(same and rel_g <= max(2e-3, 4 * noise_g) and worst_t <= max(5e-2, 8 * noise_t) and rel_d <= 1e-3
                      and max_dp <= 2.02 * lr)}
    del ref
    torch.cuda.empty_cache()
    return out
