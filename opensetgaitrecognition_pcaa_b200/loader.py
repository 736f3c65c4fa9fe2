"""Crop store and batch feeding for the PCAA hot path (the step before the path).

The reference feeds the networks through ``MSRadarDataset`` (``datasets.py:381-479``): one ``np.load`` of a float64
``(30, nmax, 4)`` crop per sample, cast to fp32, permuted to ``(4, 30, nmax)``, collated by a ``DataLoader`` with
``num_workers=0`` -- a few hundred crops per second, three orders of magnitude below what the B200 step consumes.
Here the split is read ONCE into a packed, pinned ``(M, 4, 30, nmax)`` fp32 store (same values bit for bit:
float64 -> float32 cast, then the permutation) and batches come from it in one of two ways:

* ``PackedCrops.batches`` + ``DevicePrefetcher``: pinned host batches, host->device copies on a copy stream one
  batch ahead of the consumer (what ``bench.py``'s end-to-end number measures);
* ``PackedCrops.to_device`` + ``DeviceCrops.batch``: the whole store resident in HBM (mmGait10 is a few GB, the GPU
  has 180), batches assembled by the ``pcaa_gather_rows`` kernel from a device index vector -- no per-step H2D.

File naming, label mapping and the sequential (track-ordered) listing follow the reference exactly
(``datasets.py:62-76, 163-180, 392-462``).
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops

SCENARIOS = ("free_walk", "hands_in_pockets", "smartphone")        # constants.py:13-16


# ---- file-name parsers (datasets.py:62-76)
def filename2crop(filename: str) -> int:
    return int(filename.split("_")[0][4:])


def filename2subj(filename: str) -> int:
    return int(filename.split("_")[1][4:])


def filename2track(filename: str) -> str:
    return filename.split("_")[-1][5:].split(".")[0]


def filename2scenario(filename: str) -> str:
    return str.join("_", filename.split("_")[2:-1])


def sorted_seq(all_files: Sequence[str], subject_id, track_id) -> List[str]:
    """Crops of one (subject, track) in crop order; substring matching as in datasets.py:163-180."""
    sel = [f for f in all_files if f"subj{subject_id}" in f]
    sel = [c for c in sel if f"track{track_id}" in c]
    ids = np.array([filename2crop(c) for c in sel])
    return [sel[i] for i in np.argsort(ids)]


def list_split(dataset_dir: str, scenarios: Sequence[str] = SCENARIOS, sequential: bool = False) -> List[str]:
    """The file list MSRadarDataset.__init__ builds (datasets.py:392-423): directory order, or -- sequential -- the
    crops of every (subject, track) consecutively in crop order; then the scenario filter."""
    all_crops = os.listdir(dataset_dir)
    if sequential:
        track_dict = {}
        for crop in all_crops:
            track_dict.setdefault(filename2subj(crop), set()).add(filename2track(crop))
        names: List[str] = []
        for subj in track_dict.keys():
            for track in track_dict[subj]:
                names.extend(sorted_seq(all_crops, subj, track))
    else:
        names = list(all_crops)
    return [f for f in names if filename2scenario(f) in scenarios]


def labels_of(filenames: Sequence[str]) -> Tuple[np.ndarray, np.ndarray]:
    """(labels 0..n_classes-1, original subject ids): the mapping of datasets.py:425-462 (enumerate(list(set(...))))."""
    original = [filename2subj(f) for f in filenames]
    selected = list(set(original))
    lab = {c: i for i, c in enumerate(selected)}
    return np.array([lab[j] for j in original], dtype=np.int64), np.array(original, dtype=np.int64)


class PackedCrops:
    """A split as one pinned fp32 tensor ``pcs (M, 4, 30, nmax)`` + ``labels (M,) int64`` (+ file names, subjects)."""

    def __init__(self, pcs: torch.Tensor, labels: torch.Tensor, filenames: Optional[List[str]] = None,
                 subjects: Optional[np.ndarray] = None, pin: bool = True):
        if pcs.dim() != 4 or pcs.shape[1] != 4 or pcs.dtype != torch.float32:
            raise ValueError("PackedCrops: pcs must be (M, 4, T, nmax) float32")
        if labels.shape != (pcs.shape[0],) or labels.dtype != torch.int64:
            raise ValueError("PackedCrops: labels must be (M,) int64")
        pin = pin and torch.cuda.is_available()
        self.pcs = pcs.contiguous().pin_memory() if pin else pcs.contiguous()
        self.labels = labels.contiguous().pin_memory() if pin else labels.contiguous()
        self.filenames = filenames
        self.subjects = subjects

    @classmethod
    def from_directory(cls, dataset_dir: str, scenarios: Sequence[str] = SCENARIOS, sequential: bool = False,
                       workers: int = 8, pin: bool = True) -> "PackedCrops":
        names = list_split(dataset_dir, scenarios, sequential)
        if not names:
            raise FileNotFoundError(f"no crops of scenarios {tuple(scenarios)} under {dataset_dir}")
        first = np.load(os.path.join(dataset_dir, names[0]), allow_pickle=True)
        T, nmax, F = first.shape
        if F != 4:
            raise ValueError(f"{names[0]}: expected (T, nmax, 4) crops, got {first.shape}")
        pcs = torch.empty((len(names), 4, T, nmax), dtype=torch.float32)
        out = pcs.numpy()

        def load(i):
            a = first if i == 0 else np.load(os.path.join(dataset_dir, names[i]), allow_pickle=True)
            if a.shape != (T, nmax, 4):
                raise ValueError(f"{names[i]}: shape {a.shape} differs from {(T, nmax, 4)}")
            # __getitem__: torch.from_numpy(a).to(torch.float).permute(2, 0, 1)   (datasets.py:467-472)
            out[i] = np.transpose(a.astype(np.float32), (2, 0, 1))

        with ThreadPoolExecutor(max_workers=max(1, workers)) as ex:
            list(ex.map(load, range(len(names))))
        labels, subjects = labels_of(names)
        return cls(pcs, torch.from_numpy(labels), names, subjects, pin)

    def __len__(self) -> int:
        return self.pcs.shape[0]

    def __getitem__(self, idx: int):
        """(pc_seq (4, T, nmax) fp32, label int64 scalar) -- what MSRadarDataset.__getitem__ returns."""
        return self.pcs[idx], self.labels[idx]

    def batches(self, batch_size: int, shuffle: bool = False, drop_last: bool = False,
                generator: Optional[torch.Generator] = None) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        """Host batches in DataLoader order (sequential, or a torch.randperm drawn from `generator`).  Unshuffled
        batches are views of the pinned store; shuffled ones are gathered into two alternating pinned staging
        buffers: a yielded batch stays valid until the next-but-one is requested (DevicePrefetcher waits for its
        asynchronous copy of a batch before it asks for the batch two later)."""
        M = len(self)
        order = torch.randperm(M, generator=generator) if shuffle else None
        stage = None
        if shuffle:
            pin = self.pcs.is_pinned()
            mk = (lambda *s, dtype: torch.empty(*s, dtype=dtype).pin_memory()) if pin else (lambda *s, dtype: torch.empty(*s, dtype=dtype))
            stage = [(mk((batch_size,) + tuple(self.pcs.shape[1:]), dtype=torch.float32), mk((batch_size,), dtype=torch.int64))
                     for _ in range(2)]
        nb = M // batch_size if drop_last else (M + batch_size - 1) // batch_size
        for b in range(nb):
            lo, hi = b * batch_size, min(M, (b + 1) * batch_size)
            if order is None:
                yield self.pcs[lo:hi], self.labels[lo:hi]
            else:
                sp, sl = stage[b % 2]
                ix = order[lo:hi]
                torch.index_select(self.pcs, 0, ix, out=sp[:hi - lo])
                torch.index_select(self.labels, 0, ix, out=sl[:hi - lo])
                yield sp[:hi - lo], sl[:hi - lo]

    def to_device(self, device="cuda") -> "DeviceCrops":
        return DeviceCrops(self.pcs.to(device, non_blocking=True), self.labels.to(device, non_blocking=True))


class DeviceCrops:
    """The packed store resident in HBM; batches are gathered on the device (pcaa_gather_rows)."""

    def __init__(self, pcs: torch.Tensor, labels: torch.Tensor):
        if not pcs.is_cuda:
            raise RuntimeError("DeviceCrops needs CUDA tensors (no CPU fallback)")
        self.pcs, self.labels = pcs, labels
        self._lab2d = labels.view(-1, 1)

    def __len__(self) -> int:
        return self.pcs.shape[0]

    def batch(self, idx: torch.Tensor, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
        """(pcs[idx], labels[idx]) for a device int64 index vector, written into `out` when given (e.g. the static
        input buffers of a captured train step)."""
        idx = idx.to(self.pcs.device, torch.int64).contiguous()
        pcs = ops.gather_rows(self.pcs, idx, None if out is None else out[0])
        lab = self.labels.index_select(0, idx) if out is None else torch.index_select(self.labels, 0, idx, out=out[1])
        return pcs, lab

    def epoch(self, batch_size: int, shuffle: bool = True, drop_last: bool = True,
              generator: Optional[torch.Generator] = None) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        M = len(self)
        order = (torch.randperm(M, generator=generator) if shuffle else torch.arange(M)).to(self.pcs.device)
        nb = M // batch_size if drop_last else (M + batch_size - 1) // batch_size
        for b in range(nb):
            yield self.batch(order[b * batch_size:(b + 1) * batch_size])


_COPY_STREAMS = {}


def _copy_stream(dev: torch.device) -> torch.cuda.Stream:
    """One copy stream per device, shared by all prefetchers: the caching allocator pools blocks per stream, so the
    device slots a finished prefetcher frees are reused by the next one instead of being cudaMalloc'ed again."""
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _COPY_STREAMS[key]


class DevicePrefetcher:
    """Iterate device copies of pinned host batches, copying one batch ahead on a dedicated stream.

    `batches` yields tuples of (pinned) host tensors.  `depth` device slots rotate; a slot is reused only after the
    consumer stream passed the point where it asked for the next-but-`depth-1` batch, i.e. a yielded batch is valid
    until `depth - 1` further batches have been requested.  The consumer stream waits (on the device, not the host)
    for the copy of the batch it receives."""

    def __init__(self, batches: Iterable[Sequence[torch.Tensor]], device, depth: int = 2):
        self.it = iter(batches)
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("DevicePrefetcher copies to a CUDA device (no CPU fallback)")
        self.depth = max(2, depth)
        self.stream = _copy_stream(self.dev)
        self.slots: List[Optional[Tuple[torch.Tensor, ...]]] = [None] * self.depth
        self.ready = [torch.cuda.Event() for _ in range(self.depth)]
        self.free = [None] * self.depth
        self.n_issued = 0
        self.h2d_bytes = 0
        self._queue: List[int] = []
        self._fill()

    def _issue(self) -> bool:
        s = self.n_issued % self.depth
        if self.n_issued >= self.depth:
            # the producer may reuse a host staging buffer `depth` batches later (PackedCrops.batches alternates two):
            # the asynchronous copy issued `depth` batches ago must have left the host buffer before the producer runs
            self.ready[s].synchronize()
        try:
            host = next(self.it)
        except StopIteration:
            return False
        slot = self.slots[s]
        with torch.cuda.stream(self.stream):
            if slot is None or any(d.shape != h.shape or d.dtype != h.dtype for d, h in zip(slot, host)):
                # allocated ON the copy stream: the caching allocator never hands a block to another stream than the one
                # it was allocated on, so the copies below cannot land in memory that kernels still queued on the
                # consumer stream are using (freed-but-in-flight temporaries of the previous step)
                slot = tuple(torch.empty(h.shape, dtype=h.dtype, device=self.dev) for h in host)
                self.slots[s] = slot
            if self.free[s] is not None:
                self.stream.wait_event(self.free[s])          # the consumer is done with this slot
            for d, h in zip(slot, host):
                d.copy_(h, non_blocking=True)
                self.h2d_bytes += h.numel() * h.element_size()
            self.ready[s].record(self.stream)
        self._queue.append(s)
        self.n_issued += 1
        return True

    def _fill(self):
        while len(self._queue) < self.depth - 1 and self._issue():
            pass

    def __iter__(self):
        return self

    def __next__(self) -> Tuple[torch.Tensor, ...]:
        if not self._queue:
            raise StopIteration
        s = self._queue.pop(0)
        cur = torch.cuda.current_stream(self.dev)
        # everything the consumer enqueued so far used older slots: the slot that the next copy will overwrite is free
        # once the consumer stream reaches this point
        nxt = self.n_issued % self.depth
        ev = torch.cuda.Event()
        ev.record(cur)
        self.free[nxt] = ev
        self._issue()
        cur.wait_event(self.ready[s])
        for t in self.slots[s]:
            t.record_stream(cur)          # used on the consumer stream: its block is not recycled before that work is done
        return self.slots[s]
