"""Batching behind the per-env Habitat sensor API (SURVEY.md 7 "Batching across a per-env API", 8(b)).

Habitat calls ``Sensor.get_observation`` once per env (``ss_baselines/common/sync_vector_env.py:186-199``); the
reference renders inside that call, one env at a time.  Here the call only ENQUEUES a request and returns a
:class:`DeferredObservation` -- a ``(slot, generation)`` handle into a device ring owned by the
:class:`AudioObservationBatcher` -- and the whole step's requests are rendered by ONE ``render`` when the first
handle is resolved, normally by ``sensors.batch_obs`` (the replacement of ``ss_baselines/common/utils.py:126-153``),
which returns the ``(N, 65, T', 2)`` CUDA view of the ring without a copy.

Ownership: the ring belongs to the batcher; a row stays valid until its slot is reused, which
happens after ``capacity`` further rows.  Handles that are still referenced at that point (the memoised
observations in a simulator's ``_spectrogram_cache``, simulator.py:696-699) are detached first -- they take a
private copy of their row -- so a handle NEVER shows another observation's data.
"""
from __future__ import annotations

import weakref
from typing import List, Optional, Sequence

import numpy as np
import torch

from .renderer import AudioRequest


class DeferredObservation:
    """Handle to one env's ``(65, T', 2)`` spectrogram: pending until its batcher flushes, then a device row.

    ``resolve()`` gives the CUDA tensor; ``np.asarray(handle)`` / ``handle.numpy()`` give the host array the
    reference's sensor would have returned (compat: this forces a flush and one device->host copy); pickling
    (multiprocess ``habitat.VectorEnv`` workers) materialises the host array the same way."""

    __slots__ = ("_batcher", "slot", "generation", "_row", "shape", "__weakref__")
    dtype = np.dtype(np.float32)

    def __init__(self, batcher, slot: int, generation: int, shape):
        self._batcher, self.slot, self.generation, self._row, self.shape = batcher, slot, generation, None, tuple(shape)

    @property
    def pending(self) -> bool:
        return self._row is None and self.generation > self._batcher.flushed_generation

    def resolve(self) -> torch.Tensor:
        if self._row is not None:                      # detached (private copy)
            return self._row
        b = self._batcher
        if self.generation > b.flushed_generation:
            b.flush()
        return b.ring[self.slot]

    def detach(self):
        """Take a private copy of the row (called by the batcher before the slot is reused)."""
        if self._row is None:
            self._row = self.resolve().clone()
        return self

    def numpy(self) -> np.ndarray:
        return self.resolve().cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None and a.dtype != dtype else a

    def __reduce__(self):
        return (np.asarray, (self.numpy(),))


class AudioObservationBatcher:
    """Queue of one step's audio requests + the device ring their spectrograms are rendered into."""

    def __init__(self, service, capacity: int = 1024):
        self.service = service
        self.renderer = service.renderer
        self.capacity = int(capacity)
        self.ring: Optional[torch.Tensor] = None       # (capacity, 65, T', 2), allocated on first flush
        self._owners: List[Optional[weakref.ref]] = [None] * self.capacity
        self._pending: List[tuple] = []                # (handle, request, crossfade_from)
        self._next = 0                                 # next free slot
        self.generation = 1                            # generation of the handles being enqueued now
        self.flushed_generation = 0
        self.flushes = 0

    # ------------------------------------------------------------------ enqueue
    def enqueue(self, request: AudioRequest, crossfade_from: Optional[AudioRequest] = None) -> DeferredObservation:
        if len(self._pending) >= self.capacity:
            self.flush()
        if self._next + len(self._pending) >= self.capacity:       # keep a step's rows contiguous: wrap to slot 0
            if self._pending:
                self.flush()
            self._next = 0
        slot = self._next + len(self._pending)
        old = self._owners[slot]
        if old is not None:
            h = old()
            if h is not None and h._row is None:
                h.detach()                              # still referenced (memoised): give it its own copy first
        handle = DeferredObservation(self, slot, self.generation, self.renderer.spec_shape)
        self._owners[slot] = weakref.ref(handle)
        self._pending.append((handle, request, crossfade_from))
        return handle

    # -------------------------------------------------------------------- flush
    def flush(self, out: Optional[torch.Tensor] = None):
        """Render every pending request with ONE launch chain.  ``out``: render straight into this
        ``(n_pending, 65, T', 2)`` CUDA tensor instead of the ring (rollout-storage slot); the ring
        rows are then filled by one device copy so that the handles stay valid."""
        if not self._pending:
            return None
        r = self.renderer
        if self.ring is None:
            self.ring = torch.zeros((self.capacity,) + r.spec_shape, dtype=torch.float32, device=r.device)
        pending, self._pending = self._pending, []
        n = len(pending)
        s0 = pending[0][0].slot
        rows = self.ring[s0: s0 + n]
        target = out if out is not None else rows
        reqs = [p[1] for p in pending]
        fades = [p[2] for p in pending]
        if any(f is not None for f in fades):          # continuous simulator with CROSSFADE (continuous_simulator.py:422-424)
            with r.transient_windows():
                target.copy_(r.render_crossfade(reqs, fades))
        elif any(q.rir_array is not None or q.wrap or q.offset % r.sr for q in reqs):
            with r.transient_windows():                # per-step RIRs / sample offsets: nothing worth keeping in the pool
                with r._inline_rirs(reqs):
                    r.execute(r.prepare(reqs), out=target)
        else:
            r.execute(r.prepare(reqs), out=target)
        if out is not None:
            rows.copy_(out)
        # between steps: LRU clock, landed prefetches, bank compaction (the requests above already hold their bank
        # ids, so this must not run before they have been turned into launches)
        self.service.maybe_trim()
        self._next = s0 + n
        self.flushed_generation = self.generation
        self.generation += 1
        self.flushes += 1
        return target

    # ------------------------------------------------------------------ resolve
    def gather(self, handles: Sequence[DeferredObservation], out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``(N, 65, T', 2)`` CUDA tensor of the given handles' rows: flushes once if needed; when the handles are
        the consecutive rows of one flush (the normal step) the result is a VIEW of the ring -- no copy at all."""
        n = len(handles)
        if n == 0:
            return torch.empty((0,) + self.renderer.spec_shape, dtype=torch.float32, device=self.renderer.device)
        pend = [h for h in handles if h._row is None and h.generation > self.flushed_generation]
        if pend:
            all_pending_in_order = (out is not None and len(pend) == n == len(self._pending)
                                    and all(p[0] is h for p, h in zip(self._pending, handles)))
            self.flush(out=out if all_pending_in_order else None)
            if all_pending_in_order:
                return out
        if all(h._row is None for h in handles):
            s0 = handles[0].slot
            if all(h.slot == s0 + i for i, h in enumerate(handles)):
                view = self.ring[s0: s0 + n]
                return view if out is None else out.copy_(view)
            idx = torch.tensor([h.slot for h in handles], device=self.ring.device)
            res = self.ring.index_select(0, idx)
        else:
            res = torch.stack([h.resolve() for h in handles], dim=0)
        return res if out is None else out.copy_(res)
