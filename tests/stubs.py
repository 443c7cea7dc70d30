"""Stand-ins for the CPU host-logic tests: a renderer without CUDA that records what it was asked to render."""
import contextlib

import numpy as np
import torch


class StubRenderer:
    """Implements the slice of BatchedAudioRenderer the service / batcher touch.  A "spectrogram" row is filled
    with ``1000 * rir_id + offset_seconds`` so that tests can tell rows apart; silent requests give zeros."""

    def __init__(self, sr=16000, spec_shape=(65, 26, 2)):
        self.sr, self.spec_shape, self.device = sr, tuple(spec_shape), torch.device("cpu")
        self._rir_len, self._rir_off, self._data = [], [], []
        self.renders = []                      # list of lists of requests, one per execute()
        self.compactions = 0

    # bank
    def add_rirs(self, rirs):
        ids = []
        for r in rirs:
            self._rir_off.append(sum(self._rir_len))
            self._rir_len.append(0 if r is None else len(r))
            self._data.append(None if r is None else np.asarray(r))
            ids.append(len(self._rir_len) - 1)
        return ids

    @property
    def bank_bytes(self):
        return 8 * sum(self._rir_len)

    def compact_bank(self, keep):
        self._data = [self._data[i] for i in keep]
        self._rir_len = [self._rir_len[i] for i in keep]
        self._rir_off = list(np.cumsum([0] + self._rir_len[:-1])) if keep else []
        self.compactions += 1
        return list(range(len(keep)))

    def add_source(self, samples):
        return 0

    # render
    def prepare(self, reqs):
        return list(reqs)

    def execute(self, batch, out=None, want_wave=False, channels_first=False):
        self.renders.append(list(batch))
        res = out if out is not None else torch.empty((len(batch),) + self.spec_shape)
        for i, q in enumerate(batch):
            res[i] = 0.0 if q.silent else float(1000 * q.rir + q.offset // self.sr)
        return res

    def render_crossfade(self, cur, prev, want_wave=False):
        return self.execute(cur)

    @contextlib.contextmanager
    def transient_windows(self):
        yield

    @contextlib.contextmanager
    def _inline_rirs(self, *lists):
        yield
