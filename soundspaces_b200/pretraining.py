"""Batched GPU version of the SAVi pre-training dataset's audio path
(``ss_baselines/savi/pretraining/audiogoal_dataset.py:100-156``): random ``(rir, sound, index)`` triples ->
spectrogram batch, with the reference's quirk preserved: in the steady-state branch the 1-s slice starts
one sample EARLIER than the simulator's (``source[index*sr - L :]`` + ``mode='valid'`` + drop the last
sample  =>  ``y_full[index*sr - 1 : (index+1)*sr - 1]``)."""
from __future__ import annotations

import random
from typing import Sequence

import torch

from .renderer import AudioRequest, BatchedAudioRenderer


def dataset_request(renderer: BatchedAudioRenderer, rir_id: int, source_id: int, index: int) -> AudioRequest:
    """The request equivalent to ``compute_audiogoal`` (audiogoal_dataset.py:114-140) for a given index."""
    sr = renderer.sr
    taps = renderer._rir_len[rir_id] if rir_id >= 0 else 0
    if taps == 0:
        taps = sr                                   # unreadable / empty file -> zeros((sr, 2)) (:118-123)
    if index * sr - taps < 0:
        return AudioRequest(rir=rir_id, source=source_id, offset=index * sr)
    return AudioRequest(rir=rir_id, source=source_id, offset=index * sr - 1)


class BatchedAudioGoalDataset:
    """``files``: sequence of ``(rir_id, source_id)``; clips are registered in the renderer."""

    def __init__(self, renderer: BatchedAudioRenderer, files: Sequence, rng: random.Random | None = None):
        self.r, self.files, self.rng = renderer, list(files), rng or random.Random()

    def __len__(self):
        return len(self.files)

    def audio_length(self, source_id: int) -> int:
        return self.r._sources[source_id].numel() // self.r.sr            # audiogoal_dataset.py:89-90

    def render(self, items: Sequence[int], indices: Sequence[int] | None = None, want_wave: bool = False):
        reqs = []
        for k, item in enumerate(items):
            rir_id, source_id = self.files[item]
            index = indices[k] if indices is not None else self.rng.randint(0, self.audio_length(source_id) - 2)
            reqs.append(dataset_request(self.r, rir_id, source_id, index))
        return self.r.render(reqs, want_wave=want_wave)
