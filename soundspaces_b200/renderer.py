"""Batched device-side audio renderer: the host half of the B200 audio path.

Owns the device-resident RIR bank, the source clips with their
cached overlap-save window spectra (the per-sound cache of
``soundspaces/simulator.py:595-600`` moved to the device), and the scratch the
kernels need.  One ``render()`` call replaces, for a whole batch of envs, one
cache-missing ``get_current_spectrogram_observation`` each
(``soundspaces/simulator.py:690-701`` -> ``_compute_audiogoal`` ``:608-666`` ->
``SpectrogramSensor.compute_spectrogram`` ``soundspaces/tasks/nav.py:86-100``).

PyTorch is used for device memory and streams only; all arithmetic runs in
``libssb200.so`` (hand-written sm_100a CUDA) through the C ABI.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from .planning import window_layout
from ._lib import PAD_MODES, REQ_DTYPE, SSB_FLAG_SILENT

N_FFT, HOP, WIN, POOL, SPEC_ROWS = 512, 160, 400, 4, 65


def spectrogram_shape(sr: int):
    """(65, ceil((1 + sr//160)/4), 2): (65, 26, 2) @16 kHz, (65, 69, 2) @44.1 kHz."""
    frames = 1 + sr // HOP
    return (SPEC_ROWS, -(-frames // POOL), 2)


@dataclass
class AudioRequest:
    """One env-step.  ``out[m] = sum_k rir[k] * src[offset + m - k]`` for m < out_samples.

    offset: 0 for 1-s clips (simulator.py:629-632); ``index * sr`` for multi-second clips
    (simulator.py:634-647, both the early and the ``valid`` branch); the running
    ``_current_sample_index`` for the continuous simulator (continuous_simulator.py:428-456).
    """
    rir: int                       # id in the RIR bank; -1 => unreadable/empty file => zero RIR (simulator.py:617-624)
    source: int                    # id of the source clip
    offset: int = 0
    out_samples: Optional[int] = None      # default sr; int(sr*STEP_TIME) for the continuous simulator
    wrap: bool = False             # continuous_simulator.py:443-445 wrap-around of the clip
    silent: bool = False           # simulator.py:610-612
    distractor_rir: Optional[int] = None   # simulator.py:649-664
    distractor_source: Optional[int] = None
    # transient RIR supplied inline ((L, 2) array; overrides ``rir``): the RIR habitat-sim renders every step
    # (simulator.py:626, continuous_simulator.py:419).  Uploaded for the one call and dropped afterwards.
    rir_array: Optional[object] = None


class _PoolReset(Exception):
    """The window-spectra pool was recycled while a batch was being prepared."""


@dataclass
class PreparedBatch:
    n: int
    reqs_host: np.ndarray
    reqs_dev: torch.Tensor
    plan: object = None            # the ssb_plan the requests were resolved for (partitioned or single-block)


def _dev_index(device):
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("BatchedAudioRenderer needs a CUDA device; there is no CPU fallback")
    return torch.cuda.current_device() if device.index is None else device.index


class WaveformOps:
    """Kernels that only need a context -- the STFT front end and the AV-WaN intensity reduction on waveforms the
    caller already holds -- without a convolution plan, RIR bank or window-spectra pool.  One per device; this is
    what ``SpectrogramSensor.compute_spectrogram`` (a static method other modules import and call on host arrays of
    any length, savi/ppo/ppo_trainer.py:52,374, belief_predictor.py:15,124) runs on."""

    _instances: dict = {}

    def __init__(self, device="cuda:0"):
        self.device = torch.device("cuda", _dev_index(device))
        self.ctx = _lib.Context(self.device.index)
        self.lib = self.ctx.lib

    @classmethod
    def get(cls, device="cuda:0") -> "WaveformOps":
        key = _dev_index(device)
        if key not in cls._instances:
            cls._instances[key] = cls(torch.device("cuda", key))
        return cls._instances[key]

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def spectrogram(self, wave: torch.Tensor, pad_mode: str = "reflect", out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(n, 2, S) CUDA float32 -> (n, 65, T', 2) for any S >= 512 (nav.py:86-100)."""
        if wave.ndim == 2:
            wave = wave[None]
        wave = wave.to(self.device, torch.float32).contiguous()
        n, _, S = wave.shape
        spec = out if out is not None else torch.empty((n,) + spectrogram_shape(S), dtype=torch.float32, device=self.device)
        self.ctx.check(self.lib.ssb_spectrogram_batch(self.ctx.handle, n, wave.data_ptr(), S, S, PAD_MODES[pad_mode],
                                                      spec.data_ptr(), self._stream()), "ssb_spectrogram_batch")
        return spec

    def intensity(self, wave: torch.Tensor, num_frame: int = 150) -> torch.Tensor:
        """AV-WaN ``Intensity`` (avwan_sensors.py:91-100) for a (n, 2, S) CUDA batch -> (n,) mean squares."""
        if wave.ndim == 2:
            wave = wave[None]
        wave = wave.to(self.device, torch.float32).contiguous()
        out = torch.empty(wave.shape[0], dtype=torch.float32, device=self.device)
        self.ctx.check(self.lib.ssb_intensity_batch(self.ctx.handle, wave.shape[0], wave.data_ptr(), wave.shape[2],
                                                    wave.shape[2], int(num_frame), out.data_ptr(), self._stream()),
                       "ssb_intensity_batch")
        return out

    @torch.no_grad()
    def audio_conv1(self, spec: torch.Tensor, conv: "torch.nn.Conv2d", relu: bool = True) -> torch.Tensor:
        """``relu(conv(spec.permute(0, 3, 1, 2)))`` for the first layer of ``AudioCNN``
        (audio_cnn.py:51-58,86) straight from the (n, 65, T', 2) observation: no permuted copy, one launch.
        Inference only (rollout collection); ``conv`` must be an unpadded, undilated 2-input-channel ``Conv2d``."""
        if not (spec.is_cuda and spec.dtype == torch.float32 and spec.ndim == 4 and spec.shape[3] == 2 and spec.is_contiguous()):
            raise ValueError("expected a contiguous CUDA float32 (n, H, W, 2) observation")
        if (conv.in_channels != 2 or tuple(conv.padding) != (0, 0) or tuple(conv.dilation) != (1, 1) or conv.groups != 1
                or conv.padding_mode != "zeros"):
            raise ValueError("audio_conv1 fuses Conv2d(2 -> OC, k, stride) without padding / dilation / groups")
        n, H, W, _ = spec.shape
        (kh, kw), (sh, sw) = conv.kernel_size, conv.stride
        out = torch.empty((n, conv.out_channels, (H - kh) // sh + 1, (W - kw) // sw + 1), dtype=torch.float32, device=self.device)
        w = conv.weight.detach().to(self.device, torch.float32).contiguous()
        b = conv.bias.detach().to(self.device, torch.float32).contiguous() if conv.bias is not None else None
        self.ctx.check(self.lib.ssb_audio_conv1_batch(
            self.ctx.handle, n, spec.data_ptr(), H, W, w.data_ptr(), b.data_ptr() if b is not None else None,
            conv.out_channels, kh, kw, sh, sw, int(bool(relu)), out.data_ptr(), self._stream()), "ssb_audio_conv1_batch")
        return out


class BatchedAudioRenderer:
    def __init__(self, sr: int, max_taps: int, device="cuda:0", n_terms: int = 1, log2n: int = 0,
                 pad_mode: str = "reflect", xpool_bytes: int = 512 << 20, prefer_block64: Optional[bool] = None):
        self.sr = int(sr)
        self.device = torch.device("cuda", _dev_index(device))
        self.pad_mode = PAD_MODES[pad_mode]
        # No global side effect: the caller's current CUDA device is left alone.  Every C-ABI entry selects the
        # context's device itself (and restores the caller's); tensors are allocated with an explicit device and
        # streams are taken from torch's current stream OF THAT device.
        self.ctx = _lib.Context(self.device.index)
        self.lib = self.ctx.lib
        self.max_taps = int(max_taps)
        # Two plans (include/ssb200.h).  The partitioned overlap-save plan handles every request and is the FASTER one
        # on B200 (config 2: 0.091 ms per 128-env step); the single-block plan (one fused 65536-point cluster kernel per
        # env, no H / Y intermediates: 33 MB of DRAM traffic per step instead of 180 MB) is the LOW-TRAFFIC one
        # (0.118 ms) and serves batches whose requests all have one term and at most 65536 - sr + 1 effective taps.
        # prefer_block64=True (or SSB200_BLOCK64=1) takes it whenever a batch is eligible -- chosen per batch in
        # _prepare_columns; log2n = 16 forces it (requests that do not fit raise).  DESIGN.md section 4 has the numbers.
        self.plan64 = None
        self.block64_taps = 65536 - self.sr + 1
        if prefer_block64 is None:
            prefer_block64 = os.environ.get("SSB200_BLOCK64", "0") == "1"
        if (log2n == 16 or (log2n == 0 and prefer_block64)) and self.sr <= 65536 - 4096:
            self.plan64 = self.ctx.make_plan(self.sr, min(self.max_taps, self.block64_taps), 1, 16)
        self.force_block64 = log2n == 16
        if self.force_block64 and (self.plan64 is None or self.max_taps > self.block64_taps or n_terms != 1):
            raise ValueError(f"log2n=16 (single-block plan) needs n_terms=1 and max_taps <= {self.block64_taps} at sr={self.sr}")
        self.plan = self.ctx.make_plan(self.sr, self.max_taps, n_terms, 0 if log2n == 16 else log2n)
        self.N = 1 << self.plan.log2n
        self.P = self.plan.block
        self.spec_shape = spectrogram_shape(self.sr)
        assert self.spec_shape[1] == self.lib.ssb_spec_cols(self.sr)
        # RIR bank: (total_taps, 2) float32, interleaved ears == one complex signal
        self._bank = torch.zeros((1, 2), dtype=torch.float32, device=self.device)
        self._bank_used = 0
        self._rir_off: list[int] = []
        self._rir_len: list[int] = []
        self._bank_index = None               # numpy copies of the two lists, rebuilt when the bank changes
        # sources + window-spectra pool
        self._sources: list[torch.Tensor] = []
        # window-spectra pool: allocated on first use (a renderer that only runs the STFT / intensity kernels on
        # waveforms it is handed -- SpectrogramSensor.compute_spectrogram on a host array -- never needs it)
        self._xpool_elems = max(xpool_bytes // 8, 64 * self.N, 8 * 65536 if self.plan64 is not None else 0)   # float2 elements
        self._xpool_t = None
        self._xpool_used = 0                  # float2 elements
        self._xcache: dict = {}
        self._hscratch = None
        self._wave = None
        self._prev_wave = None
        self._hbank = None

    @property
    def _xpool(self) -> torch.Tensor:
        if self._xpool_t is None:
            self._xpool_t = torch.empty(self._xpool_elems * 2, dtype=torch.float32, device=self.device)
        return self._xpool_t

    def set_conv_mode(self, mode: int):
        """0 = per-bin partition sums then inverse FFTs (default); 1 = fused into the inverse-FFT kernel."""
        self.ctx.check(self.lib.ssb_set_conv_mode(self.ctx.handle, int(mode)), "ssb_set_conv_mode")

    def set_streams(self, n: int):
        """Run render() as n sub-batches on n internal streams (overlaps kernel tails)."""
        self.ctx.check(self.lib.ssb_set_streams(self.ctx.handle, int(n)), "ssb_set_streams")

    def set_chunks(self, n: int):
        """Each internal stream works through n sub-batches in turn (smaller L2 footprint per sub-batch)."""
        self.ctx.check(self.lib.ssb_set_chunks(self.ctx.handle, int(n)), "ssb_set_chunks")

    # ------------------------------------------------------------------ banks
    def add_rirs(self, rirs: Sequence) -> list:
        """Append RIRs ((L, 2) float32 arrays / tensors; None or empty => zero-RIR fallback).
        All-or-nothing: every entry is validated and converted BEFORE the index lists or the bank are touched,
        so a bad entry in the middle of the list leaves the bank exactly as it was."""
        offs, lens, chunks, total = [], [], [], 0
        for r in rirs:
            if r is None or len(r) == 0:
                offs.append(0)
                lens.append(0)
                continue
            t = torch.as_tensor(r, dtype=torch.float32)
            if t.ndim != 2 or t.shape[1] != 2:
                raise ValueError(f"RIR must be (L, 2), got {tuple(t.shape)}")
            offs.append(self._bank_used + total)
            lens.append(t.shape[0])
            chunks.append(t)
            total += t.shape[0]
        if total:
            need = self._bank_used + total
            if need > self._bank.shape[0]:
                grown = torch.empty((max(need, 2 * self._bank.shape[0]), 2), dtype=torch.float32, device=self.device)
                grown[: self._bank_used] = self._bank[: self._bank_used]
                self._bank = grown
            if all(c.is_cuda for c in chunks):
                host = torch.cat(chunks)
            else:
                host = torch.cat([c.cpu() if c.is_cuda else c for c in chunks])
            self._bank[self._bank_used: need].copy_(host, non_blocking=True)
            self._bank_used = need
        first = len(self._rir_off)
        self._rir_off.extend(offs)
        self._rir_len.extend(lens)
        self._bank_index = None
        return list(range(first, first + len(offs)))

    def compact_bank(self, keep_ids: Sequence[int]) -> list:
        """Keep only the given RIRs (one device-side gather into a fresh arena) and return their new ids, in the
        order given; every other id becomes invalid.  The file-backed service calls this when the resident bank
        outgrows its budget, keeping the recently used part of the working set instead of re-reading it."""
        keep_ids = [int(i) for i in keep_ids]
        lens = [self._rir_len[i] for i in keep_ids]
        total = sum(lens)
        arena = torch.empty((max(total, 1), 2), dtype=torch.float32, device=self.device)
        offs, pos = [], 0
        for i, n in zip(keep_ids, lens):
            offs.append(pos if n else 0)
            if n:
                arena[pos: pos + n].copy_(self._bank[self._rir_off[i]: self._rir_off[i] + n])
                pos += n
        self._bank, self._bank_used = arena, total
        self._rir_off, self._rir_len = offs, lens
        self._bank_index = None
        return list(range(len(keep_ids)))

    def reset_bank(self):
        """Forget every RIR (ids become invalid).  Used by the file-backed service when the resident
        bank outgrows its budget: the working set of a scene is reloaded on demand."""
        self._bank_used = 0
        self._bank_index = None
        self._rir_off.clear()
        self._rir_len.clear()

    @property
    def bank_bytes(self) -> int:
        return self._bank_used * 8

    def bank_mark(self):
        """Stack mark for transient RIRs (the continuous simulator renders a new RIR every step)."""
        return (self._bank_used, len(self._rir_off))

    def bank_release(self, mark):
        """Drop every RIR added since ``mark``.  Safe while kernels are in flight: later uploads
        are ordered after them on the same stream."""
        self._bank_used, n = mark
        self._bank_index = None
        del self._rir_off[n:]
        del self._rir_len[n:]

    def set_dense_rir_bank(self, rirs: torch.Tensor) -> list:
        """Adopt an already device-resident (n, L, 2) float32 tensor as the bank (no copy)."""
        if not (rirs.is_cuda and rirs.dtype == torch.float32 and rirs.is_contiguous() and rirs.ndim == 3
                and rirs.shape[2] == 2):
            raise ValueError("expected a contiguous CUDA float32 (n, L, 2) tensor")
        n, L, _ = rirs.shape
        self._bank = rirs.view(n * L, 2)
        self._bank_used = n * L
        self._rir_off = [i * L for i in range(n)]
        self._rir_len = [L] * n
        self._bank_index = None
        return list(range(n))

    def add_source(self, samples) -> int:
        """Register a mono clip already at ``sr`` (float32).  Returns its id."""
        t = torch.as_tensor(samples, dtype=torch.float32).reshape(-1).to(self.device).contiguous()
        if t.numel() == 0:
            raise ValueError("empty source clip")
        self._sources.append(t)
        return len(self._sources) - 1

    def add_source_pcm16(self, pcm) -> int:
        """Register an int16 clip; decoded on the device as float32(x)/32768 (bit-exact
        with the soundfile decode behind ``librosa.load``, simulator.py:597)."""
        p = torch.as_tensor(np.ascontiguousarray(pcm, dtype=np.int16)).to(self.device)
        out = torch.empty(p.numel(), dtype=torch.float32, device=self.device)
        self.ctx.check(self.lib.ssb_pcm16_decode(self.ctx.handle, p.data_ptr(), p.numel(), out.data_ptr(),
                                                 self._stream()), "ssb_pcm16_decode")
        self._sources.append(out)
        return len(self._sources) - 1

    def encode_pcm16(self, wave: torch.Tensor, mode: str = "round") -> torch.Tensor:
        w = wave.contiguous()
        out = torch.empty(w.shape, dtype=torch.int16, device=self.device)
        self.ctx.check(self.lib.ssb_pcm16_encode(self.ctx.handle, w.data_ptr(), w.numel(),
                                                 {"round": 0, "demo": 1}[mode], out.data_ptr(), self._stream()),
                       "ssb_pcm16_encode")
        return out

    def intensity(self, wave: torch.Tensor, num_frame: int = 150) -> torch.Tensor:
        """AV-WaN ``Intensity`` (avwan_sensors.py:91-100) for a (n, 2, sr) CUDA batch -> (n,) mean squares."""
        if wave.ndim == 2:
            wave = wave[None]
        wave = wave.to(self.device, torch.float32).contiguous()
        out = torch.empty(wave.shape[0], dtype=torch.float32, device=self.device)
        self.ctx.check(self.lib.ssb_intensity_batch(self.ctx.handle, wave.shape[0], wave.data_ptr(), wave.shape[2],
                                                    wave.shape[2], int(num_frame), out.data_ptr(), self._stream()),
                       "ssb_intensity_batch")
        return out

    # ------------------------------------------------------------ SH decode
    def sh_decode(self, amb: torch.Tensor, azimuth_deg, hbank=None) -> torch.Tensor:
        """Ambisonic (n, L, 9) -> binaural (n, L, 2) RIRs on the device: SH rotation about the
        vertical axis + the 9 x 2 x 256-tap HRTF bank (scripts/ambisonic_to_binaural.py:14-19).
        The result can be handed to :meth:`set_dense_rir_bank` / :meth:`add_rirs`."""
        if self._hbank is None or hbank is not None:
            if hbank is None:
                import os
                hbank = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "sh_hrtf_bank.npy"))
            hb = torch.as_tensor(np.ascontiguousarray(hbank, dtype=np.float32))
            if tuple(hb.shape) != (9, 2, 256):
                raise ValueError("HRTF bank must be (9, 2, 256)")
            self._hbank = hb.to(self.device)
        amb = torch.as_tensor(amb, dtype=torch.float32).to(self.device).contiguous()
        if amb.ndim != 3 or amb.shape[2] != 9:
            raise ValueError(f"ambisonic RIRs must be (n, L, 9), got {tuple(amb.shape)}")
        n, L, _ = amb.shape
        az = torch.as_tensor(azimuth_deg, dtype=torch.float32).reshape(-1).to(self.device)
        if az.numel() != n:
            raise ValueError("one azimuth per RIR")
        filt = torch.empty((n, 9, 256, 2), dtype=torch.float32, device=self.device)
        out = torch.empty((n, L, 2), dtype=torch.float32, device=self.device)
        self.ctx.check(self.lib.ssb_sh_decode_batch(self.ctx.handle, n, amb.data_ptr(), L, az.data_ptr(),
                                                    self._hbank.data_ptr(), filt.data_ptr(), out.data_ptr(),
                                                    self._stream()), "ssb_sh_decode_batch")
        return out

    # ---------------------------------------------------------------- helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _windows(self, source: int, offset: int, wrap: bool, out_samples: int, block64: bool = False):
        """Cached source spectra of (clip, offset): the overlap-save window set of the partitioned plan, or the one
        65536-point spectrum of the single-block plan.  Returns (x_offset, nw, wofs)."""
        if block64:
            nblk, wofs, nw = 1, 0, 1
            key = (source, offset, bool(wrap), "block64")
            plan, need = self.plan64, 65536
        else:
            nblk, wofs, nw = window_layout(self.P, self.plan.max_parts, offset, out_samples)
            key = (source, offset, bool(wrap), nblk, wofs)
            plan, need = self.plan, nw * self.N
        hit = self._xcache.get(key)
        if hit is not None:
            return hit
        if self._xpool_used + need > self._xpool.numel() // 2:
            if need > self._xpool.numel() // 2:
                raise RuntimeError("window-spectra pool too small for one clip; raise xpool_bytes")
            # simplest policy: drop every cached set (they are recomputed on demand).  Offsets handed out
            # earlier in the current prepare() are stale now: signal it to restart.
            torch.cuda.current_stream(self.device).synchronize()
            self._xcache.clear()
            self._xpool_used = 0
            raise _PoolReset()
        x_off = self._xpool_used
        src = self._sources[source]
        self.ctx.check(self.lib.ssb_source_windows(
            self.ctx.handle, C.byref(plan), src.data_ptr(), src.numel(), int(offset), int(bool(wrap)),
            nw, wofs, self._xpool.data_ptr() + 8 * x_off, self._stream()), "ssb_source_windows")
        self._xpool_used += need
        self._xcache[key] = (x_off, nw, wofs)
        return self._xcache[key]

    def _fill_terms(self, terms, rows, rir_ids, sources, offsets, wraps, out_samples, block64=False):
        """Vectorised fill of one convolution term for the requests ``rows`` (indices into the batch):
        bank offsets and effective tap counts by array indexing, one window-set lookup per distinct
        ``(source, offset, wrap, out_samples)`` instead of one per request."""
        if rows.size == 0:
            return
        n_bank = len(self._rir_len)
        if rir_ids.size and int(rir_ids.max()) >= n_bank:
            raise ValueError(f"unknown RIR id {int(rir_ids.max())}")
        if self._bank_index is None:
            self._bank_index = (np.asarray(self._rir_off, dtype=np.int64), np.asarray(self._rir_len, dtype=np.int64))
        off_arr, len_arr = self._bank_index
        valid = rir_ids >= 0
        lens = np.where(valid, len_arr[np.where(valid, rir_ids, 0)] if n_bank else 0, 0)
        has = lens > 0                                          # None / -1 / empty file => zero RIR: rir_taps stays 0
        if not has.any():
            return
        rows, rir_ids, lens = rows[has], rir_ids[has], lens[has]
        sources, offsets, wraps, out_samples = sources[has], offsets[has], wraps[has], out_samples[has]
        taps = np.minimum(lens, offsets + out_samples)          # planning.effective_taps
        too_long = taps > (self.block64_taps if block64 else self.plan.max_parts * self.P)
        if too_long.any():
            k = int(np.argmax(too_long))
            raise ValueError(f"RIR {int(rir_ids[k])} needs {int(taps[k])} taps > max_taps={self.max_taps} the renderer was sized for")
        keys = np.stack([sources, offsets, wraps.astype(np.int64), out_samples], axis=1)
        if (keys == keys[0]).all():                             # the usual step: every env plays the same clip window
            xs = np.asarray([self._windows(int(keys[0, 0]), int(keys[0, 1]), bool(keys[0, 2]), int(keys[0, 3]), block64)], dtype=np.int64)
            inverse = np.zeros(rows.size, dtype=np.int64)
        else:
            uniq, first, inverse = np.unique(keys, axis=0, return_index=True, return_inverse=True)
            inverse = inverse.reshape(-1)
            xs = np.empty((uniq.shape[0], 3), dtype=np.int64)
            for u in np.argsort(first):                         # allocate window sets in request order
                src, off, wrap, outs = (int(v) for v in uniq[u])
                xs[u] = self._windows(src, off, bool(wrap), outs, block64)
        terms["rir_offset"][rows] = off_arr[rir_ids]
        terms["x_offset"][rows] = xs[inverse, 0]
        terms["rir_taps"][rows] = taps
        terms["x_nw"][rows] = xs[inverse, 1]
        terms["x_wofs"][rows] = xs[inverse, 2]

    # ----------------------------------------------------------------- render
    def prepare(self, requests: Sequence[AudioRequest]) -> PreparedBatch:
        """Resolve requests into the device request array.  A PreparedBatch stays valid until the
        window-spectra pool is recycled (only when it overflows) or the RIR bank is reset."""
        try:
            return self._prepare(requests)
        except _PoolReset:
            try:
                return self._prepare(requests)            # pool is empty now: the whole batch must fit
            except _PoolReset:
                raise RuntimeError("window-spectra pool too small for this batch; raise xpool_bytes") from None

    def _prepare(self, requests: Sequence[AudioRequest]) -> PreparedBatch:
        """Host cost matters here: a step of 128 envs is 92 us on the device, so the request array is filled
        with array operations (one pass over the request objects, then :meth:`_prepare_columns`)."""
        n = len(requests)
        if n == 0:
            return self._prepare_columns(np.zeros((0, 8), dtype=np.int64))
        sr = self.sr
        f = np.array([(-1 if r.rir is None else r.rir, r.source, r.offset, r.wrap, r.silent,
                       sr if r.out_samples is None else r.out_samples,
                       -1 if r.distractor_rir is None else r.distractor_rir,
                       -1 if r.distractor_source is None else r.distractor_source) for r in requests], dtype=np.int64)
        return self._prepare_columns(f)

    def prepare_arrays(self, rir, source, offset=0, out_samples=None, wrap=False, silent=False,
                       distractor_rir=-1, distractor_source=-1) -> PreparedBatch:
        """:meth:`prepare` for callers that already hold their requests as arrays (one entry per env, scalars
        broadcast; -1 = no RIR / no distractor): skips the per-request Python objects altogether."""
        rir = np.atleast_1d(np.asarray(rir, dtype=np.int64))
        f = np.empty((rir.shape[0], 8), dtype=np.int64)
        cols = (rir, source, offset, wrap, silent, self.sr if out_samples is None else out_samples,
                distractor_rir, distractor_source)
        for k, c in enumerate(cols):
            f[:, k] = np.asarray(c, dtype=np.int64)
        try:
            return self._prepare_columns(f)
        except _PoolReset:
            try:
                return self._prepare_columns(f)
            except _PoolReset:
                raise RuntimeError("window-spectra pool too small for this batch; raise xpool_bytes") from None

    def _prepare_columns(self, f: np.ndarray) -> PreparedBatch:
        """f: (n, 8) int64 columns rir, source, offset, wrap, silent, out_samples, distractor_rir, distractor_source."""
        n = f.shape[0]
        reqs = np.zeros(n, dtype=REQ_DTYPE)
        if n:
            out_samples = f[:, 5]
            if out_samples.min() <= 0 or out_samples.max() > self.sr:
                raise ValueError("out_samples must be in (0, sr]")
            silent = f[:, 4] != 0
            reqs["out_samples"] = out_samples
            reqs["flags"][silent] = SSB_FLAG_SILENT
            rows = np.flatnonzero(~silent)
            g = f[rows]
            terms = reqs["term"]
            sel = g[:, 7] >= 0
            block64 = self._block64_eligible(g, sel)
            self._fill_terms(terms[:, 0], rows, g[:, 0], g[:, 1], g[:, 2], g[:, 3] != 0, g[:, 5], block64)
            if sel.any():
                if self.plan.n_terms < 2:
                    raise ValueError("renderer was created with n_terms=1; distractors need n_terms=2")
                # the whole distractor clip is convolved in full mode and cut to [:sr] (simulator.py:661-664)
                d = g[sel]
                k = d.shape[0]
                self._fill_terms(terms[:, 1], rows[sel], d[:, 6], d[:, 7], np.zeros(k, np.int64), np.zeros(k, bool), d[:, 5])
        else:
            block64 = self.plan64 is not None
        dev = self._upload_requests(reqs)
        return PreparedBatch(n, reqs, dev, self.plan64 if block64 else self.plan)

    def _upload_requests(self, reqs: np.ndarray) -> torch.Tensor:
        """Request array -> device.  A plain (synchronous, pageable) copy of the 9 KB array: measured faster on the API
        path than staging through a ring of pinned buffers with completion events (835 k vs 745 k frames/s through
        ``execute(prepare(list))``, profiles/bench_r02_mid.json vs bench_r02c_final.json) -- the bookkeeping costs more
        than the copy."""
        return torch.from_numpy(reqs.view(np.uint8).reshape(-1)).to(self.device, non_blocking=False)

    def _block64_eligible(self, g: np.ndarray, has_distractor: np.ndarray) -> bool:
        """Single-block plan for this batch?  Every non-silent request must have one term and at most
        65536 - sr + 1 EFFECTIVE taps (min(file taps, offset + out_samples), planning.effective_taps)."""
        if self.plan64 is None:
            return False
        ok = not has_distractor.any()
        if ok and g.shape[0]:
            if self._bank_index is None:
                self._bank_index = (np.asarray(self._rir_off, dtype=np.int64), np.asarray(self._rir_len, dtype=np.int64))
            len_arr = self._bank_index[1]
            ids = g[:, 0]
            if ids.size and int(ids.max()) >= len(len_arr):
                raise ValueError(f"unknown RIR id {int(ids.max())}")
            lens = np.where(ids >= 0, len_arr[np.where(ids >= 0, ids, 0)] if len(len_arr) else 0, 0)
            ok = bool((np.minimum(lens, g[:, 2] + g[:, 5]) <= self.block64_taps).all())
        if self.force_block64 and not ok:
            raise ValueError(f"log2n=16: a request has a distractor or more than {self.block64_taps} effective taps")
        return ok

    def _scratch(self, n, plan=None):
        h_elems = n * (self.plan if plan is None else plan).h_elems_per_env
        if self._hscratch is None or self._hscratch.numel() < 2 * h_elems:
            self._hscratch = torch.empty(max(2 * h_elems, 2), dtype=torch.float32, device=self.device)
        if self._wave is None or self._wave.shape[0] < n:
            self._wave = torch.empty((n, 2, self.sr), dtype=torch.float32, device=self.device)
        return self._hscratch, self._wave[:n]

    def execute(self, batch: PreparedBatch, want_wave: bool = False, out: Optional[torch.Tensor] = None,
                channels_first: bool = False):
        """Run convolution + spectrogram for a prepared batch on the current stream.
        Returns spec (n, 65, T', 2) [, wave (n, 2, sr)] as CUDA tensors.  The waveform buffer is
        owned by the renderer and overwritten by the next call (clone to keep).

        ``out`` may be any contiguous CUDA tensor of the right shape, e.g. the rollout-storage slot
        ``rollouts.observations["spectrogram"][step + 1]`` (ss_baselines/common/rollout_storage.py:88-91):
        the observation is then written in place and ``insert`` has nothing to copy.
        ``channels_first=True`` emits (n, 2, 65, T') -- the layout ``AudioCNN.forward`` asks for with
        ``permute(0, 3, 1, 2)`` (ss_baselines/av_nav/models/audio_cnn.py:86); ``spec.permute(0, 2, 3, 1)`` is
        then the reference-shaped view and the CNN's permute of it is contiguous for free."""
        n = batch.n
        shape = (n, 2, self.spec_shape[0], self.spec_shape[1]) if channels_first else (n,) + self.spec_shape
        spec = out if out is not None else torch.empty(shape, dtype=torch.float32, device=self.device)
        if n == 0:
            return (spec, self._scratch(0)[1]) if want_wave else spec
        if not (spec.is_cuda and spec.is_contiguous() and tuple(spec.shape) == shape and spec.dtype == torch.float32):
            raise ValueError("bad output tensor")
        hs, wave = self._scratch(n, batch.plan)
        self.ctx.check(self.lib.ssb_render_batch(
            self.ctx.handle, C.byref(batch.plan), n, batch.reqs_dev.data_ptr(), self._bank.data_ptr(),
            self._xpool.data_ptr(), hs.data_ptr(), wave.data_ptr(), self.sr,
            self.pad_mode | (0x100 if channels_first else 0), spec.data_ptr(), self._stream()), "ssb_render_batch")
        return (spec, wave) if want_wave else spec

    @contextlib.contextmanager
    def transient_windows(self):
        """Window-spectra sets created inside the block are dropped at its end (stack discipline, like
        :meth:`bank_mark` / :meth:`bank_release` for RIRs).  The continuous simulator asks for a new
        ``(clip, sample offset)`` every step; without this each step would leave a dead set behind until the pool
        overflowed and was reset wholesale with a stream synchronise.  Safe while kernels are in flight: the next
        ``ssb_source_windows`` into the reclaimed space is ordered after them on the same stream."""
        used, keys = self._xpool_used, set(self._xcache)
        try:
            yield
        finally:
            if self._xpool_used >= used:                  # not recycled meanwhile
                for k in [k for k in self._xcache if k not in keys]:
                    del self._xcache[k]
                self._xpool_used = used

    @contextlib.contextmanager
    def _inline_rirs(self, *request_lists):
        """Upload the inline (transient) RIRs of one call behind a bank mark and drop them afterwards."""
        inline, seen = [], set()
        for reqs in request_lists:
            for r in reqs:
                if r is not None and r.rir_array is not None and not getattr(r, "_inline_live", False) \
                        and id(r) not in seen:
                    inline.append(r)
                    seen.add(id(r))
        if not inline:
            yield
            return
        mark = self.bank_mark()
        saved = [r.rir for r in inline]
        try:
            ids = self.add_rirs([np.asarray(r.rir_array, dtype=np.float32) for r in inline])
            for r, i in zip(inline, ids):
                r.rir = i
                r._inline_live = True                     # nested calls (render_crossfade -> convolve) reuse the upload
            yield
        finally:
            for r, old in zip(inline, saved):
                r.rir = old
                r._inline_live = False
            self.bank_release(mark)                       # later uploads are stream-ordered after the kernels

    def render(self, requests: Sequence[AudioRequest], want_wave: bool = False):
        with self._inline_rirs(requests):
            return self.execute(self.prepare(requests), want_wave=want_wave)

    def convolve_prepared(self, batch: PreparedBatch) -> torch.Tensor:
        """Waveforms (n, 2, sr) of a prepared batch (renderer-owned buffer, overwritten by the next call)."""
        hs, wave = self._scratch(batch.n, batch.plan)
        if batch.n:
            self.ctx.check(self.lib.ssb_convolve_batch(
                self.ctx.handle, C.byref(batch.plan), batch.n, batch.reqs_dev.data_ptr(), self._bank.data_ptr(),
                self._xpool.data_ptr(), hs.data_ptr(), wave.data_ptr(), self.sr, self._stream()),
                "ssb_convolve_batch")
        return wave

    def convolve(self, requests: Sequence[AudioRequest]) -> torch.Tensor:
        """Waveforms only: (n, 2, sr) -- ``get_current_audiogoal_observation`` for a batch."""
        with self._inline_rirs(requests):
            return self.convolve_prepared(self.prepare(requests))

    def render_crossfade(self, cur: Sequence[AudioRequest], prev: Sequence[Optional[AudioRequest]],
                         want_wave: bool = False):
        """Continuous simulator with CROSSFADE (continuous_simulator.py:422-424): render with the
        previous RIR and with the current one, blend the first int(0.05*sr)+1 samples."""
        n = len(cur)
        enable = torch.tensor([p is not None for p in prev], dtype=torch.uint8, device=self.device)
        prev_reqs = [p if p is not None else c for p, c in zip(prev, cur)]
        with self._inline_rirs(cur, prev):
            return self._render_crossfade(cur, prev_reqs, enable, n, want_wave)

    def _render_crossfade(self, cur, prev_reqs, enable, n, want_wave):
        w_prev = self.convolve(prev_reqs)
        if self._prev_wave is None or self._prev_wave.shape[0] < n:
            self._prev_wave = torch.empty((n, 2, self.sr), dtype=torch.float32, device=self.device)
        pw = self._prev_wave[:n]
        pw.copy_(w_prev)
        wave = self.convolve(cur)
        self.ctx.check(self.lib.ssb_crossfade_batch(self.ctx.handle, n, pw.data_ptr(), wave.data_ptr(), self.sr,
                                                    self.sr, enable.data_ptr(), self._stream()), "ssb_crossfade_batch")
        spec = self.spectrogram(wave)
        return (spec, wave) if want_wave else spec

    def spectrogram(self, wave: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``SpectrogramSensor.compute_spectrogram`` (nav.py:86-100) for a (n, 2, sr) CUDA batch."""
        if wave.ndim == 2:
            wave = wave[None]
        if not (wave.is_cuda and wave.dtype == torch.float32 and wave.shape[1] == 2 and wave.shape[2] == self.sr):
            raise ValueError(f"expected CUDA float32 (n, 2, {self.sr})")
        wave = wave.contiguous()
        n = wave.shape[0]
        spec = out if out is not None else torch.empty((n,) + self.spec_shape, dtype=torch.float32, device=self.device)
        self.ctx.check(self.lib.ssb_spectrogram_batch(self.ctx.handle, n, wave.data_ptr(), self.sr, self.sr,
                                                      self.pad_mode, spec.data_ptr(), self._stream()),
                       "ssb_spectrogram_batch")
        return spec

    # ------------------------------------------------------------- log-mel (extension)
    def logmel_shape(self, n_mels: int = 64):
        """(n_mels, 1 + sr//160, 2)."""
        return (int(n_mels), self.lib.ssb_logmel_frames(self.sr), 2)

    def logmel(self, wave: torch.Tensor, n_mels: int = 64, power: int = 2, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """EXTENSION (the reference has no mel front end; BASELINE.json configs[2] names one): per ear
        ``log1p(librosa.feature.melspectrogram(y, sr, n_fft=512, hop_length=160, win_length=400, n_mels, power))``
        on the reference's STFT geometry (nav.py:89-92) for a (n, 2, sr) CUDA batch -> (n, n_mels, 1 + sr//160, 2)."""
        if wave.ndim == 2:
            wave = wave[None]
        if not (wave.is_cuda and wave.dtype == torch.float32 and wave.shape[1] == 2 and wave.shape[2] == self.sr):
            raise ValueError(f"expected CUDA float32 (n, 2, {self.sr})")
        wave = wave.contiguous()
        n = wave.shape[0]
        shape = (n,) + self.logmel_shape(n_mels)
        res = out if out is not None else torch.empty(shape, dtype=torch.float32, device=self.device)
        if not (res.is_cuda and res.is_contiguous() and tuple(res.shape) == shape and res.dtype == torch.float32):
            raise ValueError("bad output tensor")
        self.ctx.check(self.lib.ssb_logmel_batch(self.ctx.handle, n, wave.data_ptr(), self.sr, self.sr, int(n_mels),
                                                 int(power), self.pad_mode, res.data_ptr(), self._stream()),
                       "ssb_logmel_batch")
        return res

    def render_logmel(self, requests: Sequence[AudioRequest], n_mels: int = 64, power: int = 2,
                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Convolution (every reference branch, as :meth:`convolve`) followed by :meth:`logmel`."""
        return self.logmel(self.convolve(requests), n_mels=n_mels, power=power, out=out)

    # --------------------------------------------------------- host-buffer path
    def make_host_session(self, n: int, taps: int, want_wave: bool = False, n_chunks: int = 4):
        """Pinned host buffers + device staging for the host-buffer entry (the e2e path)."""
        return HostSession(self, n, taps, want_wave, n_chunks)


class HostSession:
    """End-to-end entry with HOST buffers: per step the (n, taps, 2) RIRs are copied from pinned
    host memory, rendered, and the spectrograms (optionally waveforms) are copied back --
    what the reference's per-env numpy API hands over and gets back."""

    def __init__(self, r: BatchedAudioRenderer, n: int, taps: int, want_wave: bool, n_chunks: int = 4):
        self.r, self.n, self.taps, self.n_chunks = r, n, taps, n_chunks
        self.h_rir = torch.empty((n, taps, 2), dtype=torch.float32).pin_memory()
        self.h_spec = torch.empty((n,) + r.spec_shape, dtype=torch.float32).pin_memory()
        self.h_wave = torch.empty((n, 2, r.sr), dtype=torch.float32).pin_memory() if want_wave else None
        self.h_reqs = torch.empty(n * REQ_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
        # two device staging sets, alternated per step: the copy of step k+1 overlaps the kernels of step k
        self.d_rir = [torch.empty((n, taps, 2), dtype=torch.float32, device=r.device) for _ in range(2)]
        self.d_reqs = [torch.empty(n * REQ_DTYPE.itemsize, dtype=torch.uint8, device=r.device) for _ in range(2)]
        self._step = 0
        self.d_spec = torch.empty((n,) + r.spec_shape, dtype=torch.float32, device=r.device)
        self.h2d_bytes = self.h_rir.numel() * 4 + self.h_reqs.numel()
        self.d2h_bytes = self.h_spec.numel() * 4 + (self.h_wave.numel() * 4 if want_wave else 0)

    def set_requests(self, source: int, silent=None, taps=None):
        """All envs play ``source`` through their own RIR row (the 1-s clip branch)."""
        r = self.r
        reqs = np.zeros(self.n, dtype=REQ_DTYPE)
        eff = np.minimum(self.taps if taps is None else np.asarray(taps), r.sr)
        self.block64 = r.plan64 is not None and int(np.max(eff)) <= r.block64_taps
        self.plan = r.plan64 if self.block64 else r.plan
        for i in range(self.n):
            reqs[i]["out_samples"] = r.sr
            if silent is not None and silent[i]:
                reqs[i]["flags"] = SSB_FLAG_SILENT
                continue
            L = self.taps if taps is None else int(taps[i])
            if L == 0:
                continue
            x_off, nw, wofs = r._windows(source, 0, False, r.sr, self.block64)
            t = reqs[i]["term"][0]
            t["rir_offset"], t["x_offset"] = i * self.taps, x_off
            t["rir_taps"] = min(L, r.sr, r.block64_taps if self.block64 else r.plan.max_parts * r.P)
            t["x_nw"], t["x_wofs"] = nw, wofs
        self.h_reqs.numpy()[:] = reqs.view(np.uint8).reshape(-1)

    def run(self):
        """Enqueue H2D + kernels + D2H (asynchronous; the current stream completes when the results
        have landed in ``h_spec`` / ``h_wave``).  ``h_rir`` must not be refilled before that."""
        r = self.r
        hs, wave = r._scratch(self.n, self.plan)
        d_rir, d_reqs = self.d_rir[self._step & 1], self.d_reqs[self._step & 1]
        self._step += 1
        r.ctx.check(r.lib.ssb_render_batch_host(
            r.ctx.handle, C.byref(self.plan), self.n, self.h_reqs.data_ptr(), self.h_rir.data_ptr(),
            self.h_rir.numel() * 4, d_rir.data_ptr(), d_reqs.data_ptr(), r._xpool.data_ptr(),
            hs.data_ptr(), wave.data_ptr(), r.sr, r.pad_mode, self.d_spec.data_ptr(), self.h_spec.data_ptr(),
            self.h_wave.data_ptr() if self.h_wave is not None else None, self.n_chunks, r._stream()),
            "ssb_render_batch_host")
        h2d, d2h = C.c_int64(), C.c_int64()
        r.lib.ssb_host_copy_bytes(r.ctx.handle, C.byref(h2d), C.byref(d2h))
        self.h2d_bytes, self.d2h_bytes = h2d.value, d2h.value      # what this step really moved (silent envs need no RIR)
        return self.h_spec
