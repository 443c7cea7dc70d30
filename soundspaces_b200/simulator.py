"""Drop-in audio slice for the SoundSpaces simulators.

Mirrors, method for method, the audio part of the reference simulator wrappers:

* ``SoundSpacesSim._compute_audiogoal``                 soundspaces/simulator.py:608-666
* ``SoundSpacesSim.get_current_audiogoal_observation``  soundspaces/simulator.py:678-688
* ``SoundSpacesSim.get_current_spectrogram_observation`` soundspaces/simulator.py:690-701
* ``ContinuousSoundSpacesSim._compute_audiogoal`` / ``_convolve_with_rir`` / getters
                                                        soundspaces/continuous_simulator.py:413-462

``B200AudioMixin`` / ``B200ContinuousAudioMixin`` read exactly the attributes the
reference methods read (``config.AUDIO.*``, ``_receiver_position_index``,
``azimuth_angle``, ``_audio_index`` ... see SURVEY.md App. D) and keep the
reference's memoisation and ``_audio_index`` quirks, but run the arithmetic on
the GPU through :class:`AudioRenderService`.  Use either

    class SoundSpacesSim(B200AudioMixin, Simulator): ...        # in the reference
    patch_simulator(SoundSpacesSim)                             # or monkey-patch

There is no CPU fallback: without a CUDA device / libssb200.so these raise.
"""
from __future__ import annotations

import logging
import os
from typing import Callable, Dict, Optional

import numpy as np
import torch

from .renderer import AudioRequest, BatchedAudioRenderer


class AudioRenderService:
    """Per-(device, sr) owner of the renderer, the RIR-file bank and the clip bank.

    RIR files are read once with ``scipy.io.wavfile.read`` exactly like
    soundspaces/simulator.py:615-624 (unreadable -> warning + zero RIR; empty ->
    zero RIR) and stay resident on the device."""

    _instances: Dict[tuple, "AudioRenderService"] = {}

    def __init__(self, sr: int, device="cuda:0", max_taps: Optional[int] = None, pad_mode: str = "reflect",
                 n_terms: int = 2, log2n: int = 0, max_bank_bytes: int = 16 << 30):
        self.sr = sr
        self.max_bank_bytes = int(max_bank_bytes)
        # 1-s clips only ever use the first sr taps; multi-second clips need the whole RIR
        self.max_taps = int(max_taps) if max_taps else 4 * sr
        self.renderer = BatchedAudioRenderer(sr, self.max_taps, device=device, n_terms=n_terms, log2n=log2n,
                                             pad_mode=pad_mode)
        self._rir_ids: Dict[str, int] = {}
        self._src_ids: Dict[tuple, int] = {}
        self._mem_rirs = 0

    @classmethod
    def get(cls, sr: int, device="cuda:0", **kw) -> "AudioRenderService":
        key = (int(sr), str(device), kw.get("pad_mode", "reflect"))
        if key not in cls._instances:
            cls._instances[key] = cls(sr, device=device, **kw)
        return cls._instances[key]

    # -- banks ---------------------------------------------------------------
    def maybe_trim(self):
        """Call between steps (never while requests are being built): when the resident RIR bank exceeds
        its budget, drop it; files are re-read on demand (the full dataset is 867 GB, a scene's working
        set is what has to stay resident)."""
        if self.renderer.bank_bytes > self.max_bank_bytes:
            self.renderer.reset_bank()
            self._rir_ids.clear()

    def rir_from_file(self, path: str) -> int:
        rid = self._rir_ids.get(path)
        if rid is None:
            from scipy.io import wavfile
            try:
                _, rir = wavfile.read(path)                   # float32 (L, 2)
            except ValueError:
                logging.warning("{} file is not readable".format(path))
                rir = None
            if rir is not None and len(rir) == 0:
                logging.debug("Empty RIR file at {}".format(path))
                rir = None
            if rir is not None and rir.dtype != np.float32:
                rir = rir.astype(np.float32)
            rid = self.renderer.add_rirs([rir])[0]
            self._rir_ids[path] = rid
        return rid

    def source(self, key, samples) -> int:
        """Device copy of a decoded clip, memoised per array object (the reference memoises the
        decoded clip per sound name in ``_source_sound_dict``, simulator.py:595-600; keying on the
        array keeps two simulators with different clips under one name apart)."""
        k = (key, len(samples), id(samples))
        hit = self._src_ids.get(k)
        if hit is None:
            hit = (self.renderer.add_source(np.asarray(samples, dtype=np.float32)), samples)  # keep the array alive
            self._src_ids[k] = hit
        return hit[0]


SPECTROGRAM_NATIVE_ATTR = "_b200_native_spectrogram"


def _is_native_spectrogram(fn: Callable) -> bool:
    return bool(getattr(fn, SPECTROGRAM_NATIVE_ATTR, False))


class B200AudioMixin:
    """Audio slice of ``SoundSpacesSim`` (discrete, SoundSpaces 1.0)."""

    b200_device = "cuda:0"
    b200_pad_mode = "reflect"          # librosa < 0.10 behaviour (published checkpoints); see SURVEY.md #5

    # -- plumbing --------------------------------------------------------------
    def _b200_service(self) -> AudioRenderService:
        svc = getattr(self, "_b200_svc", None)
        if svc is None:
            svc = AudioRenderService.get(self.config.AUDIO.RIR_SAMPLING_RATE, self.b200_device,
                                         pad_mode=self.b200_pad_mode)
            self._b200_svc = svc
        return svc

    def _b200_rir_path(self, source_index) -> str:
        # simulator.py:303-305 + :615-616 (azimuth_angle: :568-573)
        return os.path.join(self.binaural_rir_dir, str(self.azimuth_angle),
                            "{}_{}.wav".format(self._receiver_position_index, source_index))

    def _b200_request(self) -> AudioRequest:
        """The request equivalent to one call of ``_compute_audiogoal`` (simulator.py:608-666),
        including the ``_audio_index`` advance at :635."""
        svc = self._b200_service()
        sr = self.config.AUDIO.RIR_SAMPLING_RATE
        if self._episode_step_count > self._duration:
            return AudioRequest(rir=-1, source=0, silent=True)
        rid, inline = -1, None
        if not self.config.USE_RENDERED_OBSERVATIONS:
            # simulator.py:626: RIR rendered by habitat-sim for this step -> transient, supplied inline
            inline = np.transpose(np.array(self._sim.get_sensor_observations()["audio_sensor"]))
        else:
            rid = svc.rir_from_file(self._b200_rir_path(self._source_position_index))
        clip = self.current_source_sound
        sid = svc.source(self._current_sound, clip)
        offset = 0
        if clip.shape[0] != sr:
            index = self._audio_index
            self._audio_index = (self._audio_index + 1) % self._audio_length
            offset = index * sr
        req = AudioRequest(rir=rid, source=sid, offset=offset, rir_array=inline)
        if self.config.AUDIO.HAS_DISTRACTOR_SOUND:
            dclip = self._source_sound_dict[self._current_distractor_sound]
            req.distractor_source = svc.source(self._current_distractor_sound, dclip)
            req.distractor_rir = svc.rir_from_file(self._b200_rir_path(self._distractor_position_index))
        return req

    # -- reference API ---------------------------------------------------------
    def _compute_audiogoal(self):
        sr = self.config.AUDIO.RIR_SAMPLING_RATE
        self._b200_service().maybe_trim()
        req = self._b200_request()
        if req.silent:
            logging.debug('Step count is greater than duration. Empty spectrogram.')
            return np.zeros((2, sr))                       # float64, as simulator.py:612
        wave = self._b200_service().renderer.convolve([req])
        return wave[0].cpu().numpy()

    def get_current_audiogoal_observation(self):
        if self.config.AUDIO.HAS_DISTRACTOR_SOUND:
            audiogoal = self._compute_audiogoal()
        else:
            joint_index = (self._source_position_index, self._receiver_position_index, self.azimuth_angle)
            if joint_index not in self._audiogoal_cache:
                self._audiogoal_cache[joint_index] = self._compute_audiogoal()
            audiogoal = self._audiogoal_cache[joint_index]
        return audiogoal

    def _b200_spectrogram(self, audiogoal2spectrogram):
        """One cache-missing spectrogram.  The fused device path is taken when the callable is this
        package's ``SpectrogramSensor.compute_spectrogram``; any other callable receives the host
        waveform exactly as in the reference."""
        if not _is_native_spectrogram(audiogoal2spectrogram):
            return audiogoal2spectrogram(self.get_current_audiogoal_observation())
        joint_index = (self._source_position_index, self._receiver_position_index, self.azimuth_angle)
        use_cache = not self.config.AUDIO.HAS_DISTRACTOR_SOUND
        r = self._b200_service().renderer
        if use_cache and joint_index in self._audiogoal_cache:
            wave = self._audiogoal_cache[joint_index]
            if not np.any(wave):
                return np.zeros(r.spec_shape)              # float64 zeros, as the reference on silence
            spec = r.spectrogram(torch.from_numpy(np.asarray(wave, dtype=np.float32)).to(r.device))
            return spec[0].cpu().numpy()
        self._b200_service().maybe_trim()
        req = self._b200_request()
        if req.silent:
            if use_cache:
                self._audiogoal_cache[joint_index] = np.zeros((2, r.sr))
            return np.zeros(r.spec_shape)
        spec, wave = r.render([req], want_wave=True)
        if use_cache:
            self._audiogoal_cache[joint_index] = wave[0].cpu().numpy()
        return spec[0].cpu().numpy()

    def get_current_spectrogram_observation(self, audiogoal2spectrogram):
        if self.config.AUDIO.HAS_DISTRACTOR_SOUND:
            spectrogram = self._b200_spectrogram(audiogoal2spectrogram)
        else:
            joint_index = (self._source_position_index, self._receiver_position_index, self.azimuth_angle)
            if joint_index not in self._spectrogram_cache:
                self._spectrogram_cache[joint_index] = self._b200_spectrogram(audiogoal2spectrogram)
            spectrogram = self._spectrogram_cache[joint_index]
        return spectrogram


class B200ContinuousAudioMixin:
    """Audio slice of ``ContinuousSoundSpacesSim`` (SoundSpaces 2.0)."""

    b200_device = "cuda:0"
    b200_pad_mode = "reflect"

    def _b200_service(self) -> AudioRenderService:
        svc = getattr(self, "_b200_svc", None)
        if svc is None:
            svc = AudioRenderService.get(self.config.AUDIO.RIR_SAMPLING_RATE, self.b200_device,
                                         pad_mode=self.b200_pad_mode)
            self._b200_svc = svc
        return svc

    def _b200_requests(self):
        svc = self._b200_service()
        sr = self.config.AUDIO.RIR_SAMPLING_RATE
        num_sample = int(sr * self.config.STEP_TIME)
        clip = self.current_source_sound
        sid = svc.source(self._current_sound, clip)
        kw = dict(source=sid, offset=int(self._current_sample_index), out_samples=num_sample, wrap=True)
        cur_rir = np.transpose(np.array(self._prev_sim_obs["audio_sensor"]))
        cur = AudioRequest(rir=-1, rir_array=cur_rir, **kw)
        prev = None
        if self.config.AUDIO.CROSSFADE and self._last_rir is not None:
            prev = AudioRequest(rir=-1, rir_array=self._last_rir, **kw)
        return cur, prev

    def _compute_audiogoal(self):
        sr = self.config.AUDIO.RIR_SAMPLING_RATE
        if self._episode_step_count > self._duration:
            logging.debug('Step count is greater than duration. Empty spectrogram.')
            return np.zeros((2, sr))
        r = self._b200_service().renderer
        cur, prev = self._b200_requests()
        if prev is None:
            return r.convolve([cur])[0].cpu().numpy()
        _, wave = r.render_crossfade([cur], [prev], want_wave=True)
        return wave[0].cpu().numpy()

    def get_current_audiogoal_observation(self):
        return self._compute_audiogoal()

    def get_current_spectrogram_observation(self, audiogoal2spectrogram):
        if not _is_native_spectrogram(audiogoal2spectrogram):
            return audiogoal2spectrogram(self.get_current_audiogoal_observation())
        r = self._b200_service().renderer
        if self._episode_step_count > self._duration:
            return np.zeros(r.spec_shape)
        cur, prev = self._b200_requests()
        spec = r.render([cur]) if prev is None else r.render_crossfade([cur], [prev])
        return spec[0].cpu().numpy()


_DISCRETE = ("_compute_audiogoal", "get_current_audiogoal_observation", "get_current_spectrogram_observation",
             "_b200_service", "_b200_rir_path", "_b200_request", "_b200_spectrogram")
_CONTINUOUS = ("_compute_audiogoal", "get_current_audiogoal_observation", "get_current_spectrogram_observation",
               "_b200_service", "_b200_requests")


def patch_simulator(sim_cls, continuous: bool = False, device: str = "cuda:0", pad_mode: str = "reflect"):
    """Replace the audio methods of an existing (reference) simulator class in place."""
    mixin = B200ContinuousAudioMixin if continuous else B200AudioMixin
    for name in (_CONTINUOUS if continuous else _DISCRETE):
        setattr(sim_cls, name, getattr(mixin, name))
    sim_cls.b200_device = device
    sim_cls.b200_pad_mode = pad_mode
    return sim_cls
