"""CPU oracle for the SoundSpaces per-step audio observation.

TEST INFRASTRUCTURE ONLY.  Nothing under ``soundspaces_b200/`` may import this
module; it exists so that ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` can check and time
the reference's algorithm on the CPU.

It is a restatement (numpy + the reference's own third-party leaf
``scipy.signal.fftconvolve``) of these reference functions
(paths relative to the upstream repository):

* ``soundspaces/simulator.py:608-666``  ``SoundSpacesSim._compute_audiogoal``
* ``soundspaces/continuous_simulator.py:47-53``   ``crossfade``
* ``soundspaces/continuous_simulator.py:413-456`` ``_compute_audiogoal`` /
  ``_convolve_with_rir``
* ``soundspaces/tasks/nav.py:86-100``   ``SpectrogramSensor.compute_spectrogram``
* ``ss_baselines/savi/pretraining/audiogoal_dataset.py:114-140`` (N4 quirk)

Pinning status
--------------
* Branch logic + convolution: PINNED.  ``tests/golden/make_golden.py`` imports
  the unmodified reference modules (``oracle/ref_harness.py``) in the build
  container and the committed fixtures hold the reference's own outputs;
  ``tests/test_oracle_golden.py`` checks this file against them bit-for-bit.
* ``librosa.stft`` / ``skimage.measure.block_reduce`` leaves: the reference
  imports them from un-vendored, unpinned third-party packages
  (``setup.py:34,43``) that are NOT installable here, and the reference ships no
  test that pins their output => **parity unpinned by the reference itself** at
  that leaf.  The restatement below follows librosa ``core/spectrum.py::stft``
  and skimage ``measure/block.py::block_reduce`` as published, is executed
  *through* the reference's own ``compute_spectrogram`` by the harness, and is
  triangulated against two independent STFT implementations
  (``torch.stft`` and ``scipy.signal.ShortTimeFFT``) in
  ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import numpy as np
from scipy.signal import fftconvolve, get_window

N_FFT = 512
HOP_LENGTH = 160
WIN_LENGTH = 400
POOL = 4


# --------------------------------------------------------------------------
# leaves restated from third-party packages the reference imports
# --------------------------------------------------------------------------
def librosa_stft(y, n_fft=N_FFT, hop_length=HOP_LENGTH, win_length=WIN_LENGTH,
                 pad_mode="reflect"):
    """``librosa.stft(y, n_fft, hop_length, win_length)`` with ``window='hann'``,
    ``center=True`` as called at ``soundspaces/tasks/nav.py:92``.

    ``pad_mode='reflect'`` is librosa < 0.10's default, ``'constant'`` is
    librosa >= 0.10's.  Window is float64 (``scipy.signal.get_window('hann',
    400, fftbins=True)`` centre-padded to 512); the product with the frames and
    the rFFT run in float64; the result is stored as complex64 when ``y`` is
    float32 (librosa ``util.dtype_r2c``).
    """
    y = np.asarray(y)
    w = get_window("hann", win_length, fftbins=True)
    lpad = (n_fft - win_length) // 2
    w = np.pad(w, (lpad, n_fft - win_length - lpad))
    yp = np.pad(y, n_fft // 2, mode=pad_mode)
    n_frames = 1 + (yp.shape[0] - n_fft) // hop_length
    frames = np.lib.stride_tricks.as_strided(
        yp, shape=(n_fft, n_frames),
        strides=(yp.strides[0], yp.strides[0] * hop_length), writeable=False)
    out_dtype = np.complex64 if y.dtype == np.float32 else np.complex128
    return np.fft.rfft(w[:, None] * frames, axis=0).astype(out_dtype)


def block_reduce_mean(a, block=(POOL, POOL)):
    """``skimage.measure.block_reduce(a, block, np.mean)``: trailing zero-pad
    to a block multiple, then mean over each block (always divides by the full
    block size)."""
    a = np.asarray(a)
    pr = (-a.shape[0]) % block[0]
    pc = (-a.shape[1]) % block[1]
    a = np.pad(a, ((0, pr), (0, pc)), mode="constant")
    # skimage: view_as_blocks -> (R, C, br, bc); func(blocked, axis=(2, 3))
    v = a.reshape(a.shape[0] // block[0], block[0], a.shape[1] // block[1], block[1])
    v = v.transpose(0, 2, 1, 3)
    return np.mean(v, axis=(2, 3))


# --------------------------------------------------------------------------
# nav.py:86-100
# --------------------------------------------------------------------------
def compute_spectrogram(audio_data, pad_mode="reflect"):
    """``SpectrogramSensor.compute_spectrogram`` (nav.py:86-100): per ear
    ``log1p(block_reduce(abs(stft(x)), (4,4), mean))`` stacked on the last
    axis -> ``(65, ceil((1+sr//160)/4), 2)``."""
    def compute_stft(signal):
        stft = np.abs(librosa_stft(signal, pad_mode=pad_mode))
        return block_reduce_mean(stft, (POOL, POOL))

    c1 = np.log1p(compute_stft(audio_data[0]))
    c2 = np.log1p(compute_stft(audio_data[1]))
    return np.stack([c1, c2], axis=-1)


def spectrogram_shape(sr):
    t = 1 + sr // HOP_LENGTH
    return (N_FFT // 2 // POOL + 1, -(-t // POOL), 2)


# --------------------------------------------------------------------------
# simulator.py:608-666
# --------------------------------------------------------------------------
def fallback_rir(rir, sr):
    """simulator.py:617-624: unreadable or empty RIR file -> zeros((sr, 2))."""
    if rir is None or len(rir) == 0:
        return np.zeros((sr, 2), dtype=np.float32)
    return rir


def compute_audiogoal(source, rir, sr, *, silent=False, audio_index=0,
                      distractor=None, distractor_rir=None):
    """``SoundSpacesSim._compute_audiogoal`` (simulator.py:608-666).

    source: (S,) mono clip already at ``sr``; rir: (L, 2) or None/empty (=>
    zero fallback).  ``audio_index`` is ``self._audio_index`` *before* the call
    (only used when S != sr); the caller advances it as ``(index+1) %
    (S//sr)`` (simulator.py:635).  Returns the (2, sr) waveform: float32 for
    float32 inputs, float64 zeros when silent (simulator.py:612).
    """
    if silent:
        return np.zeros((2, sr))
    rir = fallback_rir(rir, sr)
    if source.shape[0] == sr:                                   # :629-632
        conv = np.array([fftconvolve(source, rir[:, ch]) for ch in range(rir.shape[-1])])
        audiogoal = conv[:, :sr]
    else:
        index = audio_index
        if index * sr - rir.shape[0] < 0:                       # :636-640
            seg = source[: (index + 1) * sr]
            conv = np.array([fftconvolve(seg, rir[:, ch]) for ch in range(rir.shape[-1])])
            audiogoal = conv[:, index * sr: (index + 1) * sr]
        else:                                                   # :641-647
            seg = source[index * sr - rir.shape[0] + 1: (index + 1) * sr]
            audiogoal = np.array([fftconvolve(seg, rir[:, ch], mode="valid")
                                  for ch in range(rir.shape[-1])])
    if distractor is not None:                                  # :649-664
        drir = fallback_rir(distractor_rir, sr)
        dconv = np.array([fftconvolve(distractor, drir[:, ch]) for ch in range(drir.shape[-1])])
        audiogoal = audiogoal + dconv[:, :sr]
    return audiogoal


def next_audio_index(audio_index, source_len, sr):
    """simulator.py:635 (only taken when the clip is not exactly 1 s)."""
    if source_len == sr:
        return audio_index
    return (audio_index + 1) % (source_len // sr)


# --------------------------------------------------------------------------
# continuous_simulator.py
# --------------------------------------------------------------------------
def crossfade(x1, x2, sr):
    """continuous_simulator.py:47-53."""
    n = int(0.05 * sr)
    w2 = np.arange(n + 1) / n
    w1 = np.flip(w2)
    return np.concatenate([x1[:, :n + 1] * w1 + x2[:, :n + 1] * w2, x2[:, n + 1:]], axis=1)


def continuous_convolve_with_rir(source, rir, sr, step_time, sample_index):
    """``ContinuousSoundSpacesSim._convolve_with_rir`` (continuous_simulator.py:428-456)."""
    num_sample = int(sr * step_time)
    index = sample_index
    if index - rir.shape[0] < 0:
        seg = source[: index + num_sample]
        conv = np.array([fftconvolve(seg, rir[:, ch]) for ch in range(rir.shape[-1])])
        audiogoal = conv[:, index: index + num_sample]
    else:
        if index + num_sample < source.shape[0]:
            seg = source[index - rir.shape[0] + 1: index + num_sample]
        else:
            wrap = index + num_sample - source.shape[0]
            seg = np.concatenate([source[index - rir.shape[0] + 1:], source[:wrap]])
        audiogoal = np.array([fftconvolve(seg, rir[:, ch], mode="valid")
                              for ch in range(rir.shape[-1])])
    return np.pad(audiogoal, [(0, 0), (0, sr - audiogoal.shape[1])])


def continuous_compute_audiogoal(source, rir, sr, step_time, sample_index, *,
                                 silent=False, last_rir=None, crossfade_on=False):
    """``ContinuousSoundSpacesSim._compute_audiogoal`` (continuous_simulator.py:413-426)."""
    if silent:
        return np.zeros((2, sr))
    audiogoal = continuous_convolve_with_rir(source, rir, sr, step_time, sample_index)
    if crossfade_on and last_rir is not None:
        prev = continuous_convolve_with_rir(source, last_rir, sr, step_time, sample_index)
        audiogoal = crossfade(prev, audiogoal, sr)
    return audiogoal


def continuous_next_sample_index(sample_index, sr, step_time, source_len):
    """continuous_simulator.py:389-390."""
    return int(sample_index + sr * step_time) % source_len


# --------------------------------------------------------------------------
# side consumers of the waveform
# --------------------------------------------------------------------------
def intensity(audiogoal, num_frame=150):
    """``Intensity.get_observation`` (ss_baselines/av_wan/avwan_sensors.py:91-100)."""
    nonzero_idx = np.min((audiogoal > 0.1 * audiogoal.max()).argmax(axis=1))
    impulse = audiogoal[:, nonzero_idx: nonzero_idx + num_frame]
    return np.mean(impulse ** 2)


def savi_dataset_audiogoal(source, rir, sr, index):
    """``AudioGoalDataset.compute_audiogoal`` (ss_baselines/savi/pretraining/audiogoal_dataset.py:114-140)
    for a given ``index`` (the reference draws it with random.randint).  Its steady-state slice starts
    one sample earlier than the simulator's (``[index*sr - L :]`` + 'valid' + drop last)."""
    rir = fallback_rir(rir, sr)
    if index * sr - rir.shape[0] < 0:
        seg = source[: (index + 1) * sr]
        conv = np.array([fftconvolve(seg, rir[:, ch]) for ch in range(rir.shape[-1])])
        return conv[:, index * sr: (index + 1) * sr]
    seg = source[index * sr - rir.shape[0]: (index + 1) * sr]
    conv = np.array([fftconvolve(seg, rir[:, ch], mode="valid") for ch in range(rir.shape[-1])])
    return conv[:, :-1]


# --------------------------------------------------------------------------
# log-mel EXTENSION (no reference code: BASELINE.json configs[2] names a log-mel front end, the reference has
# none -- SURVEY.md 8(d)).  Restates librosa as published: ``filters.mel`` (Slaney scale, norm='slaney') and
# ``feature.melspectrogram`` (mel_basis @ abs(stft)**power), followed by the reference's own ``np.log1p``
# compression (nav.py:97).  PARITY UNPINNED (librosa is not installable here); the filterbank is
# triangulated against ``torchaudio.functional.melscale_fbanks`` in tests/test_logmel.py.
# --------------------------------------------------------------------------
def hz_to_mel(f):
    """librosa.hz_to_mel(htk=False): linear below 1 kHz (200/3 Hz per mel), log above (step log(6.4)/27)."""
    f = np.asanyarray(f, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, f / f_sp)


def mel_to_hz(m):
    """librosa.mel_to_hz(htk=False)."""
    m = np.asanyarray(m, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr, n_fft=N_FFT, n_mels=64):
    """``librosa.filters.mel(sr=sr, n_fft=n_fft, n_mels=n_mels)`` (fmin=0, fmax=sr/2, htk=False, norm='slaney',
    dtype=float32) -> (n_mels, 1 + n_fft//2)."""
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(sr / 2.0), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2: n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


def logmel(audio_data, sr, n_mels=64, power=2, pad_mode="reflect"):
    """(2, S) waveform -> (n_mels, 1 + S//160, 2): per ear ``log1p(mel_filterbank @ abs(stft)**power)``, i.e.
    ``np.log1p(librosa.feature.melspectrogram(y=ear, sr=sr, n_fft=512, hop_length=160, win_length=400,
    n_mels=n_mels, power=power))`` on the STFT geometry of nav.py:89-92."""
    fb = mel_filterbank(sr, N_FFT, n_mels)
    ears = []
    for ch in range(2):
        S = np.abs(librosa_stft(np.asarray(audio_data[ch]), pad_mode=pad_mode)) ** power
        ears.append(np.log1p(fb @ S))
    return np.stack(ears, axis=-1)


# --------------------------------------------------------------------------
# PCM helpers (A1: librosa.load -> soundfile decode; interactive_demo.py:110)
# --------------------------------------------------------------------------
def pcm16_to_float32(x):
    """soundfile/librosa int16 decode: ``float32(x) / 32768`` (exact)."""
    return (np.asarray(x, dtype=np.int16).astype(np.float32) / np.float32(32768.0))


def float32_to_pcm16_round(x):
    """inverse of :func:`pcm16_to_float32` (x*32768, round-to-nearest, saturate)."""
    v = np.rint(np.asarray(x, dtype=np.float32).astype(np.float64) * 32768.0)
    return np.clip(v, -32768, 32767).astype(np.int16)


def float32_to_pcm16_demo(x):
    """``np.int16(audio * 32767)`` (scripts/interactive_demo.py:110): truncation
    toward zero; inputs outside int16 are saturated here (numpy wraps/UB)."""
    v = np.trunc(np.asarray(x, dtype=np.float32) * np.float32(32767.0))
    return np.clip(v, -32768, 32767).astype(np.int16)


# --------------------------------------------------------------------------
# whole reference sensor path for one env (used by the CPU baseline)
# --------------------------------------------------------------------------
def render_frame(source, rir, sr, pad_mode="reflect", **kw):
    """_compute_audiogoal -> compute_spectrogram, i.e. one cache-missing call of
    ``get_current_spectrogram_observation`` (simulator.py:690-701)."""
    wave = compute_audiogoal(source, rir, sr, **kw)
    return wave, compute_spectrogram(wave, pad_mode=pad_mode)
