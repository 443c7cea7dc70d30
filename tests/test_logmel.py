"""Log-mel EXTENSION (BASELINE.json configs[2] names a log-mel front end; the reference has none).

CPU: the oracle's librosa restatement (``oracle/audio_oracle.py::mel_filterbank`` / ``logmel``) is triangulated
against torchaudio's independent Slaney filterbank and MelSpectrogram and against ``transformers.audio_utils.mel_filter_bank``
(a reimplementation of ``librosa.filters.mel``), and the filterbank the CUDA library builds
(``ssb_mel_filterbank``, host only) is bit-identical to the restatement.  GPU: ``ssb_logmel_batch`` against the
oracle, ``allclose(rtol=1e-4, atol=1e-5)`` -- the spectrogram tolerance of SURVEY.md 8(c)."""
import numpy as np
import pytest

from oracle import audio_oracle as ao
from synth import make_rir, make_source

CASES = [(16000, 64), (44100, 64), (48000, 40), (22050, 13), (16000, 1)]


@pytest.mark.parametrize("sr,n_mels", CASES)
def test_filterbank_restatement_vs_torchaudio(sr, n_mels):
    import torchaudio.functional as F
    fb = ao.mel_filterbank(sr, 512, n_mels)
    assert fb.shape == (n_mels, 257) and fb.dtype == np.float32
    ta = F.melscale_fbanks(257, 0.0, sr / 2.0, n_mels, sr, norm="slaney", mel_scale="slaney").numpy().T
    assert np.abs(fb - ta).max() <= 1e-5 * np.abs(ta).max()             # torchaudio computes the triangles in float32
    # Slaney area normalisation: every triangle integrates to ~1 over frequency (bin width sr/512);
    # exact for triangles that are wide against the bin spacing
    if n_mels <= 40:
        area = fb.sum(1) * (sr / 512.0)
        assert np.allclose(area[n_mels // 2:], 1.0, atol=0.05)


@pytest.mark.parametrize("sr,n_mels", CASES)
def test_filterbank_restatement_vs_transformers_audio_utils(sr, n_mels):
    """A third, independent implementation: ``transformers.audio_utils.mel_filter_bank`` (written to reproduce
    ``librosa.filters.mel``; float64 triangles) with Slaney scale + Slaney normalisation agrees to 1e-6 of peak."""
    au = pytest.importorskip("transformers.audio_utils")
    fb = ao.mel_filterbank(sr, 512, n_mels)
    hf = au.mel_filter_bank(num_frequency_bins=257, num_mel_filters=n_mels, min_frequency=0.0, max_frequency=sr / 2.0,
                            sampling_rate=sr, norm="slaney", mel_scale="slaney").T
    assert hf.shape == fb.shape
    assert np.abs(fb - hf).max() <= 1e-6 * np.abs(hf).max()


@pytest.mark.parametrize("sr,n_mels", CASES)
def test_library_filterbank_is_the_restatement(sr, n_mels):
    from soundspaces_b200 import _lib
    lib = _lib.load_library()
    out = np.full((n_mels, 257), np.nan, dtype=np.float32)
    assert lib.ssb_mel_filterbank(sr, n_mels, out.ctypes.data) == 0
    assert np.array_equal(out, ao.mel_filterbank(sr, 512, n_mels))
    assert lib.ssb_mel_filterbank(sr, 65, out.ctypes.data) < 0          # n_mels > 64 is rejected
    assert lib.ssb_logmel_frames(sr) == 1 + sr // 160


@pytest.mark.parametrize("power", [1, 2])
def test_oracle_logmel_vs_torchaudio(power):
    import torch
    import torchaudio
    sr = 16000
    y = np.stack([make_source(3, sr), make_source(4, sr)])
    ref = ao.logmel(y, sr, n_mels=64, power=power, pad_mode="reflect")
    assert ref.shape == (64, 101, 2)
    ms = torchaudio.transforms.MelSpectrogram(sample_rate=sr, n_fft=512, win_length=400, hop_length=160, f_min=0.0,
                                              f_max=sr / 2, n_mels=64, power=float(power), center=True,
                                              pad_mode="reflect", norm="slaney", mel_scale="slaney")
    ta = torch.log1p(ms(torch.from_numpy(y).double().float())).numpy().transpose(1, 2, 0)
    assert np.allclose(ref, ta, rtol=2e-4, atol=2e-5)


def test_oracle_logmel_silence_and_linearity():
    sr = 16000
    assert np.all(ao.logmel(np.zeros((2, sr), np.float32), sr) == 0)
    y = np.stack([make_source(1, sr), make_source(2, sr)])
    a = np.expm1(ao.logmel(y, sr, power=2))
    b = np.expm1(ao.logmel(2 * y, sr, power=2))
    assert np.allclose(b, 4 * a, rtol=1e-4)                              # power spectrum scales with amplitude^2


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("sr,n_mels", [(16000, 64), (44100, 64), (48000, 40), (22050, 13), (16000, 1)])
@pytest.mark.parametrize("pad_mode", ["reflect", "constant"])
def test_cuda_logmel_matches_oracle(sr, n_mels, pad_mode):
    import torch
    from soundspaces_b200 import BatchedAudioRenderer
    r = BatchedAudioRenderer(sr, 1024, pad_mode=pad_mode)
    waves = np.stack([np.stack([make_source(10 + i, sr), 0.3 * make_source(20 + i, sr)]) for i in range(3)])
    waves[2, 1] = 0.0                                                     # one dead ear next to a live one
    waves[1] *= np.exp(-np.arange(sr) / (sr / 8.0)).astype(np.float32)    # decaying clip
    for power in (1, 2):
        got = r.logmel(torch.from_numpy(waves).cuda(), n_mels=n_mels, power=power).cpu().numpy()
        assert got.shape == (3, n_mels, 1 + sr // 160, 2)
        for i in range(3):
            ref = ao.logmel(waves[i], sr, n_mels=n_mels, power=power, pad_mode=pad_mode)
            assert np.allclose(got[i], ref, rtol=1e-4, atol=1e-5), (power, i, np.abs(got[i] - ref).max())


@pytest.mark.gpu
def test_cuda_logmel_singing_fixture_and_silence(golden):
    """Music (res/singing.wav, the reference's only audio fixture): loud harmonics next to quiet bands."""
    import torch
    from soundspaces_b200 import BatchedAudioRenderer
    sr = 48000
    mono = ao.pcm16_to_float32(golden["singing/pcm16"])
    wave = np.stack([mono, 0.5 * mono[::-1]]).astype(np.float32)
    r = BatchedAudioRenderer(sr, 1024)
    for power in (1, 2):
        got = r.logmel(torch.from_numpy(wave).cuda(), power=power).cpu().numpy()[0]
        ref = ao.logmel(wave, sr, power=power)
        assert np.allclose(got, ref, rtol=1e-4, atol=1e-5), (power, np.abs(got - ref).max())
    z = r.logmel(torch.zeros((2, 2, sr), device="cuda"))
    assert torch.count_nonzero(z).item() == 0                            # silence => exact zeros


@pytest.mark.gpu
def test_cuda_render_logmel_config3_shape():
    """BASELINE.json configs[2]: 16 kHz, long RIRs, overlap-save convolution + log-mel."""
    import torch
    from soundspaces_b200 import AudioRequest, BatchedAudioRenderer
    sr, taps, n = 16000, 48000, 6
    r = BatchedAudioRenderer(sr, taps)
    src = make_source(5, 4 * sr)
    sid = r.add_source(src)
    rirs = [make_rir(40 + i, taps - 1000 * i) for i in range(n)]
    ids = r.add_rirs(rirs)
    reqs = [AudioRequest(rir=ids[i], source=sid, offset=(i % 4) * sr, silent=(i == 4)) for i in range(n)]
    out = torch.empty((n, 64, 101, 2), device="cuda")
    got = r.render_logmel(reqs, out=out)
    assert got.data_ptr() == out.data_ptr()
    got = got.cpu().numpy()
    for i in range(n):
        wave = ao.compute_audiogoal(src, rirs[i], sr, silent=(i == 4), audio_index=i % 4)
        ref = ao.logmel(wave.astype(np.float32), sr)
        assert np.allclose(got[i], ref, rtol=1e-4, atol=1e-5), (i, np.abs(got[i] - ref).max())
    assert np.all(got[4] == 0)
    with pytest.raises(RuntimeError):
        r.logmel(torch.zeros((1, 2, sr), device="cuda"), n_mels=65)
