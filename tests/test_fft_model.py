"""CPU checks of the kernels' index / twiddle algebra through the numpy model in tests/fft_model.py
(the thread-and-register schedule of csrc/fft16.cuh and the spectrogram kernel's epilogue)."""
import numpy as np
import pytest

import fft_model as fm
from oracle import audio_oracle as ao
from synth import make_source


@pytest.mark.parametrize("log2n", [9, 12, 13, 14])
def test_register_fft_schedule_matches_numpy(log2n):
    N = 1 << log2n
    rng = np.random.default_rng(log2n)
    x = (rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(np.complex64)
    X = fm.forward(x, log2n)                                             # slot order
    ref = np.fft.fft(x.astype(np.complex128))
    k = fm.freq_of_slot(log2n)
    assert sorted(k.reshape(-1).tolist()) == list(range(N))             # slot order is a permutation of the bins
    assert np.abs(X - ref[k]).max() <= 2e-6 * np.abs(ref).max()
    y = fm.inverse(X, log2n) / N                                         # the inverse consumes slot order as is
    assert np.abs(y - x).max() <= 2e-6 * np.abs(x).max()
    # high dynamic range (a loud tone next to quiet noise): the correctly rounded twiddle tables keep the leak small
    x2 = (100 * np.exp(2j * np.pi * 37.3 * np.arange(N) / N) + 1e-3 * rng.standard_normal(N)).astype(np.complex64)
    X2, ref2 = fm.forward(x2, log2n), np.fft.fft(x2.astype(np.complex128))
    assert np.abs(X2 - ref2[k]).max() <= 2e-6 * np.abs(ref2).max()


def test_overlap_save_product_needs_no_reordering():
    """Forward spectra in slot order, pointwise product, inverse: circular convolution (the reason fwd_rir /
    mac_bins / mac_ifft never permute a spectrum)."""
    log2n, N = 12, 4096
    rng = np.random.default_rng(1)
    a = rng.standard_normal(N).astype(np.complex64)
    b = (rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(np.complex64)
    y = fm.inverse((fm.forward(a, log2n) * fm.forward(b, log2n)).astype(np.complex64), log2n) / N
    ref = np.fft.ifft(np.fft.fft(a.astype(np.complex128)) * np.fft.fft(b.astype(np.complex128)))
    assert np.abs(y - ref).max() <= 5e-6 * np.abs(ref).max()


def test_stft_frame_with_folded_last_stage():
    """SPEC_FOLD_R2: the last radix-2 stage of the 512-point transform is folded into the magnitude reads
    (Z[k] = P[k] + Q[k], Z[N-k] = P[256-k] - Q[256-k], Z[N-0] = Z[0], Z[256] = P[0] - Q[0])."""
    rng = np.random.default_rng(5)
    zl, zr = rng.standard_normal(512).astype(np.float32), rng.standard_normal(512).astype(np.float32)
    ml, mr, l64, r64 = fm.frame_magnitudes_folded(zl, zr)
    w = fm.hann_padded().astype(np.float64)
    XL, XR = np.fft.rfft(w * zl), np.fft.rfft(w * zr)
    lane = np.arange(32)
    for m in range(8):
        k = lane + 32 * m
        assert np.allclose(0.5 * ml[:, m], np.abs(XL[k]), rtol=0, atol=2e-5 * np.abs(XL).max())
        assert np.allclose(0.5 * mr[:, m], np.abs(XR[k]), rtol=0, atol=2e-5 * np.abs(XR).max())
    assert abs(l64 - abs(XL[256])) <= 2e-5 * np.abs(XL).max() and abs(r64 - abs(XR[256])) <= 2e-5 * np.abs(XR).max()


@pytest.mark.parametrize("pad_mode", ["reflect", "constant"])
def test_pooled_column_lane_mapping_matches_oracle(pad_mode):
    """Accumulate 4 frames, pool 4 bins by two xor-shuffles in 4-lane groups, write row (lane>>2) + 8(2q+j):
    equals the reference's block_reduce + log1p, including the edge columns and the half-empty last column."""
    sr = 16000
    y = np.stack([make_source(11, sr), 0.5 * make_source(12, sr)])
    ref = ao.compute_spectrogram(y, pad_mode=pad_mode)                   # (65, 26, 2)
    for col in (0, 1, 12, 24, 25):
        got = fm.pooled_column(y[0], y[1], col, pad_mode=pad_mode)
        assert np.allclose(got, ref[:, col, :], rtol=1e-4, atol=2e-5), col
