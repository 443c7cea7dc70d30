"""CPU oracle for the ambisonic -> binaural decode.

TEST INFRASTRUCTURE ONLY.  The reference does this offline with a closed-source binary
(``scripts/AmbisonicBinauralizer``, driven by ``scripts/ambisonic_to_binaural.py:14-19``).  Probing it
(``oracle/sh_elf.py``, ``tests/golden/make_sh_golden.py``) shows a linear system:

    out[n, ear] = sum_k sum_tau (R(az) a)_k[n - 128 - tau] * h[k, ear, tau],   tau < 256

* ``a`` is the 9-channel second-order ambisonic signal in ACN order, ``R(az)`` the rotation of real
  spherical harmonics about the vertical axis: the (m=-|m|, m=+|m|) pairs (1,3), (5,7) rotate by
  ``az`` and (4,8) by ``2 az`` (fitted from the tool's impulse responses to 1e-7);
* ``h`` is a fixed bank of 9 x 2 FIR filters of 256 taps (HRTF, identical at every sample rate)
  behind a 128-sample bulk delay; the output has the input's length (tail cut).

The tool itself is only approximately time invariant (its block convolver deviates by ~1.2e-3
absolute for impulses that are not aligned to its 128-sample blocks), so parity is defined as
SURVEY.md section 7 says: the CUDA path must match THIS LTI model to 1e-4 of peak, and the model
matches the tool to 5e-3 of peak on arbitrary signals and ~1e-7 on block-aligned impulses.
"""
from __future__ import annotations

import numpy as np
from scipy.signal import fftconvolve

SH_CHANNELS = 9
SH_TAPS = 256
SH_DELAY = 128


def rotation_matrix(azimuth_deg):
    """R such that the rotated signal is R @ a (ACN order, rotation about the vertical axis)."""
    al = np.deg2rad(float(azimuth_deg))
    R = np.eye(SH_CHANNELS)
    for i_neg, i_pos, m in ((1, 3, 1), (5, 7, 1), (4, 8, 2)):
        c, s = np.cos(m * al), np.sin(m * al)
        R[i_pos, i_pos] = c
        R[i_pos, i_neg] = -s
        R[i_neg, i_pos] = s
        R[i_neg, i_neg] = c
    return R


def sh_decode(amb, azimuth_deg, hbank):
    """amb: (n, 9); hbank: (9, 2, 256).  Returns the (n, 2) float32 binaural signal."""
    amb = np.asarray(amb, dtype=np.float64)
    n = amb.shape[0]
    rot = amb @ rotation_matrix(azimuth_deg).T                      # (n, 9): (R a)_j
    out = np.zeros((n + SH_DELAY + SH_TAPS, 2))
    for k in range(SH_CHANNELS):
        for e in range(2):
            y = fftconvolve(rot[:, k], np.asarray(hbank[k, e], dtype=np.float64))
            out[SH_DELAY: SH_DELAY + y.shape[0], e] += y
    return out[:n].astype(np.float32)
