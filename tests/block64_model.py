"""numpy (complex64) model of the single-block 65536-point convolution kernel (``conv64k_kernel`` in
``soundspaces_b200/csrc/conv64k.cuh``): which CTA of the 4-CTA cluster, which 256-thread group and which
thread holds which element in every phase, the modulation / radix-4 / twiddle algebra of the pre- and post-stages
around the 4096-point sub-transforms (``fft_model.forward/inverse`` = ``fft16.cuh``), the in-place exchange buffers,
the cross-CTA radix-4 combine and the shift ``D`` that makes the circular convolution alias-free on the wanted
samples.  Test infrastructure: pins the index / twiddle ALGEBRA on the CPU (tests/test_block64_model.py); the
arithmetic itself is checked against the oracle on the GPU."""
import numpy as np

import fft_model as fm

c64 = np.complex64
M, NS, R, NCTA, NGRP = 65536, 4096, 16, 4, 4
TPB = 1024                      # threads per CTA = 4 groups x 256


def wM(k):
    return np.exp(-2j * np.pi * (np.asarray(k) % M) / M).astype(c64)


def shift_of(sr):
    """D: output sample m sits at circular index m + D; alias-free for taps <= D + 1."""
    return M - sr


def source_segment(src, offset, wrap, sr):
    """xs[j] = x_ext[offset - D + j], j < M (x_ext: zeros before the clip; one wrap past its end if ``wrap``)."""
    S, D = len(src), shift_of(sr)
    n = offset - D + np.arange(M)
    xs = np.zeros(M, dtype=np.float32)
    ok = (n >= 0) & (n < S)
    xs[ok] = src[n[ok]]
    if wrap:
        w = (n >= S) & (n - S < S)
        xs[w] = src[n[w] - S]
    return xs


def sub_input(x, r):
    """a_r[n1] = w_M^(n1 r) * sum_q x[n1 + 4096 q] w16^(q r)   (DIF radix-16 first stage, output r)."""
    n1 = np.arange(NS)
    acc = np.zeros(NS, dtype=np.complex128)
    for q in range(R):
        acc += x[n1 + NS * q] * np.exp(-2j * np.pi * ((q * r) % 16) / 16)
    return (acc * np.exp(-2j * np.pi * ((n1 * r) % M) / M)).astype(c64)


def source_spectrum(xs):
    """src64k_kernel: SX[r][slot] with slot = i * 256 + t  <->  bin 16 * freq_of_slot(t, i) + r."""
    SX = np.zeros((R, NS), dtype=c64)
    for r in range(R):
        v = fm.forward(sub_input(xs.astype(np.float64), r), 12)          # (256, 16) slots (t, i)
        SX[r] = v.T.reshape(-1)                                          # [i * 256 + t]
    return SX


def tw_table():
    """TWM[r][tau] = w_M^(tau r), r < 16, tau < 1024 (correctly rounded once, as the library builds it)."""
    r, tau = np.arange(R)[:, None], np.arange(TPB)[None]
    return np.exp(-2j * np.pi * ((r * tau) % M) / M).astype(c64)


W64 = np.exp(-2j * np.pi * np.arange(64) / 64).astype(c64)


def bfly4(a, inv):
    return fm.bfly4(a, inv)


def conv64k(rir, taps, SX, sr, nvalid):
    """One env.  rir: (L, 2) float32; returns wave (2, sr) float32 exactly as the kernel's data flow produces it."""
    D = shift_of(sr)
    assert taps <= D + 1 and nvalid <= sr
    h = np.zeros(M, dtype=c64)
    h[:taps] = (rir[:taps, 0] + 1j * rir[:taps, 1]).astype(c64)
    TWM = tw_table()
    tau = np.arange(TPB)
    bufs = np.zeros((NCTA, NGRP, NS), dtype=c64)                          # [cta][buffer][n1]
    for c in range(NCTA):
        # ---- phase A: modulate, radix-4 over q1 (and q2 when taps > 16384), twiddle; thread tau, n1 = tau + 1024 i
        for i in range(4):
            n1 = tau + 1024 * i
            u = []
            for q1 in range(4):
                acc = np.zeros(TPB, dtype=c64)
                for q2 in range(4):
                    if 16384 * q2 >= taps:
                        break
                    w4 = c64(np.exp(-2j * np.pi * ((q2 * c) % 4) / 4))
                    acc = (acc + h[n1 + NS * (q1 + 4 * q2)] * w4).astype(c64)
                u.append((acc * W64[((4 * q1 + i) * c) & 63]).astype(c64))     # w16^(q1 c) * w64^(i c)
            s = bfly4(u, False)                                            # s[m] = sum_q1 w4^(q1 m) u[q1]
            for m in range(4):
                r = c + 4 * m
                tw = (TWM[r][tau] * W64[(i * 4 * m) & 63]).astype(c64)     # w_M^(tau r) * w16^(i m);  w64^(i c) is in u
                bufs[c, m, n1] = (s[m] * tw).astype(c64)
        # ---- phases B-D: per group m: forward 4096, multiply by the source spectrum, inverse 4096 (unscaled)
        for m in range(4):
            r = c + 4 * m
            v = fm.forward(bufs[c, m].copy(), 12)                          # v[t, i] = slot (t, i)
            sx = SX[r].reshape(16, 256).T                                  # [t, i]
            v = (v * sx).astype(c64)
            bufs[c, m] = fm.inverse(v, 12)                                 # e_r[n1], n1 = t + 256 j
        # ---- phase E: conj twiddle, inverse radix-4 over m, demodulate; in place: buffer q1 <- v_c[q1][n1]
        for i in range(4):
            n1 = tau + 1024 * i
            b = []
            for m in range(4):
                r = c + 4 * m
                tw = (TWM[r][tau] * W64[(i * 4 * m) & 63]).astype(c64)
                b.append((bufs[c, m, n1] * np.conj(tw)).astype(c64))
            g = bfly4(b, True)                                             # g[q1] = sum_m conj(w4)^(q1 m) b[m]
            for q1 in range(4):
                bufs[c, q1, n1] = (g[q1] * np.conj(W64[((4 * q1 + i) * c) & 63])).astype(c64)
    # ---- phase F: CTA k combines n1 in [1024 k, 1024 k + 1024): radix-4 over the CTAs (DSMEM reads)
    wave = np.zeros((2, sr), dtype=np.float32)
    scale = np.float32(1.0 / M)
    for k in range(NCTA):
        n1 = 1024 * k + tau
        for q1 in range(4):
            val = [bufs[cc, q1, n1] for cc in range(NCTA)]
            y = bfly4(val, True)                                           # y[q2] = sum_c conj(w4)^(q2 c) val[c]
            for q2 in range(4):
                m_out = n1 + NS * (q1 + 4 * q2) - D
                ok = (m_out >= 0) & (m_out < nvalid)
                wave[0, m_out[ok]] = y[q2][ok].real * scale
                wave[1, m_out[ok]] = y[q2][ok].imag * scale
    return wave
