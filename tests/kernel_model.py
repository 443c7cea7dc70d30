"""numpy model of the CUDA kernels' DATAFLOW (windows, partitions, valid halves) with np.fft in
place of the register FFT.  Test infrastructure: lets the CPU suite check the host-side plan
arithmetic (soundspaces_b200/planning.py) against the oracle without a GPU."""
import numpy as np

from soundspaces_b200.planning import ceil_div, effective_taps, partition_range, window_layout


def source_windows(src, S, m0, wrap, nw, wofs, P):
    """fwd_src_kernel: window j covers [m0 + (j-wofs-1)P, m0 + (j-wofs+1)P)."""
    N = 2 * P
    X = np.zeros((nw, N), dtype=np.complex128)
    for j in range(nw):
        n = m0 + (j - wofs - 1) * P + np.arange(N)
        x = np.zeros(N)
        ok = (n >= 0) & (n < S)
        x[ok] = src[n[ok]]
        if wrap:
            w = (n >= S) & (n - S < S)
            x[w] = src[n[w] - S]
        X[j] = np.fft.fft(x)
    return X


def rir_partitions(rir, taps, P, max_parts):
    """fwd_rir_kernel: partition p = taps [pP, (p+1)P) of (L + i R), zero padded to 2P."""
    N = 2 * P
    z = (rir[:taps, 0] + 1j * rir[:taps, 1]).astype(np.complex128)
    nparts = min(ceil_div(taps, P), max_parts)
    H = np.zeros((nparts, N), dtype=np.complex128)
    for p in range(nparts):
        seg = z[p * P:(p + 1) * P]
        H[p, :len(seg)] = seg
        H[p] = np.fft.fft(H[p])
    return H


def render_model(src, rir, sr, P, max_parts, offset=0, out_samples=None, wrap=False):
    out_samples = sr if out_samples is None else out_samples
    wave = np.zeros((2, sr))
    if rir is None or len(rir) == 0:
        return wave
    nblk, wofs, nw = window_layout(P, max_parts, offset, out_samples)
    taps = effective_taps(len(rir), offset, out_samples)
    assert taps <= max_parts * P
    X = source_windows(src, len(src), offset, wrap, nw, wofs, P)
    H = rir_partitions(rir, taps, P, max_parts)
    for b in range(ceil_div(sr, P)):
        n0 = b * P
        if n0 >= out_samples:
            continue
        lo, hi = partition_range(b, H.shape[0], nw, wofs)
        acc = np.zeros(2 * P, dtype=np.complex128)
        for p in range(lo, hi + 1):
            acc += X[b - p + wofs] * H[p]
        y = np.fft.ifft(acc)[P:]
        n = np.arange(n0, min(n0 + P, sr))
        ok = n < out_samples
        wave[0, n[ok]] = y.real[: len(n)][ok]
        wave[1, n[ok]] = y.imag[: len(n)][ok]
    return wave
