"""numpy emulation (complex64) of the planned per-thread FFT schedule, to validate
index math before writing CUDA. Thread t holds 16 complex values."""
import numpy as np

def plan(log2n):
    N = 1 << log2n
    T = N // 16
    strides = []
    M = N
    while M >= 16:
        strides.append(M // 16)
        M //= 16
    return N, T, strides, M   # M leftover radix (1,2,4,8)

def positions(t, i, stride):
    return (t // stride) * 16 * stride + (t % stride) + i * stride

def dft16(v, sign):
    # v: (T,16) -> (T,16) 16-point DFT along axis 1, complex64
    k = np.arange(16)
    W = np.exp(sign * 2j * np.pi * np.outer(k, k) / 16).astype(np.complex64)
    return (v @ W.T).astype(np.complex64)  # out[s] = sum_q v[q] w^{qs}

def tw_powers(w1):
    # emulate power tree in complex64
    w = [None] * 16
    w[1] = w1
    w[2] = w[1] * w[1]; w[3] = w[2] * w[1]; w[4] = w[2] * w[2]
    w[5] = w[4] * w[1]; w[6] = w[3] * w[3]; w[7] = w[4] * w[3]; w[8] = w[4] * w[4]
    for k in range(9, 16):
        w[k] = w[8] * w[k - 8]
    w[0] = np.ones_like(w1)
    return np.stack(w, axis=1).astype(np.complex64)

def forward(x, log2n):
    N, T, strides, M = plan(log2n)
    t = np.arange(T)
    table = np.exp(-2j * np.pi * np.arange(N) / N).astype(np.complex64)
    buf = x.astype(np.complex64).copy()
    for p, st in enumerate(strides):
        pos = positions(t[:, None], np.arange(16)[None, :], st)   # (T,16)
        v = buf[pos]
        u = dft16(v, -1)
        j = t % st
        Mp = 16 * st
        w1 = table[j * (N // Mp)]
        u = u * tw_powers(w1)
        buf[pos] = u.astype(np.complex64)
    # leftover radix M across M adjacent lanes (here: in buffer, positions seg*M + j)
    st = strides[-1]
    assert st == M or (M == 1 and st == 1)
    if M > 1:
        # in-place DIF radix-2 stages over groups of M contiguous elements
        b = buf.reshape(-1, M)
        m = M
        while m >= 2:
            h = m // 2
            bb = b.reshape(b.shape[0], M // m, m)
            a0 = bb[:, :, :h].copy(); a1 = bb[:, :, h:].copy()
            tw = np.exp(-2j * np.pi * np.arange(h) / m).astype(np.complex64)
            bb[:, :, :h] = a0 + a1
            bb[:, :, h:] = (a0 - a1) * tw
            m = h
        buf = b.reshape(-1)
    return buf

def freq_of_position(log2n):
    """frequency index stored at each final buffer position"""
    N, T, strides, M = plan(log2n)
    pos = np.arange(N)
    k = np.zeros(N, dtype=np.int64)
    mult = 1
    rem = pos.copy()
    size = N
    for st in strides:
        s = rem // st          # digit
        rem = rem % st
        k += s * mult
        mult *= 16
    # leftover: rem in [0,M): bit-reversed
    if M > 1:
        bits = int(np.log2(M))
        r = np.zeros_like(rem)
        for bpos in range(bits):
            r |= ((rem >> bpos) & 1) << (bits - 1 - bpos)
        k += r * mult
    return k

for log2n in (8, 9, 10, 12, 13, 14):
    N = 1 << log2n
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(np.complex64)
    X = forward(x, log2n)
    ref = np.fft.fft(x.astype(np.complex128))
    k = freq_of_position(log2n)
    err = np.abs(X - ref[k]).max() / np.abs(ref).max()
    print(log2n, plan(log2n)[2:], "rel err", err)
