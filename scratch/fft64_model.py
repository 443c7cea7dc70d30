"""Register-level numpy emulation of scratch/fft64_probe.cu (fft8 / fft64 codelets, twiddles, the two transposes, the
store layouts): forward == np.fft.fft in natural order, inverse(forward(x)) / N == x.  Run on the CPU."""
import numpy as np


def bfly4(a, inv):
    a0, a1, a2, a3 = a
    t0, t1, t2, t3 = a0 + a2, a0 - a2, a1 + a3, a1 - a3
    r = (1j if inv else -1j) * t3
    return [t0 + t2, t1 + r, t0 - t2, t1 - r]


def fft8(a, inv):
    R2 = np.sqrt(0.5)
    mi = lambda z: (1j if inv else -1j) * z
    b0, b4 = a[0] + a[4], a[0] - a[4]; b1, b5 = a[1] + a[5], a[1] - a[5]
    b2, b6 = a[2] + a[6], a[2] - a[6]; b3, b7 = a[3] + a[7], a[3] - a[7]
    b5 = (b5 + mi(b5)) * R2; b6 = mi(b6); b7 = (mi(b7) - b7) * R2
    b0, b1, b2, b3 = bfly4([b0, b1, b2, b3], inv); b4, b5, b6, b7 = bfly4([b4, b5, b6, b7], inv)
    return [b0, b4, b1, b5, b2, b6, b3, b7]


def fft64(v, inv):
    v = list(v)
    for q1 in range(8):
        idx = [q1 + 8 * j for j in range(8)]
        for i, o in zip(idx, fft8([v[i] for i in idx], inv)): v[i] = o
    for q1 in range(1, 8):
        for s1 in range(1, 8):
            k = (q1 * s1) & 63
            v[q1 + 8 * s1] = v[q1 + 8 * s1] * np.exp((2j if inv else -2j) * np.pi * k / 64)
    for s1 in range(8):
        idx = [8 * s1 + j for j in range(8)]
        for i, o in zip(idx, fft8([v[i] for i in idx], inv)): v[i] = o
    return v


N = 4096
t = np.arange(64)
tw = np.exp(-2j * np.pi * np.outer(t, np.arange(64)) / N)


def fwd_r64(x):
    v = fft64([x[t + 64 * q] for q in range(64)], False)
    buf = np.zeros((64, 65), complex)
    for s1 in range(8):
        for s2 in range(8):
            s = s1 + 8 * s2
            buf[s, t] = v[8 * s1 + s2] * (tw[t, s] if s else 1.0)
    v = fft64([buf[t, i] for i in range(64)], False)
    dst = np.zeros(N, complex)
    for r1 in range(8):
        for r2 in range(8): dst[(r1 + 8 * r2) * 64 + t] = v[8 * r1 + r2]
    return dst


def inv_r64(X):
    s = t
    v = fft64([X[r * 64 + s] for r in range(64)], True)
    buf = np.zeros((64, 65), complex)
    for t1 in range(8):
        for t2 in range(8):
            tt = t1 + 8 * t2
            buf[tt, s] = v[8 * t1 + t2] * (np.conj(tw[tt, s]) if tt else 1.0)
    v = fft64([buf[s, i] for i in range(64)], True)
    out = np.zeros(N, complex)
    for q1 in range(8):
        for q2 in range(8): out[s + 64 * (q1 + 8 * q2)] = v[8 * q1 + q2]
    return out


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    X = fwd_r64(x)
    print("forward vs np.fft.fft (natural order): %.2e" % np.abs(X - np.fft.fft(x)).max())
    print("inverse(forward(x)) / N vs x:          %.2e" % np.abs(inv_r64(X) / N - x).max())
