"""numpy (complex64) emulation of fft16.cuh v2: factored twiddles inside the 4x4 butterfly,
forward DIF + inverse DIT, to validate the index/twiddle algebra before running on the GPU."""
import numpy as np
c64 = np.complex64

def plan(log2n):
    N = 1 << log2n; npass = log2n // 4
    strides = [N >> (4 * (p + 1)) for p in range(npass)]
    return N, N // 16, strides, N >> (4 * npass)

def pass_pos(t, i, st): return (t // st) * 16 * st + (t % st) + i * st

def tw6(st, j):
    e = np.array([1, 2, 3, 4, 8, 12])
    return np.exp(-2j * np.pi * np.outer(j, e) / (16 * st)).astype(c64)   # (T,6)

def bfly4(a, inv):
    a0, a1, a2, a3 = a
    t0, t1, t2, t3 = a0 + a2, a0 - a2, a1 + a3, a1 - a3
    r = (1j if inv else -1j) * t3
    return [t0 + t2, t1 + r, t0 - t2, t1 - r]

W16 = lambda k, inv: c64(np.exp((2j if inv else -2j) * np.pi * k / 16))

def fft16(v, w, inv, tw):
    v = [v[:, k].astype(c64) for k in range(16)]
    if inv and tw:
        for m in range(1, 4):
            for r in range(4): v[r + 4*m] = (v[r + 4*m] * np.conj(w[:, 3 + m - 1])).astype(c64)
    for c in range(4):
        v[c], v[c+4], v[c+8], v[c+12] = bfly4([v[c], v[c+4], v[c+8], v[c+12]], inv)
    for c in range(1, 4):
        for r in range(1, 4): v[c + 4*r] = (v[c + 4*r] * W16(c*r, inv)).astype(c64)
    if tw:
        if not inv:
            for r in range(1, 4):
                for c in range(4): v[c + 4*r] = (v[c + 4*r] * w[:, r - 1]).astype(c64)
        else:
            for c in range(1, 4):
                for r in range(4): v[c + 4*r] = (v[c + 4*r] * np.conj(w[:, c - 1])).astype(c64)
    for r in range(4):
        v[4*r], v[4*r+1], v[4*r+2], v[4*r+3] = bfly4([v[4*r], v[4*r+1], v[4*r+2], v[4*r+3]], inv)
    if tw and not inv:
        for m in range(1, 4):
            for r in range(4): v[4*r + m] = (v[4*r + m] * w[:, 3 + m - 1]).astype(c64)
    return np.stack([v[4*(s & 3) + (s >> 2)] for s in range(16)], axis=1).astype(c64)

def lanes_fwd(v, t, M):
    if M == 2:
        o = v.reshape(-1, 2, 16)[:, ::-1].reshape(-1, 16); up = (t & 1)[:, None].astype(bool)
        return np.where(up, o - v, v + o).astype(c64)
    if M == 4:
        j = t & 3
        o = v[t ^ 2]; r = np.where((j & 2)[:, None].astype(bool), o - v, v + o); r = np.where((j == 3)[:, None], -1j * r, r).astype(c64)
        o = r[t ^ 1]; return np.where((j & 1)[:, None].astype(bool), o - r, r + o).astype(c64)
    return v

def lanes_inv(v, t, M):
    if M == 2: return lanes_fwd(v, t, 2)
    if M == 4:
        j = t & 3
        o = v[t ^ 1]; r = np.where((j & 1)[:, None].astype(bool), o - v, v + o); r = np.where((j == 3)[:, None], 1j * r, r).astype(c64)
        o = r[t ^ 2]; return np.where((j & 2)[:, None].astype(bool), o - r, r + o).astype(c64)
    return v

def forward(x, log2n):
    N, T, strides, M = plan(log2n); t = np.arange(T); i = np.arange(16)
    buf = x.astype(c64).copy()
    v = buf[pass_pos(t[:, None], i[None], strides[0])]
    v = fft16(v, tw6(strides[0], t), False, True)
    for p in range(1, len(strides)):
        buf[pass_pos(t[:, None], i[None], strides[p-1])] = v
        st = strides[p]; v = buf[pass_pos(t[:, None], i[None], st)]
        v = fft16(v, tw6(st, t % st), False, st > 1)
    return lanes_fwd(v, t, M)          # slots (t, i)

def inverse(v, log2n):
    N, T, strides, M = plan(log2n); t = np.arange(T); i = np.arange(16)
    buf = np.zeros(N, c64)
    v = lanes_inv(v, t, M)
    for p in range(len(strides) - 1, 0, -1):
        st = strides[p]
        v = fft16(v, tw6(st, t % st), True, st > 1)
        buf[pass_pos(t[:, None], i[None], st)] = v
        v = buf[pass_pos(t[:, None], i[None], strides[p-1])]
    v = fft16(v, tw6(strides[0], t), True, True)
    out = np.zeros(N, c64); out[pass_pos(t[:, None], i[None], strides[0])] = v
    return out

def freq_of_slot(log2n):
    N, T, strides, M = plan(log2n); t = np.arange(T)[:, None]; i = np.arange(16)[None]
    pos = pass_pos(t, i, strides[-1]); k = np.zeros_like(pos); mult = 1; rem = pos.copy()
    for st in strides:
        k += (rem // st) * mult; rem = rem % st; mult *= 16
    if M == 2: k += rem * mult
    if M == 4: k += (((rem & 1) << 1) | (rem >> 1)) * mult
    return k

for log2n in (9, 12, 13, 14):
    N = 1 << log2n; rng = np.random.default_rng(0)
    x = (rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(c64)
    X = forward(x, log2n); ref = np.fft.fft(x.astype(np.complex128)); k = freq_of_slot(log2n)
    e1 = np.abs(X - ref[k]).max() / np.abs(ref).max()
    y = inverse(X, log2n) / N
    e2 = np.abs(y - x).max() / np.abs(x).max()
    # high dynamic range check: strong tone + weak noise
    x2 = (100 * np.exp(2j * np.pi * 37.3 * np.arange(N) / N) + 1e-3 * rng.standard_normal(N)).astype(c64)
    X2 = forward(x2, log2n); ref2 = np.fft.fft(x2.astype(np.complex128))
    e3 = np.abs(X2 - ref2[k]).max() / np.abs(ref2).max()
    print(log2n, "fwd rel err", f"{e1:.2e}", "roundtrip", f"{e2:.2e}", "tone leak", f"{e3:.2e}")
