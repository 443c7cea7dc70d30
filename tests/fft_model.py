"""numpy (complex64) model of the per-thread schedule in ``soundspaces_b200/csrc/fft16.cuh`` and of the
spectrogram kernel's per-frame epilogue in ``csrc/ssb200.cu``: which thread holds which element in which pass,
the factored twiddles inside the 4x4 butterfly, the digit-permuted "slot order" of the spectrum, the radix-2/4
stage across lanes, the folded last radix-2 stage of the STFT, the ear unpacking and the 4x4 pooling by lane
groups.  Test infrastructure (tests/test_fft_model.py): it pins the index/twiddle ALGEBRA of the kernels on the
CPU; the arithmetic itself is checked against the oracle by the GPU parity tests."""
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

def forward(x, log2n, lane_stage=True):
    """fft_forward<LOG2N, LANE_STAGE>: returns v[t, i] = spectrum slot (t, i) (or, without the lane stage and
    M == 2, the P/Q halves: lane 2a holds P_a[i], lane 2a+1 holds Q_a[i])."""
    N, T, strides, M = plan(log2n); t = np.arange(T); i = np.arange(16)
    buf = x.astype(c64).copy()
    v = buf[pass_pos(t[:, None], i[None], strides[0])]
    v = fft16(v, tw6(strides[0], t), False, True)
    for p in range(1, len(strides)):
        buf[pass_pos(t[:, None], i[None], strides[p-1])] = v
        st = strides[p]; v = buf[pass_pos(t[:, None], i[None], st)]
        v = fft16(v, tw6(st, t % st), False, st > 1)
    return lanes_fwd(v, t, M) if lane_stage else v          # slots (t, i)

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



# ---------------------------------------------------------------------------------------------
# spectrogram kernel (csrc/ssb200.cu: spec_column + spectrogram_kernel epilogue)
# ---------------------------------------------------------------------------------------------
N_FFT, HOP, WIN, POOL = 512, 160, 400, 4


def nat_idx(k):
    return k + ((k >> 8) << 3)


def hann_padded():
    w = np.zeros(N_FFT)
    w[(N_FFT - WIN) // 2:(N_FFT - WIN) // 2 + WIN] = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(WIN) / WIN)
    return w.astype(np.float32)


def frame_magnitudes_folded(zl, zr):
    """One STFT frame as the kernel computes it with SPEC_FOLD_R2: packed FFT-512 of w*(L + iR) WITHOUT the last
    lane stage, P/Q halves stored to the exchange buffer, magnitudes of both ears from four reads per bin.
    zl, zr: the 512 (already padded) samples of the frame.  Returns (|XL[0..255]|, |XR[0..255]|, |XL[256]|, |XR[256]|)
    in the lane/register layout of the kernel: element [lane, m] is bin k = lane + 32 m."""
    w = hann_padded()
    x = (w * zl + 1j * (w * zr)).astype(c64)
    lane = np.arange(32)
    v = forward(x, 9, lane_stage=False)                                  # (32, 16)
    buf = np.zeros(N_FFT + 32, c64)
    for i in range(16):
        buf[nat_idx((lane >> 1) + 16 * i + 256 * (lane & 1))] = v[:, i]
    ml, mr = np.zeros((32, 8)), np.zeros((32, 8))
    for m in range(8):
        k = lane + 32 * m
        kk = (256 - k) & 255
        a = buf[nat_idx(k)] + buf[nat_idx(256 + k)]
        sgn = np.where((lane == 0) & (m == 0), 1.0, -1.0)
        bb = buf[nat_idx(kk)] + sgn * buf[nat_idx(256 + kk)]
        l, r = a + np.conj(bb), a - np.conj(bb)
        ml[:, m], mr[:, m] = np.abs(l), np.abs(r)                         # = 2 |XL[k]|, 2 |XR[k]|
    z256 = buf[nat_idx(0)] - buf[nat_idx(256)]
    return ml, mr, abs(z256.real), abs(z256.imag)


def pooled_column(yl, yr, col, pad_mode="reflect"):
    """One pooled spectrogram column (4 frames) -> (65, 2), following the kernel's accumulation, the two
    xor-shuffle pooling steps over 4-lane groups and its row mapping row = (lane >> 2) + 8 * (2 q + j)."""
    sr = len(yl)
    n_frames = 1 + sr // HOP
    pl, pr = (np.pad(y, N_FFT // 2, mode=pad_mode) for y in (yl, yr))
    accl, accr, a64l, a64r = np.zeros((32, 8)), np.zeros((32, 8)), 0.0, 0.0
    for fr in range(POOL):
        f = col * POOL + fr
        if f >= n_frames:
            continue                                                      # block_reduce pads with zeros
        ml, mr, l64, r64 = frame_magnitudes_folded(pl[f * HOP: f * HOP + N_FFT], pr[f * HOP: f * HOP + N_FFT])
        accl += ml; accr += mr; a64l += l64; a64r += r64
    out = np.zeros((65, 2))
    lane = np.arange(32)
    for acc, ear in ((accl, 0), (accr, 1)):
        s = acc.copy()
        s = s + s[lane ^ 1]
        s = s + s[lane ^ 2]                                               # every lane of a 4-lane group: group sum
        for ln in range(32):
            q = ln & 3
            for j in range(2):
                out[(ln >> 2) + 8 * (2 * q + j), ear] = np.log1p(s[ln, 2 * q + j] * (0.5 / 16.0))
    out[64] = np.log1p(np.array([a64l, a64r]) / 16.0)
    return out
