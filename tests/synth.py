"""Seeded synthetic inputs shared by the oracle tests, the GPU parity tests and
bench.py: source = 0.1*N(0,1) float32; RIR =
N(0,1)*exp(-t/tau) float32 (L, 2), tau = L/6, scaled to max|rir| = 0.5."""
import numpy as np


def make_source(seed, n):
    rng = np.random.default_rng(1000 + seed)
    return (0.1 * rng.standard_normal(n)).astype(np.float32)


def make_rir(seed, taps, tau=None):
    rng = np.random.default_rng(1234 + seed)
    tau = taps / 6.0 if tau is None else tau
    r = rng.standard_normal((taps, 2)) * np.exp(-np.arange(taps) / tau)[:, None]
    r *= 0.5 / np.abs(r).max()
    return r.astype(np.float32)
