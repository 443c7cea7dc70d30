"""Record tests/golden/sh_golden.npz and soundspaces_b200/data/sh_hrtf_bank.npy by running the
reference's closed-source ``scripts/AmbisonicBinauralizer`` (build container only):

* the 9 x 2 x 256 filter bank = the tool's responses (lags 128..383) to a unit impulse in each
  spherical-harmonic channel at a block-aligned position, azimuth 0;
* its outputs on seeded random 9-channel signals at several azimuths (pins oracle/sh_oracle.py).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import sh_elf  # noqa: E402
from oracle.sh_oracle import SH_DELAY, SH_TAPS  # noqa: E402

CASES = {"az0": (0.0, 11), "az90": (90.0, 12), "az180": (180.0, 13), "az270": (270.0, 14), "az37": (37.5, 15)}
N = 3072


def make_amb(seed, n=N):
    rng = np.random.default_rng(seed)
    env = np.exp(-np.arange(n) / (n / 5.0))[:, None]
    return (rng.standard_normal((n, 9)) * env * 0.3).astype(np.float32)


def main():
    p = 256
    bank = np.zeros((9, 2, SH_TAPS), dtype=np.float32)
    for k in range(9):
        imp = np.zeros((1024, 9), np.float32)
        imp[p, k] = 1.0
        y = sh_elf.binauralize(imp, 0.0)
        bank[k] = y[p + SH_DELAY: p + SH_DELAY + SH_TAPS].T
        assert not y[: p + SH_DELAY].any()
    out = {"bank": bank}
    for name, (az, seed) in CASES.items():
        out[f"{name}/out"] = sh_elf.binauralize(make_amb(seed), az)
        out[f"{name}/meta"] = np.array([az, seed])
    # block-aligned impulse at another position / azimuth: exact LTI check
    imp = np.zeros((1024, 9), np.float32)
    imp[384, 3] = 1.0
    out["imp_az90/out"] = sh_elf.binauralize(imp, 90.0)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sh_golden.npz"), **out)
    os.makedirs(os.path.join(ROOT, "soundspaces_b200", "data"), exist_ok=True)
    np.save(os.path.join(ROOT, "soundspaces_b200", "data", "sh_hrtf_bank.npy"), bank)
    print("bank energy per channel", (bank[:, 0] ** 2).sum(axis=1).round(3))


if __name__ == "__main__":
    main()
