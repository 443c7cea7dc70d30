"""Driver for the reference's closed-source ``scripts/AmbisonicBinauralizer`` (x86-64 ELF, Oculus
HRTF library linked in; ``scripts/ambisonic_to_binaural.py:14-19``).

TEST INFRASTRUCTURE ONLY, build container only (the GPU box has no /root/reference).  Used by
``tests/golden/make_sh_golden.py`` to record (a) the tool's 9 x 2 bank of 256-tap filters as its
responses to unit impulses in each spherical-harmonic channel and (b) its outputs on seeded random
ambisonic signals, which pin ``oracle/sh_oracle.py``.
"""
from __future__ import annotations

import os
import struct
import subprocess
import tempfile

import numpy as np

ELF = os.path.join(os.environ.get("SOUNDSPACES_REFERENCE", "/root/reference"), "scripts", "AmbisonicBinauralizer")


def elf_available():
    return os.path.isfile(ELF) and os.access(ELF, os.X_OK)


def write_wav_f32(path, data, sr):
    """Plain 16-byte fmt chunk, format tag 3 (IEEE float): the tool rejects WAVE_FORMAT_EXTENSIBLE,
    which scipy writes for > 2 channels."""
    data = np.ascontiguousarray(data, dtype="<f4")
    n, ch = data.shape
    payload = data.tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(payload)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, 3, ch, sr, sr * ch * 4, ch * 4, 32))
        f.write(b"data" + struct.pack("<I", len(payload)) + payload)


def read_wav_f32(path):
    from scipy.io import wavfile
    sr, x = wavfile.read(path)
    return sr, np.asarray(x, dtype=np.float32)


def binauralize(amb, azimuth_deg, sr=16000):
    """Run the tool on one (n, 9) float32 ambisonic signal; returns (n, 2) float32."""
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "in")
        dst = os.path.join(d, "out")
        os.makedirs(src)
        os.makedirs(dst)
        write_wav_f32(os.path.join(src, "x.wav"), amb, sr)
        res = subprocess.run([ELF, "-i", src, "-o", dst, "-a", str(azimuth_deg)], capture_output=True, text=True)
        out = os.path.join(dst, "x.wav")
        if not os.path.exists(out):
            raise RuntimeError("AmbisonicBinauralizer produced no output: " + res.stdout + res.stderr)
        _, y = read_wav_f32(out)
    return y
