"""ctypes binding of ``libssb200.so`` (the C ABI declared in ``include/ssb200.h``).

The product path has NO CPU fallback: if the shared library is missing or a
call fails, a ``RuntimeError`` is raised.  ``build_library()`` compiles it
in-tree with ``nvcc`` for sm_100a (cross-compiles without a GPU).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("SSB200_LIB", os.path.join(_HERE, "libssb200.so"))   # override: A/B builds while tuning
SOURCES = [os.path.join(_HERE, "csrc", "ssb200.cu")]
HEADERS = [os.path.join(_HERE, "csrc", "fft16.cuh"), os.path.join(_HERE, "csrc", "conv64k.cuh"),
           os.path.join(ROOT, "include", "ssb200.h")]

NVCC_FLAGS = ["-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
              "-lineinfo", "-O3", "-std=c++17"]

SSB_FLAG_SILENT = 1
PAD_MODES = {"reflect": 0, "constant": 1}

# numpy mirror of ssb_conv_term / ssb_req (include/ssb200.h)
TERM_DTYPE = np.dtype([("rir_offset", "<i8"), ("x_offset", "<i8"), ("rir_taps", "<i4"),
                       ("x_nw", "<i4"), ("x_wofs", "<i4"), ("reserved", "<i4")], align=True)
REQ_DTYPE = np.dtype([("term", TERM_DTYPE, (2,)), ("out_samples", "<i4"), ("flags", "<u4")], align=True)
assert TERM_DTYPE.itemsize == 32 and REQ_DTYPE.itemsize == 72


class Plan(C.Structure):
    _fields_ = [("log2n", C.c_int32), ("block", C.c_int32), ("sr", C.c_int32), ("n_blocks", C.c_int32),
                ("max_parts", C.c_int32), ("n_terms", C.c_int32), ("h_elems_per_env", C.c_int64)]


EXPORTS = {
    # name: (restype, argtypes)
    "ssb_version": (C.c_int, []),
    "ssb_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "ssb_destroy": (None, [C.c_void_p]),
    "ssb_last_error": (C.c_char_p, [C.c_void_p]),
    "ssb_launch_count": (C.c_int64, [C.c_void_p]),
    "ssb_set_conv_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "ssb_set_streams": (C.c_int, [C.c_void_p, C.c_int]),
    "ssb_set_chunks": (C.c_int, [C.c_void_p, C.c_int]),
    "ssb_set_debug": (C.c_int, [C.c_void_p, C.c_int]),
    "ssb_set_kernel_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "ssb_get_kernel_timing": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "ssb_make_plan": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Plan)]),
    "ssb_spec_cols": (C.c_int, [C.c_int]),
    "ssb_source_windows": (C.c_int, [C.c_void_p, C.POINTER(Plan), C.c_void_p, C.c_int, C.c_int64, C.c_int,
                                     C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ssb_convolve_batch": (C.c_int, [C.c_void_p, C.POINTER(Plan), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ssb_crossfade_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int64,
                                      C.c_void_p, C.c_void_p]),
    "ssb_spectrogram_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p]),
    "ssb_render_batch": (C.c_int, [C.c_void_p, C.POINTER(Plan), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "ssb_render_batch_host": (C.c_int, [C.c_void_p, C.POINTER(Plan), C.c_int, C.c_void_p, C.c_void_p, C.c_int64,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                        C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ssb_host_copy_bytes": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "ssb_sh_decode_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "ssb_intensity_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ssb_audio_conv1_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ssb_logmel_frames": (C.c_int, [C.c_int]),
    "ssb_mel_filterbank": (C.c_int, [C.c_int, C.c_int, C.c_void_p]),
    "ssb_logmel_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p]),
    "ssb_pcm16_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "ssb_pcm16_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
}

_lib = None


def build_library(force=False, verbose=False):
    """Compile libssb200.so in-tree for sm_100a.  Returns the path."""
    deps = SOURCES + HEADERS
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(p) for p in deps)):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return LIB_PATH


def load_library():
    """dlopen libssb200.so and declare prototypes.  Fails loudly when missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the audio observation path)")
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in EXPORTS.items():
        fn = getattr(lib, name)        # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


class Context:
    """Owns one ``ssb_ctx`` (one per process and device)."""

    def __init__(self, device_index=0):
        self.lib = load_library()
        handle = C.c_void_p()
        rc = self.lib.ssb_create(int(device_index), C.byref(handle))
        self.handle = handle
        if rc != 0:
            msg = self.lib.ssb_last_error(handle).decode() if handle else "allocation failed"
            if handle:
                self.lib.ssb_destroy(handle)
                self.handle = None
            raise RuntimeError(f"ssb_create(device={device_index}) failed ({rc}): {msg}")

    def check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self.lib.ssb_last_error(self.handle).decode()}")

    def make_plan(self, sr, max_taps, n_terms=1, log2n=0):
        plan = Plan()
        self.check(self.lib.ssb_make_plan(self.handle, sr, max_taps, n_terms, log2n, C.byref(plan)), "ssb_make_plan")
        return plan

    KERNEL_NAMES = ("fwd_rir_kernel", "mac_ifft_kernel", "spectrogram_kernel", "fwd_src_kernel", "mac_bins_kernel",
                    "conv64k_kernel")

    def set_kernel_timing(self, enable):
        self.check(self.lib.ssb_set_kernel_timing(self.handle, int(bool(enable))), "ssb_set_kernel_timing")

    def get_kernel_timing(self):
        """{kernel: (summed ms, launches)} since the last call (synchronises)."""
        ms = (C.c_double * len(self.KERNEL_NAMES))()
        cnt = (C.c_int64 * len(self.KERNEL_NAMES))()
        self.check(self.lib.ssb_get_kernel_timing(self.handle, ms, cnt), "ssb_get_kernel_timing")
        return {n: (ms[i], cnt[i]) for i, n in enumerate(self.KERNEL_NAMES)}

    @property
    def launch_count(self):
        return int(self.lib.ssb_launch_count(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.ssb_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
