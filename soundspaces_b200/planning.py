"""Pure host-side arithmetic of the overlap-save plan (no CUDA, no torch): which source windows
a request needs and how many RIR taps can influence its output.  Shared by the renderer and by
the CPU tests that model the kernels' dataflow (tests/kernel_model.py)."""
from __future__ import annotations


def ceil_div(a: int, b: int) -> int:
    return -(-a // b)


def window_layout(block: int, max_parts: int, offset: int, out_samples: int):
    """Windows of the source needed to render ``out_samples`` outputs starting at source time
    ``offset`` (``out[m] = sum_k h[k] x[offset + m - k]``).

    Overlap-save block b (outputs [b*P, (b+1)*P)) and RIR partition p use the window with block
    index beta = b - p, covering source samples [offset + (beta-1)P, offset + (beta+1)P).  Windows
    with beta < -ceil(offset/P) lie entirely before the clip and are never stored.
    Returns (n_out_blocks, wofs, nw): stored window j holds beta = j - wofs.
    """
    nblk = ceil_div(out_samples, block)
    wofs = min(max_parts - 1, ceil_div(offset, block))
    return nblk, wofs, nblk + wofs


def effective_taps(rir_len: int, offset: int, out_samples: int) -> int:
    """Only taps k <= offset + m can multiply a non-zero sample: simulator.py:629-632 keeps the
    first sr outputs of the full convolution, so a 1-s clip uses min(L, sr) taps (exact)."""
    return min(rir_len, offset + out_samples)


def partition_range(b: int, nparts: int, nw: int, wofs: int):
    """Inclusive range of RIR partitions that block b accumulates (mac_ifft_kernel)."""
    return max(0, b + wofs - (nw - 1)), min(nparts - 1, b + wofs)


def shard_envs(n_envs: int, rank: int, world: int):
    """env i -> rank i mod world; returns this rank's env indices."""
    return list(range(rank, n_envs, world))
