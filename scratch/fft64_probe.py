"""Round-2 prototype harness (needs a GPU): numerics and transforms/s of the 64 x 64 schedule of a 4096-point FFT
(scratch/fft64_probe.cu) against the radix-16 schedule of csrc/fft16.cuh, at the size of one inverse-FFT launch of the
bench step (64 envs x 22 blocks = 1408 transforms; 46 MB in + 46 MB out, L2 resident).

    nvcc -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
         -o scratch/libfft64_probe.so scratch/fft64_probe.cu        # in the build container
    gpurun -- 'python scratch/fft64_probe.py'
"""
import ctypes as C
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
lib = C.CDLL(os.path.join(HERE, "libfft64_probe.so"))
lib.probe_run.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
assert lib.probe_init() == 0
N, n = 4096, 1408
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.view_as_complex(torch.randn((n, N, 2), device="cuda", generator=g))
x[1] = 100 * torch.exp(2j * torch.pi * 37.3 * torch.arange(N, device="cuda") / N) + 1e-3 * x[1]     # loud tone + quiet noise
ref = torch.fft.fft(x.to(torch.complex128), dim=1)
y = torch.empty_like(x)
z = torch.empty_like(x)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def run(which, a, b):
    assert lib.probe_run(which, a.data_ptr(), b.data_ptr(), n, st) == 0


# ---- numerics
run(0, x, y)
err64 = ((y.to(torch.complex128) - ref).abs().amax(1) / ref.abs().amax(1)).max().item()
run(1, y, z)
rt64 = ((z / N - x).abs().amax(1) / x.abs().amax(1)).max().item()
freq = np.zeros(N, dtype=np.int32)
lib.probe_r16_slot_freq(freq.ctypes.data_as(C.c_void_p))
run(2, x, y)
perm = torch.from_numpy(freq.astype(np.int64)).cuda()
err16 = ((y.to(torch.complex128) - ref[:, perm]).abs().amax(1) / ref.abs().amax(1)).max().item()
run(3, y, z)
rt16 = ((z / N - x).abs().amax(1) / x.abs().amax(1)).max().item()
print("max error / peak: r64 fwd %.2e roundtrip %.2e | r16 fwd %.2e roundtrip %.2e" % (err64, rt64, err16, rt16))
assert err64 < 2e-6 and rt64 < 2e-6, "64 x 64 schedule is wrong"

# ---- throughput (L2-resident operands, like the live step)
for name, which, a, b in (("fwd r64", 0, x, y), ("inv r64", 1, y, z), ("fwd r16", 2, x, y), ("inv r16", 3, y, z)):
    for _ in range(20):
        run(which, a, b)
    torch.cuda.synchronize()
    ts = []
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100):
            run(which, a, b)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 100)
    t = sorted(ts)[2]
    print("%s: %.2f us per launch of %d transforms (%.1f M transforms/s)" % (name, 1e3 * t, n, n / t / 1e3))
