"""Resident-input throughput of the other BASELINE.json configs (not bench lines; recorded in DESIGN.md)."""
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
from synth import make_source
from soundspaces_b200 import AudioRequest, BatchedAudioRenderer

def timeit(fn, steps=100, warm=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps

rng = np.random.default_rng(0)
for B in (64, 512):
    sr, L = 16000, 48000
    r = BatchedAudioRenderer(sr, L)
    rirs = torch.from_numpy((rng.standard_normal((B, L, 2)) * 0.1).astype(np.float32)).cuda()
    ids = r.set_dense_rir_bank(rirs)
    s1, s4 = r.add_source(make_source(1, sr)), r.add_source(make_source(2, 4 * sr))
    out = torch.empty((B,) + r.spec_shape, device="cuda")
    head = r.prepare([AudioRequest(rir=i, source=s1) for i in ids])
    valid = r.prepare([AudioRequest(rir=i, source=s4, offset=3 * sr) for i in ids])
    for name, b in (("C3 head  (1-s clip, 16000 of 48000 taps matter)", head), ("C3 valid (4-s clip, all 48000 taps)", valid)):
        ms = timeit(lambda: r.execute(b, out=out))
        print(f"B={B:4d} {name}: {ms*1e3:8.1f} us/step  {B/ms*1e3:10.0f} frames/s", flush=True)
    del r
# C4: decode + conv chain, 64 envs (per-GPU share at 4 GPUs) and 256
for B in (64, 256):
    sr, L = 16000, 48000
    r = BatchedAudioRenderer(sr, L)
    amb = torch.randn((B, L, 9), device="cuda") * 0.05
    az = torch.tensor([0., 90., 180., 270.] * (B // 4))
    s4 = r.add_source(make_source(2, 4 * sr))
    out = torch.empty((B,) + r.spec_shape, device="cuda")
    def step():
        rirs = r.sh_decode(amb, az)
        ids = r.set_dense_rir_bank(rirs)
        r.execute(step.batch, out=out)
    rirs = r.sh_decode(amb, az); ids = r.set_dense_rir_bank(rirs)
    step.batch = r.prepare([AudioRequest(rir=i, source=s4, offset=3 * sr) for i in ids])
    ms = timeit(step, steps=20, warm=3)
    ms_dec = timeit(lambda: r.sh_decode(amb, az), steps=20, warm=3)
    print(f"B={B:4d} C4 (9-ch decode + valid-mode conv + spectrogram): {ms:8.3f} ms/step  {B/ms*1e3:9.0f} frames/s  (decode alone {ms_dec:.3f} ms)", flush=True)
    del r
