import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
from synth import make_source
from soundspaces_b200 import AudioRequest, BatchedAudioRenderer
rng = np.random.default_rng(0)
sr, L = 16000, 48000
for B in (64, 512):
    for log2n in (12, 13):
        r = BatchedAudioRenderer(sr, L, log2n=log2n)
        rirs = torch.from_numpy((rng.standard_normal((B, L, 2)) * 0.1).astype(np.float32)).cuda()
        ids = r.set_dense_rir_bank(rirs)
        s1, s4 = r.add_source(make_source(1, sr)), r.add_source(make_source(2, 4 * sr))
        out = torch.empty((B,) + r.spec_shape, device="cuda")
        for name, b in (("head", r.prepare([AudioRequest(rir=i, source=s1) for i in ids])),
                        ("valid", r.prepare([AudioRequest(rir=i, source=s4, offset=3 * sr) for i in ids]))):
            for _ in range(5): r.execute(b, out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50): r.execute(b, out=out)
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 50 * 1e3
            r.ctx.set_kernel_timing(True)
            for _ in range(20): r.execute(b, out=out)
            kt = r.ctx.get_kernel_timing(); r.ctx.set_kernel_timing(False)
            print(f"B={B} log2n={log2n} {name}: {us:.1f} us/step {B/us*1e6:.0f} frames/s", {k: round(v[0] / 20 * 1e3, 1) for k, v in kt.items() if v[1]}, flush=True)
        del r
