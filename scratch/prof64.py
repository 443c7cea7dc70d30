"""Short steady-state run for ncu: argv[1] = streams (default 2), argv[2] = envs (default 128)."""
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
from synth import make_source
from soundspaces_b200 import AudioRequest, BatchedAudioRenderer
sr, L, NB = 44100, 16384, 4
streams = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
g = torch.Generator(device="cuda").manual_seed(1)
env = torch.exp(-torch.arange(L, device="cuda") / (L / 6.0))[None, :, None]
bank = (torch.randn((NB * B, L, 2), device="cuda", generator=g) * env * 0.1).contiguous()
r = BatchedAudioRenderer(sr, L)
r.set_streams(streams)
sid = r.add_source(make_source(7, sr))
ids = r.set_dense_rir_bank(bank)
batches = [r.prepare([AudioRequest(rir=ids[k * B + i], source=sid) for i in range(B)]) for k in range(NB)]
out = torch.empty((B,) + r.spec_shape, device="cuda")
for i in range(int(os.environ.get("PROF_STEPS", 8))):
    r.execute(batches[i % NB], out=out)
torch.cuda.synchronize()
if os.environ.get("PROF_TIME"):
    for st in (1, 2, 3):
        r.set_streams(st)
        for i in range(20):
            r.execute(batches[i % NB], out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(200):
            r.execute(batches[i % NB], out=out)
        e1.record(); torch.cuda.synchronize()
        r.ctx.set_kernel_timing(True)
        for i in range(50):
            r.execute(batches[i % NB], out=out)
        kt = r.ctx.get_kernel_timing(); r.ctx.set_kernel_timing(False)
        print("streams=%d B=%d: %.2f us/step | per-step kernel sums (us): %s" % (st, B, e0.elapsed_time(e1) * 5, {k: round(1e3 * v[0] / 50, 1) for k, v in kt.items() if v[1]}), flush=True)
