"""Steady-state run of the bench workload (rotating RIR banks) for traffic profiling: argv = streams chunks."""
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
from synth import make_source
from soundspaces_b200 import AudioRequest, BatchedAudioRenderer
sr, L, B, NB = 44100, 16384, 128, 16
g = torch.Generator(device="cuda").manual_seed(1)
env = torch.exp(-torch.arange(L, device="cuda") / (L / 6.0))[None, :, None]
bank = (torch.randn((NB * B, L, 2), device="cuda", generator=g) * env * 0.1).contiguous()
r = BatchedAudioRenderer(sr, L)
r.set_streams(int(sys.argv[1]) if len(sys.argv) > 1 else 2)
r.set_chunks(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
sid = r.add_source(make_source(7, sr))
ids = r.set_dense_rir_bank(bank)
batches = [r.prepare([AudioRequest(rir=ids[k * B + i], source=sid) for i in range(B)]) for k in range(NB)]
out = torch.empty((B,) + r.spec_shape, device="cuda")
for i in range(12):
    r.execute(batches[i % NB], out=out)
torch.cuda.synchronize()
