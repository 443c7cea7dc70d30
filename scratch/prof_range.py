"""Live-step DRAM traffic: one steady-state step of the bench workload inside a cudaProfilerStart/Stop range.

    ncu --replay-mode range --cache-control none --clock-control none \
        --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum \
        --csv --log-file gpurun_out/live_traffic.csv python scratch/prof_range.py [steps_in_range]

Range replay keeps the launches of the range together (both internal streams run concurrently, caches are not
flushed), and with three metrics one pass suffices, so the bytes are those of a live step -- unlike the per-kernel
captures, whose replays serialise producer and consumer through DRAM (profiles/README.md)."""
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
from synth import make_source
from soundspaces_b200 import AudioRequest, BatchedAudioRenderer
sr, L, B, NB = 44100, 16384, 128, 16
n_in = int(sys.argv[1]) if len(sys.argv) > 1 else 1
g = torch.Generator(device="cuda").manual_seed(1)
env = torch.exp(-torch.arange(L, device="cuda") / (L / 6.0))[None, :, None]
bank = (torch.randn((NB * B, L, 2), device="cuda", generator=g) * env * 0.1).contiguous()
r = BatchedAudioRenderer(sr, L)
sid = r.add_source(make_source(7, sr))
ids = r.set_dense_rir_bank(bank)
sil = np.random.default_rng(99).random(B) < 0.05
batches = [r.prepare([AudioRequest(rir=ids[k * B + i], source=sid, silent=bool(sil[i])) for i in range(B)]) for k in range(NB)]
out = torch.empty((B,) + r.spec_shape, device="cuda")
for i in range(40):
    r.execute(batches[i % NB], out=out)
torch.cuda.synchronize()
rt = torch.cuda.cudart()
rt.cudaProfilerStart()
for i in range(40, 40 + n_in):
    r.execute(batches[i % NB], out=out)
torch.cuda.synchronize()
rt.cudaProfilerStop()
print("range done: %d step(s), algorithmic bytes per step = %d" % (n_in, B * (8 * L + 4 * sr // B + 8 * 65 * 69)))
