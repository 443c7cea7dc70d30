import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
from synth import make_source
from soundspaces_b200 import AudioRequest, BatchedAudioRenderer
sr, L, B, NB = 44100, 16384, 128, 8
rng = np.random.default_rng(0)
bank = torch.from_numpy((rng.standard_normal((NB * B, L, 2)) * 0.1).astype(np.float32)).cuda()
r = BatchedAudioRenderer(sr, L)
sid = r.add_source(make_source(7, sr))
ids = r.set_dense_rir_bank(bank)
batches = [r.prepare([AudioRequest(rir=ids[k * B + i], source=sid) for i in range(B)]) for k in range(NB)]
out = torch.empty((B,) + r.spec_shape, device="cuda")
for streams in (1, 2):
    r.set_streams(streams)
    for i in range(20): r.execute(batches[i % NB], out=out)
    torch.cuda.synchronize()
    steps = 300
    t0 = time.perf_counter()
    for i in range(steps): r.execute(batches[i % NB], out=out)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"streams={streams}: CPU enqueue {1e6*(t1-t0)/steps:.1f} us/step, wall incl. drain {1e6*(t2-t0)/steps:.1f} us/step")
