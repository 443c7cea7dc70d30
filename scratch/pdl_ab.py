"""A/B of programmatic dependent launch in the render chain: whole-step time with CUDA events,
debug flag 64 = plain stream-ordered launches.  Also checks the two produce identical output."""
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
from synth import make_rir, make_source
from soundspaces_b200 import AudioRequest, BatchedAudioRenderer
sr, L, B = 44100, 16384, 128
r = BatchedAudioRenderer(sr, L)
sid = r.add_source(make_source(7, sr))
rng = np.random.default_rng(0)
NB = 16
bank = torch.from_numpy((rng.standard_normal((NB * B, L, 2)) * 0.1).astype(np.float32)).cuda()
ids = r.set_dense_rir_bank(bank)
batches = [r.prepare([AudioRequest(rir=ids[k * B + i], source=sid) for i in range(B)]) for k in range(NB)]
def run(flags, streams, n=400):
    r.lib.ssb_set_debug(r.ctx.handle, flags)
    r.set_streams(streams)
    for i in range(30): r.execute(batches[i % NB])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n): out = r.execute(batches[i % NB])
    b.record(); torch.cuda.synchronize()
    spec = r.execute(batches[0])
    spec = (spec[0] if isinstance(spec, tuple) else spec).clone()
    return a.elapsed_time(b) / n * 1e3, spec
for streams in (1, 2):
    res = {}
    for rep in range(2):
        for flags, name in ((64, "plain"), (0, "pdl")):
            us, spec = run(flags, streams)
            res[name] = spec
            print(f"streams={streams} {name:6s} {us:7.1f} us/step", flush=True)
    print("identical:", bool(torch.equal(res["plain"], res["pdl"])), flush=True)
