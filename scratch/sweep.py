import sys, os, itertools
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
from synth import make_source
from soundspaces_b200 import AudioRequest, BatchedAudioRenderer
sr, L, B, NB = 44100, 16384, 128, 8
rng = np.random.default_rng(0)
bank = torch.from_numpy((rng.standard_normal((NB * B, L, 2)) * 0.1).astype(np.float32)).cuda()
src = make_source(7, sr)
def run(log2n, mode, streams, steps=200):
    r = BatchedAudioRenderer(sr, L, log2n=log2n)
    r.set_conv_mode(mode); r.set_streams(streams)
    sid = r.add_source(src)
    ids = r.set_dense_rir_bank(bank)
    batches = [r.prepare([AudioRequest(rir=ids[k * B + i], source=sid) for i in range(B)]) for k in range(NB)]
    out = torch.empty((B,) + r.spec_shape, device="cuda")
    for i in range(20): r.execute(batches[i % NB], out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps): r.execute(batches[i % NB], out=out)
    e1.record(); torch.cuda.synchronize()
    ref = out.clone()
    del r
    return e0.elapsed_time(e1) / steps * 1e3, ref
base = None
for log2n, mode in ((13, 1), (12, 0), (13, 0), (12, 1)):
    for streams in (1, 2, 3, 4, 8):
        us, out = run(log2n, mode, streams)
        if base is None: base = out
        print(f"log2n={log2n} mode={mode} streams={streams}: {us:7.1f} us/step  {B/us*1e6:9.0f} frames/s  maxdiff={float((out-base).abs().max()):.2e}", flush=True)
