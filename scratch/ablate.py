import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
from synth import make_rir, make_source
from soundspaces_b200 import AudioRequest, BatchedAudioRenderer
sr, L, B = 44100, 16384, 128
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 13
r = BatchedAudioRenderer(sr, L, log2n=log2n)
r.set_conv_mode(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
r.lib.ssb_set_sub_batch(r.ctx.handle, int(sys.argv[3]) if len(sys.argv) > 3 else 0)
sid = r.add_source(make_source(7, sr))
rng = np.random.default_rng(0)
bank = torch.from_numpy((rng.standard_normal((8 * B, L, 2)) * 0.1).astype(np.float32)).cuda()
ids = r.set_dense_rir_bank(bank)
batches = [r.prepare([AudioRequest(rir=ids[k * B + i], source=sid) for i in range(B)]) for k in range(8)]
def run(flags, n=100):
    r.lib.ssb_set_debug(r.ctx.handle, flags)
    for i in range(10): r.execute(batches[i % 8])
    torch.cuda.synchronize()
    r.ctx.set_kernel_timing(True)
    for i in range(n): r.execute(batches[i % 8])
    kt = r.ctx.get_kernel_timing(); r.ctx.set_kernel_timing(False)
    return {k: round(v[0] / n * 1e3, 1) for k, v in kt.items() if v[1]}, 'launches/step', sum(v[1] for v in kt.values()) / n
for flags, name in [(0, "full"), (1, "no MAC"), (2, "no IFFT"), (3, "no MAC no IFFT"), (4, "spec: no loads"), (8, "spec: no FFT"), (12, "spec: no loads no FFT")]:
    print(f"{name:24s}", run(flags), flush=True)
