import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
from synth import make_rir, make_source
from soundspaces_b200 import AudioRequest, BatchedAudioRenderer
sr, L, B = 44100, 16384, 128
r = BatchedAudioRenderer(sr, L)
sid = r.add_source(make_source(7, sr))
rng = np.random.default_rng(0)
bank = torch.from_numpy((rng.standard_normal((B, L, 2)) * 0.1).astype(np.float32)).cuda()
ids = r.set_dense_rir_bank(bank)
batch = r.prepare([AudioRequest(rir=ids[i], source=sid) for i in range(B)])
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
r.lib.ssb_set_debug(r.ctx.handle, flags)
for i in range(4): r.execute(batch)
spec, wave = r.execute(batch, want_wave=True)
r.logmel(wave[:64])                      # the log-mel extension kernel, same 64-env launch size as the others
torch.cuda.synchronize()
