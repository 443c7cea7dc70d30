"""Small run of every kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
from synth import make_rir, make_source
from soundspaces_b200 import AudioRequest, BatchedAudioRenderer
for log2n in (12, 13, 14):
    for mode in (0, 1):
        sr = 16000
        r = BatchedAudioRenderer(sr, 20000, n_terms=2, log2n=log2n)
        r.set_conv_mode(mode); r.set_streams(1)
        s1 = r.add_source(make_source(1, sr)); s2 = r.add_source(make_source(2, 3 * sr))
        ids = r.add_rirs([make_rir(i, L) for i, L in enumerate((100, 4097, 9000, 20000))] + [None])
        reqs = [AudioRequest(rir=ids[0], source=s1), AudioRequest(rir=ids[1], source=s2, offset=sr),
                AudioRequest(rir=ids[2], source=s2, offset=2 * sr, distractor_rir=ids[1], distractor_source=s1),
                AudioRequest(rir=ids[3], source=s2, offset=8000, out_samples=4000, wrap=True),
                AudioRequest(rir=ids[4], source=s1), AudioRequest(rir=ids[0], source=s1, silent=True)]
        spec, wave = r.render(reqs, want_wave=True)
        r.render_crossfade(reqs[:2], [reqs[1], None])
        r.intensity(wave)
        torch.cuda.synchronize()
r = BatchedAudioRenderer(16000, 4096)
r.sh_decode(torch.randn(2, 1500, 9), [10.0, 200.0])          # direct form
r.sh_decode(torch.randn(2, 7600, 9), [10.0, 200.0])          # overlap-save FFT form
r.set_streams(2)
sid = r.add_source(make_source(3, 16000))
ids = r.add_rirs([make_rir(i, 3000) for i in range(64)])
r.render([AudioRequest(rir=i, source=sid) for i in ids])
r.set_chunks(2)
spec, wave = r.render([AudioRequest(rir=i, source=sid, silent=(i % 9 == 0)) for i in ids], want_wave=True)   # 2 streams x 2 chunks
r.set_chunks(1)
for pm in ("reflect", "constant"):                                       # log-mel extension kernel, both pad modes / powers
    rm = BatchedAudioRenderer(16000, 1024, pad_mode=pm)
    for n_mels, power in ((64, 2), (13, 1), (1, 2)):
        rm.logmel(wave[:5].clone(), n_mels=n_mels, power=power)
torch.cuda.synchronize()
# single-block plan: the 65536-point cluster kernel (named group barriers, in-place staging, DSMEM combine) + its
# source-spectrum kernel: short / long (> 16384 taps: q2 loop) / zero RIRs, offsets, wrap, truncated outputs, silent
for sr in (16000, 44100):
    rb = BatchedAudioRenderer(sr, min(48000, 65536 - sr + 1), prefer_block64=True)
    s1 = rb.add_source(make_source(1, sr)); s2 = rb.add_source(make_source(2, 3 * sr))
    ids = rb.add_rirs([make_rir(i, L) for i, L in enumerate((100, 4097, 16384, min(48000, 65536 - sr + 1)))] + [None])
    reqs = [AudioRequest(rir=ids[0], source=s1), AudioRequest(rir=ids[1], source=s2, offset=sr),
            AudioRequest(rir=ids[3], source=s2, offset=2 * sr), AudioRequest(rir=ids[2], source=s2, offset=8000, out_samples=4000, wrap=True),
            AudioRequest(rir=ids[4], source=s1), AudioRequest(rir=ids[0], source=s1, silent=True)]
    batch = rb.prepare(reqs)
    assert batch.plan.log2n == 16
    rb.execute(batch, want_wave=True)
    torch.cuda.synchronize()
# fused first layer of the audio encoder
from soundspaces_b200.renderer import WaveformOps
for shape, k, st in (((3, 65, 69, 2), 8, 4), ((3, 65, 26, 2), 5, 2)):
    conv = torch.nn.Conv2d(2, 32, k, st).cuda()
    WaveformOps.get("cuda:0").audio_conv1(torch.rand(shape, device="cuda"), conv)
torch.cuda.synchronize()
hs = r.make_host_session(8, 3000, want_wave=True, n_chunks=2)
hs.h_rir.numpy()[:] = np.stack([make_rir(i, 3000) for i in range(8)]); hs.set_requests(sid); hs.run(); hs.run()
torch.cuda.synchronize()
print("done")
