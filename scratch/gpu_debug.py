"""First-contact debug script: prints error statistics piece by piece."""
import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
from oracle import audio_oracle as ao
from synth import make_rir, make_source
from soundspaces_b200 import AudioRequest, BatchedAudioRenderer

def stats(name, got, ref):
    got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
    err = np.abs(got - ref)
    print(f"{name}: max|err|={err.max():.3e} peak={np.abs(ref).max():.3e} rel={err.max()/max(np.abs(ref).max(),1e-30):.3e} argmax={np.unravel_index(err.argmax(), err.shape)} nan={np.isnan(got).any()}", flush=True)

print(torch.cuda.get_device_name(0))
for sr in (16000, 44100):
    r = BatchedAudioRenderer(sr, 48000, n_terms=2)
    rng = np.random.default_rng(1)
    w = rng.standard_normal((2, 2, sr)).astype(np.float32)
    got = r.spectrogram(torch.from_numpy(w).cuda()).cpu().numpy()
    for pm in ("reflect",):
        ref = ao.compute_spectrogram(w[0], pad_mode=pm)
        stats(f"spec sr={sr}", got[0], ref)
        # column-wise error to localise
        e = np.abs(got[0] - ref).max(axis=(0, 2)); print("  col err", np.round(e[:4], 6), np.round(e[-3:], 6))
        e = np.abs(got[0] - ref).max(axis=(1, 2)); print("  row err", np.round(e[:4], 6), np.round(e[-3:], 6))
for log2n in (13, 12, 14):
    sr = 16000
    r = BatchedAudioRenderer(sr, 48000, n_terms=2, log2n=log2n)
    src = make_source(0, sr); sid = r.add_source(src)
    for L in (100, 4096, 9000, 20000):
        rir = make_rir(L, L); rid = r.add_rirs([rir])[0]
        wave = r.convolve([AudioRequest(rir=rid, source=sid)]).cpu().numpy()
        ref = ao.compute_audiogoal(src, rir, sr)
        stats(f"conv log2n={log2n} L={L}", wave[0], ref)
        e = np.abs(wave[0] - ref).max(axis=0)
        blk = r.P
        print("   per-block err", [float(f"{e[i*blk:(i+1)*blk].max():.2e}") for i in range(-(-sr // blk))])
# timing quick look (C2)
sr, L, B = 44100, 16384, 128
r = BatchedAudioRenderer(sr, L)
sid = r.add_source(make_source(7, sr))
rirs = torch.from_numpy(np.stack([make_rir(i, L) for i in range(B)])).cuda()
ids = r.set_dense_rir_bank(rirs)
batch = r.prepare([AudioRequest(rir=i, source=sid) for i in ids])
for _ in range(5): r.execute(batch)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(50): r.execute(batch)
ev[1].record(); torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 50
print(f"C2 B=128: {ms*1e3:.1f} us/step -> {B/ms*1e3:.0f} frames/s")
