"""A/B of alternative builds of libssb200.so on the bench workload (C2: 128 envs x 44.1 kHz x 16384 taps).

    python scratch/ab_libs.py lib_a.so lib_b.so ...        # parent: one subprocess per library (SSB200_LIB)
    python scratch/ab_libs.py --child                      # child: measures the library named by SSB200_LIB

Per library: ms/step (median of 5 x 200 steps, CUDA events) for 1..4 internal streams, the per-kernel event
sums with 2 streams, and a checksum of the spectrograms (all builds must agree to rounding)."""
import os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child():
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np, torch
    from synth import make_source
    from soundspaces_b200 import AudioRequest, BatchedAudioRenderer
    sr, L, B, NB = 44100, 16384, 128, 16
    g = torch.Generator(device="cuda").manual_seed(1)
    env = torch.exp(-torch.arange(L, device="cuda") / (L / 6.0))[None, :, None]
    bank = (torch.randn((NB * B, L, 2), device="cuda", generator=g) * env * 0.1).contiguous()
    r = BatchedAudioRenderer(sr, L)
    sid = r.add_source(make_source(7, sr))
    ids = r.set_dense_rir_bank(bank)
    sil = np.random.default_rng(99).random(B) < 0.05
    batches = [r.prepare([AudioRequest(rir=ids[k * B + i], source=sid, silent=bool(sil[i])) for i in range(B)])
               for k in range(NB)]
    out = torch.empty((B,) + r.spec_shape, device="cuda")
    res = {}
    cfgs = os.environ.get("AB_CONFIGS")
    if cfgs:                                   # "2x1,2x2,3x2": streams x chunks-per-stream
        out_s = []
        for c in cfgs.split(","):
            st, ch = (int(v) for v in c.split("x"))
            r.set_streams(st); r.set_chunks(ch)
            for i in range(30):
                r.execute(batches[i % NB], out=out)
            torch.cuda.synchronize()
            ts = []
            for rep in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(200):
                    r.execute(batches[i % NB], out=out)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) / 200)
            out_s.append("%s=%.4f" % (c, sorted(ts)[2]))
        print("AB %-24s ms/step %s | chk=%.6f" % (os.path.basename(os.environ.get("SSB200_LIB", "default")), " ".join(out_s),
                                                  float(out.double().sum())), flush=True)
        return
    for streams in (2, 1, 3, 4):
        r.set_streams(streams)
        for i in range(30):
            r.execute(batches[i % NB], out=out)
        torch.cuda.synchronize()
        ts = []
        for rep in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(200):
                r.execute(batches[i % NB], out=out)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 200)
        res[streams] = sorted(ts)[2]
    r.set_streams(2)
    r.execute(batches[0], out=out)
    torch.cuda.synchronize()
    chk = float(out.double().sum())
    r.ctx.set_kernel_timing(True)
    for i in range(100):
        r.execute(batches[i % NB], out=out)
    kt = r.ctx.get_kernel_timing()
    r.ctx.set_kernel_timing(False)
    ks = " ".join("%s=%.1f" % (k.replace("_kernel", ""), 1e3 * v[0] / 100) for k, v in kt.items() if v[1])
    print("AB %-28s ms/step s2=%.4f s1=%.4f s3=%.4f s4=%.4f | us: %s | chk=%.6f" % (
        os.path.basename(os.environ.get("SSB200_LIB", "default")), res[2], res[1], res[3], res[4], ks, chk), flush=True)


if __name__ == "__main__":
    if "--child" in sys.argv:
        child()
    else:
        for lib in sys.argv[1:]:
            env = dict(os.environ, SSB200_LIB=os.path.abspath(lib))
            subprocess.call([sys.executable, os.path.abspath(__file__), "--child"], env=env)
