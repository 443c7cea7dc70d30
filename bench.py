#!/usr/bin/env python
"""bench.py -- binaural audio frames/s (RIR conv + spectrogram) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm

A "step" is one pass of the hot path over one batch of synthetic input: every
env's 1-s source clip is convolved with its binaural RIR and turned into the
(65, T', 2) log-magnitude spectrogram (reference: soundspaces/simulator.py:608-701
+ soundspaces/tasks/nav.py:86-100).  Workload = BASELINE.json configs[1]: 128 envs
per GPU, 44.1 kHz, 16384-tap RIRs, output (128, 65, 69, 2); weak scaling (each
rank renders its own 128 envs; no data-path collective, as in the reference's
DD-PPO where observations never cross ranks).

Prints ONE JSON line on rank 0 (see the field notes in DESIGN.md section "Measurement").
"""
from __future__ import annotations

import argparse
import json
import os

# worker processes of the CPU arm are single-threaded (one env per process, like the reference's
# VectorEnv workers): must be set before numpy/scipy are imported in the spawned children
for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ.setdefault(_k, "1")
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "binaural audio frames/sec (RIR conv+STFT)"
UNIT = "frames/s"

# workload (BASELINE.json configs[1])
SR = 44100
TAPS = 16384
ENVS_PER_GPU = 128
N_BANKS = 16            # RIR banks rotated between steps: 16 x 16.8 MB = 268 MB > 126 MB L2
# bytes = 8*L_eff + 4*S/B_share + 8*65*T'  (spectrogram-only output)
ALG_BYTES_PER_FRAME = 8 * TAPS + 4 * SR // ENVS_PER_GPU + 8 * 65 * 69
# dram__bytes_read.sum + dram__bytes_write.sum per 64-env launch from the committed capture
# profiles/prof_r01d.ncu-rep (ncu --set full, cold-cache replay), by kernel
NCU_DRAM_BYTES_PER_LAUNCH = {"fwd_rir_kernel": 8.42e6, "mac_bins_kernel": 17.79e6, "mac_ifft_kernel": 46.86e6,
                             "spectrogram_kernel": 22.62e6}
# same capture: FMA-pipe busy cycles per SM (sm__pipe_fma_cycles_active.avg) and executed warp-instructions
# (smsp__inst_executed.sum) of one 64-env launch of each kernel -- the compute roofline this FP32 path really has
NCU_FMA_PIPE_CYCLES_PER_LAUNCH = {"fwd_rir_kernel": 4663.0, "mac_bins_kernel": 8912.0, "mac_ifft_kernel": 13389.0,
                                  "spectrogram_kernel": 21105.0}
NCU_WARP_INST_PER_LAUNCH = {"fwd_rir_kernel": 2.359e6, "mac_bins_kernel": 5.456e6, "mac_ifft_kernel": 8.000e6,
                            "spectrogram_kernel": 11.864e6}
ALG_FLOP_PER_FRAME = 17.3e6   # 2 packed 65536-pt FFT equivalents + mul + 276 packed 512-pt FFTs
FP32_PEAK_TFLOPS = 75.0       # nominal B200 FP32 SIMT, SURVEY.md 8(d)


def workload_config(n_gpus):
    return {
        "workload": "C2: 128 envs/GPU x (1-s 44.1 kHz source (*) 16384-tap binaural RIR -> (65,69,2) log-spectrogram)",
        "sr": SR, "rir_taps": TAPS, "envs_per_gpu": ENVS_PER_GPU, "global_envs": ENVS_PER_GPU * n_gpus,
        "spectrogram": [65, 69, 2], "stft_pad_mode": "reflect", "source": "one shared clip per batch",
        "l2_policy": f"inputs larger than L2: {N_BANKS} RIR banks ({N_BANKS * ENVS_PER_GPU * TAPS * 8 / 1e6:.0f} MB) rotated per step",
        "parallelism": f"env-sharded x{n_gpus}, no data-path collective",
    }


def make_bank_host(n_envs, seed0=0):
    """(n_envs, TAPS, 2) float32: N(0,1)*exp(-t/tau), tau = L/6, max|rir| = 0.5;
    2 % of envs get the zero-RIR fallback."""
    rng = np.random.default_rng(1234 + seed0)
    env = np.exp(-np.arange(TAPS) / (TAPS / 6.0)).astype(np.float32)
    out = np.empty((n_envs, TAPS, 2), dtype=np.float32)
    for i in range(n_envs):
        r = rng.standard_normal((TAPS, 2), dtype=np.float32) * env[:, None]
        r *= np.float32(0.5) / np.abs(r).max()
        out[i] = r
    out[rng.random(n_envs) < 0.02] = 0.0
    return out


def silent_mask(n_envs, seed0=0):
    return np.random.default_rng(99 + seed0).random(n_envs) < 0.05      # 5 % silent envs


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.samples = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [l for (t, l) in self.samples if t0 - 0.05 <= t <= t1 + 0.15] or [l for (_, l) in self.samples]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in rows:
            p = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except Exception:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- CPU arm
def usable_cpus():
    """Host threads this process may really use: min(sched affinity, cgroup CPU quota)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    quota = None
    try:                                     # cgroup v2
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            quota = float(q) / float(per)
    except Exception:
        try:                                 # cgroup v1
            q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                quota = q / per
        except Exception:
            pass
    if quota is not None:
        n = max(1, min(n, int(quota + 0.5)))
    return n


def _cpu_init():
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"


def _cpu_frames(args):
    """Worker: render `count` frames of the C2 workload with the oracle (the reference's algorithm:
    scipy.signal.fftconvolve x2 ears + librosa.stft restatement + 4x4 mean + log1p)."""
    seed, count = args
    from oracle import audio_oracle as ao
    from synth import make_rir, make_source
    src = make_source(7, SR)
    rirs = [make_rir(seed * 4 + j, TAPS) for j in range(4)]
    t0 = time.perf_counter()
    acc = 0.0
    for i in range(count):
        _, spec = ao.render_frame(src, rirs[i % 4], SR)
        acc += float(spec[0, 0, 0])
    return time.perf_counter() - t0, acc


def cpu_throughput(frames_total, procs):
    """frames/s of the oracle on `procs` host processes (one env per task, single-threaded BLAS/FFT
    per process -- the reference's process-per-env model, ss_baselines/common/env_utils.py:41-106)."""
    import multiprocessing as mp
    per = max(1, frames_total // procs)
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs, initializer=_cpu_init) as pool:
        pool.map(_cpu_frames, [(i, 2) for i in range(procs)])          # warm-up / import
        t0 = time.perf_counter()
        pool.map(_cpu_frames, [(i, per) for i in range(procs)])
        dt = time.perf_counter() - t0
    return per * procs / dt, per * procs, dt


def run_reference(args, rank):
    """--impl reference: the reference's own CPU algorithm for the path, all host threads."""
    if rank != 0:
        return
    cores = usable_cpus()
    per_step = max(cores * 32, 64)       # bounded sample per step (about 0.15 s of wall time on 16 cores)
    vals = []
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    per = max(1, per_step // cores)
    with ctx.Pool(cores, initializer=_cpu_init) as pool:
        for _ in range(max(1, min(args.warmup, 2))):
            pool.map(_cpu_frames, [(i, per) for i in range(cores)])
        t0 = time.perf_counter()
        steps = max(1, min(args.steps, 20))   # keep the whole run within minutes
        for _ in range(steps):
            pool.map(_cpu_frames, [(i, per) for i in range(cores)])
        dt = time.perf_counter() - t0
    value = steps * per * cores / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 2), "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{steps} steps x {per * cores} frames of the C2 workload, {cores} processes x 1 thread "
                                   "(oracle/audio_oracle.py: scipy.signal.fftconvolve + restated librosa.stft/block_reduce)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- GPU arm
def run_gpu(args, rank, local_rank, world):
    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL prints its version there)
    # are sent to stderr for the duration of the run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from soundspaces_b200 import AudioRequest, BatchedAudioRenderer
    from synth import make_source

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B = ENVS_PER_GPU
    r = BatchedAudioRenderer(SR, TAPS, device=dev, log2n=args.log2n)
    r.set_conv_mode(args.conv_mode)
    sid = r.add_source(make_source(7, SR))
    bank_host = make_bank_host(N_BANKS * B, seed0=rank)
    bank = torch.from_numpy(bank_host).to(dev)
    ids = r.set_dense_rir_bank(bank)
    sil = silent_mask(B, seed0=rank)
    batches = [r.prepare([AudioRequest(rir=ids[k * B + i], source=sid, silent=bool(sil[i])) for i in range(B)])
               for k in range(N_BANKS)]
    spec_out = torch.empty((B,) + r.spec_shape, dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i):
        r.execute(batches[i % N_BANKS], out=spec_out)

    # ---- device-resident throughput ("value")
    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    launches0 = r.ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    t_wall1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = r.ctx.launch_count - launches0
    # a long enough region for the clock sampler: keep the GPU under the same load for >= 1.5 s
    if sampler:
        t_end = time.time() + max(0.0, 1.5 - (t_wall1 - t_wall0))
        i = 0
        while time.time() < t_end:
            for _ in range(50):
                step(i); i += 1
            torch.cuda.synchronize()
        t_wall1 = time.time()
        clocks = sampler.stop(t_wall0, t_wall1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())

    # ---- optional centralised-policy mode: every step ends with ONE in-place all-gather of the
    # observation batch; the spectrogram kernel writes straight into this rank's slice of the gather buffer
    gather_info = None
    if world > 1 and args.gather:
        try:
            from soundspaces_b200.distributed import GatheredObservations
            gobs = GatheredObservations(B * world, r.spec_shape, rank, world, dev)
            assert gobs.n_local == B

            def gstep(i):
                r.execute(batches[i % N_BANKS], out=gobs.local)
                return gobs.gather()

            for i in range(max(3, args.warmup // 2)):
                gstep(i)
            barrier()
            ge0, ge1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ge0.record()
            for i in range(args.steps):
                flat = gstep(i)
            ge1.record()
            barrier()
            tg = torch.tensor([ge0.elapsed_time(ge1)], dtype=torch.float64, device=dev)
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
            # every rank must now hold every rank's rows: compare a checksum of the gathered batch across ranks
            chk = flat.double().sum().reshape(1)
            lo, hi = chk.clone(), chk.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            gather_info = {"value": B * world * args.steps / (float(tg.item()) * 1e-3), "unit": UNIT,
                           "ms_per_step": float(tg.item()) / args.steps,
                           "bytes_sent_per_rank_per_step": int(gobs.local.numel() * 4),
                           "collective": "one in-place ncclAllGather per step (all_gather_into_tensor on the buffer the "
                                         "spectrogram kernel wrote into)",
                           "identical_on_all_ranks": bool(float(lo.item()) == float(hi.item()))}
        except Exception as e:          # informational leg: never take the headline measurement down with it
            gather_info = {"error": repr(e)[:200]}

    # ---- per-kernel durations for the roofline (same K steps, events around every launch)
    r.ctx.set_kernel_timing(True)
    for i in range(args.steps):
        step(i)
    ktimes = r.ctx.get_kernel_timing()
    r.ctx.set_kernel_timing(False)

    # ---- end to end through the host-buffer C-ABI entry ("e2e")
    hs = r.make_host_session(B, TAPS, want_wave=False, n_chunks=args.chunks)
    hs.h_rir.numpy()[:] = bank_host[:B]
    hs.set_requests(sid, silent=sil)
    for _ in range(max(3, args.warmup // 4)):
        hs.run()
    barrier()
    e2e_steps = args.steps
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(e2e_steps):
        hs.run()
    g1.record()
    barrier()
    e2e_ms = g0.elapsed_time(g1)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms_max = float(t.item())
    checksum = float(hs.h_spec.double().sum())

    if rank == 0:
        frames = B * world * args.steps
        value = frames / (ms_max * 1e-3)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (of measured)"
        else:
            peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
        hot = [k for k in ("fwd_rir_kernel", "mac_bins_kernel", "mac_ifft_kernel", "spectrogram_kernel") if ktimes[k][1]]
        # per STEP: a kernel may be launched several times per step (sub-batches on internal streams)
        per_step_ms = {k: ktimes[k][0] / args.steps for k in hot}
        launches_per_step = {k: ktimes[k][1] / args.steps for k in hot}
        dom = max(hot, key=lambda k: per_step_ms[k])
        dom_ms = per_step_ms[dom]
        alg_bytes_launch = ALG_BYTES_PER_FRAME * B            # all of the step's frames pass through every kernel
        achieved = alg_bytes_launch / (dom_ms * 1e-3) / 1e9
        step_ms = ms_max / args.steps
        kernel_sum = sum(per_step_ms.values())
        # compute roofline of the kernels as built (default plan only): the time the FMA pipes / the issue slots need
        # for one step's instructions (counts from profiles/prof_r01d.ncu-rep) over the measured step time
        fma_frac = issue_frac = None
        try:
            if args.log2n in (0, 12) and args.conv_mode == 0:
                sm_mhz = (clocks.get("sm_mhz") or 1965.0)
                sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
                fma_frac = (sum(NCU_FMA_PIPE_CYCLES_PER_LAUNCH[k] * launches_per_step[k] for k in hot)
                            / (sm_mhz * 1e6) / (step_ms * 1e-3))
                issue_frac = (sum(NCU_WARP_INST_PER_LAUNCH[k] * launches_per_step[k] for k in hot)
                              / (sm_count * 4) / (sm_mhz * 1e6) / (step_ms * 1e-3))
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(world),
            "roofline": {
                "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak,
                "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get(dom, 0.0) * launches_per_step[dom] if args.log2n in (0, 12) else None,
                "traffic_source": "profiles/prof_r01d.ncu-rep: dram__bytes_read.sum + dram__bytes_write.sum of the "
                                  "dominant kernel (cold-cache replay, 64-env launch) x its launches per step",
                "peak_source": peak_src,
                "algorithmic_bytes_per_step": alg_bytes_launch, "kernel_ms": dom_ms,
                "kernel_ms_all": per_step_ms, "kernel_launches_per_step": launches_per_step,
                "kernel_timing": "CUDA events around every launch, summed per step; sub-batches run on 2 internal streams, "
                                 "so the per-kernel sums overlap in wall time",
                "kernel_share_of_step": dom_ms / max(kernel_sum, 1e-9),
                "path_achieved_gbs": alg_bytes_launch / (step_ms * 1e-3) / 1e9,
                "path_frac_hbm": alg_bytes_launch / (step_ms * 1e-3) / 1e9 / peak,
                "fp32_frac": (value / world) * ALG_FLOP_PER_FRAME / (FP32_PEAK_TFLOPS * 1e12),
                "fma_pipe_frac": fma_frac, "issue_slot_frac": issue_frac,
                "note": "FFT work is FP32-pipe bound (about 100 flop/B at algorithmic traffic): fma_pipe_frac / issue_slot_frac are the "
                        "fractions that measure kernel quality; the launch sizes of the capture are 64-env sub-batches, the same "
                        "as the 2 launches per kernel per step here; see DESIGN.md",
            },
            "e2e": {"value": B * world * e2e_steps / (e2e_ms_max * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": hs.h2d_bytes, "d2h_bytes_per_step": hs.d2h_bytes,
                    "ms_per_step": e2e_ms_max / e2e_steps, "api": f"ssb_render_batch_host (pinned host RIRs in, host spectrograms out), {args.chunks} pipelined chunks",
                    "checksum": checksum},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if gather_info is not None:
            line["with_allgather"] = gather_info
        if world == 1 and not args.no_cpu:
            cores = usable_cpus()
            v, n, dt = cpu_throughput(cores * args.cpu_frames_per_core, cores)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{n} frames of the same C2 workload in {dt:.1f} s, {cores} processes x 1 thread "
                                              "(oracle/audio_oracle.py: scipy.signal.fftconvolve + restated librosa.stft/block_reduce)"}
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2n", type=int, default=0)
    ap.add_argument("--conv-mode", type=int, default=0, help="0: mac_bins + ifft kernels, 1: fused mac_ifft")
    ap.add_argument("--chunks", type=int, default=2, help="pipeline depth of the host-buffer (e2e) entry")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--gather", action="store_true",
                    help="N > 1: also time the centralised-policy mode (one in-place all-gather of the observations per step)")
    ap.add_argument("--cpu-frames-per-core", type=int, default=150)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            # convenience: re-launch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"),
                   os.path.abspath(__file__)] + sys.argv[1:]
            sys.exit(subprocess.call(cmd))
        raise SystemExit(f"WORLD_SIZE={world} does not match --gpus {args.gpus}")
    run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
