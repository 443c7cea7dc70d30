"""Worker of tests/test_gpu_multirank.py (launched under torchrun, one rank per GPU, NCCL).

Gate of BASELINE.md section 5 / SURVEY.md 8(c): "sharded-vs-single-GPU bit-identical".  A 128-env batch is sharded
``env i -> rank i mod G``; every rank renders its shard straight into its slice of the gather
buffer, ONE in-place NCCL all-gather reassembles the batch, and rank 0 compares it -- ``torch.equal`` -- with the
same 128 envs rendered on one GPU.  Every rank additionally checks that it holds the identical gathered batch."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from soundspaces_b200 import AudioRequest, BatchedAudioRenderer
    from soundspaces_b200.distributed import GatheredObservations
    from soundspaces_b200.planning import shard_envs
    from synth import make_rir, make_source

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    sr, taps, n_envs = int(os.environ.get("MR_SR", 16000)), int(os.environ.get("MR_TAPS", 6000)), int(os.environ.get("MR_ENVS", 128))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    src = make_source(3, sr)
    rirs = [make_rir(1000 + i, taps - 13 * (i % 17)) for i in range(n_envs)]      # ragged, seeded: the same on every rank
    silent = np.random.default_rng(5).random(n_envs) < 0.05

    def render(envs, out):
        r = BatchedAudioRenderer(sr, taps, device=dev)
        sid = r.add_source(src)
        ids = r.add_rirs([rirs[i] for i in envs])
        r.execute(r.prepare([AudioRequest(rir=ids[k], source=sid, silent=bool(silent[i])) for k, i in enumerate(envs)]), out=out)
        return r

    mine = shard_envs(n_envs, rank, world)
    gobs = GatheredObservations(n_envs, BatchedAudioRenderer(sr, taps, device=dev).spec_shape, rank, world, dev)
    render(mine, gobs.local)
    gobs.gather()                                                  # the ONE collective: in-place ncclAllGather
    full = gobs.in_env_order()
    # every rank holds the same batch
    chk = full.double().sum().reshape(1)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    res = {"rank": rank, "world": world, "identical_on_all_ranks": bool(lo.item() == hi.item())}
    if rank == 0:
        single = torch.empty_like(full)
        render(list(range(n_envs)), single)
        torch.cuda.synchronize()
        res["bit_identical_to_single_gpu"] = bool(torch.equal(full, single))
        res["max_abs_diff"] = float((full - single).abs().max())
        res["n_envs"], res["nonzero_rows"] = n_envs, int((full.flatten(1).abs().sum(1) > 0).sum())
        print("MULTIRANK " + json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
