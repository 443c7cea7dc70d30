"""Multi-GPU parity gate on NCCL (needs >= 2 GPUs on the box; skipped otherwise): env-sharded rendering + the one
all-gather must be BIT-IDENTICAL to the single-GPU batch."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("sr,taps", [(16000, 6000), (44100, 16384)])
def test_sharded_render_plus_nccl_allgather_is_bit_identical(sr, taps):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 4 if n >= 4 else 2
    env = dict(os.environ, MR_SR=str(sr), MR_TAPS=str(taps), MR_ENVS="128")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29650 + (sr % 7)), os.path.join(HERE, "multirank_worker.py")]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("MULTIRANK ")][-1]
    res = json.loads(line[len("MULTIRANK "):])
    assert res["world"] == world and res["identical_on_all_ranks"]
    assert res["bit_identical_to_single_gpu"], res
    assert res["nonzero_rows"] > 100
