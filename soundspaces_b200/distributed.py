"""Multi-GPU plumbing: one process per GPU, envs sharded ``i -> rank i mod G``.

Every (env, ear) is independent, so the data path needs NO collective: in DD-PPO each rank's
policy consumes its own shard -- as in the reference, whose collectives only touch gradients and
statistics (ss_baselines/av_nav/ddppo/ddppo.py:35-39, ddppo_trainer.py:315-324).
``gather_observations`` is the one optional collective: an all-gather that reassembles the
observation batch for a centralised policy (NCCL on CUDA tensors; gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .planning import shard_envs  # noqa: F401


def gpu_numa_cpus(device_index: int):
    """(NUMA node, CPU list) of the host socket a GPU hangs off, from sysfs
    (``/sys/bus/pci/devices/<domain:bus:dev.fn>/numa_node`` and ``/sys/devices/system/node/node<N>/cpulist``), or
    ``(None, [])`` when the box does not expose it.  Eight ranks each pushing their RIRs through pinned buffers that
    live on the wrong socket share one inter-socket link: binding rank -> local CPUs before the pinned buffers are
    allocated (first touch) keeps every rank's host traffic on its own socket."""
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return None, []
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.extend(range(int(lo), int(hi or lo) + 1))
        return node, cpus
    except Exception:          # noqa: BLE001 - containers often hide sysfs
        return None, []


def gather_observations(local: torch.Tensor, n_envs: int, rank: int, world: int) -> torch.Tensor:
    """All-gather the per-rank rows (rank r holds envs r, r+G, r+2G, ...) into the
    (n_envs, ...) batch in env order.  Ragged shards are padded to the largest shard."""
    if world == 1:
        return local
    per = -(-n_envs // world)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    # rank r, slot k  ->  env k*world + r
    out = out.view(world, per, *local.shape[1:]).transpose(0, 1).reshape(world * per, *local.shape[1:])
    return out[:n_envs]


class GatheredObservations:
    """The all-gather target as a pre-allocated buffer the kernels write into ("fuse by having
    the last kernel write directly into the rank's slice of the pre-registered gather buffer").

    ``buffer`` is ``(world, per_rank, *row_shape)``; rank r passes ``local`` -- the contiguous view
    ``buffer[r]`` -- as ``out=`` to ``BatchedAudioRenderer.execute`` so the spectrogram kernel's stores ARE the
    send buffer, and ``gather()`` is then one in-place ``all_gather_into_tensor`` (NCCL's in-place form:
    send pointer = receive pointer + rank * count; no staging copy, no padding copy).  Rows come back in
    rank-major order; ``env_ids`` maps row -> env for the ``i -> rank i mod G`` sharding and ``in_env_order()``
    gives the permuted copy when a consumer insists on env order."""

    def __init__(self, n_envs: int, row_shape, rank: int, world: int, device, dtype=torch.float32):
        self.n_envs, self.rank, self.world = int(n_envs), int(rank), int(world)
        self.per = -(-self.n_envs // self.world)
        self.buffer = torch.zeros((self.world, self.per) + tuple(row_shape), dtype=dtype, device=device)
        ids = torch.arange(self.world * self.per).view(self.world, self.per)
        self.env_ids = (ids % self.per) * self.world + ids // self.per          # row (r, k) holds env k*world + r
        self.n_local = len(shard_envs(self.n_envs, self.rank, self.world))

    @property
    def local(self) -> torch.Tensor:
        """This rank's slice, ``(n_local, *row_shape)``: hand it to the renderer as ``out=``."""
        return self.buffer[self.rank, : self.n_local]

    def gather(self) -> torch.Tensor:
        """In-place all-gather; returns the ``(world * per_rank, *row_shape)`` rank-major view of the buffer."""
        flat = self.buffer.view((self.world * self.per,) + tuple(self.buffer.shape[2:]))
        if self.world > 1:
            dist.all_gather_into_tensor(flat, self.buffer[self.rank])
        return flat

    def in_env_order(self) -> torch.Tensor:
        flat = self.buffer.view((self.world * self.per,) + tuple(self.buffer.shape[2:]))
        order = torch.argsort(self.env_ids.reshape(-1))[: self.n_envs]
        return flat[order.to(flat.device)]
