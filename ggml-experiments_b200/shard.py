"""Multi-GPU sharding of the batch: images are independent (attention never crosses images,
main.cpp:976-983), so a batch splits into contiguous per-rank sub-batches; weights are replicated; there is NO device
collective on the data path -- only a final host-side gather of the per-image logits.

`torch.distributed` is used as plumbing (gloo on CPU for the tests, nccl/gloo under torchrun on the GPU box)."""
from __future__ import annotations

import numpy as np


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split: the first (n_total % world) ranks get one extra image."""
    base, extra = divmod(n_total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_rows(local: np.ndarray, n_total: int, rank: int, world: int, dist=None, dst: int = 0):
    """Host-side gather of per-image rows ([n_local, ...]) to rank `dst`; returns the full array there, None elsewhere.
    Ragged shards are padded to the largest shard for the collective and trimmed afterwards."""
    if world == 1:
        return local
    import torch
    counts = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    mx = max(counts)
    pad = np.zeros((mx,) + local.shape[1:], dtype=local.dtype)
    pad[: local.shape[0]] = local
    t = torch.from_numpy(pad)
    out = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
    dist.gather(t, out, dst=dst)
    if rank != dst:
        return None
    return np.concatenate([o.numpy()[:c] for o, c in zip(out, counts)], axis=0)


class HostGather:
    """The "final host-side gather of logits" (BASELINE.json north_star): ONE host buffer [n_total, width] f32 in POSIX
    shared memory; every rank (one process per GPU) writes the rows of its own images into its disjoint slice right after its
    device->host copy; rank 0 reads the whole array.  No collective, no extra hop: the data path of a step ends with a memcpy.

    `dist` (torch.distributed, any backend) is used only for two barriers: create -> attach, and before unlink."""

    def __init__(self, name: str, n_total: int, width: int, rank: int, world: int, dist=None):
        from multiprocessing import shared_memory
        self.rank, self.world, self.dist = rank, world, dist
        self.lo, self.hi = shard_range(n_total, rank, world)
        nbytes = max(1, n_total * width * 4)
        if rank == 0:
            try:  # a stale segment of a crashed run
                old = shared_memory.SharedMemory(name=name)
                old.close()
                old.unlink()
            except FileNotFoundError:
                pass
            self.shm = shared_memory.SharedMemory(name=name, create=True, size=nbytes)
        if world > 1:
            dist.barrier()
        if rank != 0:
            self.shm = shared_memory.SharedMemory(name=name)
            try:  # only the creator owns the segment: keep this process's resource tracker from unlinking (and warning about) it at exit
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        self.full = np.ndarray((n_total, width), dtype=np.float32, buffer=self.shm.buf)
        self.mine = self.full[self.lo:self.hi]

    def write(self, rows: np.ndarray) -> None:
        """this rank's rows ([n_local, width]) -> its slice of the shared buffer"""
        np.copyto(self.mine, rows)

    def close(self) -> None:
        if self.world > 1:
            self.dist.barrier()
        self.full = self.mine = None
        self.shm.close()
        if self.rank == 0:
            self.shm.unlink()
