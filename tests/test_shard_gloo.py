"""CPU test of the N>1 host logic (world_size 2, gloo): shard the batch, run each shard, gather the logits on the host,
compare with the single-process result.  The per-shard compute is the CPU oracle here (tests may use it); on the GPU box
bench.py runs the same sharding with the CUDA path."""
import os
import socket

import numpy as np
import pytest

from ggml_experiments_b200 import shard


def test_shard_range_is_a_partition():
    for n in (1, 5, 7, 256, 257):
        for world in (1, 2, 3, 4, 8):
            spans = [shard.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, weight_path, n_total, hw, q):
    import torch.distributed as dist
    from ggml_experiments_b200 import weights as W
    from oracle import binding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    imgs = W.synthetic_images(n_total, hw, hw, seed=7)
    a, b = shard.shard_range(n_total, rank, world)
    _, pooled = binding.OracleModel(weight_path).forward(imgs[a:b], 0, 1)
    full = shard.gather_rows(pooled, n_total, rank, world, dist, dst=0)
    # the bench's data path: every rank writes its rows into its slice of ONE shared host buffer (no collective)
    g = shard.HostGather(f"mvit_test_gather_{port}", n_total, pooled.shape[1], rank, world, dist)
    g.write(pooled)
    dist.barrier()
    if rank == 0:
        q.put((full, g.full.copy()))
    g.close()
    dist.destroy_process_group()


def test_two_rank_sharded_run_equals_single_process(weight_files, oracle):
    import torch.multiprocessing as mp
    from ggml_experiments_b200 import weights as W
    n_total, hw, world = 5, 64, 2  # ragged: 3 + 2 images
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, weight_files["xxs"], n_total, hw, q)) for r in range(world)]
    for p in procs:
        p.start()
    full, shared = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    _, ref = oracle.OracleModel(weight_files["xxs"]).forward(W.synthetic_images(n_total, hw, hw, seed=7), 0, 1)
    assert full.shape == ref.shape and np.array_equal(full, ref)
    assert shared.shape == ref.shape and np.array_equal(shared, ref)
