"""Multi-GPU correctness on real devices (skipped with fewer than 2 GPUs): a 2-rank CUDA run -- one process per GPU, each on its
contiguous sub-batch, logits written into the shared host gather buffer -- must equal the 1-rank CUDA run of the whole batch bit for bit
(images are independent, main.cpp:976-983; no collective on the data path)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, weight_path, n_total, hw, q):
    import torch.distributed as dist
    import ggml_experiments_b200 as G
    from ggml_experiments_b200 import mobilevit as MV
    from ggml_experiments_b200 import shard
    from ggml_experiments_b200 import weights as W
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)  # plumbing only: two barriers
    G.lib_ggml().ggml_b200_set_device(rank)                        # one process per GPU
    MV.set_mode(MV.FAST)
    imgs = W.synthetic_images(n_total, hw, hw, seed=7)
    lo, hi = shard.shard_range(n_total, rank, world)
    m = G.MobileViT(weight_path)
    feat, pooled = m.extract_features(imgs[lo:hi])
    info = m.plan_info(hi - lo, hw, hw)
    g = shard.HostGather(f"mvit_test_gpu_gather_{port}", n_total, pooled.shape[1], rank, world, dist)
    g.write(pooled)
    dist.barrier()
    if rank == 0:
        full_f, full_p = m.extract_features(imgs)  # the same GPU, the whole batch: the 1-rank result
        q.put((g.full.copy(), full_p, info["mode"]))
    m.close()
    g.close()
    dist.destroy_process_group()


def test_two_rank_cuda_run_equals_one_rank(weight_files):
    import ggml_experiments_b200 as G
    if G.lib_ggml().ggml_b200_device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    from ggml_experiments_b200 import mobilevit as MV
    n_total, hw, world = 7, 128, 2  # ragged: 4 + 3 images
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, weight_files["xs"], n_total, hw, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, single, mode = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert mode == MV.FAST
    assert gathered.shape == single.shape == (n_total, 384)
    np.testing.assert_array_equal(gathered, single)
