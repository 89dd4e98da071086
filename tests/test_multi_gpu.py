"""
NCCL parity test of the sharded sweeps on real GPUs (skipped with fewer than two): a 2-rank sweep over a
small fixed corpus must equal the 1-rank sweep of the same corpus -- k-means means bit for bit, boundaries
and assignments identical (the diffuse start exercises inactive-slot wins, the device clamp with its
fixed-size all-gather and the compaction), frozen FBGMM decisions identical and statistics to 1e-13.
The same gate runs inside bench.py at every N ("parity" key).
"""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import bench
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    try:
        res = bench.multi_gpu_parity(None, world, rank, dev, None, None)
        if rank == 0:
            q.put(res)
    finally:
        dist.destroy_process_group()


def test_nccl_two_ranks_equal_one_rank():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res["nrank_equals_1rank_kmeans_small_corpus"], res
    assert res["nrank_equals_1rank_fbgmm_small_corpus"], res
