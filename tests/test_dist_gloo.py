"""
CPU multi-process tests (gloo, world_size 2): the sharded frozen sweep equals the unsharded one.

Each rank scores + segments its contiguous range of utterances with the oracle's pure functions
(standing in for the CUDA phases), the product's host-side logic (segmentalist_b200/sharding.py:
shard_ranges, reduce_stats, clamp_plan, compaction_plan, means_from_stats) combines the shards, and
the result must equal the reference-derived golden fixture of the unsharded sweep.
"""
import os
import random
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import seg_oracle as so
from tests import _golden as G


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, init, out):
    from segmentalist_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        z = G.load("kmeans_wordseg.npz")
        mats, vids, durs, lms = G.unpack_dicts(z)
        random.seed(4)
        np.random.seed(4)
        seg = so.SegmentalKMeansWordseg(5, mats, vids, durs, lms, p_boundary_init=0.5, n_slices_max=6,
                                        init_am_assignments=init, wip=0)
        comps = seg.acoustic_model.components
        K_before, K_max, D = comps.K, comps.K_max, comps.D
        lo, hi = sharding.shard_ranges(seg.utterances.D, world)[rank]
        totals, _, plan = so.frozen_kmeans_phase1(seg, range(lo, hi))           # this rank's shard
        # tokens won by inactive slots: global order = rank order (contiguous shards)
        ks_local = [int(k) for _, ks in plan for k in ks]
        ids_local = [int(e) for embeds, _ in plan for e in embeds]
        sel = [i for i, k in enumerate(ks_local) if k >= K_before]
        parts = [None] * world
        dist.all_gather_object(parts, [ks_local[i] for i in sel])
        all_new, K = sharding.clamp_plan([k for p in parts for k in p], K_before)
        off = sum(len(p) for p in parts[:rank])
        for n, i in enumerate(sel):
            ks_local[i] = int(all_new[off + n])
        # local sufficient statistics -> all-reduce -> identical means on every rank
        sum_x = torch.zeros(K_max, D, dtype=torch.float64)
        cnt = torch.zeros(K_max, dtype=torch.int64)
        for e, k in zip(ids_local, ks_local):
            sum_x[k] += torch.from_numpy(comps.X[e].astype(np.float64))
            cnt[k] += 1
        obj = torch.tensor([float(np.sum(totals))], dtype=torch.float64)
        # the product's collective: one flat float64 buffer [sum_x | counts] (batch.py)
        red = torch.zeros(K_max * D + K_max, dtype=torch.float64)
        red[:K_max * D] = sum_x.reshape(-1)
        sum_x = red[:K_max * D].view(K_max, D)
        sharding.reduce_packed(red, red[K_max * D:], cnt)
        sharding.reduce_stats(torch.zeros(1, 1, dtype=torch.float64), torch.zeros(1, dtype=torch.int64), obj)
        means = sharding.means_from_stats(sum_x.numpy(), cnt.numpy(), comps.means)
        K_new, dst, src = sharding.compaction_plan(cnt.numpy(), K)
        counts = cnt.numpy().copy()
        num = sum_x.numpy().copy()
        means[dst], counts[dst], num[dst] = means[src], counts[src], num[src]
        means[K_new:K] = comps.random_means[K_new:K]
        counts[K_new:K] = 0
        num[K_new:K] = 0
        if rank == 0:
            np.savez(out, means=means, counts=counts, num=num, K=K_new, obj=obj.numpy(),
                     bounds_lo=seg.utterances.boundaries[lo:hi])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("init", ["spread", "rand"])
def test_sharded_frozen_sweep_equals_unsharded(tmp_path, init):
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, _free_port(), init, out), nprocs=2, join=True)
    r = np.load(out)
    z = G.load("kmeans_wordseg.npz")
    p = init + "_"
    K = int(z[p + "frozen_K"])
    assert int(r["K"]) == K
    np.testing.assert_array_equal(r["counts"], z[p + "frozen_counts"])
    np.testing.assert_array_equal(r["means"], z[p + "frozen_means"])
    np.testing.assert_allclose(r["num"], z[p + "frozen_mean_numerators"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(float(r["obj"][0]), float(z[p + "frozen_total"]), rtol=1e-13)
    n0 = r["bounds_lo"].shape[0]
    np.testing.assert_array_equal(r["bounds_lo"], z[p + "frozen_boundaries"][:n0])


def test_shard_ranges_and_plans():
    from segmentalist_b200 import sharding
    assert sharding.shard_ranges(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert sharding.shard_ranges(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    # compaction: same result as the oracle's sequential clean_components
    rng = np.random.RandomState(0)
    for _ in range(50):
        K = int(rng.randint(1, 12))
        cnt = rng.randint(0, 3, size=K)
        K_new, dst, src = sharding.compaction_plan(cnt, K)
        labels = list(range(K))                      # simulate del_component on a label list
        KK = K
        for k in np.where(cnt == 0)[0][::-1]:
            KK -= 1
            if k != KK:
                labels[k] = labels[KK]
        assert K_new == KK
        for d_, s_ in zip(dst, src):
            assert labels[d_] == s_
        assert all(cnt[labels[j]] > 0 for j in range(K_new))
    ks, K = sharding.clamp_plan([7, 3, 9, 4, 4, 1], 3)
    assert list(ks) == [3, 3, 4, 4, 4, 1] and K == 5
