#!/usr/bin/env python
"""Where does a frozen sweep spend time beyond its kernels?  Times variants of the sweep loop on one
GPU (development aid)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                                    # noqa: E402
from segmentalist_b200 import _lib                              # noqa: E402
from segmentalist_b200.batch import FrozenKMeansSweep           # noqa: E402
from segmentalist_b200.kmeans_components import KMeansComponents  # noqa: E402
from segmentalist_b200.utterances import DeviceCorpus           # noqa: E402

n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
fused = None if (len(sys.argv) <= 2 or sys.argv[2] != "two-kernel") else False
K = 5000
dev = torch.device("cuda", 0)
lengths, seg_id, seg_dur, bounds0, n_emb = bench.corpus_structure(n_utt, seed=1000)
X, Z = bench.make_embeddings_gpu(n_emb, torch.from_numpy(bench.centres_cpu(K)).to(dev), seed=2000, device=dev)
corpus = DeviceCorpus(lengths, seg_id, seg_dur, bounds0, 0, bench.S_MAX, bench.S_MAX)
perm = torch.randperm(n_emb, device=dev, generator=torch.Generator(device=dev).manual_seed(7))[:K]
comps = KMeansComponents.from_device(X, K, X[perm].clone())
tok = corpus.tok_id[corpus.tok_id >= 0].long()
comps._assign[tok] = Z[tok]
sw = FrozenKMeansSweep(comps, corpus, wip=0.0, scorer="mma", fused=fused)
sw.init_means_from_assignments()
for _ in range(3):
    sw.sweep()


def timed(fn, reps=5):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, (time.perf_counter() - t0) * 1e3 / reps


def kernels_only():
    sw.score(); sw.segment(); sw.collect(); sw.reduce_and_update()


def kernels_sync():
    kernels_only(); torch.cuda.current_stream().synchronize()


def host_only():                      # how long the host needs to enqueue one sweep
    t0 = time.perf_counter()
    kernels_only()
    return (time.perf_counter() - t0) * 1e3


print("utterances", n_utt, "embeddings", n_emb)
print("full sweep()          dev %.3f ms  wall %.3f ms" % timed(sw.sweep))
print("kernels only          dev %.3f ms  wall %.3f ms" % timed(kernels_only))
print("kernels + sync        dev %.3f ms  wall %.3f ms" % timed(kernels_sync))
torch.cuda.synchronize()
print("host enqueue time of one sweep: %.3f ms" % host_only())
print("phases", sw.profile_phases())

# ---- wall-clock of every statement group of sweep() (host view, device synced between groups)
import torch.distributed as dist  # noqa: E402


def step(name, fn, acc):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
    return r


acc = {}
for _ in range(5):
    K_before = sw.K_host
    step("score", sw.score, acc)
    step("segment", sw.segment, acc)
    step("summarize", sw.summarize, acc)
    if K_before < comps.K_max:
        step("clamp", lambda: sw._clamp_inactive_winners(K_before), acc)
    step("collect", sw.collect, acc)
    step("reduce", sw.reduce_and_update, acc)
    step("clean", sw._clean_components, acc)

    def tail():
        sw.flags[1:2].copy_(sw.mma.n_fallback)
        sw.flags[2:3].copy_(comps._K)
        torch.cuda.current_stream().wait_stream(sw.side)
        sw.flags_h.copy_(sw.flags, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return [int(v) for v in sw.flags_h.tolist()]
    n_bad, n_fb, K_now = step("tail", tail, acc)[:3]
    step("cumsum", lambda: float(np.cumsum(sw.log_prob_h.numpy())[-1]), acc)
print({k: round(v / 5, 3) for k, v in acc.items()}, "K_host", sw.K_host, "K_max", comps.K_max)
