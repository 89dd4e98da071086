#!/usr/bin/env python
"""Where does the host synchronisation at the end of a frozen sweep cost device time?  (development aid)"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                                    # noqa: E402
from segmentalist_b200.batch import FrozenKMeansSweep           # noqa: E402
from segmentalist_b200.kmeans_components import KMeansComponents  # noqa: E402
from segmentalist_b200.utterances import DeviceCorpus           # noqa: E402

n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
K = 5000
dev = torch.device("cuda", 0)
lengths, seg_id, seg_dur, bounds0, n_emb = bench.corpus_structure(n_utt, seed=1000)
X, Z = bench.make_embeddings_gpu(n_emb, torch.from_numpy(bench.centres_cpu(K)).to(dev), seed=2000, device=dev)
corpus = DeviceCorpus(lengths, seg_id, seg_dur, bounds0, 0, bench.S_MAX, bench.S_MAX)
perm = torch.randperm(n_emb, device=dev, generator=torch.Generator(device=dev).manual_seed(7))[:K]
comps = KMeansComponents.from_device(X, K, X[perm].clone())
tok = corpus.tok_id[corpus.tok_id >= 0].long()
comps._assign[tok] = Z[tok]
sw = FrozenKMeansSweep(comps, corpus, wip=0.0, scorer="mma", fused=False)
sw.init_means_from_assignments()
for _ in range(3):
    sw.sweep()


def timed(fn, reps=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def kernels_only():
    sw.score(); sw.segment(); sw.collect(); sw.reduce_and_update()


flag_d = torch.zeros(1, dtype=torch.int64, device=dev)
flag_h = torch.zeros(1, dtype=torch.int64).pin_memory()
seq = [0]


def kernels_sync():
    kernels_only(); torch.cuda.current_stream().synchronize()


def kernels_event_sync():
    kernels_only()
    e = torch.cuda.Event()
    e.record()
    e.synchronize()


def kernels_poll():
    kernels_only()
    seq[0] += 1
    flag_d.fill_(seq[0])
    flag_h.copy_(flag_d, non_blocking=True)
    a = flag_h.numpy()
    while a[0] != seq[0]:
        pass


def kernels_query():
    kernels_only()
    e = torch.cuda.Event()
    e.record()
    while not e.query():
        pass


for name, fn in (("kernels only", kernels_only), ("stream.synchronize", kernels_sync), ("event.synchronize", kernels_event_sync),
                 ("pinned flag poll", kernels_poll), ("event.query poll", kernels_query), ("full sweep()", sw.sweep),
                 ("kernels only", kernels_only), ("stream.synchronize", kernels_sync)):
    print("%-22s dev %.3f ms" % (name, timed(fn)))

# host timeline after a sync: when is each launch of the next sweep enqueued?
torch.cuda.synchronize()
ts = []
for _ in range(5):
    kernels_only()
    torch.cuda.current_stream().synchronize()
    t0 = time.perf_counter()
    sw.mma.pack_means()
    t1 = time.perf_counter()
    sw.mma.filter()
    t2 = time.perf_counter()
    sw.mma.refine(sw.best_val, sw.best_k)
    t3 = time.perf_counter()
    sw.segment(); sw.collect(); sw.reduce_and_update()
    t4 = time.perf_counter()
    ts.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3))
    torch.cuda.synchronize()
print("host ms (pack_means, filter, refine, rest) per sweep:", np.round(np.array(ts), 3).tolist())

# tiny kernel + sync round trip
x = torch.zeros(1, device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    x.add_(1)
    torch.cuda.current_stream().synchronize()
print("tiny kernel + stream.synchronize round trip: %.1f us" % ((time.perf_counter() - t0) / 200 * 1e6))
