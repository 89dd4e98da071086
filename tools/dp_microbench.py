#!/usr/bin/env python
"""Time the batched DP kernel alone on the bench corpus shape (200k utterances, N~U{15..25}, S=6)
with random finite scores.  Development aid for profiles/ (ncu target); not part of the product."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from segmentalist_b200 import _lib                      # noqa: E402
from segmentalist_b200.utterances import DeviceCorpus   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--utts", type=int, default=200000)
ap.add_argument("--mode", type=int, default=2)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--scores", default="real", choices=["real", "uniform"])
args = ap.parse_args()
S = 6
rng = np.random.RandomState(0)
lengths = rng.randint(15, 26, size=args.utts).astype(np.int64)
n_pos = int(lengths.sum())
corpus = DeviceCorpus(lengths, np.full((n_pos, S), -1, np.int32), np.full((n_pos, S), np.nan),
                      np.zeros(n_pos, np.uint8), 0, S, S)
if args.scores == "uniform":
    scores = -torch.rand(n_pos * S, dtype=torch.float64, device="cuda") * 40.0
else:
    # like the k-means sweep: -(distance ~ 0.3) * duration in frames, -inf where the span leaves the utterance
    gaps = rng.randint(3, 15, size=n_pos).astype(np.float64)
    B = np.concatenate([[0.0], np.cumsum(gaps)])
    pos_off = np.concatenate([[0], np.cumsum(lengths)])
    within = np.arange(n_pos) - np.repeat(pos_off[:-1], lengths)          # t - 1
    sc = np.full((n_pos, S), -np.inf)
    for l in range(1, S + 1):
        ok = within >= l - 1
        idx = np.nonzero(ok)[0]
        dur = B[idx + 1] - B[idx + 1 - l]
        sc[idx, l - 1] = -(0.3 + 0.02 * rng.standard_normal(len(idx))) * dur
    scores = torch.from_numpy(sc.reshape(-1)).cuda()
uni = torch.rand(n_pos, dtype=torch.float64, device="cuda")
lp = torch.zeros(args.utts, dtype=torch.float64, device="cuda")
st = torch.zeros(args.utts, dtype=torch.int32, device="cuda")
lib, sp, cs = _lib.lib(), _lib.stream_ptr(), corpus.struct()


def run():
    _lib.check(lib.segb_dp_banded(cs, 0, corpus.n_utt, _lib.ptr(scores), args.mode, 0.0, 1.0, _lib.ptr(uni), None,
                                  _lib.ptr(corpus.bounds), _lib.ptr(lp), None, None, _lib.ptr(st), sp))


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.reps):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.reps
nbytes = 8.0 * n_pos * S + n_pos + 8.0 * args.utts + 8.0 * (args.utts + 1) + 4.0 * args.utts
print("dp mode %d: %.4f ms  %.1f GB/s (%.1f MB)  status ok=%s" % (args.mode, ms, nbytes / ms / 1e6, nbytes / 1e6,
                                                                bool((st == 0).all().item())))
