#!/usr/bin/env python
"""Time the fused k-means score kernel alone, with parts of its work switched off (SEGB_FUSED_DBG bit 0: no
operand conversion, bit 1: no refine) and against the two-kernel path.  Development aid."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from segmentalist_b200.batch import MmaScorer                      # noqa: E402
from segmentalist_b200.kmeans_components import KMeansComponents   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=6 * 1024 * 1024)
ap.add_argument("--K", type=int, default=5000)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
D, K = 130, args.K
g = torch.Generator(device="cuda").manual_seed(1)
centres = torch.randn(K, D, generator=g, device="cuda")
centres /= centres.norm(dim=1, keepdim=True)
X = torch.empty(args.rows, D, dtype=torch.float32, device="cuda")
for lo in range(0, args.rows, 1 << 20):
    hi = min(args.rows, lo + (1 << 20))
    z = torch.randint(0, K, (hi - lo,), generator=g, device="cuda")
    x = centres[z] + 0.05 * torch.randn(hi - lo, D, generator=g, device="cuda")
    X[lo:hi] = x / x.norm(dim=1, keepdim=True)
comps = KMeansComponents.from_device(X, K, centres.clone())
comps._means.copy_(centres)
comps._meansT.copy_(centres.t())
comps._K.fill_(K)
val = torch.empty(args.rows, dtype=torch.float32, device="cuda")
arg = torch.empty(args.rows, dtype=torch.int32, device="cuda")


def timed(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.reps


fl = 2.0 * D * args.rows * K
mma = MmaScorer(comps, fused=True)
mma.pack_means()
m2 = MmaScorer(comps, fused=False)
m2.pack_means()
res = {}


def note(name, ms):
    res.setdefault(name, []).append(ms)


for rnd in range(4):                                   # alternate, so clock / power drift hits every variant alike
    for dbg in (0, 3, 1, 2):
        os.environ["SEGB_FUSED_DBG"] = str(dbg)
        note("fused dbg=%d" % dbg, timed(lambda: mma.fused_score(val, arg)))
    os.environ["SEGB_FUSED_DBG"] = "0"
    note("two-kernel filter", timed(m2.filter))
    note("two-kernel refine", timed(lambda: m2.refine(val, arg)))
for k, v in res.items():
    print("%-20s min %.3f  median %.3f ms   (%.0f TFLOP/s at min)" % (k, min(v), sorted(v)[len(v) // 2], fl / min(v) / 1e9))
os.environ["SEGB_FUSED_DBG"] = "4"
mma.fused_score(val, arg)
torch.cuda.synchronize()
