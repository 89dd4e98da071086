#!/usr/bin/env python
"""Calibrate the accumulation-error constant of the e4m3 filter pass (mma_common.cuh: C_ACC8_PER_STEP).

The kernel's best-chunk maximum m1 (scaled space) is compared with the float64 dot products of the QUANTISED operands
(e4m3 values are exact in float64, so the difference is the tensor core's accumulation error alone), relative to
|x^||mu^| + |bias| -- the quantity filter_tau8 multiplies by c_acc.  Prints the largest ratio seen, per K = 32 step.
usage: fp8_acc_microbench.py [n_emb] [K_max] [D]"""
import sys

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__file__), ".."))
from segmentalist_b200 import _lib, synth                     # noqa: E402
from segmentalist_b200.batch import MmaScorer                 # noqa: E402
from segmentalist_b200.kmeans_components import KMeansComponents   # noqa: E402


def main():
    n_emb = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    K_max = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    D = int(sys.argv[3]) if len(sys.argv) > 3 else 130
    rng = np.random.RandomState(5)
    worst = 0.0
    for kind in ("clustered", "gauss", "wide"):
        if kind == "clustered":
            centres = synth.cluster_centres(K_max, D, rng)
            X = synth._unit_rows(centres[rng.randint(0, K_max, n_emb)] + 0.05 * rng.standard_normal((n_emb, D)).astype(np.float32))
        elif kind == "gauss":
            X = rng.standard_normal((n_emb, D)).astype(np.float32)
        else:
            X = (rng.standard_normal((n_emb, D)) * np.exp(rng.standard_normal((n_emb, D)))).astype(np.float32)
        assign = -np.ones(n_emb, dtype=np.int64)
        assign[:K_max * 4] = np.arange(K_max * 4) % K_max
        np.random.seed(1)
        comps = KMeansComponents(X, assign, K_max)
        mma = MmaScorer(comps, precision="fp8")
        mma.pack_means()
        mma.filter()
        torch.cuda.synchronize()
        s = mma.scale
        rec = mma.cand.cpu().numpy().view(np.float32).reshape(n_emb, 8)
        m1 = rec[:, 0].astype(np.float64)
        i1 = mma.cand.cpu().numpy().view(np.int32).reshape(n_emb, 8)[:, 3]
        # quantised operands in float64
        Xq = (torch.from_numpy(X).cuda() * s).to(torch.float8_e4m3fn).double()
        M = comps._means * s
        Mq = M.to(torch.float8_e4m3fn).double()
        bias = -0.5 * (M.double() ** 2).sum(1)
        t0 = bias / 256.0
        b0 = t0.float().to(torch.float8_e4m3fn).double()
        r1 = (t0.float() - b0.float())
        b1 = r1.to(torch.float8_e4m3fn).double()
        r2 = (r1 - b1.float())
        b2 = r2.to(torch.float8_e4m3fn).double()
        bq = 256.0 * (b0 + b1 + b2)
        ratios = []
        for lo in range(0, n_emb, 4096):
            sc = Xq[lo:lo + 4096] @ Mq.t() + bq[None, :]                 # exact scores of the quantised operands
            ch = torch.from_numpy(i1[lo:lo + 4096].astype(np.int64)).cuda()
            idx = ch[:, None] * 16 + torch.arange(16, device="cuda")[None, :]
            ok = idx < K_max
            vals = torch.where(ok, sc.gather(1, idx.clamp(max=K_max - 1)), torch.full_like(idx, -1e300, dtype=torch.float64))
            exact = vals.max(1).values.cpu().numpy()
            mag = (Xq[lo:lo + 4096].norm(dim=1) * Mq.norm(dim=1).max() + bq.abs().max()).cpu().numpy()
            ratios.append(np.abs(m1[lo:lo + 4096] - exact) / mag)
        r = np.concatenate(ratios)
        steps = ((D + 3 + 31) // 32)
        print("%-10s scale %g  max |m1 - exact| / (|x||mu| + |bias|) = %.3e  (per K=32 step: %.3e; C_ACC8_PER_STEP = %.3e)"
              % (kind, s, r.max(), r.max() / steps, 2.0 ** -20))
        worst = max(worst, r.max() / steps)
    print("worst per step %.3e -> margin %.1fx" % (worst, 2.0 ** -20 / worst))


if __name__ == "__main__":
    main()
