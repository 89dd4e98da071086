#!/usr/bin/env python
"""Time the filter-and-refine log_marg_i (segb_fvf_*: one fp16 tcgen05 pass + exact float64 refine) phase by
phase (rows x K_max, D=130) and check it against the exact float64 kernel on a sample.  Development aid
for profiles/ (ncu target)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from segmentalist_b200 import fbgmm                                                # noqa: E402
from segmentalist_b200.batch import FvScorer                                       # noqa: E402
from segmentalist_b200.gaussian_components_fixedvar import FixedVarPrior, GaussianComponentsFixedVar  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=4 * 1024 * 1024)
ap.add_argument("--K", type=int, default=5000)
ap.add_argument("--K-true", type=int, default=0, help="generating clusters (default K)")
ap.add_argument("--tokens-per-k", type=int, default=4)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--aniso", action="store_true")
ap.add_argument("--noise", type=float, default=0.05)
ap.add_argument("--fused", action="store_true")
ap.add_argument("--fp8", action="store_true", help="e4m3 first-level filter (isotropic variances)")
args = ap.parse_args()
D, K = 130, args.K
K_true = args.K_true or K
g = torch.Generator(device="cuda").manual_seed(1)
centres = torch.randn(K_true, D, generator=g, device="cuda")
centres /= centres.norm(dim=1, keepdim=True)
z = torch.randint(0, K_true, (args.rows,), generator=g, device="cuda")
X = torch.empty(args.rows, D, dtype=torch.float32, device="cuda")
for lo in range(0, args.rows, 1 << 20):
    hi = min(args.rows, lo + (1 << 20))
    x = centres[z[lo:hi]] + args.noise * torch.randn(hi - lo, D, generator=g, device="cuda")
    X[lo:hi] = x / x.norm(dim=1, keepdim=True)
rng = np.random.RandomState(0)
var = 0.002 * (0.5 + rng.rand(D)) if args.aniso else 0.002 * np.ones(D)
var_0 = 0.04 * (0.5 + rng.rand(D)) if args.aniso else var / 0.05
am = fbgmm.FBGMM.__new__(fbgmm.FBGMM)
am.alpha, am.lms, am.covariance_type = 10., 1.0, "fixed"
am.components = GaussianComponentsFixedVar.from_device(X, FixedVarPrior(var, np.zeros(D), var_0), K, alpha=10., lms=1.0)
n_tok = min(args.rows, args.tokens_per_k * K)
zh = z[:n_tok].cpu().numpy()
_, first = np.unique(zh, return_index=True)        # component = cluster, labels in order of first appearance
rank = np.empty(K_true, dtype=np.int64)
rank[zh[np.sort(first)]] = np.arange(len(first))
am.components._add_many(np.arange(n_tok), rank[zh])
fv = FvScorer(am.components, fused=bool(args.fused), precision="fp8" if args.fp8 else "fp16")
fv.score()
torch.cuda.synchronize()


def timed(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


t_pack = timed(fv.pack_model, args.reps)
if fv.fused:
    t_filter, t_refine = timed(fv.fused_score, args.reps), 0.0
else:
    t_filter, t_refine = timed(fv.filter, args.reps), timed(fv.refine, args.reps)

fl = (4.0 if args.aniso else 2.0) * D * args.rows * K
out = fv.log_marg.cpu().numpy()
ids = np.arange(n_tok, n_tok + 4096) if args.rows >= n_tok + 4096 else np.arange(min(4096, args.rows))
exact = am.log_marg_items(ids)
err = np.abs(out[ids] - exact)
print(json.dumps({"fused": fv.fused, "e4m3_first_level": bool(fv.fp8), "rows": args.rows, "K": K, "K_act": am.components.K, "aniso": bool(args.aniso),
                  "pack_model_ms": t_pack, "filter_ms": t_filter, "refine_ms": t_refine,
                  "filter_tflops_algorithmic": fl / t_filter / 1e9,
                  "score_tflops_algorithmic": fl / (t_pack + t_filter + t_refine) / 1e9,
                  "fallback_rows": int(fv.n_fallback.item()), "max_abs_err": float(err.max()),
                  "max_rel_err": float((err / np.abs(exact)).max()), "sample": len(ids)}))
