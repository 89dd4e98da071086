#!/usr/bin/env python
"""Time the tensor-core log_marg_i kernel alone (rows x K_max, D=130) and check it against the
exact float64 kernel on a sample.  Development aid for profiles/ (ncu target)."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from segmentalist_b200 import _lib, fbgmm                                        # noqa: E402
from segmentalist_b200.gaussian_components_fixedvar import FixedVarPrior, GaussianComponentsFixedVar  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=4 * 1024 * 1024)
ap.add_argument("--K", type=int, default=5000)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--model", default="clusters", choices=["clusters", "roundrobin"])
args = ap.parse_args()
D, K = 130, args.K
g = torch.Generator(device="cuda").manual_seed(1)
centres = torch.randn(K, D, generator=g, device="cuda")
centres /= centres.norm(dim=1, keepdim=True)
z = torch.randint(0, K, (args.rows,), generator=g, device="cuda")
X = torch.empty(args.rows, D, dtype=torch.float32, device="cuda")
for lo in range(0, args.rows, 1 << 20):
    hi = min(args.rows, lo + (1 << 20))
    x = centres[z[lo:hi]] + 0.05 * torch.randn(hi - lo, D, generator=g, device="cuda")
    X[lo:hi] = x / x.norm(dim=1, keepdim=True)
var = 0.002 * np.ones(D)
am = fbgmm.FBGMM.__new__(fbgmm.FBGMM)
am.alpha, am.lms, am.covariance_type = 10., 1.0, "fixed"
am.components = GaussianComponentsFixedVar.from_device(X, FixedVarPrior(var, np.zeros(D), var / 0.05), K, alpha=10., lms=1.0)
n_tok = min(args.rows, 4 * K)
zh = z[:n_tok].cpu().numpy()
if args.model == "clusters":                 # component = cluster (labels in order of first appearance)
    _, first = np.unique(zh, return_index=True)
    rank = np.empty(K, dtype=np.int64)
    rank[zh[np.sort(first)]] = np.arange(len(first))
    ks = rank[zh]
else:
    ks = np.arange(n_tok) % K
am.components._add_many(np.arange(n_tok), ks)
out = am.log_marg_all(tensor_cores=True)
x_t, w_t, out_t = am._tc
lib = _lib.lib()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.reps):
    _lib.check(lib.segb_fvmma_log_marg(am.components.struct(), _lib.ptr(x_t), _lib.ptr(w_t), args.rows, _lib.ptr(out_t),
                                       _lib.stream_ptr()))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.reps
fl = 2.0 * D * args.rows * K
ids = np.arange(n_tok, n_tok + 4096) if args.rows >= n_tok + 4096 else np.arange(min(4096, args.rows))
exact = am.log_marg_items(ids)
rel = np.abs(out[ids] - exact) / np.abs(exact)
print("fv log_marg: %.3f ms  %.1f TFLOP/s algorithmic  K_act=%d  max rel err %.2e (sample of %d)"
      % (ms, fl / ms / 1e9, am.components.K, rel.max(), len(ids)))
