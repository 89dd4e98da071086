#!/usr/bin/env python
"""Full-array comparison of the two first levels of the FBGMM log_marg_i filter (development aid): every row's
log_marg_i and MAP slot from the e4m3 cascade against the fp16 one on the bench's trained-like model."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__file__), ".."))
import bench                                                          # noqa: E402
from segmentalist_b200 import fbgmm as fbgmm_mod                      # noqa: E402
from segmentalist_b200.batch import FvScorer                          # noqa: E402
from segmentalist_b200.gaussian_components_fixedvar import FixedVarPrior, GaussianComponentsFixedVar   # noqa: E402

n, K, D = int(sys.argv[1]) if len(sys.argv) > 1 else 4 * 1024 * 1024, 5000, bench.D
dev = torch.device("cuda", 0)
centres = torch.from_numpy(bench.centres_cpu(K)).to(dev)
X, Z = bench.make_embeddings_gpu(n, centres, seed=2000, device=dev)
var = 0.002 * np.ones(D)
am = fbgmm_mod.FBGMM.__new__(fbgmm_mod.FBGMM)
am.alpha, am.lms, am.covariance_type = 10., 1.0, "fixed"
am.components = GaussianComponentsFixedVar.from_device(X, FixedVarPrior(var, np.zeros(D), var / 0.05), K, alpha=10., lms=1.0)
n_tok = 20 * K
zh = Z[:n_tok].cpu().numpy()
_, first = np.unique(zh, return_index=True)
rank_of = np.empty(K, dtype=np.int64)
rank_of[zh[np.sort(first)]] = np.arange(len(first))
am.components._add_many(np.arange(n_tok), rank_of[zh])
out = {}
res = {}
for prec in ("fp16", "fp8"):
    fv = FvScorer(am.components, precision=prec)
    fv.score()
    torch.cuda.synchronize()
    res[prec] = (fv.log_marg.clone(), fv.map_k.clone(), int(fv.n_fallback.item()))
    del fv
d = (res["fp8"][0] - res["fp16"][0]).abs()
out = {"rows": n, "K": K, "max_abs_diff_log_marg": float(d.max()), "max_rel_diff": float((d / res["fp16"][0].abs()).max()),
       "map_k_identical": bool(torch.equal(res["fp8"][1], res["fp16"][1])),
       "undecided_rows": {"fp16": res["fp16"][2], "e4m3": res["fp8"][2]}}
print(json.dumps(out))
