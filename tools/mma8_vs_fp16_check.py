#!/usr/bin/env python
"""Full-array comparison of the k-means scorer's two first levels (development aid): best value (float32 bit patterns)
and component of EVERY row from the e4m3 cascade against the fp16 one, at the benchmark size."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__file__), ".."))
import bench                                                          # noqa: E402
from segmentalist_b200.batch import MmaScorer                         # noqa: E402
from segmentalist_b200.kmeans_components import KMeansComponents      # noqa: E402

n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
K, D = 5000, bench.D
dev = torch.device("cuda", 0)
lengths, seg_id, seg_dur, bounds0, n_emb, n_head = bench.shard_structure(n_utt, 0)
centres = torch.from_numpy(bench.centres_cpu(K)).to(dev)
X, Z = bench.make_embeddings_gpu(n_emb, centres, seed=2000, device=dev)
out = {"rows": int(n_emb), "K": K}
for model in ("trained", "random"):
    rnd = (centres + 0.01 * torch.randn_like(centres)) if model == "trained" else X[torch.randperm(n_emb, device=dev)[:K]].clone()
    comps = KMeansComponents.from_device(X, K, rnd.contiguous())
    comps._K.fill_(K)
    res = {}
    for prec in ("fp16", "fp8"):
        mma = MmaScorer(comps, precision=prec)
        val = torch.empty(n_emb, dtype=torch.float32, device=dev)
        arg = torch.empty(n_emb, dtype=torch.int32, device=dev)
        mma.score(val, arg)
        torch.cuda.synchronize()
        res[prec] = (val, arg, int(mma.n_fallback.item()))
        del mma
    out[model] = {"values_bit_identical": bool(torch.equal(res["fp8"][0].view(torch.int32), res["fp16"][0].view(torch.int32))),
                  "components_identical": bool(torch.equal(res["fp8"][1], res["fp16"][1])),
                  "undecided_rows": {"fp16": res["fp16"][2], "e4m3": res["fp8"][2]}}
print(json.dumps(out))
