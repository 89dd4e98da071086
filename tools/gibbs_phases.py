#!/usr/bin/env python
"""Per-phase clock totals of the cooperative Gibbs sweep (CTA 0's view, barrier waits included):
where the ~170 us per utterance of BASELINE configs[1] go.  Development aid for profiles/."""
import ctypes
import os
import random
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from segmentalist_b200 import _lib, fbgmm, gaussian_components_fixedvar as gcf, synth            # noqa: E402
from segmentalist_b200 import unigram_acoustic_wordseg as uaw                                    # noqa: E402

D, K, U, S = 130, 1000, int(sys.argv[1]) if len(sys.argv) > 1 else 2000, 6
mats, vids, durs, lms = synth.make_corpus_dicts(U, D=D, K_true=K, n_min=15, n_max=25, n_slices_max=S, noise=0.05, seed=31)
var = 0.002 * np.ones(D)
random.seed(3)
np.random.seed(3)
seg = uaw.UnigramAcousticWordseg(fbgmm.FBGMM, 10., K, gcf.FixedVarPrior(var, np.zeros(D), var / 0.05), mats, vids, durs,
                                 lms, p_boundary_init=0.5, beta_sent_boundary=-1, n_slices_max=S)
order = list(range(U))
seg._sweep(order, 1, False)
torch.cuda.synchronize()
lib = ctypes.CDLL(_lib.SO_PATH)
out = (ctypes.c_ulonglong * 16)()
lib.segb_debug_gibbs_prof(out, 1)
t0 = time.perf_counter()
seg._sweep(order, 1, False)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
lib.segb_debug_gibbs_prof(out, 0)
names = {0: "remove tokens", 1: "score: slot consts / loop exit", 11: "score: stage embeddings", 12: "score: (segment, component) pairs",
         13: "score: partial LSE + log_prior", 9: "score: grid barrier", 2: "score: combine",
         10: "score: grid barrier 2", 3: "DP (FFBS)", 4: "assign: stage x + owners publish", 5: "assign: grid barrier",
         6: "assign: read K_max values + decide", 7: "assign: owner update", 8: "loop head"}
tot = float(sum(out))
n_tok = seg.acoustic_model.get_n_assigned()
print("sweep %.1f ms, %.1f us/utt, %d tokens (%.1f per utt), K_act %d" % (wall * 1e3, wall / U * 1e6, n_tok, n_tok / U,
                                                                        seg.acoustic_model.components.K))
for ph, name in sorted(names.items(), key=lambda kv: -out[kv[0]]):
    print("%-40s %6.1f %%  %8.2f us/utt" % (name, 100 * out[ph] / tot, out[ph] / tot * wall / U * 1e6))
