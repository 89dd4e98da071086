#!/usr/bin/env python
"""Sequential (online-update) k-means segmenter sweep vs the frozen-state sweep on the same corpus
(D=130, K=1000, 2000 utterances): SegmentalKMeansWordseg.segment vs segment_frozen.  Development aid."""
import sys, time, random
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from segmentalist_b200 import kmeans_acoustic_wordseg as kaw, synth
U=2000
mats, vids, durs, lms = synth.make_corpus_dicts(U, D=130, K_true=1000, n_min=15, n_max=25, n_slices_max=6, noise=0.05, seed=31)
random.seed(3); np.random.seed(3)
seg = kaw.KMeansAcousticWordseg(1000, mats, vids, durs, lms, n_slices_max=6, init_am_assignments="spread")
for it in range(2):
    torch.cuda.synchronize(); t0=time.perf_counter()
    rec = seg.segment(1)
    torch.cuda.synchronize(); dt=time.perf_counter()-t0
    print("sequential k-means sweep: %.1f ms, %.0f utt/s, K=%d" % (dt*1e3, U/dt, seg.acoustic_model.components.K))
t0=time.perf_counter(); rec = seg.segment_frozen(1); torch.cuda.synchronize(); print("frozen sweep (incl. setup): %.1f ms" % ((time.perf_counter()-t0)*1e3))
t0=time.perf_counter(); rec = seg.segment_frozen(1); torch.cuda.synchronize(); print("frozen sweep: %.1f ms" % ((time.perf_counter()-t0)*1e3))
