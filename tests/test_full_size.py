"""
Parity at BASELINE.json's FULL sizes (run on the B200 box: pytest -m gpu).

The oracle cannot finish these sizes (config 3 would take ~21 h on one core), so the CUDA path is
checked through properties that do not depend on the size, plus the oracle on a bounded sample cut
out of the full problem:
  configs[2]  k-means Viterbi sweep, D=130, K=5000, 200k utterances (21M candidate segments)
  configs[1]  sequential FBGMM Gibbs sweep, D=130, K=1000, 2000 utterances
"""
import random

import numpy as np
import numpy.testing as npt
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_frozen_kmeans_sweep_full_size_properties():
    import bench
    from oracle import seg_oracle as so
    from segmentalist_b200 import _lib
    from segmentalist_b200.batch import FrozenKMeansSweep
    from segmentalist_b200.kmeans_components import KMeansComponents
    from segmentalist_b200.utterances import DeviceCorpus
    lib = _lib.lib()
    dev = torch.device("cuda", 0)
    U, K, S, D = bench.TOTAL_UTTS, bench.K_MAX, bench.S_MAX, bench.D
    lengths, seg_id, seg_dur, bounds0, n_emb = bench.corpus_structure(U, seed=1000)
    assert n_emb > 20_000_000
    X, Z = bench.make_embeddings_gpu(n_emb, torch.from_numpy(bench.centres_cpu(K)).to(dev), seed=2000, device=dev)
    corpus = DeviceCorpus(lengths, seg_id, seg_dur, bounds0, 0, S, S)
    perm = torch.randperm(n_emb, device=dev, generator=torch.Generator(device=dev).manual_seed(7))[:K]
    comps = KMeansComponents.from_device(X, K, X[perm].clone())
    tok = corpus.tok_id[corpus.tok_id >= 0].long()
    comps._assign[tok] = Z[tok]
    del Z
    sweep = FrozenKMeansSweep(comps, corpus, wip=0.0, scorer="mma")
    sweep.init_means_from_assignments()
    means_before = comps._means.clone()

    # --- 1. tensor-core scorer (filter GEMM + exact refine) == exact float32 kernel, bit for bit, on a sample
    sweep.score()
    rng = np.random.RandomState(11)
    ids = torch.from_numpy(rng.randint(0, n_emb, size=60000).astype(np.int32)).cuda()
    val = torch.empty(len(ids), dtype=torch.float32, device=dev)
    arg = torch.empty(len(ids), dtype=torch.int32, device=dev)
    _lib.check(lib.segb_kmeans_best(comps.struct(), _lib.ptr(ids), len(ids), _lib.ptr(val), _lib.ptr(arg),
                                    _lib.stream_ptr()))
    assert torch.equal(sweep.best_k[ids.long()], arg)
    assert torch.equal(sweep.best_val[ids.long()].view(torch.int32), val.view(torch.int32))
    assert int(sweep.mma.n_fallback.item()) < n_emb // 1000        # the filter decides (almost) every row itself

    # --- 2. Viterbi: structure, objective = sum of the chosen scores, optimal against other segmentations
    sweep.segment()
    torch.cuda.synchronize()
    assert int((sweep.status != _lib.DP_OK).sum().item()) == 0
    b = corpus.bounds.cpu().numpy().astype(bool)
    scores = sweep.scores.cpu().numpy().reshape(-1, S)
    obj = sweep.log_prob.cpu().numpy()
    pos_off = corpus.pos_off_h
    assert b[pos_off[1:] - 1].all()                                 # boundaries[N-1] is always set (:686)
    idx = np.where(b)[0]
    utt = np.searchsorted(pos_off, idx, "right") - 1
    first = np.concatenate([[True], utt[1:] != utt[:-1]])
    start = np.where(first, pos_off[utt], np.concatenate([[0], idx[:-1] + 1]))
    span = idx - start + 1
    assert span.min() >= 1 and span.max() <= S                      # no segment longer than n_slices_max
    chosen = scores[idx, span - 1]
    assert np.isfinite(chosen).all()
    npt.assert_allclose(np.bincount(utt, weights=chosen, minlength=U), obj, rtol=1e-12)

    def objective_of(bounds):
        i2 = np.where(bounds)[0]
        u2 = np.searchsorted(pos_off, i2, "right") - 1
        f2 = np.concatenate([[True], u2[1:] != u2[:-1]])
        st2 = np.where(f2, pos_off[u2], np.concatenate([[0], i2[:-1] + 1]))
        return np.bincount(u2, weights=scores[i2, i2 - st2], minlength=U)
    for alt in (np.ones_like(b), bounds0.astype(bool)):             # every landmark a boundary; the initial segmentation
        assert (obj >= objective_of(alt) - 1e-9 * np.abs(obj)).all()

    # --- 3. token collection conserves mass: counts and per-dimension sums
    sweep.collect()
    torch.cuda.synchronize()
    n_tok = int(b.sum())
    assert int(sweep.cnt.sum().item()) == n_tok
    tok = corpus.tok_id[corpus.tok_id >= 0].long()
    assert len(tok) == n_tok
    npt.assert_allclose(sweep.sum_x.sum(dim=0).cpu().numpy(), X[tok].double().sum(dim=0).cpu().numpy(), rtol=1e-9)
    k_tok = comps._assign[tok].long()
    assert torch.equal(torch.bincount(k_tok, minlength=K), sweep.cnt)
    assert torch.equal(k_tok.int(), sweep.best_k[tok])              # assignments = argmax under the frozen means

    # --- 4. the oracle on a bounded sample cut out of the full problem (first 48 utterances, all 5000 means)
    n_s = 48
    n_pos_s = int(pos_off[n_s])
    ids_s = seg_id[:n_pos_s]
    hi = int(ids_s.max()) + 1
    oseg = bench._cpu_segmenter(so, X[:hi].cpu().numpy(), lengths[:n_s], ids_s, seg_dur[:n_pos_s],
                                   means_before.cpu().numpy())
    totals, _, plan = so.frozen_kmeans_phase1(oseg)
    for u in range(n_s):
        N = int(lengths[u])
        npt.assert_array_equal(oseg.utterances.boundaries[u, :N], b[pos_off[u]:pos_off[u + 1]])
    assert np.array_equal(np.asarray(totals), obj[:n_s])            # float64 objectives, bit for bit
    got_tok = corpus.tok_id[:n_pos_s].cpu().numpy()
    got_k = comps._assign.cpu().numpy()
    want = [k for _, ks in plan for k in ks]
    assert [int(got_k[i]) for i in got_tok[got_tok >= 0]] == [int(k) for k in want]

    # --- 5. update: every rank-local step done, means = sum_x / counts in X's dtype (kmeans_components.py:110)
    sweep.reduce_and_update()
    torch.cuda.synchronize()
    cnt = sweep.cnt.cpu().numpy()
    live = cnt > 0
    want_means = (sweep.sum_x.cpu().numpy()[live] / cnt[live, None]).astype(np.float32)
    npt.assert_array_equal(comps._means.cpu().numpy()[live], want_means)


def test_gibbs_sweep_full_size_properties():
    from oracle import seg_oracle as so
    from segmentalist_b200 import fbgmm, gaussian_components_fixedvar as gcf, synth
    from segmentalist_b200 import unigram_acoustic_wordseg as uaw
    D, K, U, S = 130, 1000, 2000, 6
    mats, vids, durs, lms = synth.make_corpus_dicts(U, D=D, K_true=K, n_min=15, n_max=25, n_slices_max=S,
                                                    noise=0.05, seed=31)
    var = 0.002 * np.ones(D)
    prior = gcf.FixedVarPrior(var, np.zeros(D), var / 0.05)

    def make(mod, am, pr):
        random.seed(3)
        np.random.seed(3)
        return mod.UnigramAcousticWordseg(am, 10., K, pr, mats, vids, durs, lms, p_boundary_init=0.5,
                                          beta_sent_boundary=-1, n_slices_max=S)
    seg = make(uaw, fbgmm.FBGMM, prior)
    st = random.getstate()
    seg._sweep(list(range(U)), 1, False)                            # one full sweep, utterances in index order
    torch.cuda.synchronize()
    c = seg.acoustic_model.components
    assign, counts, Kact = c.assignments, c.counts, c.K

    # --- structure: one token per boundary that carries an embedding, labels 0..K-1 all populated
    n_tok = int((assign >= 0).sum())
    assert counts.sum() == n_tok and (counts[:Kact] > 0).all() and not counts[Kact:].any()
    assert assign.max() == Kact - 1
    npt.assert_array_equal(np.bincount(assign[assign >= 0], minlength=c.K_max), counts)
    bnd = seg.utterances.boundaries
    for u in range(U):
        assert bnd[u, seg.utterances.lengths[u] - 1]
    n_bound = int(sum(bnd[u, :seg.utterances.lengths[u]].sum() for u in range(U)))
    assert n_tok == n_bound                                         # max_span 6 always has an embedding

    # --- the incrementally maintained statistics equal a from-scratch rebuild (:153-188, :317-325)
    X = c.X.astype(np.float64)
    prec, prec0, mu0 = 1. / prior.var, 1. / prior.var_0, prior.mu_0
    sum_x = np.zeros((c.K_max, D))
    np.add.at(sum_x, assign[assign >= 0], X[assign >= 0])
    npt.assert_allclose(c.mu_N_numerators[:Kact], (prec0 * mu0)[None] + prec[None] * sum_x[:Kact], rtol=1e-9, atol=1e-9)
    npt.assert_allclose(c.precision_Ns[:Kact], prec0[None] + counts[:Kact, None] * prec[None], rtol=1e-12)
    pN = c.precision_Ns[:Kact]
    npt.assert_allclose(c.precision_preds[:Kact], pN * prec / (pN + prec), rtol=1e-14)

    # --- the oracle on the first utterances of the same sweep: identical samples (sequential chain,
    #     so the first n utterances do not depend on the rest)
    n_cpu = 16
    oseg = make(so, so.FBGMM, so.FixedVarPrior(var, np.zeros(D), var / 0.05))
    assert random.getstate() == st                                   # both constructions consumed the same draws
    for u in range(n_cpu):
        oseg.gibbs_sample_i(u)
    npt.assert_array_equal(bnd[:n_cpu], oseg.utterances.boundaries[:n_cpu])
