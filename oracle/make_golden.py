"""
oracle/make_golden.py -- TEST INFRASTRUCTURE.  Generates tests/golden/*.npz by
RUNNING THE REFERENCE ITSELF (kamperh/segmentalist at /root/reference) in the
build container.  Run once by hand:   python oracle/make_golden.py

The reference is Python 2.  It is copied to a throw-away temp directory (never
into this repo), passed through the mechanical py2->py3 text shim listed in
SURVEY.md 8c (syntax / removed NumPy aliases only -- no numerics are touched),
its Cython extension is compiled, and the scenarios below are executed through
the reference's own public classes and functions.  Inputs, the recorded uniform
stream / utterance orders, and the outputs are stored side by side so the GPU
box (which has no /root/reference) can replay them.
"""
import importlib
import os
import random
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

OWN = ["_cython_utils", "utterances", "utils", "niw", "wishart", "bigram_lms", "bigram_fbgmm",
       "fbgmm", "kmeans", "kmeans_components", "gaussian_components", "gaussian_components_diag",
       "gaussian_components_fixedvar", "unigram_acoustic_wordseg", "kmeans_acoustic_wordseg",
       "bigram_acoustic_wordseg", "plot_utils"]


def shim_text(src, is_pyx=False):
    s = src
    s = s.replace("xrange", "range").replace("basestring", "str")
    s = s.replace("(int, long)", "(int, np.integer)")
    s = re.sub(r"\bnp\.float\b(?!\d|_)", "np.float64", s)
    s = re.sub(r"\bnp\.int\b(?!\d|_|c)", "np.int_", s)
    s = s.replace("from scipy.misc import logsumexp", "from scipy.special import logsumexp")
    for m in OWN:
        s = re.sub(r"^(\s*)import %s\s*$" % m, r"\1from segmentalist import %s" % m, s, flags=re.M)
        s = re.sub(r"^(\s*)from %s import" % m, r"\1from segmentalist.%s import" % m, s, flags=re.M)
    # print statements -> functions (single-line ones; one multi-line case handled below)
    s = re.sub(r"^(\s*)print (?!\()(.*)$", r"\1print(\2)", s, flags=re.M)
    s = re.sub(r"^(\s*)print\s*$", r"\1print()", s, flags=re.M)
    # integer division used for sizes / indices
    s = s.replace("(N**2 + N)/2", "(N**2 + N)//2")
    s = s.replace("(n_slices**2 + n_slices)/2", "(n_slices**2 + n_slices)//2")
    s = s.replace("self.N_max*(self.N_max + 1)/2", "self.N_max*(self.N_max + 1)//2")
    s = s.replace("t*(t - 1)/2", "t*(t - 1)//2")
    s = s.replace("i = 0.5*(t - 1)*t", "i = int(0.5*(t - 1)*t)")
    s = s.replace("(range(K)*int(", "(list(range(K))*int(")
    s = s.replace("(range(am_K)*int(", "(list(range(am_K))*int(")
    s = s.replace("utt_order = range(self.utterances.D)", "utt_order = list(range(self.utterances.D))")
    if is_pyx:
        s = s.replace("np.int_t", "long")
    return s


def build_shimmed_reference(target=None):
    """Copy + shim + build the reference.  target=None: a throw-away temp directory (fixtures);
    otherwise `target` (the git-ignored baseline/_ref, which travels to the GPU box like the built
    .so files so that `bench.py --impl reference` can time the reference itself)."""
    if target is None:
        tmp = tempfile.mkdtemp(prefix="segref_")
    else:
        tmp = target
        if os.path.isdir(tmp):
            shutil.rmtree(tmp)
        os.makedirs(tmp)
    pkg = os.path.join(tmp, "segmentalist")
    shutil.copytree(os.path.join(REF, "segmentalist"), pkg)
    files = [os.path.join(dp, fn) for dp, _, fns in os.walk(pkg) for fn in fns]
    for p in files:
        fn = os.path.basename(p)
        if fn.endswith(".py") or fn.endswith(".pyx"):
            with open(p) as f:
                src = f.read()
            out = shim_text(src, fn.endswith(".pyx"))
            if fn == "gaussian_components_diag.py":
                # the one multi-line print statement (diag main(); off the hot path)
                out = re.sub(r"print\(\((.*?)\n(.*?)\n(.*?)\n(\s*)\)\)", r"print((\1\n\2\n\3\n\4))", out, flags=re.S)
            with open(p, "w") as f:
                f.write(out)
    setup = os.path.join(tmp, "setup.py")
    with open(setup, "w") as f:
        f.write(
            "from setuptools import setup, Extension\n"
            "from Cython.Build import cythonize\nimport numpy\n"
            "setup(ext_modules=cythonize([Extension('segmentalist._cython_utils',"
            "['segmentalist/_cython_utils.pyx'], include_dirs=[numpy.get_include()])],"
            "language_level=3))\n")
    subprocess.check_call([sys.executable, "setup.py", "-q", "build_ext", "--inplace"], cwd=tmp)
    shutil.rmtree(os.path.join(tmp, "build"), ignore_errors=True)
    if target is None:
        sys.path.insert(0, tmp)
    return tmp


class Tap(object):
    """Records every random.random() and random.shuffle() the reference makes."""

    def __init__(self):
        self.uniforms, self.orders = [], []
        self._rr, self._sh = random.random, random.shuffle

    def __enter__(self):
        def rr():
            u = self._rr()
            self.uniforms.append(u)
            return u

        def sh(x):
            self._sh(x)
            self.orders.append(list(x))
        random.random, random.shuffle = rr, sh
        return self

    def __exit__(self, *a):
        random.random, random.shuffle = self._rr, self._sh


def pack_dicts(prefix, mats, vids, durs, lms):
    labels = sorted(mats)
    d = {prefix + "labels": np.array(labels)}
    for i, l in enumerate(labels):
        d["%smat_%d" % (prefix, i)] = mats[l]
        d["%svid_%d" % (prefix, i)] = np.asarray(vids[l])
        d["%sdur_%d" % (prefix, i)] = np.asarray(durs[l])
        d["%slm_%d" % (prefix, i)] = np.asarray(lms[l])
    return d


# ---------------------------------------------------------------------------

def golden_dp(ref_uni, ref_km):
    """Random DP cases through the three reference DP functions."""
    rng = np.random.RandomState(1234)
    cases = []
    for case in range(400):
        N = int(rng.randint(1, 14))
        S = int(rng.choice([0, 1, 2, 3, 6]))
        mode = case % 3
        temp = 2.5 if (mode == 0 and case % 12 == 0) else 1
        vec = -np.inf * np.ones(N * (N + 1) // 2)
        for t in range(1, N + 1):
            for j in range(t):
                if S and t - j > S:
                    continue
                if rng.rand() < 0.1 and N > 1:
                    continue
                vec[t * (t - 1) // 2 + j] = rng.randn() * 5 - 3
        us = rng.rand(N + 1)
        pos = [0]

        def fake():
            u = us[pos[0]]
            pos[0] += 1
            return u
        old = random.random
        random.random = fake
        try:
            if mode == 0:
                lp, b = ref_uni.forward_backward(vec.copy(), 0.0, N, 0, S, None, temp)
            elif mode == 1:
                lp, b = ref_uni.forward_backward_viterbi(vec.copy(), 0.0, N, 0, S, None, None)
            else:
                lp, b = ref_km.forward_backward_kmeans_viterbi(vec.copy(), N, 0, S, None)
            ok = 1
        except Exception:
            lp, b, ok = np.nan, np.zeros(N, dtype=bool), 0
        finally:
            random.random = old
        cases.append((N, S, mode, temp, vec, us, pos[0], ok, lp, b))
    d = {"n_cases": np.array(len(cases))}
    for i, (N, S, mode, temp, vec, us, used, ok, lp, b) in enumerate(cases):
        d["c%d_meta" % i] = np.array([N, S, mode, temp, used, ok], dtype=np.float64)
        d["c%d_vec" % i] = vec
        d["c%d_u" % i] = us
        d["c%d_lp" % i] = np.array(lp)
        d["c%d_b" % i] = b
    np.savez_compressed(os.path.join(OUT, "dp_cases.npz"), **d)
    print("dp_cases:", len(cases), "cases,", sum(c[7] for c in cases), "ok")


def golden_fixedvar(ref_fv, ref_fbgmm):
    """log_post_pred / log_prior / log_marg_i on D=130 float32 data, isotropic and
    anisotropic priors, K_act < K_max, after deletions incl. a whole component."""
    d = {}
    for tag, aniso in (("iso", False), ("aniso", True)):
        np.random.seed(7 if aniso else 3)
        random.seed(5)
        D, N, K_max = 130, 60, 12
        X = np.random.randn(N, D).astype(np.float32)
        X /= np.linalg.norm(X, axis=1, keepdims=True)
        if aniso:
            var = 0.001 + 0.004 * np.random.rand(D)
            mu_0 = 0.1 * np.random.randn(D)
            var_0 = 0.02 + 0.05 * np.random.rand(D)
        else:
            var = 0.002 * np.ones(D)
            mu_0 = np.zeros(D)
            var_0 = var / 0.05
        prior = ref_fv.FixedVarPrior(var, mu_0, var_0)
        assignments = np.random.randint(0, 7, N)
        assignments[50:] = -1
        # make labels consecutive the way the reference requires
        uniq = sorted(set(assignments) - {-1})
        remap = {k: i for i, k in enumerate(uniq)}
        remap[-1] = -1
        assignments = np.array([remap[a] for a in assignments])
        am = ref_fbgmm.FBGMM(X, prior, 10., K_max, assignments.copy(), covariance_type="fixed", lms=0.7)
        c = am.components
        # delete every member of component 2 (exercises del_component swap + relabel)
        for i in np.where(c.assignments == 2)[0]:
            c.del_item(i)
        items = np.arange(50, 60)
        d[tag + "_X"] = X
        d[tag + "_var"], d[tag + "_mu_0"], d[tag + "_var_0"] = var, mu_0, var_0
        d[tag + "_assign_in"] = assignments
        d[tag + "_assign_out"] = c.assignments.copy()
        d[tag + "_counts"] = c.counts.copy()
        d[tag + "_K"] = np.array(c.K)
        d[tag + "_mu_N_numerators"] = c.mu_N_numerators.copy()
        d[tag + "_precision_Ns"] = c.precision_Ns.copy()
        d[tag + "_precision_preds"] = c.precision_preds.copy()
        d[tag + "_log_prod_precision_preds"] = c.log_prod_precision_preds.copy()
        d[tag + "_items"] = items
        d[tag + "_log_post_pred"] = np.array([c.log_post_pred(i) for i in items])
        d[tag + "_log_prior"] = np.array([c.log_prior(i) for i in items])
        d[tag + "_log_marg_i"] = np.array([am.log_marg_i(i) for i in items])
        d[tag + "_log_marg"] = np.array(am.log_marg())
        d[tag + "_log_prob_z"] = np.array(am.log_prob_z())
    np.savez_compressed(os.path.join(OUT, "fixedvar_scoring.npz"), **d)
    print("fixedvar_scoring done")


def golden_kmeans_scoring(ref_kc, ref_kmeans):
    """float32 neg_sqrd_norm bit patterns, argmax incl. inactive (random-row)
    slots, KMeans.fit trace."""
    np.random.seed(11)
    random.seed(11)
    D, N, K_max = 130, 400, 24
    centres = np.random.randn(10, D)
    X = centres[np.random.randint(0, 10, N)] + 0.3 * np.random.randn(N, D)
    X = (X / np.linalg.norm(X, axis=1, keepdims=True)).astype(np.float32)
    assignments = -1 * np.ones(N, dtype=int)
    assignments[:300] = np.random.randint(0, 16, 300)
    assignments[:16] = np.arange(16)
    state = np.random.get_state()
    km = ref_kmeans.KMeans(X, K_max, assignments.copy())
    c = km.components
    d = {"X": X, "assign_in": assignments, "K_max": np.array(K_max),
         "np_state_keys": state[1], "np_state_pos": np.array(state[2]),
         "random_means": c.random_means.copy(), "means0": c.means.copy(),
         "counts0": c.counts.copy(), "K0": np.array(c.K)}
    items = np.arange(280, 340)
    d["items"] = items
    d["neg_sqrd_norm"] = np.array([c.neg_sqrd_norm(i) for i in items])
    d["argmax"] = np.array([c.argmax_neg_sqrd_norm_i(i) for i in items])
    d["max"] = np.array([c.max_neg_sqrd_norm_i(i) for i in items])
    rec = km.fit(5, consider_unassigned=False)
    d["fit_assign"] = c.assignments.copy()
    d["fit_counts"] = c.counts.copy()
    d["fit_K"] = np.array(c.K)
    d["fit_means"] = c.means.copy()
    d["fit_mean_numerators"] = c.mean_numerators.copy()
    d["fit_n_mean_updates"] = np.array(rec["n_mean_updates"])
    d["fit_sum_neg_sqrd_norm"] = np.array(rec["sum_neg_sqrd_norm"])
    np.savez_compressed(os.path.join(OUT, "kmeans_scoring.npz"), **d)
    print("kmeans_scoring done; fit updates", rec["n_mean_updates"])


def golden_unigram(ref_uni, ref_fbgmm, ref_fv):
    """UnigramAcousticWordseg.gibbs_sample on a small synthetic corpus, FFBS and
    Viterbi flavours, with the uniform stream and utterance orders recorded."""
    from segmentalist_b200 import synth
    for tag, fb_type, n_iter, kw in (
            ("ffbs", "standard", 3, {}),
            ("ffbs_anneal", "standard", 3, {"anneal_schedule": "linear", "anneal_start_temp_inv": 0.5,
                                            "anneal_gibbs_am": True}),
            ("viterbi", "viterbi", 2, {})):
        mats, vids, durs, lms = synth.make_corpus_dicts(
            14, D=16, K_true=5, n_min=3, n_max=9, n_slices_max=4, noise=0.08, seed=21)
        random.seed(2)
        np.random.seed(2)
        D = 16
        prior = ref_fv.FixedVarPrior(0.002 * np.ones(D), np.zeros(D), 0.002 * np.ones(D) / 0.05)
        seg = ref_uni.UnigramAcousticWordseg(
            ref_fbgmm.FBGMM, 10., 9, prior, mats, vids, durs, lms, p_boundary_init=0.5,
            beta_sent_boundary=-1, n_slices_max=4, lms=1.0, wip=-0.3, fb_type=fb_type,
            time_power_term=1.1)
        ref_uni.i_debug_monitor = -1
        d = pack_dicts("in_", mats, vids, durs, lms)
        d["init_boundaries"] = seg.utterances.boundaries.copy()
        d["init_assignments"] = seg.acoustic_model.components.assignments.copy()
        with Tap() as tap:
            rec = seg.gibbs_sample(n_iter, **kw)
        c = seg.acoustic_model.components
        d["uniforms"] = np.array(tap.uniforms)
        d["orders"] = np.array(tap.orders)
        d["anneal_temp"] = np.array(rec["anneal_temp"], dtype=np.float64)
        for key in ("log_marg", "log_marg*length", "log_prob_z", "log_prob_X_given_z", "components", "n_tokens"):
            d["rec_" + key] = np.array(rec[key], dtype=np.float64)
        d["boundaries"] = seg.utterances.boundaries.copy()
        d["assignments"] = c.assignments.copy()
        d["counts"] = c.counts.copy()
        d["K"] = np.array(c.K)
        d["mu_N_numerators"] = c.mu_N_numerators.copy()
        # frozen scores of utterance 0 under the final model (pure function)
        for e in seg.utterances.get_segmented_embeds_i(0):
            if e != -1:
                c.del_item(e)
        N0 = seg.utterances.lengths[0]
        d["u0_scores"] = seg.get_vec_embed_log_probs(
            seg.utterances.vec_ids[0, :(N0 ** 2 + N0) // 2], seg.utterances.durations[0, :(N0 ** 2 + N0) // 2])
        np.savez_compressed(os.path.join(OUT, "unigram_%s.npz" % tag), **d)
        print("unigram", tag, "log_marg", rec["log_marg"], "K", c.K, "uniforms", len(tap.uniforms))


def golden_kmeans_wordseg(ref_km):
    """BASELINE config 1 (synthesised: D=10, K=5, 50 utterances): sequential
    segment() sweeps with in-between KMeans.fit, plus one frozen-state sweep
    assembled from the reference's pure functions."""
    from segmentalist_b200 import synth
    mats, vids, durs, lms = synth.make_corpus_dicts(
        50, D=10, K_true=5, n_min=5, n_max=12, n_slices_max=6, noise=0.15, seed=33)
    d = pack_dicts("in_", mats, vids, durs, lms)
    for init in ("spread", "rand"):
        random.seed(4)
        np.random.seed(4)
        seg = ref_km.SegmentalKMeansWordseg(
            5, mats, vids, durs, lms, p_boundary_init=0.5, n_slices_max=6,
            init_am_assignments=init, wip=0)
        ref_km.i_debug_monitor = -1
        c = seg.acoustic_model.components
        p = init + "_"
        d[p + "init_boundaries"] = seg.utterances.boundaries.copy()
        d[p + "init_assignments"] = c.assignments.copy()
        d[p + "random_means"] = c.random_means.copy()
        # ---- frozen sweep from pure reference calls, on a deep copy of the state
        import copy
        fz = copy.deepcopy(seg)
        fc = fz.acoustic_model.components
        total, plan, old = 0.0, [], []
        scores_all = []
        for u in range(fz.utterances.D):
            N = fz.utterances.lengths[u]
            n_packed = (N ** 2 + N) // 2
            old.extend(e for e in fz.utterances.get_segmented_embeds_i(u) if e != -1)
            sc = fz.get_vec_embed_neg_len_sqrd_norms(fz.utterances.vec_ids[u, :n_packed],
                                                     fz.utterances.durations[u, :n_packed])
            scores_all.append(sc)
            obj, b = ref_km.forward_backward_kmeans_viterbi(sc, N, 0, 6, u)
            total += obj
            fz.utterances.boundaries[u, :N] = b
            emb = fz.utterances.get_segmented_embeds_i(u)
            plan.append((emb, fc.get_max_assignments(emb)))
        for e in old:
            fc.del_item(e)
        for emb, ks in plan:
            for e, k in zip(emb, ks):
                fc.add_item(e, k)
        fc.clean_components()
        d[p + "frozen_total"] = np.array(total)
        d[p + "frozen_scores_u0"] = scores_all[0]
        d[p + "frozen_boundaries"] = fz.utterances.boundaries.copy()
        d[p + "frozen_assignments"] = fc.assignments.copy()
        d[p + "frozen_counts"] = fc.counts.copy()
        d[p + "frozen_K"] = np.array(fc.K)
        d[p + "frozen_means"] = fc.means.copy()
        d[p + "frozen_mean_numerators"] = fc.mean_numerators.copy()
        # ---- sequential reference sweeps
        with Tap() as tap:
            rec = seg.segment(3, n_iter_inbetween_kmeans=2)
        d[p + "orders"] = np.array(tap.orders)
        d[p + "rec_sum_neg_sqrd_norm"] = np.array(rec["sum_neg_sqrd_norm"])
        d[p + "rec_sum_neg_len_sqrd_norm"] = np.array(rec["sum_neg_len_sqrd_norm"])
        d[p + "rec_components"] = np.array(rec["components"])
        d[p + "rec_n_tokens"] = np.array(rec["n_tokens"])
        d[p + "boundaries"] = seg.utterances.boundaries.copy()
        d[p + "assignments"] = c.assignments.copy()
        d[p + "counts"] = c.counts.copy()
        d[p + "K"] = np.array(c.K)
        d[p + "means"] = c.means.copy()
        d[p + "mean_numerators"] = c.mean_numerators.copy()
        print("kmeans_wordseg", init, rec["sum_neg_len_sqrd_norm"], rec["components"])
    np.savez_compressed(os.path.join(OUT, "kmeans_wordseg.npz"), **d)


def golden_fbgmm_gibbs(ref_fbgmm, ref_fv, ref_uni):
    """FBGMM.gibbs_sample (fbgmm.py:288-420) on fixed-variance components -- the in-between
    acoustic-model resampling of UnigramAcousticWordseg.gibbs_sample(am_n_iter > 0)
    (unigram_acoustic_wordseg.py:440-443) -- with the uniform stream recorded; and one segmenter
    run that interleaves it with the segmentation sweeps."""
    from segmentalist_b200 import synth
    D = 16
    for tag, n_iter, consider_unassigned, kw in (
            ("plain", 3, False, {}),
            ("all_anneal", 2, True, {"anneal_schedule": "linear", "anneal_start_temp_inv": 0.4})):
        rng = np.random.RandomState(11)
        centres = synth.cluster_centres(6, D, rng)
        n = 220
        X = synth._unit_rows(centres[rng.randint(0, 6, n)] + 0.1 * rng.standard_normal((n, D)).astype(np.float32))
        assign = rng.randint(0, 9, n)
        assign[rng.rand(n) < 0.3] = -1                                  # unassigned candidate segments
        random.seed(4)
        np.random.seed(4)
        prior = ref_fv.FixedVarPrior(0.002 * np.ones(D), np.zeros(D), 0.002 * np.ones(D) / 0.05)
        am = ref_fbgmm.FBGMM(X, prior, 5., 12, assign.copy(), covariance_type="fixed", lms=0.8)
        d = {"X": X, "init_assignments": am.components.assignments.copy(), "K_max": np.array(12),
             "alpha": np.array(5.), "lms": np.array(0.8), "consider_unassigned": np.array(consider_unassigned)}
        with Tap() as tap:
            rec = am.gibbs_sample(n_iter, consider_unassigned=consider_unassigned, **kw)
        c = am.components
        d["uniforms"] = np.array(tap.uniforms)
        d["anneal_temp"] = np.array(rec["anneal_temp"], dtype=np.float64)
        for key in ("log_marg", "log_prob_z", "log_prob_X_given_z", "components"):
            d["rec_" + key] = np.array(rec[key], dtype=np.float64)
        d["assignments"] = c.assignments.copy()
        d["counts"] = c.counts.copy()
        d["K"] = np.array(c.K)
        d["mu_N_numerators"] = c.mu_N_numerators.copy()
        d["precision_Ns"] = c.precision_Ns.copy()
        d["log_prod_precision_preds"] = c.log_prod_precision_preds.copy()
        np.savez_compressed(os.path.join(OUT, "fbgmm_gibbs_%s.npz" % tag), **d)
        print("fbgmm_gibbs", tag, "log_marg", rec["log_marg"], "K", c.K, "uniforms", len(tap.uniforms))
    # segmenter with in-between acoustic-model resampling
    mats, vids, durs, lms = synth.make_corpus_dicts(12, D=D, K_true=5, n_min=3, n_max=8, n_slices_max=4, noise=0.08, seed=33)
    random.seed(6)
    np.random.seed(6)
    prior = ref_fv.FixedVarPrior(0.002 * np.ones(D), np.zeros(D), 0.002 * np.ones(D) / 0.05)
    seg = ref_uni.UnigramAcousticWordseg(
        ref_fbgmm.FBGMM, 10., 9, prior, mats, vids, durs, lms, p_boundary_init=0.5,
        beta_sent_boundary=-1, n_slices_max=4, lms=1.0, wip=0.0, fb_type="standard")
    ref_uni.i_debug_monitor = -1
    d = pack_dicts("in_", mats, vids, durs, lms)
    with Tap() as tap:
        rec = seg.gibbs_sample(2, am_n_iter=2)
    c = seg.acoustic_model.components
    d["uniforms"] = np.array(tap.uniforms)
    d["orders"] = np.array(tap.orders)
    for key in ("log_marg", "log_marg*length", "log_prob_z", "log_prob_X_given_z", "components", "n_tokens"):
        d["rec_" + key] = np.array(rec[key], dtype=np.float64)
    d["boundaries"] = seg.utterances.boundaries.copy()
    d["assignments"] = c.assignments.copy()
    d["counts"] = c.counts.copy()
    d["K"] = np.array(c.K)
    np.savez_compressed(os.path.join(OUT, "unigram_am_iter.npz"), **d)
    print("unigram_am_iter log_marg", rec["log_marg"], "K", c.K, "uniforms", len(tap.uniforms))


def golden_diag(ref_diag, ref_niw, ref_fbgmm, ref_uni):
    """Diagonal-covariance components (gaussian_components_diag.py): statistics after add/del incl.
    a component deletion, predictive scores, whole-model Gibbs and a segmenter run (BASELINE config 5
    family), all produced by the reference."""
    from segmentalist_b200 import synth
    D = 12
    rng = np.random.RandomState(5)
    m_0 = 0.2 * rng.rand(D) - 0.1
    prior_args = dict(m_0=m_0, k_0=0.05, v_0=D + 3, S_0=0.02 * rng.rand(D) + 0.01)
    centres = synth.cluster_centres(5, D, rng)
    n = 90
    X = synth._unit_rows(centres[rng.randint(0, 5, n)] + 0.1 * rng.standard_normal((n, D)).astype(np.float32))
    assign = rng.randint(0, 6, n)
    assign[rng.rand(n) < 0.25] = -1
    assign[assign == 5] = -1
    assign[:2] = 5                                     # a two-item component that will be deleted
    d = {"X": X, "m_0": m_0, "k_0": np.array(prior_args["k_0"]), "v_0": np.array(prior_args["v_0"]),
         "S_0": prior_args["S_0"], "init_assignments": assign.copy()}
    c = ref_diag.GaussianComponentsDiag(X, ref_niw.NIW(**prior_args), assign.copy(), K_max=9)
    probe = np.where(assign == -1)[0][:6]
    d["probe"] = probe
    d["post_pred0"] = np.array([c.log_post_pred(int(i)) for i in probe])
    d["prior0"] = np.array([c.log_prior(int(i)) for i in probe])
    d["log_marg0"] = np.array(c.log_marg())
    c.del_item(0)
    c.del_item(1)                                      # deletes component 5 (the last one)
    c.del_item(int(np.where(assign == 0)[0][0]))
    c.add_item(int(probe[0]), c.K)                     # new component
    c.add_item(int(probe[1]), 2)
    d["assignments1"] = c.assignments.copy()
    d["counts1"] = c.counts.copy()
    d["K1"] = np.array(c.K)
    d["m_N_numerators1"] = c.m_N_numerators.copy()
    d["S_N_partials1"] = c.S_N_partials.copy()
    d["log_prod_vars1"] = c.log_prod_vars.copy()
    d["inv_vars1"] = c.inv_vars.copy()
    d["post_pred1"] = np.array([c.log_post_pred(int(i)) for i in probe[2:]])
    d["log_marg1"] = np.array(c.log_marg())
    # whole-model Gibbs
    random.seed(8)
    np.random.seed(8)
    am = ref_fbgmm.FBGMM(X, ref_niw.NIW(**prior_args), 3., 9, assign.copy(), covariance_type="diag", lms=0.9)
    d["gs_init_assignments"] = am.components.assignments.copy()
    with Tap() as tap:
        rec = am.gibbs_sample(2, consider_unassigned=False)
    cc = am.components
    d["gs_uniforms"] = np.array(tap.uniforms)
    d["gs_assignments"] = cc.assignments.copy()
    d["gs_counts"] = cc.counts.copy()
    d["gs_K"] = np.array(cc.K)
    d["gs_log_marg"] = np.array(rec["log_marg"], dtype=np.float64)
    d["gs_m_N_numerators"] = cc.m_N_numerators.copy()
    d["gs_log_marg_i"] = np.array([am.log_marg_i(int(i)) for i in probe])
    np.savez_compressed(os.path.join(OUT, "diag_components.npz"), **d)
    print("diag components K", int(d["K1"]), "gibbs log_marg", rec["log_marg"], "uniforms", len(tap.uniforms))
    # segmenter with diagonal covariance
    mats, vids, durs, lms = synth.make_corpus_dicts(10, D=D, K_true=4, n_min=3, n_max=8, n_slices_max=4, noise=0.1, seed=41)
    random.seed(9)
    np.random.seed(9)
    seg = ref_uni.UnigramAcousticWordseg(
        ref_fbgmm.FBGMM, 5., 8, ref_niw.NIW(**prior_args), mats, vids, durs, lms, p_boundary_init=0.5,
        beta_sent_boundary=-1, n_slices_max=4, lms=1.0, wip=0.0, fb_type="standard", covariance_type="diag")
    ref_uni.i_debug_monitor = -1
    e = pack_dicts("in_", mats, vids, durs, lms)
    for k_, v_ in (("m_0", m_0), ("k_0", np.array(prior_args["k_0"])), ("v_0", np.array(prior_args["v_0"])),
                   ("S_0", prior_args["S_0"])):
        e[k_] = v_
    e["init_boundaries"] = seg.utterances.boundaries.copy()
    e["init_assignments"] = seg.acoustic_model.components.assignments.copy()
    with Tap() as tap:
        rec = seg.gibbs_sample(3)
    cc = seg.acoustic_model.components
    e["uniforms"] = np.array(tap.uniforms)
    e["orders"] = np.array(tap.orders)
    for key in ("log_marg", "log_marg*length", "log_prob_z", "log_prob_X_given_z", "components", "n_tokens"):
        e["rec_" + key] = np.array(rec[key], dtype=np.float64)
    e["boundaries"] = seg.utterances.boundaries.copy()
    e["assignments"] = cc.assignments.copy()
    e["counts"] = cc.counts.copy()
    e["K"] = np.array(cc.K)
    np.savez_compressed(os.path.join(OUT, "unigram_diag.npz"), **e)
    print("unigram_diag log_marg", rec["log_marg"], "K", cc.K, "uniforms", len(tap.uniforms))


def golden_bigram(ref_bi, ref_fv, ref_lm):
    """BigramAcousticWordseg (bigram_acoustic_wordseg.py) with fb_type="unigram": unigram segmentation,
    component assignments sampled under the smoothed bigram LM (bigram_lms.py) -- SURVEY 8f rank 4 /
    BASELINE configs[3].  Plus known-answer vectors of BigramSmoothLM itself."""
    from segmentalist_b200 import synth
    # --- the LM alone (the data of the reference's own main(), bigram_lms.py:118-152)
    lm = ref_lm.BigramSmoothLM(0.1, 1., 2., 5)
    data = [[1, 1, 3, 4, 0], [4, 4], [1, 0, 2, 2, 2, 2, 3, 1], [3, 3, 1]]
    lm.counts_from_data(data)
    d = {"lm_unigram_counts": lm.unigram_counts.copy(), "lm_bigram_counts": lm.bigram_counts.copy(),
         "lm_prob_vec_i": lm.prob_vec_i(), "lm_log_prob_vec_i": lm.log_prob_vec_i(),
         "lm_prob_vec_given_j": np.array([lm.prob_vec_given_j(j) for j in range(5)]),
         "lm_log_prob_vec_given_j": np.array([lm.log_prob_vec_given_j(j) for j in range(5)])}
    lm.remove_counts_from_utterance(data[2])
    d["lm_unigram_counts_removed"] = lm.unigram_counts.copy()
    d["lm_bigram_counts_removed"] = lm.bigram_counts.copy()
    np.savez_compressed(os.path.join(OUT, "bigram_lm.npz"), **d)
    # --- the segmenter
    for tag, n_iter, lms_, kw in (("plain", 3, 1.0, {}),
                                  ("anneal", 3, 0.8, {"anneal_schedule": "linear", "anneal_start_temp_inv": 0.5,
                                                      "anneal_gibbs_am": True}),
                                  ("assign_only", 2, 1.0, {"assignments_only": True})):
        mats, vids, durs, lms = synth.make_corpus_dicts(
            14, D=16, K_true=5, n_min=3, n_max=9, n_slices_max=4, noise=0.08, seed=23)
        random.seed(6)
        np.random.seed(6)
        D = 16
        prior = ref_fv.FixedVarPrior(0.002 * np.ones(D), np.zeros(D), 0.002 * np.ones(D) / 0.05)
        lm_params = {"type": "smooth", "intrp_lambda": 0.15, "a": 1.5, "b": 2.5}
        seg = ref_bi.BigramAcousticWordseg(
            9, prior, lm_params, mats, vids, durs, lms, p_boundary_init=0.5, beta_sent_boundary=-1,
            n_slices_max=4, lms=lms_, wip=-0.3, fb_type="unigram", time_power_term=1.1)
        ref_bi.i_debug_monitor = -1
        d = pack_dicts("in_", mats, vids, durs, lms)
        d["lm_params"] = np.array([lm_params["intrp_lambda"], lm_params["a"], lm_params["b"]])
        d["lms"] = np.array(lms_)
        d["init_boundaries"] = seg.utterances.boundaries.copy()
        d["init_assignments"] = seg.acoustic_model.components.assignments.copy()
        d["init_unigram_counts"] = seg.lm.unigram_counts.copy()
        d["init_bigram_counts"] = seg.lm.bigram_counts.copy()
        d["init_log_prob_z"] = np.array(seg.log_prob_z())
        with Tap() as tap:
            rec = seg.gibbs_sample(n_iter, **kw)
        c = seg.acoustic_model.components
        d["uniforms"] = np.array(tap.uniforms)
        d["orders"] = np.array(tap.orders)
        for key in ("log_marg", "log_marg*length", "log_prob_z", "log_prob_X_given_z", "components", "n_tokens",
                    "anneal_temp"):
            d["rec_" + key] = np.array(rec[key], dtype=np.float64)
        d["boundaries"] = seg.utterances.boundaries.copy()
        d["assignments"] = c.assignments.copy()
        d["counts"] = c.counts.copy()
        d["K"] = np.array(c.K)
        d["mu_N_numerators"] = c.mu_N_numerators.copy()
        d["unigram_counts"] = seg.lm.unigram_counts.copy()
        d["bigram_counts"] = seg.lm.bigram_counts.copy()
        np.savez_compressed(os.path.join(OUT, "bigram_%s.npz" % tag), **d)
        print("bigram", tag, "log_marg", rec["log_marg"], "K", c.K, "uniforms", len(tap.uniforms),
              "tied", np.array_equal(c.counts, seg.lm.unigram_counts))


def main():
    only = set(sys.argv[1:])                 # e.g. `make_golden.py fbgmm_gibbs` regenerates one family
    run = lambda name: (not only) or (name in only)
    os.makedirs(OUT, exist_ok=True)
    tmp = build_shimmed_reference()
    try:
        ref_uni = importlib.import_module("segmentalist.unigram_acoustic_wordseg")
        ref_km = importlib.import_module("segmentalist.kmeans_acoustic_wordseg")
        ref_fv = importlib.import_module("segmentalist.gaussian_components_fixedvar")
        ref_fbgmm = importlib.import_module("segmentalist.fbgmm")
        ref_kc = importlib.import_module("segmentalist.kmeans_components")
        ref_kmeans = importlib.import_module("segmentalist.kmeans")
        # self-check: the reference's own tests for this path pass under the shim
        r = subprocess.call([sys.executable, "-m", "pytest", "-q", "-x",
                             "segmentalist/tests/test_unigram_acoustic_wordseg.py",
                             "segmentalist/tests/test_gaussian_components_fixedvar.py",
                             "segmentalist/tests/test_kmeans_components.py"], cwd=tmp)
        assert r == 0, "shimmed reference fails its own tests"
        if run("dp"):
            golden_dp(ref_uni, ref_km)
        if run("fixedvar"):
            golden_fixedvar(ref_fv, ref_fbgmm)
        if run("kmeans_scoring"):
            golden_kmeans_scoring(ref_kc, ref_kmeans)
        if run("unigram"):
            golden_unigram(ref_uni, ref_fbgmm, ref_fv)
        if run("kmeans_wordseg"):
            golden_kmeans_wordseg(ref_km)
        if run("fbgmm_gibbs"):
            golden_fbgmm_gibbs(ref_fbgmm, ref_fv, ref_uni)
        if run("diag"):
            golden_diag(importlib.import_module("segmentalist.gaussian_components_diag"),
                        importlib.import_module("segmentalist.niw"), ref_fbgmm, ref_uni)
        if run("bigram"):
            golden_bigram(importlib.import_module("segmentalist.bigram_acoustic_wordseg"), ref_fv,
                          importlib.import_module("segmentalist.bigram_lms"))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
