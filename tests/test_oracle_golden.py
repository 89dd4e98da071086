"""
CPU tests: pin the oracle (oracle/seg_oracle.py + seg_oracle_c.c) against
  (a) the known-answer values in the reference's own tests, and
  (b) fixtures produced by running the reference (oracle/make_golden.py).
"""
import random

import numpy as np
import numpy.testing as npt
import pytest

from oracle import seg_oracle as so
from tests import _golden as G


# --------------------------------------------------------------------------- (a) reference KATs

def _three_embedding_fixture():
    # segmentalist/tests/test_unigram_acoustic_wordseg.py:16-57
    embedding_mat = np.array([
        [-0.2702691, -0.12348549, -0.20069546, -0.10067126, -0.32822475,
         -0.24878924, -0.17988801, -0.13201745, 0.66409844, -0.44816282],
        [-0.27186683, -0.12384345, -0.20049213, -0.10272419, -0.32618827,
         -0.24660945, -0.17784701, -0.13362537, 0.66524321, -0.44805479],
        [-0.2465426, -0.06354388, -0.22458388, 0.79060942, 0.48230717,
         -0.11888564, 0.06724239, -0.04977163, 0.06908087, 0.03395205]], dtype=np.float32)
    vec_ids = np.array([0, 1, 2])
    return ({"test": embedding_mat}, {"test": vec_ids}, {"test": [1, 2, 1]}, {"test": [1, 2]},
            {"test": [2]})


def _six_embedding_fixture():
    # segmentalist/tests/test_unigram_acoustic_wordseg.py:150-194
    m1 = np.array([[1.55329044, 0.82568932, 0.56011276], [1.10640768, -0.41715366, 0.30323529],
                   [1.24183824, -2.39021548, 0.02369367], [1.26094544, -0.27567053, 1.35731148],
                   [1.59711416, -0.54917262, -0.56074459], [-0.4298405, 1.39010761, -1.2608597]],
                  dtype=np.float32)
    m2 = np.array([[1.63075195, 0.25297823, -1.75406467], [-0.59324473, 0.96613426, -0.20922202],
                   [0.97066059, -1.22315308, -0.37979187], [-0.31613254, -0.07262261, -1.04392799],
                   [-1.11535652, 0.33905751, 1.85588856], [-1.08211738, 0.88559445, 0.2924617]],
                  dtype=np.float32)
    vec_ids = np.array([0, 1, 3, 2, 4, 5])
    return ({"test1": m1, "test2": m2}, {"test1": vec_ids, "test2": vec_ids.copy()},
            {"test1": [1, 2, 1, 3, 2, 1], "test2": [1, 2, 1, 3, 2, 1]},
            {"test1": [1, 2, 3], "test2": [1, 2, 3]})


def _prior(D):
    S_0 = 0.002 * np.ones(D)
    return so.FixedVarPrior(S_0, np.zeros(D), S_0 / 0.05)


def test_ref_kat_vec_embed_log_probs():
    # reference test_unigram_acoustic_wordseg.py:60-90
    random.seed(1)
    np.random.seed(1)
    mats, vids, durs, lms, seeds = _three_embedding_fixture()
    seg = so.UnigramAcousticWordseg(so.FBGMM, 10., 2, _prior(10), mats, vids, durs, lms,
                                    seed_boundaries_dict=seeds, beta_sent_boundary=-1)
    seg.gibbs_sample_i(0)
    got = seg.get_vec_embed_log_probs(seg.utterances.vec_ids[0], seg.utterances.durations[0])
    npt.assert_almost_equal(got, np.array([17.5548998, 35.103967, 17.5548998]))


def test_ref_kat_simple_sampling():
    # reference test_unigram_acoustic_wordseg.py:93-142
    random.seed(1)
    np.random.seed(1)
    mats, vids, durs, lms, seeds = _three_embedding_fixture()
    seg = so.UnigramAcousticWordseg(so.FBGMM, 10., 2, _prior(10), mats, vids, durs, lms,
                                    seed_boundaries_dict=seeds, beta_sent_boundary=-1)
    rec = seg.gibbs_sample(6)
    npt.assert_almost_equal(rec["log_marg"], [
        -11.969040866436707, -11.969040866436707, -11.969040866436707,
        -5.9368664797514707, -11.969040866436707, -5.9368664797514707])
    npt.assert_almost_equal(rec["log_prob_z"], [
        -1.4816045409242173, -1.4816045409242173, -1.4816045409242173,
        -0.69314718055994673, -1.4816045409242173, -0.69314718055994673])
    npt.assert_almost_equal(rec["log_prob_X_given_z"], [
        -10.48743632551249, -10.48743632551249, -10.48743632551249,
        -5.2437192991915236, -10.48743632551249, -5.2437192991915236])


def test_ref_kat_simple_sampling2():
    # reference test_unigram_acoustic_wordseg.py:145-231 (n_slices_max=2)
    mats, vids, durs, lms = _six_embedding_fixture()
    random.seed(1)
    np.random.seed(1)
    seg = so.UnigramAcousticWordseg(so.FBGMM, 10., 2, _prior(3), mats, vids, durs, lms,
                                    p_boundary_init=0.5, beta_sent_boundary=-1, n_slices_max=2)
    rec = seg.gibbs_sample(3)
    npt.assert_almost_equal(rec["log_marg"], [-1520.885395538874, -435.84314783538349, -435.84314783538349])
    npt.assert_almost_equal(rec["log_prob_z"], [-3.641088790277589, -2.7937909298903829, -2.7937909298903829])
    npt.assert_almost_equal(rec["log_prob_X_given_z"],
                            [-1517.2443067485965, -433.04935690549308, -433.04935690549308])


def test_ref_kat_log_prior_uses_var_0():
    # reference test_gaussian_components_fixedvar.py:16-33
    np.random.seed(1)
    D = 10
    var = 1 * np.random.rand(D)
    mu_0 = 5 * np.random.rand(D) - 2
    var_0 = 2 * np.random.rand(D)
    x = 3 * np.random.rand(D) + 4
    gmm = so.FixedVarComponents(np.array([x]), so.FixedVarPrior(var, mu_0, var_0), K_max=D)
    expected = np.sum(-0.5 * (np.log(2 * np.pi) + np.log(var_0)) - 1. / (2 * var_0) * (x - mu_0) ** 2)
    npt.assert_almost_equal(gmm.log_prior(0), expected)


def test_ref_kat_log_post_pred_vectorised_equals_loop():
    # reference test_gaussian_components_fixedvar.py:89-108
    np.random.seed(1)
    X = np.random.rand(11, 10)
    D = 10
    prior = so.FixedVarPrior(1 * np.random.rand(D), 5 * np.random.rand(D) - 2, 2 * np.random.rand(D))
    gmm = so.FixedVarComponents(X, prior, assignments=[0, 0, 0, 1, 0, 1, 3, 4, 3, 2, -1], K_max=11)
    loop = np.array([gmm.log_post_pred_k(10, k) for k in range(gmm.K)])
    npt.assert_almost_equal(gmm.log_post_pred(10), loop)


def test_ref_kat_kmeans_neg_sqrd_norm():
    # reference test_kmeans_components.py:44-79
    np.random.seed(1)
    D, N, K_true = 4, 11, 4
    z_true = np.random.randint(0, K_true, N)
    mu = np.random.randn(D, K_true) * 4.0
    X = (mu[:, z_true] + np.random.randn(D, N) * 0.7).T
    assignments = so._consecutive(np.random.randint(0, 5, N))
    comps = so.KMeansComponents(X, assignments, 5)
    for i in range(N):
        exp = [-np.linalg.norm(X[i] - comps.mean_numerators[k] / comps.counts[k]) ** 2 for k in range(comps.K)]
        npt.assert_almost_equal(comps.neg_sqrd_norm(i)[:comps.K], exp)


# --------------------------------------------------------------------------- (b) fixtures from the reference

def test_golden_dp_python_and_c():
    n_checked = 0
    for c in G.dp_cases():
        st, lp, b, al, used = so.dp_packed_c(c["vec"], c["N"], 0, c["S"], c["mode"], c["u"], c["temp"])
        if st != 0:
            # fully infeasible utterance: the reference either raised or read through a
            # negative index (undefined); the oracle flags it instead
            continue
        assert c["ok"] == 1
        n_checked += 1
        assert np.array_equal(b, c["b"]), c
        assert used == c["used"]
        if np.isfinite(c["lp"]):
            assert lp == c["lp"]
        src = so.UniformSource(c["u"])
        lp2, b2, al2 = so.dp_packed_py(c["vec"], c["N"], 0, c["S"], c["mode"], src, c["temp"])
        assert np.array_equal(b2, c["b"]) and src.pos == c["used"]
        npt.assert_allclose(lp2, c["lp"], rtol=1e-13)
        npt.assert_array_equal(al2, al)
    assert n_checked > 350


@pytest.mark.parametrize("tag", ["iso", "aniso"])
def test_golden_fixedvar(tag):
    z = G.load("fixedvar_scoring.npz")
    X = z[tag + "_X"]
    prior = so.FixedVarPrior(z[tag + "_var"], z[tag + "_mu_0"], z[tag + "_var_0"])
    am = so.FBGMM(X, prior, 10., 12, z[tag + "_assign_in"].copy(), lms=0.7)
    c = am.components
    for i in np.where(c.assignments == 2)[0]:
        c.del_item(i)
    npt.assert_array_equal(c.assignments, z[tag + "_assign_out"])
    npt.assert_array_equal(c.counts, z[tag + "_counts"])
    assert c.K == int(z[tag + "_K"])
    npt.assert_array_equal(c.mu_N_numerators, z[tag + "_mu_N_numerators"])
    npt.assert_array_equal(c.precision_Ns, z[tag + "_precision_Ns"])
    npt.assert_array_equal(c.precision_preds, z[tag + "_precision_preds"])
    npt.assert_array_equal(c.log_prod_precision_preds, z[tag + "_log_prod_precision_preds"])
    items = z[tag + "_items"]
    npt.assert_array_equal(np.array([c.log_post_pred(i) for i in items]), z[tag + "_log_post_pred"])
    npt.assert_array_equal(np.array([c.log_prior(i) for i in items]), z[tag + "_log_prior"])
    npt.assert_array_equal(np.array([am.log_marg_i(i) for i in items]), z[tag + "_log_marg_i"])
    npt.assert_allclose(am.log_marg(), z[tag + "_log_marg"], rtol=1e-14)
    npt.assert_allclose(am.log_prob_z(), z[tag + "_log_prob_z"], rtol=1e-14)


def _np_state(z):
    np.random.set_state(("MT19937", z["np_state_keys"], int(z["np_state_pos"]), 0, 0.0))


def test_golden_kmeans_scoring_and_fit():
    z = G.load("kmeans_scoring.npz")
    X = z["X"]
    _np_state(z)
    km = so.KMeans(X, int(z["K_max"]), z["assign_in"].copy())
    c = km.components
    npt.assert_array_equal(c.random_means, z["random_means"])
    npt.assert_array_equal(c.means, z["means0"])
    assert c.means.dtype == np.float32
    items = z["items"]
    got = np.array([c.neg_sqrd_norm(i) for i in items])
    assert got.dtype == np.float32
    npt.assert_array_equal(got, z["neg_sqrd_norm"])
    # the C emulation of NumPy's float32 pairwise order is bit-exact too
    import ctypes
    out = np.empty(c.K_max, dtype=np.float32)
    fp = ctypes.POINTER(ctypes.c_float)
    means = np.ascontiguousarray(c.means)
    for n, i in enumerate(items):
        x = np.ascontiguousarray(X[i])
        so.clib().orc_kmeans_neg_sqrd_norm_f32(means.ctypes.data_as(fp), x.ctypes.data_as(fp),
                                               c.K_max, c.D, out.ctypes.data_as(fp))
        npt.assert_array_equal(out, z["neg_sqrd_norm"][n])
    npt.assert_array_equal([c.argmax_neg_sqrd_norm_i(i) for i in items], z["argmax"])
    km.fit(5, consider_unassigned=False)
    npt.assert_array_equal(c.assignments, z["fit_assign"])
    npt.assert_array_equal(c.counts, z["fit_counts"])
    assert c.K == int(z["fit_K"])
    npt.assert_array_equal(c.means, z["fit_means"])
    npt.assert_array_equal(c.mean_numerators, z["fit_mean_numerators"])


@pytest.mark.parametrize("tag,fb_type", [("ffbs", "standard"), ("ffbs_anneal", "standard"),
                                         ("viterbi", "viterbi")])
def test_golden_unigram_gibbs(tag, fb_type):
    z = G.load("unigram_%s.npz" % tag)
    mats, vids, durs, lms = G.unpack_dicts(z)
    random.seed(2)
    np.random.seed(2)
    D = 16
    prior = so.FixedVarPrior(0.002 * np.ones(D), np.zeros(D), 0.002 * np.ones(D) / 0.05)
    src = so.UniformSource(z["uniforms"])
    seg = so.UnigramAcousticWordseg(
        so.FBGMM, 10., 9, prior, mats, vids, durs, lms, p_boundary_init=0.5, beta_sent_boundary=-1,
        n_slices_max=4, lms=1.0, wip=-0.3, fb_type=fb_type, time_power_term=1.1, uniform=src)
    npt.assert_array_equal(seg.utterances.boundaries, z["init_boundaries"])
    npt.assert_array_equal(seg.acoustic_model.components.assignments, z["init_assignments"])
    n_iter = len(z["rec_log_marg"])
    temps = list(z["anneal_temp"])
    rec = seg.gibbs_sample(n_iter, anneal_temps=temps, anneal_gibbs_am=(tag == "ffbs_anneal"),
                           utt_orders=z["orders"])
    c = seg.acoustic_model.components
    assert src.pos == len(z["uniforms"])
    npt.assert_array_equal(seg.utterances.boundaries, z["boundaries"])
    npt.assert_array_equal(c.assignments, z["assignments"])
    npt.assert_array_equal(c.counts, z["counts"])
    assert c.K == int(z["K"])
    npt.assert_allclose(rec["log_marg"], z["rec_log_marg"], rtol=1e-12)
    npt.assert_allclose(rec["log_marg*length"], z["rec_log_marg*length"], rtol=1e-12)
    npt.assert_allclose(rec["log_prob_z"], z["rec_log_prob_z"], rtol=1e-12)
    npt.assert_allclose(c.mu_N_numerators, z["mu_N_numerators"], rtol=1e-13, atol=1e-12)


@pytest.mark.parametrize("tag", ["plain", "all_anneal"])
def test_golden_fbgmm_gibbs(tag):
    """FBGMM.gibbs_sample (fbgmm.py:288-420, fixed variance) as run by the reference: the oracle port
    reproduces assignments, counts, statistics and the record under the recorded uniform stream."""
    z = G.load("fbgmm_gibbs_%s.npz" % tag)
    D = z["X"].shape[1]
    prior = so.FixedVarPrior(0.002 * np.ones(D), np.zeros(D), 0.002 * np.ones(D) / 0.05)
    src = so.UniformSource(z["uniforms"])
    am = so.FBGMM(z["X"], prior, float(z["alpha"]), int(z["K_max"]), z["init_assignments"].copy(),
                  lms=float(z["lms"]), uniform=src)
    log_margs = []
    for temp in z["anneal_temp"]:
        am.gibbs_sample(1, consider_unassigned=bool(z["consider_unassigned"]), anneal_temp=float(temp))
        log_margs.append(am.log_marg())
    c = am.components
    assert src.pos == len(z["uniforms"])
    npt.assert_array_equal(c.assignments, z["assignments"])
    npt.assert_array_equal(c.counts, z["counts"])
    assert c.K == int(z["K"])
    npt.assert_allclose(log_margs, z["rec_log_marg"], rtol=1e-12)
    npt.assert_allclose(c.mu_N_numerators, z["mu_N_numerators"], rtol=1e-13, atol=1e-12)
    npt.assert_allclose(c.log_prod_precision_preds, z["log_prod_precision_preds"], rtol=1e-13)


def test_golden_unigram_with_am_resampling():
    """UnigramAcousticWordseg.gibbs_sample(2, am_n_iter=2): segmentation sweeps interleaved with
    whole-model FBGMM.gibbs_sample (unigram_acoustic_wordseg.py:440-443)."""
    z = G.load("unigram_am_iter.npz")
    mats, vids, durs, lms = G.unpack_dicts(z)
    random.seed(6)
    np.random.seed(6)
    D = 16
    prior = so.FixedVarPrior(0.002 * np.ones(D), np.zeros(D), 0.002 * np.ones(D) / 0.05)
    src = so.UniformSource(z["uniforms"])
    seg = so.UnigramAcousticWordseg(
        so.FBGMM, 10., 9, prior, mats, vids, durs, lms, p_boundary_init=0.5, beta_sent_boundary=-1,
        n_slices_max=4, lms=1.0, wip=0.0, fb_type="standard", uniform=src)
    rec = seg.gibbs_sample(2, am_n_iter=2, utt_orders=z["orders"])
    c = seg.acoustic_model.components
    assert src.pos == len(z["uniforms"])
    npt.assert_array_equal(seg.utterances.boundaries, z["boundaries"])
    npt.assert_array_equal(c.assignments, z["assignments"])
    npt.assert_array_equal(c.counts, z["counts"])
    npt.assert_allclose(rec["log_marg"], z["rec_log_marg"], rtol=1e-12)


def _diag_prior(z):
    return so.NIW(m_0=z["m_0"], k_0=float(z["k_0"]), v_0=int(z["v_0"]), S_0=z["S_0"])


def test_ref_kat_diag_students_t():
    """The reference's own analytic checks for the diagonal components
    (tests/test_gaussian_components_diag.py:17-86,229-256): prior and posterior predictive equal
    the product of univariate Student's t densities; vectorised == per-component loop."""
    np.random.seed(1)
    D = 10
    m_0 = 5 * np.random.rand(D) - 2
    k_0 = np.random.randint(15)
    v_0 = D + np.random.randint(5)
    S_0 = 2 * np.random.rand(D) + 3
    prior = so.NIW(m_0=m_0, k_0=k_0, v_0=v_0, S_0=S_0)
    x = 3 * np.random.rand(D) + 4
    gmm = so.DiagComponents(np.array([x]), prior)
    expected = np.sum([so.students_t(x[i], m_0[i], S_0[i] * (k_0 + 1) / (k_0 * v_0), v_0) for i in range(D)])
    npt.assert_almost_equal(gmm.log_prior(0), expected)
    N = 12
    X = 5 * np.random.rand(N, D) - 1
    gmm = so.DiagComponents(X, prior)
    for i in range(N):
        gmm.add_item(i, 0)
    k_N, v_N = k_0 + N, v_0 + N
    m_N = (k_0 * m_0 + N * X.mean(axis=0)) / k_N
    S_N = S_0 + np.square(X).sum(axis=0) + k_0 * np.square(m_0) - k_N * np.square(m_N)
    expected = np.sum([so.students_t(X[0, i], m_N[i], S_N[i] * (k_N + 1) / (k_N * v_N), v_N) for i in range(D)])
    npt.assert_almost_equal(gmm.log_post_pred_k(0, 0), expected)
    gmm2 = so.DiagComponents(X, prior, np.arange(N) % 3, K_max=5)
    npt.assert_almost_equal(gmm2.log_post_pred(4), [gmm2.log_post_pred_k(4, k) for k in range(gmm2.K)])


def test_golden_diag_components():
    """DiagComponents == the reference's GaussianComponentsDiag: predictive scores, statistics after
    add/del incl. a component deletion, whole-model Gibbs under the recorded uniforms."""
    z = G.load("diag_components.npz")
    c = so.DiagComponents(z["X"], _diag_prior(z), z["init_assignments"].copy(), K_max=9)
    probe = z["probe"]
    npt.assert_array_equal(np.array([c.log_post_pred(int(i)) for i in probe]), z["post_pred0"])
    npt.assert_array_equal(np.array([c.log_prior(int(i)) for i in probe]), z["prior0"])
    assert c.log_marg() == float(z["log_marg0"])
    assign = z["init_assignments"]
    c.del_item(0)
    c.del_item(1)
    c.del_item(int(np.where(assign == 0)[0][0]))
    c.add_item(int(probe[0]), c.K)
    c.add_item(int(probe[1]), 2)
    npt.assert_array_equal(c.assignments, z["assignments1"])
    npt.assert_array_equal(c.counts, z["counts1"])
    assert c.K == int(z["K1"])
    npt.assert_array_equal(c.m_N_numerators, z["m_N_numerators1"])
    npt.assert_array_equal(c.S_N_partials, z["S_N_partials1"])
    npt.assert_array_equal(c.log_prod_vars, z["log_prod_vars1"])
    npt.assert_array_equal(c.inv_vars, z["inv_vars1"])
    npt.assert_array_equal(np.array([c.log_post_pred(int(i)) for i in probe[2:]]), z["post_pred1"])
    src = so.UniformSource(z["gs_uniforms"])
    am = so.FBGMM(z["X"], _diag_prior(z), 3., 9, z["init_assignments"].copy(), covariance_type="diag", lms=0.9,
                  uniform=src)
    npt.assert_array_equal(am.components.assignments, z["gs_init_assignments"])
    lm = []
    for _ in range(2):
        am.gibbs_sample(1, consider_unassigned=False)
        lm.append(am.log_marg())
    assert src.pos == len(z["gs_uniforms"])
    npt.assert_array_equal(am.components.assignments, z["gs_assignments"])
    npt.assert_array_equal(am.components.counts, z["gs_counts"])
    npt.assert_allclose(lm, z["gs_log_marg"], rtol=1e-12)
    npt.assert_allclose([am.log_marg_i(int(i)) for i in probe], z["gs_log_marg_i"], rtol=1e-13)


def test_golden_unigram_diag():
    """UnigramAcousticWordseg with covariance_type='diag' (BASELINE config 5 family)."""
    z = G.load("unigram_diag.npz")
    mats, vids, durs, lms = G.unpack_dicts(z)
    random.seed(9)
    np.random.seed(9)
    src = so.UniformSource(z["uniforms"])
    seg = so.UnigramAcousticWordseg(
        so.FBGMM, 5., 8, _diag_prior(z), mats, vids, durs, lms, p_boundary_init=0.5, beta_sent_boundary=-1,
        n_slices_max=4, lms=1.0, wip=0.0, fb_type="standard", covariance_type="diag", uniform=src)
    npt.assert_array_equal(seg.utterances.boundaries, z["init_boundaries"])
    npt.assert_array_equal(seg.acoustic_model.components.assignments, z["init_assignments"])
    rec = seg.gibbs_sample(3, utt_orders=z["orders"])
    c = seg.acoustic_model.components
    assert src.pos == len(z["uniforms"])
    npt.assert_array_equal(seg.utterances.boundaries, z["boundaries"])
    npt.assert_array_equal(c.assignments, z["assignments"])
    npt.assert_array_equal(c.counts, z["counts"])
    npt.assert_allclose(rec["log_marg"], z["rec_log_marg"], rtol=1e-12)
    npt.assert_allclose(rec["log_marg*length"], z["rec_log_marg*length"], rtol=1e-12)


@pytest.mark.parametrize("init", ["spread", "rand"])
def test_golden_kmeans_wordseg(init):
    z = G.load("kmeans_wordseg.npz")
    mats, vids, durs, lms = G.unpack_dicts(z)
    p = init + "_"
    random.seed(4)
    np.random.seed(4)
    seg = so.SegmentalKMeansWordseg(5, mats, vids, durs, lms, p_boundary_init=0.5, n_slices_max=6,
                                    init_am_assignments=init, wip=0)
    c = seg.acoustic_model.components
    npt.assert_array_equal(seg.utterances.boundaries, z[p + "init_boundaries"])
    npt.assert_array_equal(c.assignments, z[p + "init_assignments"])
    npt.assert_array_equal(c.random_means, z[p + "random_means"])
    # frozen sweep (new mode) == the reference's pure functions applied without interleaved updates
    import copy
    fz = copy.deepcopy(seg)
    total, _ = so.frozen_kmeans_sweep(fz)
    fc = fz.acoustic_model.components
    assert total == float(z[p + "frozen_total"])
    npt.assert_array_equal(fz.utterances.boundaries, z[p + "frozen_boundaries"])
    npt.assert_array_equal(fc.assignments, z[p + "frozen_assignments"])
    npt.assert_array_equal(fc.counts, z[p + "frozen_counts"])
    npt.assert_array_equal(fc.means, z[p + "frozen_means"])
    npt.assert_array_equal(fc.mean_numerators, z[p + "frozen_mean_numerators"])
    # sequential sweeps
    rec = seg.segment(3, n_iter_inbetween_kmeans=2, utt_orders=z[p + "orders"])
    npt.assert_array_equal(seg.utterances.boundaries, z[p + "boundaries"])
    npt.assert_array_equal(c.assignments, z[p + "assignments"])
    npt.assert_array_equal(c.counts, z[p + "counts"])
    npt.assert_array_equal(c.means, z[p + "means"])
    npt.assert_array_equal(c.mean_numerators, z[p + "mean_numerators"])
    npt.assert_array_equal(rec["sum_neg_len_sqrd_norm"], z[p + "rec_sum_neg_len_sqrd_norm"])
    npt.assert_array_equal(rec["components"], z[p + "rec_components"])


# ---------------------------------------------------------------------------
# bigram LM + bigram cluster sampling (BASELINE configs[3])
# ---------------------------------------------------------------------------

def test_golden_bigram_lm():
    """BigramSmoothLM (bigram_lms.py) on the data of the reference's own main() (:118-152)."""
    z = G.load("bigram_lm.npz")
    lm = so.BigramSmoothLM(0.1, 1., 2., 5)
    data = [[1, 1, 3, 4, 0], [4, 4], [1, 0, 2, 2, 2, 2, 3, 1], [3, 3, 1]]
    lm.counts_from_data(data)
    npt.assert_array_equal(lm.unigram_counts, z["lm_unigram_counts"])
    npt.assert_array_equal(lm.bigram_counts, z["lm_bigram_counts"])
    npt.assert_array_equal(lm.prob_vec_i(), z["lm_prob_vec_i"])
    npt.assert_array_equal(lm.log_prob_vec_i(), z["lm_log_prob_vec_i"])
    for j in range(5):
        npt.assert_array_equal(lm.prob_vec_given_j(j), z["lm_prob_vec_given_j"][j])
        npt.assert_array_equal(lm.log_prob_vec_given_j(j), z["lm_log_prob_vec_given_j"][j])
    # the reference's own checks (bigram_lms.py:136-152; its literal constants assume Python 2's integer a/K)
    for i in range(5):
        assert lm.prob_vec_i()[i] == lm.prob_i(i)
        npt.assert_allclose(lm.prob_vec_given_j(3)[i], lm.prob_i_given_j(i, 3), rtol=1e-15)
    npt.assert_allclose(lm.prob_i(1), (5. + 1. / 5) / (18 + 1.), rtol=1e-15)
    lm.remove_counts_from_utterance(data[2])
    npt.assert_array_equal(lm.unigram_counts, z["lm_unigram_counts_removed"])
    npt.assert_array_equal(lm.bigram_counts, z["lm_bigram_counts_removed"])


def _bigram_kw(tag):
    if tag == "anneal":
        return 0.8, {"anneal_gibbs_am": True}, True
    if tag == "assign_only":
        return 1.0, {"assignments_only": True}, False
    return 1.0, {}, False


@pytest.mark.parametrize("tag", ["plain", "anneal", "assign_only"])
def test_golden_bigram_gibbs(tag):
    """BigramAcousticWordseg(fb_type="unigram").gibbs_sample: identical samples, LM counts and traces
    as the reference under the recorded uniform stream / utterance orders."""
    z = G.load("bigram_%s.npz" % tag)
    mats, vids, durs, lms = G.unpack_dicts(z)
    lam, a, b = (float(v) for v in z["lm_params"])
    lms_, kw, annealed = _bigram_kw(tag)
    random.seed(6)
    np.random.seed(6)
    D = 16
    prior = so.FixedVarPrior(0.002 * np.ones(D), np.zeros(D), 0.002 * np.ones(D) / 0.05)
    seg = so.BigramAcousticWordseg(
        9, prior, {"type": "smooth", "intrp_lambda": lam, "a": a, "b": b}, mats, vids, durs, lms,
        p_boundary_init=0.5, beta_sent_boundary=-1, n_slices_max=4, lms=lms_, wip=-0.3, fb_type="unigram",
        time_power_term=1.1, uniform=so.UniformSource(z["uniforms"]))
    npt.assert_array_equal(seg.utterances.boundaries, z["init_boundaries"])
    npt.assert_array_equal(seg.acoustic_model.components.assignments, z["init_assignments"])
    npt.assert_array_equal(seg.lm.unigram_counts, z["init_unigram_counts"])
    npt.assert_array_equal(seg.lm.bigram_counts, z["init_bigram_counts"])
    npt.assert_allclose(seg.log_prob_z(), z["init_log_prob_z"], rtol=1e-13)
    n_iter = len(z["rec_log_marg"])
    rec = seg.gibbs_sample(n_iter, anneal_temps=list(z["rec_anneal_temp"]) if annealed else None,
                           utt_orders=z["orders"], **kw)
    assert seg.uniform.pos == len(z["uniforms"])
    c = seg.acoustic_model.components
    npt.assert_array_equal(seg.utterances.boundaries, z["boundaries"])
    npt.assert_array_equal(c.assignments, z["assignments"])
    npt.assert_array_equal(c.counts, z["counts"])
    assert c.K == int(z["K"])
    npt.assert_array_equal(seg.lm.unigram_counts, z["unigram_counts"])
    npt.assert_array_equal(seg.lm.bigram_counts, z["bigram_counts"])
    npt.assert_array_equal(c.counts, seg.lm.unigram_counts)          # the tie (:205-221) keeps them equal
    npt.assert_allclose(c.mu_N_numerators, z["mu_N_numerators"], rtol=1e-13, atol=1e-12)
    npt.assert_allclose(rec["log_marg"], z["rec_log_marg"], rtol=1e-12)
    npt.assert_allclose(rec["log_marg*length"], z["rec_log_marg*length"], rtol=1e-12)
    npt.assert_allclose(rec["log_prob_z"], z["rec_log_prob_z"], rtol=1e-12)
