"""
oracle/seg_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (NumPy + the small C library in seg_oracle_c.c) of the
segmentalist hot path: score every candidate segment embedding of an utterance
against every mixture component, then segment with dynamic programming.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs import this module.  segmentalist_b200/ never does.

Parity status: PINNED.  tests/test_oracle_golden.py checks this module against
(a) the known-answer values the reference's own tests hold
(segmentalist/tests/test_unigram_acoustic_wordseg.py:88,127-142,225-231,
test_gaussian_components_fixedvar.py, test_kmeans_components.py) and (b) the
fixtures in tests/golden/ that oracle/make_golden.py produced by running the
(py2->py3 shimmed, otherwise unmodified) reference in the build container.

All `file:line` citations are relative to /root/reference/segmentalist/.
The arithmetic (operand order, dtypes, NumPy reductions) follows the reference
statement by statement so that results agree to the last bit where NumPy is
deterministic; the code organisation is this repo's own.
"""
import ctypes
import math
import os
import random
import subprocess
import time

import numpy as np
from scipy.special import gammaln, logsumexp as sp_logsumexp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c(force=False):
    """Compile oracle/seg_oracle_c.c (gcc) -> oracle/libseg_oracle_c.so."""
    so = os.path.join(_HERE, "libseg_oracle_c.so")
    src = os.path.join(_HERE, "seg_oracle_c.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libseg_oracle_c.so"])
    return so


def clib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_c())
        dp = ctypes.POINTER(ctypes.c_double)
        lib.orc_logsumexp.restype = ctypes.c_double
        lib.orc_logsumexp.argtypes = [dp, ctypes.c_int]
        lib.orc_sum_log.restype = ctypes.c_double
        lib.orc_sum_log.argtypes = [dp, ctypes.c_int]
        lib.orc_sum_square_a_times_b.restype = ctypes.c_double
        lib.orc_sum_square_a_times_b.argtypes = [dp, dp, ctypes.c_int]
        lib.orc_draw.restype = ctypes.c_int
        lib.orc_draw.argtypes = [dp, ctypes.c_int, ctypes.c_double]
        lib.orc_dp_packed.restype = ctypes.c_int
        lib.orc_dp_packed.argtypes = [
            dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
            ctypes.c_double, dp, ctypes.POINTER(ctypes.c_int), dp,
            ctypes.POINTER(ctypes.c_uint8), dp]
        fp = ctypes.POINTER(ctypes.c_float)
        lib.orc_kmeans_neg_sqrd_norm_f32.restype = None
        lib.orc_kmeans_neg_sqrd_norm_f32.argtypes = [fp, fp, ctypes.c_int, ctypes.c_int, fp]
        lib.orc_kmeans_best_f32.restype = None
        lib.orc_kmeans_best_f32.argtypes = [
            fp, fp, ctypes.POINTER(ctypes.c_int64), ctypes.c_int, ctypes.c_int, ctypes.c_int,
            fp, ctypes.POINTER(ctypes.c_int32)]
        _LIB = lib
    return _LIB


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


# ---------------------------------------------------------------------------
# _cython_utils.pyx restatements (thin wrappers over the C library)
# ---------------------------------------------------------------------------

def c_logsumexp(a):
    """_cython_utils.pyx:13-25."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    return clib().orc_logsumexp(_dptr(a), a.shape[0])


def c_sum_log(y):
    """_cython_utils.pyx:52-59."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    return clib().orc_sum_log(_dptr(y), y.shape[0])


def c_sum_square_a_times_b(a, b):
    """_cython_utils.pyx:63-70."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return clib().orc_sum_square_a_times_b(_dptr(a), _dptr(b), a.shape[0])


def draw(p, u):
    """_cython_utils.pyx:75-89 / utils.py:10-21 with the uniform supplied."""
    p = np.ascontiguousarray(p, dtype=np.float64)
    return clib().orc_draw(_dptr(p), p.shape[0], float(u))


class UniformSource(object):
    """Where draws come from.  Default: Python's global `random.random`, exactly
    what the reference consumes; tests pass a recorded stream instead."""

    def __init__(self, stream=None):
        self.stream = None if stream is None else np.asarray(stream, dtype=np.float64)
        self.pos = 0

    def __call__(self):
        if self.stream is None:
            self.pos += 1
            return random.random()
        u = self.stream[self.pos]
        self.pos += 1
        return u


# ---------------------------------------------------------------------------
# Fixed-variance Gaussian components   (gaussian_components_fixedvar.py)
# ---------------------------------------------------------------------------

class FixedVarPrior(object):
    """gaussian_components_fixedvar.py:349-356."""

    def __init__(self, var, mu_0, var_0):
        self.var, self.mu_0, self.var_0 = var, mu_0, var_0


class FixedVarComponents(object):
    """Sufficient statistics of K fixed-diagonal-variance Bayesian Gaussians.

    Restates GaussianComponentsFixedVar (gaussian_components_fixedvar.py:80-338).
    """

    def __init__(self, X, prior, assignments=None, K_max=None, lm=None):
        assert K_max is not None                                    # :88-89
        self.X = X
        self.lm = lm                                                # :92 (tied bigram LM counts, :205-221)
        self.precision = 1. / prior.var                             # :84
        self.mu_0 = prior.mu_0
        self.precision_0 = 1. / prior.var_0                         # :86
        self.N, self.D = X.shape
        self.K_max = K_max
        self.mu_N_numerators = np.zeros((K_max, self.D))            # :95-99
        self.precision_Ns = np.zeros((K_max, self.D))
        self.log_prod_precision_preds = np.zeros(K_max)
        self.precision_preds = np.zeros((K_max, self.D))
        self.counts = np.zeros(K_max, dtype=np.int64)
        self.neg_half_D_log_2pi = -0.5 * self.D * math.log(2. * np.pi)   # :123
        self.K = 0
        if assignments is None:
            self.assignments = -1 * np.ones(self.N, dtype=np.int64)
        else:
            assignments = np.asarray(assignments, dtype=np.int64)   # :111-120
            assert (self.N,) == assignments.shape
            assert set(assignments).difference([-1]) == set(range(assignments.max() + 1))
            self.assignments = assignments
            for k in range(self.assignments.max() + 1):
                for i in np.where(self.assignments == k)[0]:
                    self.add_item(i, k)

    def _refresh_pred(self, k):
        """:317-325."""
        pp = self.precision_Ns[k] * self.precision / (self.precision_Ns[k] + self.precision)
        self.log_prod_precision_preds[k] = np.log(pp).sum()
        self.precision_preds[k, :] = pp

    def add_item(self, i, k):
        """:153-170."""
        assert not i == -1
        if k == self.K:
            self.K += 1
            self.mu_N_numerators[k, :] = self.precision_0 * self.mu_0
            self.precision_Ns[k, :] = self.precision_0
        self.mu_N_numerators[k, :] += self.precision * self.X[i]
        self.precision_Ns[k, :] += self.precision
        self.counts[k] += 1
        self._refresh_pred(k)
        self.assignments[i] = k

    def del_item(self, i):
        """:172-188."""
        assert not i == -1
        k = self.assignments[i]
        if k != -1:
            self.counts[k] -= 1
            self.assignments[i] = -1
            if self.counts[k] == 0:
                self.del_component(k)
            else:
                self.mu_N_numerators[k, :] -= self.precision * self.X[i]
                self.precision_Ns[k, :] -= self.precision
                self._refresh_pred(k)

    def del_component(self, k):
        """:190-221, incl. the language-model tie-in (:205-208, :218-221)."""
        self.K -= 1
        last = self.K
        if k != last and self.lm is not None:
            self.lm.unigram_counts[k] = self.lm.unigram_counts[last]
            self.lm.bigram_counts[k, :] = self.lm.bigram_counts[last, :]
            self.lm.bigram_counts[:, k] = self.lm.bigram_counts[:, last]
        if k != last:
            self.mu_N_numerators[k] = self.mu_N_numerators[last]
            self.precision_Ns[k, :] = self.precision_Ns[last, :]
            self.log_prod_precision_preds[k] = self.log_prod_precision_preds[last]
            self.precision_preds[k, :] = self.precision_preds[last, :]
            self.counts[k] = self.counts[last]
            self.assignments[np.where(self.assignments == last)] = k
        self.mu_N_numerators[last].fill(0.)
        self.precision_Ns[last, :].fill(0.)
        self.log_prod_precision_preds[last] = 0.
        self.precision_preds[last, :].fill(0.)
        self.counts[last] = 0
        if self.lm is not None:
            self.lm.unigram_counts[last] = 0
            self.lm.bigram_counts[last, :].fill(0)
            self.lm.bigram_counts[:, last].fill(0)

    def _log_prod_norm(self, i, mu, log_prod_precision_pred, precision_pred):
        """:328-338 -- scalar path through the sequential C reduction."""
        delta = self.X[i, :] - mu
        return (self.neg_half_D_log_2pi + 0.5 * log_prod_precision_pred
                - 0.5 * c_sum_square_a_times_b(delta, precision_pred))

    def log_prior(self, i):
        """:224-231.  Predictive precision of an empty slot is precision_0."""
        return self._log_prod_norm(i, self.mu_0, c_sum_log(self.precision_0), self.precision_0)

    def log_post_pred_k(self, i, k):
        """:233-239."""
        mu_N = self.mu_N_numerators[k] / self.precision_Ns[k]
        return self._log_prod_norm(i, mu_N, self.log_prod_precision_preds[k], self.precision_preds[k])

    def log_post_pred(self, i):
        """:242-253 -- vectorised over the K active components."""
        K = self.K
        mu_Ns = self.mu_N_numerators[:K] / self.precision_Ns[:K]
        deltas = mu_Ns - self.X[i]
        return (self.neg_half_D_log_2pi + 0.5 * self.log_prod_precision_preds[:K]
                - 0.5 * ((deltas * deltas) * self.precision_preds[:K]).sum(axis=1))

    def log_marg_k(self, k):
        """:261-283."""
        X = self.X[np.where(self.assignments == k)]
        N = self.counts[k]
        return np.sum(
            (N - 1) / 2. * np.log(self.precision)
            - 0.5 * N * math.log(2 * np.pi)
            - 0.5 * np.log(N / self.precision_0 + 1. / self.precision)
            - 0.5 * self.precision * np.square(X).sum(axis=0)
            - 0.5 * self.precision_0 * np.square(self.mu_0)
            + 0.5 * (
                np.square(X.sum(axis=0)) * self.precision / self.precision_0
                + np.square(self.mu_0) * self.precision_0 / self.precision
                + 2 * X.sum(axis=0) * self.mu_0
            ) / (N / self.precision_0 + 1. / self.precision))

    def log_marg(self):
        """:285-296."""
        total = 0.
        for k in range(self.K):
            total += self.log_marg_k(k)
        return total

    def get_assignments(self, list_of_i):
        return self.assignments[np.asarray(list_of_i)]


# ---------------------------------------------------------------------------
# Diagonal-covariance Gaussian components   (gaussian_components_diag.py, niw.py)
# ---------------------------------------------------------------------------

class NIW(object):
    """niw.py:7-15 (v_0 is an integer >= D: it indexes the cached gammaln table)."""

    def __init__(self, m_0, k_0, v_0, S_0):
        self.m_0, self.k_0, self.S_0 = m_0, k_0, S_0
        assert v_0 >= len(m_0), "v_0 must be larger or equal to dimension of data"
        self.v_0 = v_0


class DiagComponents(object):
    """Normal-inverse-chi-squared components with a product-of-Student's-t predictive.

    Restates GaussianComponentsDiag (gaussian_components_diag.py:82-345)."""

    def __init__(self, X, prior, assignments=None, K_max=None):
        self.X, self.prior = X, prior
        self.N, self.D = X.shape
        if K_max is None:                                            # :90-91
            K_max = self.N
        self.K_max = K_max
        assert len(prior.S_0.shape) == 1, "For diagonal covariance, S_0 needs to be vector."
        self.m_N_numerators = np.zeros((K_max, self.D))              # :97-101
        self.S_N_partials = np.zeros((K_max, self.D))
        self.log_prod_vars = np.zeros(K_max)
        self.inv_vars = np.zeros((K_max, self.D))
        self.counts = np.zeros(K_max, dtype=np.int64)
        # :119-131 caches
        self._sq_m_0 = np.square(prior.m_0)
        self._sq = np.square(X)
        n = np.concatenate([[1], np.arange(1, prior.v_0 + self.N + 2)])
        self._log_v = np.log(n)
        self._gammaln_by_2 = gammaln(n / 2.)
        self._log_pi = math.log(np.pi)
        self.K = 0
        if assignments is None:
            self.assignments = -1 * np.ones(self.N, dtype=np.int64)
        else:
            assignments = np.asarray(assignments, dtype=np.int64)
            assert (self.N,) == assignments.shape
            assert set(assignments).difference([-1]) == set(range(assignments.max() + 1))
            self.assignments = assignments
            for k in range(self.assignments.max() + 1):
                for i in np.where(self.assignments == k)[0]:
                    self.add_item(i, k)

    def _refresh(self, k):
        """:332-345."""
        k_N = self.prior.k_0 + self.counts[k]
        v_N = self.prior.v_0 + self.counts[k]
        m_N = self.m_N_numerators[k] / k_N
        var = (k_N + 1.) / (k_N * v_N) * (self.S_N_partials[k] - k_N * np.square(m_N))
        self.log_prod_vars[k] = np.log(var).sum()
        self.inv_vars[k, :] = 1. / var

    def add_item(self, i, k):
        """:162-177."""
        if k == self.K:
            self.K += 1
            self.m_N_numerators[k, :] = self.prior.k_0 * self.prior.m_0
            self.S_N_partials[k, :] = self.prior.S_0 + self.prior.k_0 * self._sq_m_0
        self.m_N_numerators[k, :] += self.X[i]
        self.S_N_partials[k, :] += self._sq[i]
        self.counts[k] += 1
        self._refresh(k)
        self.assignments[i] = k

    def del_item(self, i):
        """:179-194."""
        k = self.assignments[i]
        if k != -1:
            self.counts[k] -= 1
            self.assignments[i] = -1
            if self.counts[k] == 0:
                self.del_component(k)
            else:
                self.m_N_numerators[k, :] -= self.X[i]
                self.S_N_partials[k, :] -= self._sq[i]
                self._refresh(k)

    def del_component(self, k):
        """:196-214."""
        self.K -= 1
        last = self.K
        if k != last:
            self.m_N_numerators[k] = self.m_N_numerators[last]
            self.S_N_partials[k, :] = self.S_N_partials[last, :]
            self.log_prod_vars[k] = self.log_prod_vars[last]
            self.inv_vars[k, :] = self.inv_vars[last, :]
            self.counts[k] = self.counts[last]
            self.assignments[np.where(self.assignments == last)] = k
        self.m_N_numerators[last].fill(0.)
        self.S_N_partials[last, :].fill(0.)
        self.log_prod_vars[last] = 0.
        self.inv_vars[last, :].fill(0.)
        self.counts[last] = 0

    def _log_prod_students_t(self, i, mu, log_prod_var, inv_var, v):
        """:347-360."""
        delta = self.X[i, :] - mu
        return (self.D * (self._gammaln_by_2[v + 1] - self._gammaln_by_2[v]
                          - 0.5 * self._log_v[v] - 0.5 * self._log_pi)
                - 0.5 * log_prod_var
                - (v + 1.) / 2. * (np.log(1. + 1. / v * np.square(delta) * inv_var)).sum())

    def log_prior(self, i):
        """:216-223."""
        pr = self.prior
        var = (pr.k_0 + 1.) / (pr.k_0 * pr.v_0) * pr.S_0
        return self._log_prod_students_t(i, pr.m_0, np.log(var).sum(), 1. / var, pr.v_0)

    def log_post_pred_k(self, i, k):
        """:225-234."""
        k_N = self.prior.k_0 + self.counts[k]
        v_N = self.prior.v_0 + self.counts[k]
        return self._log_prod_students_t(i, self.m_N_numerators[k] / k_N, self.log_prod_vars[k],
                                         self.inv_vars[k], v_N)

    def log_post_pred(self, i):
        """:237-259 (vectorised; the row sum is einsum("ij->i") there)."""
        K = self.K
        k_Ns = self.prior.k_0 + self.counts[:K]
        v_Ns = self.prior.v_0 + self.counts[:K]
        m_Ns = self.m_N_numerators[:K] / k_Ns[:, np.newaxis]
        studentt_gammas = self._gammaln_by_2[v_Ns + 1] - self._gammaln_by_2[v_Ns]
        deltas = m_Ns - self.X[i]
        return (self.D * (studentt_gammas - 0.5 * self._log_v[v_Ns] - 0.5 * self._log_pi)
                - 0.5 * self.log_prod_vars[:K]
                - (v_Ns + 1) / 2. * np.einsum("ij->i", np.log(
                    1 + np.square(deltas) * self.inv_vars[:K] * (1. / v_Ns[:, np.newaxis]))))

    def log_marg_k(self, k):
        """:270-288."""
        pr = self.prior
        k_N = pr.k_0 + self.counts[k]
        v_N = pr.v_0 + self.counts[k]
        m_N = self.m_N_numerators[k] / k_N
        S_N = self.S_N_partials[k] - k_N * np.square(m_N)
        return (- self.counts[k] * self.D / 2. * self._log_pi
                + self.D / 2. * math.log(pr.k_0) - self.D / 2. * math.log(k_N)
                + pr.v_0 / 2. * np.log(pr.S_0).sum()
                - v_N / 2. * np.log(S_N).sum()
                + self.D * (self._gammaln_by_2[v_N] - self._gammaln_by_2[pr.v_0]))

    def log_marg(self):
        """:290-301."""
        total = 0.
        for k in range(self.K):
            total += self.log_marg_k(k)
        return total

    def get_assignments(self, list_of_i):
        return self.assignments[np.asarray(list_of_i)]


def students_t(x, mu, var, v):
    """gaussian_components_diag.py:371-380 (test helper of the reference)."""
    c = gammaln((v + 1) / 2.) - gammaln(v / 2.) - 0.5 * (math.log(v) + math.log(np.pi) + math.log(var))
    return c - (v + 1) / 2. * math.log(1 + 1. / v * (x - mu) ** 2 / var)


# ---------------------------------------------------------------------------
# K-means components   (kmeans_components.py)
# ---------------------------------------------------------------------------

class KMeansComponents(object):
    """Restates KMeansComponents (kmeans_components.py:18-266).  `means` has X's
    dtype (float32 in practice) and scoring happens in that dtype."""

    def __init__(self, X, assignments, K_max):
        self.X = X
        self.N, self.D = X.shape
        self.K_max = K_max
        self.mean_numerators = np.zeros((K_max, self.D))            # :63
        self.counts = np.zeros(K_max, dtype=np.int64)
        self.K = 0
        assignments = np.asarray(assignments, dtype=np.int64)
        assert (self.N,) == assignments.shape
        assert set(assignments).difference([-1]) == set(range(assignments.max() + 1))
        self.assignments = -1 * np.ones(self.N, dtype=np.int64)
        self.setup_random_means()                                   # :75-76
        self.means = self.random_means.copy()
        for k in range(assignments.max() + 1):
            for i in np.where(assignments == k)[0]:
                self.add_item(i, k)

    def setup_random_means(self):
        """:90-91 (consumes np.random)."""
        self.random_means = self.X[np.random.choice(range(self.N), self.K_max, replace=True), :]

    def add_item(self, i, k):
        """:93-111 incl. the k > K clamp."""
        assert not i == -1
        assert self.assignments[i] == -1
        if k > self.K:
            k = self.K
        if k == self.K:
            self.K += 1
        self.mean_numerators[k, :] += self.X[i]
        self.counts[k] += 1
        self.means[k, :] = self.mean_numerators[k, :] / self.counts[k]
        self.assignments[i] = k

    def del_item(self, i):
        """:113-132 -- an emptied component keeps its stale mean."""
        assert not i == -1
        k = self.assignments[i]
        if k != -1:
            self.counts[k] -= 1
            self.assignments[i] = -1
            self.mean_numerators[k, :] -= self.X[i]
            if self.counts[k] != 0:
                self.means[k, :] = self.mean_numerators[k, :] / self.counts[k]

    def del_component(self, k):
        """:149-166."""
        assert k < self.K
        self.K -= 1
        last = self.K
        if k != last:
            self.mean_numerators[k] = self.mean_numerators[last]
            self.counts[k] = self.counts[last]
            self.means[k, :] = self.mean_numerators[last, :] / self.counts[last]
            self.assignments[np.where(self.assignments == last)] = k
        self.mean_numerators[last].fill(0.)
        self.counts[last] = 0
        self.means[last] = self.random_means[last]

    def neg_sqrd_norm(self, i):
        """:225-226 -- over ALL K_max rows of `means`, in means' dtype."""
        deltas = self.means - self.X[i]
        return -(deltas * deltas).sum(axis=1)

    def max_neg_sqrd_norm_i(self, i):
        return np.max(self.neg_sqrd_norm(i))                        # :228-229

    def argmax_neg_sqrd_norm_i(self, i):
        return np.argmax(self.neg_sqrd_norm(i))                     # :231-232

    def sum_neg_sqrd_norm(self):
        """:234-247."""
        objective = 0
        for k in range(self.K):
            X = self.X[np.where(self.assignments == k)]
            mean = self.mean_numerators[k, :] / self.counts[k]
            deltas = mean - X
            objective += -np.sum(deltas * deltas)
        return objective

    def get_assignments(self, list_of_i):
        return self.assignments[np.asarray(list_of_i)]

    def get_max_assignments(self, list_of_i):
        return [self.argmax_neg_sqrd_norm_i(i) for i in list_of_i]  # :256-261

    def clean_components(self):
        """:263-266."""
        for k in np.where(self.counts[:self.K] == 0)[0][::-1]:
            self.del_component(k)


def _consecutive(assignments):
    """The 'make sure we have consecutive values' loop that appears at
    fbgmm.py:124-128, kmeans.py:88-92, unigram_acoustic_wordseg.py:212-216."""
    for k in range(assignments.max()):
        while len(np.nonzero(assignments == k)[0]) == 0:
            assignments[np.where(assignments > k)] -= 1
        if assignments.max() == k:
            break
    return assignments


# ---------------------------------------------------------------------------
# FBGMM   (fbgmm.py)
# ---------------------------------------------------------------------------

class FBGMM(object):
    """Finite Bayesian GMM over FixedVarComponents (fbgmm.py:27-494); only
    covariance_type == "fixed" is on the hot path."""

    def __init__(self, X, prior, alpha, K, assignments="rand", covariance_type="fixed",
                 lms=1.0, uniform=None):
        assert covariance_type in ("fixed", "diag"), "oracle covers the fixed-variance and diagonal paths"
        self.alpha, self.prior, self.covariance_type, self.lms = alpha, prior, covariance_type, lms
        self.uniform = uniform if uniform is not None else UniformSource()
        N = X.shape[0]
        if isinstance(assignments, str) and assignments == "rand":          # :116-121
            assignments = np.random.randint(0, K, N)
        elif isinstance(assignments, str) and assignments == "each-in-own":
            assignments = np.arange(N)
        assignments = _consecutive(assignments)
        if covariance_type == "diag":                                       # fbgmm.py:130-137
            self.components = DiagComponents(X, prior, assignments, K_max=K)
        else:
            self.components = FixedVarComponents(X, prior, assignments, K_max=K)

    def _log_prior_z(self, with_norm):
        c = self.components
        lp = np.log(float(self.alpha) / c.K_max + c.counts)
        if with_norm:
            lp = lp - np.log(int(c.counts.sum()) + self.alpha)
        return lp

    def log_marg_i(self, i):
        """fbgmm.py:256-285."""
        assert i != -1
        c = self.components
        log_prob_z = self.lms * (
            np.log(float(self.alpha) / c.K_max + c.counts)
            - np.log(int(c.counts.sum()) + self.alpha))
        log_prob_z[:c.K] += c.log_post_pred(i)
        log_prob_z[c.K:] += c.log_prior(i)
        return c_logsumexp(log_prob_z)

    def _assign_scores(self, i, use_lms):
        c = self.components
        lp = np.ones(c.K_max) * np.log(float(self.alpha) / c.K_max + c.counts)
        if use_lms:
            lp = self.lms * lp
        lp[:c.K] += c.log_post_pred(i)
        lp[c.K:] += c.log_prior(i)
        return lp

    def gibbs_sample_inside_loop_i(self, i, anneal_temp=1):
        """fbgmm.py:422-463."""
        c = self.components
        log_prob_z = self._assign_scores(i, True)
        if anneal_temp != 1:
            log_prob_z = log_prob_z - sp_logsumexp(log_prob_z)
            log_prob_z_anneal = 1. / anneal_temp * log_prob_z - sp_logsumexp(1. / anneal_temp * log_prob_z)
            prob_z = np.exp(log_prob_z_anneal)
        else:
            prob_z = np.exp(log_prob_z - sp_logsumexp(log_prob_z))
        assert not np.isnan(np.sum(prob_z))
        k = draw(prob_z, self.uniform())
        if k > c.K:
            k = c.K
        c.add_item(i, k)
        return k

    def map_assign_i(self, i):
        """fbgmm.py:465-494 (no lms, argmax of the normalised vector)."""
        c = self.components
        log_prob_z = self._assign_scores(i, False)
        prob_z = np.exp(log_prob_z - sp_logsumexp(log_prob_z))
        k = int(np.argmax(prob_z))
        if k > c.K:
            k = c.K
        c.add_item(i, k)
        return k

    def gibbs_sample(self, n_iter, consider_unassigned=True, anneal_temp=1):
        """fbgmm.py:288-420 (constant temperature; annealing schedules are host
        logic outside the hot path)."""
        c = self.components
        for _ in range(n_iter):
            for i in range(c.N):
                k_old = c.assignments[i]
                if not consider_unassigned and k_old == -1:
                    continue
                K_old = c.K
                diag = self.covariance_type == "diag"
                if diag:        # gaussian_components_diag.py:137-150
                    stats_old = (c.m_N_numerators[k_old].copy(), c.S_N_partials[k_old].copy(),
                                 c.log_prod_vars[k_old], c.inv_vars[k_old].copy(), c.counts[k_old])
                else:
                    stats_old = (c.mu_N_numerators[k_old].copy(), c.precision_Ns[k_old].copy(),
                                 c.log_prod_precision_preds[k_old], c.precision_preds[k_old].copy(),
                                 c.counts[k_old])
                c.del_item(i)
                log_prob_z = self._assign_scores(i, True)
                if anneal_temp != 1:
                    log_prob_z = log_prob_z - sp_logsumexp(log_prob_z)
                    la = 1. / anneal_temp * log_prob_z - sp_logsumexp(1. / anneal_temp * log_prob_z)
                    prob_z = np.exp(la)
                else:
                    prob_z = np.exp(log_prob_z - sp_logsumexp(log_prob_z))
                k = draw(prob_z, self.uniform())
                if k > c.K:
                    k = c.K
                if k == k_old and c.K == K_old:
                    if diag:
                        (c.m_N_numerators[k_old, :], c.S_N_partials[k_old, :], c.log_prod_vars[k_old],
                         c.inv_vars[k_old, :], c.counts[k_old]) = stats_old
                    else:
                        (c.mu_N_numerators[k_old, :], c.precision_Ns[k_old, :],
                         c.log_prod_precision_preds[k_old], c.precision_preds[k_old, :],
                         c.counts[k_old]) = stats_old
                    c.assignments[i] = k_old
                else:
                    c.add_item(i, k)

    def log_prob_z(self):
        """fbgmm.py:208-225."""
        c = self.components
        return (gammaln(self.alpha) - gammaln(self.alpha + np.sum(c.counts))
                + np.sum(gammaln(c.counts + float(self.alpha) / c.K_max)
                         - gammaln(self.alpha / c.K_max)))

    def log_prob_X_given_z(self):
        return self.components.log_marg()

    def log_marg(self):
        return self.log_prob_z() + self.log_prob_X_given_z()

    def get_n_assigned(self):
        return len(np.where(self.components.assignments != -1)[0])


# ---------------------------------------------------------------------------
# KMeans   (kmeans.py)
# ---------------------------------------------------------------------------

class KMeans(object):
    """kmeans.py:24-176."""

    def __init__(self, X, K, assignments="rand"):
        N = X.shape[0]
        if isinstance(assignments, str) and assignments == "rand":          # :75-86
            assignments = np.random.randint(0, K, N)
        elif isinstance(assignments, str) and assignments == "each-in-own":
            assignments = np.arange(N)
        elif isinstance(assignments, str) and assignments == "spread":
            assignment_list = (list(range(K)) * int(np.ceil(float(N) / K)))[:N]
            random.shuffle(assignment_list)
            assignments = np.array(assignment_list)
        assignments = _consecutive(assignments)
        self.components = KMeansComponents(X, assignments, K)

    def fit(self, n_iter, consider_unassigned=True):
        """kmeans.py:97-173: frozen-means hard assignment, then apply."""
        c = self.components
        record = {"sum_neg_sqrd_norm": [], "components": [], "n_mean_updates": []}
        for _ in range(n_iter):
            updates = []
            for i in range(c.N):
                k_old = c.assignments[i]
                if not consider_unassigned and k_old == -1:
                    continue
                k = np.argmax(c.neg_sqrd_norm(i))
                if k != k_old:
                    updates.append((i, k))
            for i, k in updates:
                c.del_item(i)
                c.add_item(i, k)
            c.clean_components()
            record["sum_neg_sqrd_norm"].append(c.sum_neg_sqrd_norm())
            record["components"].append(c.K)
            record["n_mean_updates"].append(len(updates))
            if len(updates) == 0:
                break
        return record

    def get_n_assigned(self):
        return len(np.where(self.components.assignments != -1)[0])


# ---------------------------------------------------------------------------
# Utterances   (utterances.py)  and process_embeddings
# ---------------------------------------------------------------------------

def tri(t):
    """Packed-triangular offset of the block of segments ending at landmark t."""
    return t * (t - 1) // 2


class Utterances(object):
    """utterances.py:14-229 (only what the hot path touches)."""

    def __init__(self, lengths, vec_ids, durations, landmarks, seed_boundaries=None,
                 p_boundary_init=0.5, n_slices_min=0, n_slices_max=6, min_duration=0):
        assert lengths == [len(i) for i in landmarks]
        self.lengths = lengths
        self.D = len(lengths)
        assert self.D == len(vec_ids)
        self.N_max = max(lengths)
        self.landmarks = landmarks
        n_packed = self.N_max * (self.N_max + 1) // 2
        self.vec_ids = -1 * np.ones((self.D, n_packed), dtype=np.int64)
        for u, v in enumerate(vec_ids):
            self.vec_ids[u, :len(v)] = v
        self.durations = -np.nan * np.ones((self.D, n_packed))                   # :94
        for u, dv in enumerate(durations):
            if not (min_duration == 0 or len(dv) == 1):                         # :96-101
                cur = np.array(dv, dtype=np.float64)
                cur[cur < min_duration] = -np.nan
                if np.all(np.isnan(cur)):
                    cur[np.argmax(dv)] = np.max(dv)
                dv = cur
            self.durations[u, :len(dv)] = dv
        self.boundaries = np.zeros((self.D, self.N_max), dtype=bool)
        if seed_boundaries is not None:                                         # :106-115
            for u, bounds in enumerate(seed_boundaries):
                closest = [int(np.argmin([abs(b - lm) for lm in landmarks[u]])) for b in bounds]
                self.boundaries[u, closest] = True
        elif p_boundary_init == 0:                                              # :128-135
            for u in range(self.D):
                self.boundaries[u, self.lengths[u] - 1] = True
        else:                                                                   # :136-157
            for u in range(self.D):
                N = self.lengths[u]
                while True:
                    self.boundaries[u, 0:N] = (np.random.rand(N) < p_boundary_init)
                    self.boundaries[u, N - 1] = True
                    if np.all(np.asarray(self.get_segmented_embeds_i(u)) == -1):
                        continue
                    spans = [b - a for a, b in self.get_segmented_landmark_indices(u)]
                    if ((np.max(spans) <= n_slices_max and np.min(spans) >= n_slices_min)
                            or N <= n_slices_min):
                        break

    def _segments(self, u):
        j_prev = 0
        for j in range(self.lengths[u]):
            if self.boundaries[u, j]:
                yield j_prev, j + 1
                j_prev = j + 1

    def get_segmented_embeds_i(self, u):
        """:159-174."""
        return [self.vec_ids[u, tri(t) + j] for j, t in self._segments(u)]

    def get_segmented_durations_i(self, u):
        return [self.durations[u, tri(t) + j] for j, t in self._segments(u)]

    def get_segmented_landmark_indices(self, u):
        return list(self._segments(u))


def process_embeddings(embedding_mats, vec_ids_dict):
    """unigram_acoustic_wordseg.py:571-646: stack matrices in sorted-label order
    and rewrite per-utterance row indices into global embedding ids."""
    embeddings, vec_ids, labels = [], [], []
    n_seen = 0
    for utt in sorted(embedding_mats):
        labels.append(utt)
        src = vec_ids_dict[utt]
        cur = src.copy()
        for i_row, row in enumerate(embedding_mats[utt]):
            embeddings.append(row)
            cur[np.where(src == i_row)[0]] = n_seen
            n_seen += 1
        vec_ids.append(cur)
    return np.asarray(embeddings), vec_ids, labels


# ---------------------------------------------------------------------------
# Segmentation DPs -- NumPy restatements (small cases) + C twins (large cases)
# ---------------------------------------------------------------------------

def _win(a, S, cut=None):
    # a[-S:cut]; python's a[-0:] is the whole array, as in the reference
    return a[-S:cut]


def dp_packed_py(vec, N, n_slices_min, n_slices_max, mode, uniform=None, anneal_temp=1,
                 log_p_continue=0.0):
    """Pure NumPy statement of the three DPs (mode 0 FFBS, 1 Viterbi/GMM,
    2 Viterbi/k-means); follows unigram_acoustic_wordseg.py:653-864 and
    kmeans_acoustic_wordseg.py:449-555.  Returns (log_prob, bounds, alphas)."""
    S = n_slices_max
    cut = -(n_slices_min - 1) if n_slices_min > 1 else None
    bounds = np.zeros(N, dtype=bool)
    bounds[-1] = True
    al = np.ones(N)
    al[0] = 0.0
    i = 0
    for t in range(1, N):
        if np.all(_win(vec[i:i + t], S) + _win(al[:t], S) == -np.inf):
            al[t] = -np.inf
        else:
            c = _win(vec[i:i + t], S, cut) + _win(al[:t], S, cut)
            al[t] = (c_logsumexp(c) + log_p_continue) if mode == 0 else np.max(c)
        i += t
    t = N
    total = np.float64(0.)
    while True:
        i = tri(t)
        c = _win(vec[i:i + t], S, cut) + _win(al[:t], S, cut)
        if mode != 1:
            assert not np.isnan(np.sum(c))
        if np.all(c == -np.inf):
            while np.all(c == -np.inf):
                t = t - 1
                if t == 0:
                    raise FloatingPointError("utterance has no feasible segmentation")
                i = tri(t)
                c = _win(vec[i:i + t], S) + _win(al[:t], S)
            bounds[t - 1] = True
        if mode == 2:
            k = int(np.argmax(c[::-1])) + 1
        else:
            if mode == 0 and anneal_temp != 1:
                lp = c[::-1] - c_logsumexp(c)
                la = 1. / anneal_temp * lp - c_logsumexp(1. / anneal_temp * lp)
                p = np.exp(la)
            else:
                p = np.exp(c[::-1] - c_logsumexp(c))
            k = (draw(p, uniform()) if mode == 0 else int(np.argmax(p))) + 1
        if cut is not None:
            k += n_slices_min - 1
        total += vec[i + t - k]
        if t - k - 1 < 0:
            break
        bounds[t - k - 1] = True
        t = t - k
    return total, bounds, al


def dp_packed_c(vec, N, n_slices_min, n_slices_max, mode, uniforms=None, anneal_temp=1.0,
                log_p_continue=0.0):
    """C twin of dp_packed_py.  `uniforms`: array consumed left to right.
    Returns (status, log_prob, bounds[bool N], alphas[N], n_uniforms_used)."""
    assert N <= 4096
    vec = np.ascontiguousarray(vec, dtype=np.float64)
    if uniforms is None:
        uniforms = np.zeros(1)
    uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
    al = np.empty(N)
    b = np.empty(N, dtype=np.uint8)
    used = ctypes.c_int(0)
    lp = ctypes.c_double(0.0)
    st = clib().orc_dp_packed(
        _dptr(vec), N, n_slices_min, n_slices_max, mode, float(log_p_continue),
        float(anneal_temp), _dptr(uniforms), ctypes.byref(used), _dptr(al),
        b.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), ctypes.byref(lp))
    return st, lp.value, b.astype(bool), al, used.value


def forward_backward(vec, log_p_continue, N, n_slices_min=0, n_slices_max=0, i_utt=None,
                     anneal_temp=1, uniform=None):
    """unigram_acoustic_wordseg.py:653-756."""
    uniform = uniform if uniform is not None else UniformSource()
    lp, b, _ = dp_packed_py(vec, N, n_slices_min, n_slices_max, 0, uniform, anneal_temp, log_p_continue)
    assert lp != -np.inf
    return lp, b


def forward_backward_viterbi(vec, log_p_continue, N, n_slices_min=0, n_slices_max=0, i_utt=None,
                             anneal_temp=None, uniform=None):
    """unigram_acoustic_wordseg.py:759-864."""
    lp, b, _ = dp_packed_py(vec, N, n_slices_min, n_slices_max, 1)
    return lp, b


def forward_backward_kmeans_viterbi(vec, N, n_slices_min=0, n_slices_max=0, i_utt=None):
    """kmeans_acoustic_wordseg.py:449-555."""
    lp, b, _ = dp_packed_py(vec, N, n_slices_min, n_slices_max, 2)
    return lp, b


# ---------------------------------------------------------------------------
# Segmenters
# ---------------------------------------------------------------------------

def _init_corpus(self, embedding_mats, vec_ids_dict, durations_dict, landmarks_dict,
                 seed_boundaries_dict, p_boundary_init, n_slices_min, n_slices_max, min_duration):
    embeddings, vec_ids, labels = process_embeddings(embedding_mats, vec_ids_dict)
    self.ids_to_utterance_labels = labels
    seeds = None if seed_boundaries_dict is None else [seed_boundaries_dict[i] for i in labels]
    self.utterances = Utterances(
        [len(landmarks_dict[i]) for i in labels], vec_ids,
        [durations_dict[i] for i in labels], [landmarks_dict[i] for i in labels],
        seed_boundaries=seeds, p_boundary_init=p_boundary_init, n_slices_min=n_slices_min,
        n_slices_max=n_slices_max, min_duration=min_duration)
    init_embeds = []
    for u in range(self.utterances.D):
        init_embeds.extend(self.utterances.get_segmented_embeds_i(u))
    init_embeds = np.array(init_embeds, dtype=np.int64)
    return embeddings, init_embeds[np.where(init_embeds != -1)]


class UnigramAcousticWordseg(object):
    """unigram_acoustic_wordseg.py:27-564 (hot path: gibbs_sample_i, gibbs_sample,
    get_vec_embed_log_probs).  Seed-assignment initialisation is host-only setup
    and is not restated."""

    def __init__(self, am_class, am_alpha, am_K, am_param_prior, embedding_mats, vec_ids_dict,
                 durations_dict, landmarks_dict, seed_boundaries_dict=None,
                 seed_assignments_dict=None, covariance_type="fixed", n_slices_min=0,
                 n_slices_max=20, min_duration=0, p_boundary_init=0.5, beta_sent_boundary=2.0,
                 lms=1., wip=0., fb_type="standard", init_am_assignments="rand",
                 time_power_term=1., uniform=None):
        assert seed_assignments_dict is None, "not restated in the oracle"
        self.n_slices_min, self.n_slices_max = n_slices_min, n_slices_max
        self.beta_sent_boundary = beta_sent_boundary
        self.wip, self.time_power_term = wip, time_power_term
        self.fb_type = fb_type
        assert fb_type in ("standard", "viterbi")
        self.uniform = uniform if uniform is not None else UniformSource()
        embeddings, init_embeds = _init_corpus(
            self, embedding_mats, vec_ids_dict, durations_dict, landmarks_dict,
            seed_boundaries_dict, p_boundary_init, n_slices_min, n_slices_max, min_duration)
        N = embeddings.shape[0]
        assignments = -1 * np.ones(N, dtype=np.int64)
        if init_am_assignments == "rand":                                       # :206-223
            a = np.random.randint(0, am_K, len(init_embeds))
            assignments[init_embeds] = _consecutive(a)
            self.acoustic_model = am_class(embeddings, am_param_prior, am_alpha, am_K, assignments,
                                           covariance_type=covariance_type, lms=lms)
        elif init_am_assignments == "one-by-one":                               # :225-236
            self.acoustic_model = am_class(embeddings, am_param_prior, am_alpha, am_K, assignments,
                                           covariance_type=covariance_type, lms=lms)
            for e in init_embeds:
                self.acoustic_model.gibbs_sample_inside_loop_i(e)
        else:
            assert False, "invalid value for `init_am_assignments`"
        self.acoustic_model.uniform = self.uniform

    def get_vec_embed_log_probs(self, vec_ids, durations):
        """:474-511."""
        out = -np.inf * np.ones(len(vec_ids))
        for i, e in enumerate(vec_ids):
            if e == -1:
                continue
            out[i] = self.acoustic_model.log_marg_i(e)
            if np.isnan(durations[i]):
                out[i] = -np.inf
            else:
                out[i] *= durations[i] ** self.time_power_term
        return out + self.wip

    def calc_p_continue(self):
        assert self.beta_sent_boundary == -1, "to check (reference :520-521)"
        return 1.0

    def gibbs_sample_i(self, u, anneal_temp=1, anneal_gibbs_am=False):
        """:252-360."""
        utts, am = self.utterances, self.acoustic_model
        for e in utts.get_segmented_embeds_i(u):
            if e == -1:
                continue
            am.components.del_item(e)
        N = utts.lengths[u]
        n_packed = (N ** 2 + N) // 2
        scores = self.get_vec_embed_log_probs(utts.vec_ids[u, :n_packed], utts.durations[u, :n_packed])
        log_p_continue = math.log(self.calc_p_continue())
        if self.fb_type == "standard":
            log_prob, utts.boundaries[u, :N] = forward_backward(
                scores, log_p_continue, N, self.n_slices_min, self.n_slices_max, u, anneal_temp,
                uniform=self.uniform)
        else:
            log_prob, utts.boundaries[u, :N] = forward_backward_viterbi(
                scores, log_p_continue, N, self.n_slices_min, self.n_slices_max, u, anneal_temp)
        for e in utts.get_segmented_embeds_i(u):
            if e == -1:
                continue
            if self.fb_type == "standard":
                am.gibbs_sample_inside_loop_i(e, anneal_temp if anneal_gibbs_am else 1)
            else:
                am.map_assign_i(e)
        return log_prob

    def gibbs_sample(self, n_iter, am_n_iter=0, anneal_temps=None, anneal_gibbs_am=False,
                     utt_orders=None):
        """:362-472.  `utt_orders` (list of permutations) pins random.shuffle;
        `anneal_temps` is the already-expanded schedule (host logic)."""
        record = {k: [] for k in ("sample_time", "log_marg", "log_marg*length", "log_prob_z",
                                  "log_prob_X_given_z", "anneal_temp", "components", "n_tokens")}
        for it in range(n_iter):
            t0 = time.time()
            if am_n_iter > 0:
                self.acoustic_model.gibbs_sample(am_n_iter, consider_unassigned=False)
            temp = 1 if anneal_temps is None else anneal_temps[it]
            if utt_orders is None:
                order = list(range(self.utterances.D))
                random.shuffle(order)
            else:
                order = list(utt_orders[it])
            log_prob = 0
            for u in order:
                log_prob += self.gibbs_sample_i(u, temp, anneal_gibbs_am)
            record["sample_time"].append(time.time() - t0)
            record["log_marg"].append(self.acoustic_model.log_marg())
            record["log_marg*length"].append(log_prob)
            record["log_prob_z"].append(self.acoustic_model.log_prob_z())
            record["log_prob_X_given_z"].append(self.acoustic_model.log_prob_X_given_z())
            record["anneal_temp"].append(temp)
            record["components"].append(self.acoustic_model.components.K)
            record["n_tokens"].append(self.acoustic_model.get_n_assigned())
        return record


class SegmentalKMeansWordseg(object):
    """kmeans_acoustic_wordseg.py:27-447."""

    def __init__(self, am_K, embedding_mats, vec_ids_dict, durations_dict, landmarks_dict,
                 seed_boundaries_dict=None, seed_assignments_dict=None, n_slices_min=0,
                 n_slices_max=20, min_duration=0, p_boundary_init=0.5,
                 init_am_assignments="rand", wip=0):
        assert seed_assignments_dict is None                                    # :148-149
        self.n_slices_min, self.n_slices_max, self.wip = n_slices_min, n_slices_max, wip
        embeddings, init_embeds = _init_corpus(
            self, embedding_mats, vec_ids_dict, durations_dict, landmarks_dict,
            seed_boundaries_dict, p_boundary_init, n_slices_min, n_slices_max, min_duration)
        N = embeddings.shape[0]
        assignments = -1 * np.ones(N, dtype=np.int64)
        if init_am_assignments == "rand":                                       # :181-196
            a = np.random.randint(0, am_K, len(init_embeds))
            assignments[init_embeds] = _consecutive(a)
        elif init_am_assignments == "spread":                                   # :198-207
            n = len(init_embeds)
            lst = (list(range(am_K)) * int(np.ceil(float(n) / am_K)))[:n]
            random.shuffle(lst)
            assignments[init_embeds] = np.array(lst)
        else:
            assert False, "invalid value for `init_am_assignments`"
        self.acoustic_model = KMeans(embeddings, am_K, assignments)

    def get_vec_embed_neg_len_sqrd_norms(self, vec_ids, durations):
        """:334-351."""
        out = -np.inf * np.ones(len(vec_ids))
        for i, e in enumerate(vec_ids):
            if e == -1:
                continue
            out[i] = self.acoustic_model.components.max_neg_sqrd_norm_i(e)
            if np.isnan(durations[i]):
                out[i] = -np.inf
            else:
                out[i] *= durations[i]
        return out + self.wip

    def segment_i(self, u):
        """:225-332."""
        utts, comps = self.utterances, self.acoustic_model.components
        old_embeds = utts.get_segmented_embeds_i(u)
        N = utts.lengths[u]
        n_packed = (N ** 2 + N) // 2
        scores = self.get_vec_embed_neg_len_sqrd_norms(utts.vec_ids[u, :n_packed],
                                                       utts.durations[u, :n_packed])
        total, utts.boundaries[u, :N] = forward_backward_kmeans_viterbi(
            scores, N, self.n_slices_min, self.n_slices_max, u)
        new_embeds = utts.get_segmented_embeds_i(u)
        new_k = comps.get_max_assignments(new_embeds)
        for e in old_embeds:
            if e == -1:
                continue
            comps.del_item(e)
        for e, k in zip(new_embeds, new_k):
            comps.add_item(e, k)
        comps.clean_components()
        return total

    def segment(self, n_iter, n_iter_inbetween_kmeans=0, utt_orders=None):
        """:353-425."""
        record = {k: [] for k in ("sum_neg_sqrd_norm", "sum_neg_len_sqrd_norm", "components",
                                  "sample_time", "n_tokens")}
        for it in range(n_iter):
            t0 = time.time()
            if utt_orders is None:
                order = list(range(self.utterances.D))
                random.shuffle(order)
            else:
                order = list(utt_orders[it])
            total = 0
            for u in order:
                total += self.segment_i(u)
            record["sample_time"].append(time.time() - t0)
            record["sum_neg_sqrd_norm"].append(self.acoustic_model.components.sum_neg_sqrd_norm())
            record["sum_neg_len_sqrd_norm"].append(total)
            record["components"].append(self.acoustic_model.components.K)
            record["n_tokens"].append(self.acoustic_model.get_n_assigned())
            if n_iter_inbetween_kmeans > 0:
                self.acoustic_model.fit(n_iter_inbetween_kmeans, consider_unassigned=False)
        return record


# ---------------------------------------------------------------------------
# Frozen-state batch sweep (new mode; SURVEY 8c "oracle for the frozen-state
# batch mode"): every utterance is scored and segmented against the SAME means
# with the reference's pure functions, then all updates are applied.
# ---------------------------------------------------------------------------

# ---------------------------------------------------------------------------
# Bigram LM + bigram cluster sampling   (bigram_lms.py, bigram_fbgmm.py, bigram_acoustic_wordseg.py)
# ---------------------------------------------------------------------------

class BigramSmoothLM(object):
    """bigram_lms.py:18-113: smoothed, interpolated maximum-likelihood bigram LM."""

    def __init__(self, intrp_lambda, a, b, K):
        self.intrp_lambda, self.a, self.b, self.K = intrp_lambda, a, b, K
        self.unigram_counts = np.zeros(K, np.int64)                 # :46-47
        self.bigram_counts = np.zeros((K, K), np.int64)

    def prob_i(self, i):                                            # :49-54
        return (self.unigram_counts[i] + float(self.a) / self.K) / (int(self.unigram_counts.sum()) + self.a)

    def prob_i_given_j(self, i, j):                                 # :56-62
        p = (self.bigram_counts[j, i] + float(self.b) / self.K) / (self.unigram_counts[j] + float(self.b))
        return self.intrp_lambda * self.prob_i(i) + (1 - self.intrp_lambda) * p

    def log_prob_vec_i(self):                                       # :64-69
        return (np.log(self.unigram_counts + float(self.a) / self.K)
                - np.log(int(self.unigram_counts.sum()) + self.a))

    def prob_vec_i(self):                                           # :71-76
        return (self.unigram_counts + float(self.a) / self.K) / (int(self.unigram_counts.sum()) + self.a)

    def prob_vec_given_j(self, j):                                  # :84-91
        return (self.intrp_lambda * self.prob_vec_i() + (1 - self.intrp_lambda) *
                (self.bigram_counts[j, :] + float(self.b) / self.K) / (self.unigram_counts[j] + float(self.b)))

    def log_prob_vec_given_j(self, j):                              # :78-82
        return np.log(self.prob_vec_given_j(j))

    def counts_from_data(self, data):                               # :93-96
        for utterance in data:
            self.counts_from_utterance(utterance)

    def counts_from_utterance(self, utterance):                     # :98-105
        j_prev = None
        for i_cur in utterance:
            self.unigram_counts[i_cur] += 1
            if j_prev is not None:
                self.bigram_counts[j_prev, i_cur] += 1
            j_prev = i_cur

    def remove_counts_from_utterance(self, utterance):              # :107-113
        j_prev = None
        for i_cur in utterance:
            self.unigram_counts[i_cur] -= 1
            if j_prev is not None:
                self.bigram_counts[j_prev, i_cur] -= 1
            j_prev = i_cur


class BigramFBGMM(object):
    """bigram_fbgmm.py:20-100 (fixed covariance; components tied to the LM counts)."""

    def __init__(self, X, prior, K, assignments, covariance_type="fixed", lms=1.0, lm=None):
        assert covariance_type == "fixed"
        self.prior, self.covariance_type, self.lms = prior, covariance_type, lms
        assignments = _consecutive(np.asarray(assignments, dtype=np.int64))      # :75-79
        self.components = FixedVarComponents(X, prior, assignments, K_max=K, lm=lm)

    def log_prob_X_given_z(self):
        return self.components.log_marg()

    def get_n_assigned(self):
        return len(np.where(self.components.assignments != -1)[0])


class BigramAcousticWordseg(object):
    """bigram_acoustic_wordseg.py:30-749 with fb_type="unigram": segmentation as in the unigram model
    (scores from log_marg_i_embed_unigram), component assignments sampled under the bigram LM
    (gibbs_sample_inside_loop_i_embed).  fb_type="bigram" is a stub in the reference (:756-789: `pass`)."""

    def __init__(self, am_K, am_param_prior, lm_params, embedding_mats, vec_ids_dict, durations_dict,
                 landmarks_dict, seed_boundaries_dict=None, seed_assignments_dict=None, covariance_type="fixed",
                 n_slices_min=0, n_slices_max=20, min_duration=0, p_boundary_init=0.5, beta_sent_boundary=2.0,
                 lms=1., wip=0., fb_type="bigram", init_am_assignments="rand", time_power_term=1., uniform=None):
        assert seed_assignments_dict is None, "not restated in the oracle"
        assert fb_type == "unigram", "the reference's bigram forward-backward is a stub"
        self.n_slices_min, self.n_slices_max = n_slices_min, n_slices_max
        self.beta_sent_boundary, self.wip, self.lms = beta_sent_boundary, wip, lms
        self.time_power_term, self.fb_type = time_power_term, fb_type
        self.uniform = uniform if uniform is not None else UniformSource()
        embeddings, init_embeds = _init_corpus(
            self, embedding_mats, vec_ids_dict, durations_dict, landmarks_dict,
            seed_boundaries_dict, p_boundary_init, n_slices_min, n_slices_max, min_duration)
        assert lm_params["type"] == "smooth"                                      # :184-189
        self.lm = BigramSmoothLM(lm_params["intrp_lambda"], lm_params["a"], lm_params["b"], am_K)
        assignments = -1 * np.ones(embeddings.shape[0], dtype=np.int64)
        assert init_am_assignments == "rand"                                      # :228-243
        assignments[init_embeds] = _consecutive(np.random.randint(0, am_K, len(init_embeds)))
        self.acoustic_model = BigramFBGMM(embeddings, am_param_prior, am_K, assignments,
                                          covariance_type=covariance_type, lms=lms, lm=self.lm)
        self.set_lm_counts()

    def set_lm_counts(self):                                                      # :265-267
        for u in range(self.utterances.D):
            self.lm.counts_from_utterance(self.get_unsup_transcript_i(u))

    def get_unsup_transcript_i(self, u):                                          # :743-747
        return list(self.acoustic_model.components.assignments[
            np.asarray(self.utterances.get_segmented_embeds_i(u), dtype=np.int64)])

    def log_prob_z(self):                                                         # :269-305
        lm_tmp = BigramSmoothLM(self.lm.intrp_lambda, self.lm.a, self.lm.b, self.lm.K)
        log_prob_z = 0.
        for u in range(self.utterances.D):
            j_prev = None
            for i_cur in self.get_unsup_transcript_i(u):
                if j_prev is not None:
                    log_prob_z += np.log(lm_tmp.prob_i_given_j(i_cur, j_prev))
                    lm_tmp.bigram_counts[j_prev, i_cur] += 1
                else:
                    log_prob_z += np.log(lm_tmp.prob_i(i_cur))
                lm_tmp.unigram_counts[i_cur] += 1
        return log_prob_z

    def log_marg(self):                                                           # :307-311
        return self.log_prob_z() + self.acoustic_model.log_prob_X_given_z()

    def log_marg_i_embed_unigram(self, e):                                        # :314-330
        c = self.acoustic_model.components
        log_prob_z = self.lms * self.lm.log_prob_vec_i()
        log_prob_z[:c.K] += c.log_post_pred(e)
        log_prob_z[c.K:] += c.log_prior(e)
        return c_logsumexp(log_prob_z)

    def get_vec_embed_log_probs(self, vec_ids, durations):                        # :696-714
        out = -np.inf * np.ones(len(vec_ids))
        for i, e in enumerate(vec_ids):
            if e == -1:
                continue
            out[i] = self.log_marg_i_embed_unigram(e)
            if np.isnan(durations[i]):
                out[i] = -np.inf
            else:
                out[i] *= durations[i] ** self.time_power_term
        return out + self.wip

    def gibbs_sample_inside_loop_i_embed(self, e, j_prev_assignment=None, anneal_temp=1):   # :333-384
        c = self.acoustic_model.components
        if j_prev_assignment is not None:
            log_prob_z = np.log(self.lm.prob_vec_given_j(j_prev_assignment))
        else:
            log_prob_z = self.lm.log_prob_vec_i()
        log_prob_z *= self.lms
        log_prob_z[:c.K] += c.log_post_pred(e)
        log_prob_z[c.K:] += c.log_prior(e)
        if anneal_temp != 1:
            log_prob_z = log_prob_z - c_logsumexp(log_prob_z)
            log_prob_z_anneal = 1. / anneal_temp * log_prob_z - c_logsumexp(1. / anneal_temp * log_prob_z)
            prob_z = np.exp(log_prob_z_anneal)
        else:
            prob_z = np.exp(log_prob_z - c_logsumexp(log_prob_z))
        assert not np.isnan(np.sum(prob_z))
        k = draw(prob_z, self.uniform())
        if k > c.K:
            k = c.K
        c.add_item(e, k)
        return k

    def gibbs_sample_i(self, u, anneal_temp=1, anneal_gibbs_am=False, assignments_only=False):   # :386-543
        utts, c = self.utterances, self.acoustic_model.components
        self.lm.remove_counts_from_utterance(self.get_unsup_transcript_i(u))
        for e in utts.get_segmented_embeds_i(u):
            if e == -1:
                continue
            c.del_item(e)
        log_prob = 0.
        if not assignments_only:
            N = utts.lengths[u]
            n_packed = (N ** 2 + N) // 2
            scores = self.get_vec_embed_log_probs(utts.vec_ids[u, :n_packed], utts.durations[u, :n_packed])
            assert self.beta_sent_boundary == -1, "to check (reference :728-729)"
            log_prob, utts.boundaries[u, :N] = forward_backward(
                scores, math.log(1.0), N, self.n_slices_min, self.n_slices_max, u, anneal_temp,
                uniform=self.uniform)
        j_prev = None
        for e in utts.get_segmented_embeds_i(u):
            if e == -1:
                continue
            j_prev = self.gibbs_sample_inside_loop_i_embed(e, j_prev, anneal_temp if anneal_gibbs_am else 1)
        self.lm.counts_from_utterance(self.get_unsup_transcript_i(u))
        return log_prob

    def gibbs_sample(self, n_iter, anneal_temps=None, anneal_gibbs_am=False, assignments_only=False,
                     utt_orders=None):
        """:545-694.  `utt_orders` pins random.shuffle; `anneal_temps` is the expanded schedule."""
        record = {k: [] for k in ("sample_time", "log_marg", "log_marg*length", "log_prob_z",
                                  "log_prob_X_given_z", "anneal_temp", "components", "n_tokens")}
        for it in range(n_iter):
            t0 = time.time()
            temp = 1 if anneal_temps is None else anneal_temps[it]
            if utt_orders is None:
                order = list(range(self.utterances.D))
                random.shuffle(order)
            else:
                order = list(utt_orders[it])
            log_prob = 0
            for u in order:
                log_prob += self.gibbs_sample_i(u, temp, anneal_gibbs_am, assignments_only)
            record["sample_time"].append(time.time() - t0)
            record["log_marg"].append(self.log_marg())
            record["log_marg*length"].append(log_prob)
            record["log_prob_z"].append(self.log_prob_z())
            record["log_prob_X_given_z"].append(self.acoustic_model.log_prob_X_given_z())
            record["anneal_temp"].append(temp)
            record["components"].append(self.acoustic_model.components.K)
            record["n_tokens"].append(self.acoustic_model.get_n_assigned())
        return record


def frozen_kmeans_phase1(seg, utt_indices=None):
    """Pure part of the frozen sweep (means untouched): per utterance
    get_vec_embed_neg_len_sqrd_norms (kmeans_acoustic_wordseg.py:334-351) ->
    forward_backward_kmeans_viterbi (:449-555) -> get_max_assignments
    (kmeans_components.py:256-261).  Writes the new boundaries into
    seg.utterances and returns (per-utterance objectives, old tokens, plan)."""
    utts, comps = seg.utterances, seg.acoustic_model.components
    if utt_indices is None:
        utt_indices = range(utts.D)
    plan, totals, old_tokens = [], [], []
    for u in utt_indices:
        N = utts.lengths[u]
        n_packed = (N ** 2 + N) // 2
        old_tokens.extend(e for e in utts.get_segmented_embeds_i(u) if e != -1)
        scores = seg.get_vec_embed_neg_len_sqrd_norms(utts.vec_ids[u, :n_packed],
                                                      utts.durations[u, :n_packed])
        obj, bounds = forward_backward_kmeans_viterbi(scores, N, seg.n_slices_min, seg.n_slices_max, u)
        totals.append(obj)
        utts.boundaries[u, :N] = bounds
        embeds = utts.get_segmented_embeds_i(u)
        plan.append((embeds, comps.get_max_assignments(embeds)))
    return totals, old_tokens, plan


def frozen_kmeans_sweep(seg, utt_indices=None):
    """One frozen-means sweep of a SegmentalKMeansWordseg oracle object.

    Phase 1: frozen_kmeans_phase1.  Phase 2: del_item for every old token,
    add_item(new token, k) in utterance order left to right (same calls as
    segment_i :312-319), one clean_components() at the end (:320).
    Returns (sum of per-utterance objectives, list of per-utterance (embeds, ks))."""
    comps = seg.acoustic_model.components
    totals, old_tokens, plan = frozen_kmeans_phase1(seg, utt_indices)
    total = 0.0
    for t in totals:
        total += t
    for e in old_tokens:
        comps.del_item(e)
    for embeds, ks in plan:
        for e, k in zip(embeds, ks):
            comps.add_item(e, k)
    comps.clean_components()
    return total, plan


# ---------------------------------------------------------------------------
# Frozen-state FBGMM sweep (new batch mode; SURVEY 8c "oracle for the frozen-state batch mode")
# ---------------------------------------------------------------------------

def frozen_fbgmm_phase1(seg, u_fb=None, u_assign=None, utt_indices=None):
    """Pure part of the frozen FBGMM sweep (model untouched): per utterance
    get_vec_embed_log_probs (unigram_acoustic_wordseg.py:474-511, i.e. FBGMM.log_marg_i of every
    candidate against the CURRENT model) -> forward_backward / forward_backward_viterbi (:653-864)
    -> for every chosen segment the component choice of gibbs_sample_inside_loop_i (fbgmm.py:422-458,
    fb_type "standard") or map_assign_i (:465-491, "viterbi") WITHOUT its add_item.
    Draws: utterance u's i-th back-sampled segment uses u_fb[pos_off[u] + i]; the token ending at
    landmark position p uses u_assign[p].  Writes the new boundaries into seg.utterances; returns
    (per-utterance log_prob, [(embedding id, raw slot index)] in token order)."""
    utts, am = seg.utterances, seg.acoustic_model
    if utt_indices is None:
        utt_indices = range(utts.D)
    pos_off = np.concatenate([[0], np.cumsum(utts.lengths)])
    log_probs, choices = [], []
    for u in utt_indices:
        N = utts.lengths[u]
        n_packed = (N ** 2 + N) // 2
        scores = seg.get_vec_embed_log_probs(utts.vec_ids[u, :n_packed], utts.durations[u, :n_packed])
        log_p_continue = math.log(seg.calc_p_continue())
        if seg.fb_type == "standard":
            src = UniformSource(np.asarray(u_fb[pos_off[u]:pos_off[u + 1]], dtype=np.float64))
            lp, bounds = forward_backward(scores, log_p_continue, N, seg.n_slices_min, seg.n_slices_max, u, 1,
                                          uniform=src)
        else:
            lp, bounds = forward_backward_viterbi(scores, log_p_continue, N, seg.n_slices_min, seg.n_slices_max, u)
        log_probs.append(lp)
        utts.boundaries[u, :N] = bounds
        for (s, e_), emb in zip(utts.get_segmented_landmark_indices(u), utts.get_segmented_embeds_i(u)):
            if emb == -1:
                continue
            if seg.fb_type == "standard":
                lpz = am._assign_scores(emb, True)
                prob_z = np.exp(lpz - sp_logsumexp(lpz))
                j = draw(prob_z, float(u_assign[pos_off[u] + e_ - 1]))
            else:
                lpz = am._assign_scores(emb, False)
                prob_z = np.exp(lpz - sp_logsumexp(lpz))
                j = int(np.argmax(prob_z))
            choices.append((int(emb), int(j)))
    return log_probs, choices


def frozen_clamp(choices, K_before, K_max):
    """add_item's / FBGMM's `if k > K: k = K`, a choice of slot K opening a component
    (fbgmm.py:459-460, gaussian_components_fixedvar.py:162-165), applied in token order."""
    K = int(K_before)
    out = []
    for e, j in choices:
        k = min(j, K)
        if k == K and K < K_max:
            K += 1
        out.append((e, k))
    return out, K


def frozen_fbgmm_sweep(seg, u_fb=None, u_assign=None):
    """One frozen-model sweep of a UnigramAcousticWordseg oracle object (fixed-variance FBGMM).
    Phase 1: frozen_fbgmm_phase1.  Phase 2: the choices go through add_item's clamp in token order, and
    the model is rebuilt from the new assignments the way FBGMM.setup_components builds it
    (fbgmm.py:96-137: labels made consecutive, then the GaussianComponentsFixedVar constructor's add_item
    loop, gaussian_components_fixedvar.py:111-120).  Returns (sum of log_probs in utterance order, choices)."""
    am = seg.acoustic_model
    c = am.components
    log_probs, choices = frozen_fbgmm_phase1(seg, u_fb, u_assign)
    total = 0.0
    for lp in log_probs:
        total += lp
    resolved, _ = frozen_clamp(choices, c.K, c.K_max)
    assignments = -1 * np.ones(c.N, dtype=np.int64)
    for e, k in resolved:
        assignments[e] = k
    if len(resolved):
        assignments = _consecutive(assignments)
    am.components = FixedVarComponents(c.X, am.prior, assignments, K_max=c.K_max)
    return total, choices
