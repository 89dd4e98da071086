"""
Fixed-variance Bayesian Gaussian components on the device.

Same constructor, attributes and methods as the reference's
`GaussianComponentsFixedVar` / `FixedVarPrior`
(segmentalist/gaussian_components_fixedvar.py:21-356); the sufficient
statistics live in HBM (see include/segb200.h, segb_fixedvar) and every method
is a call into libsegb200.so.  NumPy-facing attributes (`mu_N_numerators`,
`precision_Ns`, `precision_preds`, `log_prod_precision_preds`, `counts`,
`assignments`, `K`) are read-only mirrors downloaded on access.
"""
import math

import numpy as np
import torch

from . import _lib


class FixedVarPrior(object):
    """gaussian_components_fixedvar.py:349-356."""

    def __init__(self, var, mu_0, var_0):
        self.var = var
        self.mu_0 = mu_0
        self.var_0 = var_0


def _sum_log_sequential(v):
    """_cython_utils.sum_log (segmentalist/_cython_utils.pyx:52-59): left-to-right."""
    s = math.log(v[0])
    for x in v[1:]:
        s += math.log(x)
    return s


class GaussianComponentsFixedVar(object):

    def __init__(self, X, prior, assignments=None, K_max=None, lm=None, alpha=1.0, lms=1.0):
        assert K_max is not None, "To-do: remove this, always require `K_max`"   # :88-89
        X = np.ascontiguousarray(X)
        if X.dtype not in (np.float32, np.float64):
            X = X.astype(np.float64)
        self.X = X
        self.N, self.D = X.shape
        self.K_max = int(K_max)
        self.precision = np.asarray(1. / prior.var, dtype=np.float64) * np.ones(self.D)
        self.mu_0 = np.asarray(prior.mu_0, dtype=np.float64) * np.ones(self.D)
        self.precision_0 = np.asarray(1. / prior.var_0, dtype=np.float64) * np.ones(self.D)
        self._cached_neg_half_D_log_2pi = -0.5 * self.D * math.log(2. * np.pi)
        self.lm = lm                  # tied bigram LM (bigram_lms.BigramSmoothLM): its counts follow del_component, :205-221
        assert lm is None or lm.K == self.K_max

        dv, z = _lib.dev, lambda *s: torch.zeros(*s, dtype=torch.float64, device="cuda")
        _lib.lib()
        self._X = dv(X)
        self._mu_N_numT, self._prec_NT = z(self.D, self.K_max), z(self.D, self.K_max)
        self._prec_predT, self._mu_NT = z(self.D, self.K_max), z(self.D, self.K_max)
        self._log_prod = z(self.K_max)
        self._counts = torch.zeros(self.K_max, dtype=torch.int32, device="cuda")
        self._assign = torch.full((self.N,), -1, dtype=torch.int32, device="cuda")
        self._K = torch.zeros(1, dtype=torch.int32, device="cuda")
        self._n_total = torch.zeros(1, dtype=torch.int64, device="cuda")
        self._precision, self._mu_0, self._precision_0 = dv(self.precision), dv(self.mu_0), dv(self.precision_0)
        self._alpha, self._lms = float(alpha), float(lms)
        self._relabel = None          # optional (tensor) live-token table used instead of a full scan
        self._scratch_row = z(self.K_max)

        if assignments is not None:
            assignments = np.asarray(assignments, dtype=np.int64)                # :111-120
            assert (self.N,) == assignments.shape
            assert set(assignments).difference([-1]) == set(range(assignments.max() + 1))
            if assignments.max() >= 0:
                # all components at once, each replaying its members' add_item calls in index order
                a_dev = _lib.dev(assignments.astype(np.int32))
                order, seg_off = _lib.members_by_component(a_dev, self.K_max)
                _lib.check(_lib.lib().segb_fixedvar_build(self.struct(), _lib.ptr(order), _lib.ptr(seg_off),
                                                          int(assignments.max()) + 1, _lib.stream_ptr()))

    @classmethod
    def from_device(cls, X_dev, prior, K_max, alpha=1.0, lms=1.0):
        """Bench-scale constructor: float32 embeddings already resident in HBM, nothing assigned."""
        self = cls.__new__(cls)
        assert X_dev.is_cuda and X_dev.dtype == torch.float32 and X_dev.is_contiguous()
        _lib.lib()
        self.X = None
        self.N, self.D = int(X_dev.shape[0]), int(X_dev.shape[1])
        self.K_max = int(K_max)
        self.precision = np.asarray(1. / prior.var, dtype=np.float64) * np.ones(self.D)
        self.mu_0 = np.asarray(prior.mu_0, dtype=np.float64) * np.ones(self.D)
        self.precision_0 = np.asarray(1. / prior.var_0, dtype=np.float64) * np.ones(self.D)
        self._cached_neg_half_D_log_2pi = -0.5 * self.D * math.log(2. * np.pi)
        self.lm = None
        dv, z = _lib.dev, lambda *s: torch.zeros(*s, dtype=torch.float64, device="cuda")
        self._X = X_dev
        self._mu_N_numT, self._prec_NT = z(self.D, self.K_max), z(self.D, self.K_max)
        self._prec_predT, self._mu_NT = z(self.D, self.K_max), z(self.D, self.K_max)
        self._log_prod = z(self.K_max)
        self._counts = torch.zeros(self.K_max, dtype=torch.int32, device="cuda")
        self._assign = torch.full((self.N,), -1, dtype=torch.int32, device="cuda")
        self._K = torch.zeros(1, dtype=torch.int32, device="cuda")
        self._n_total = torch.zeros(1, dtype=torch.int64, device="cuda")
        self._precision, self._mu_0, self._precision_0 = dv(self.precision), dv(self.mu_0), dv(self.precision_0)
        self._alpha, self._lms = float(alpha), float(lms)
        self._relabel = None
        self._scratch_row = z(self.K_max)
        return self

    # ---- C-ABI plumbing
    def struct(self):
        m = self.struct_base()
        m.sum_log_precision_0 = _sum_log_sequential(self.precision_0)
        return m

    def struct_base(self):
        m = _lib.FixedVar()
        m.D, m.K_max, m.x_is_f64, m.n_emb = self.D, self.K_max, int(self._X.dtype == torch.float64), self.N
        m.X = self._X.data_ptr()
        m.mu_N_numT, m.prec_NT = self._mu_N_numT.data_ptr(), self._prec_NT.data_ptr()
        m.prec_predT, m.mu_NT = self._prec_predT.data_ptr(), self._mu_NT.data_ptr()
        m.log_prod_prec_pred = self._log_prod.data_ptr()
        m.counts, m.assignments = self._counts.data_ptr(), self._assign.data_ptr()
        m.K, m.n_total = self._K.data_ptr(), self._n_total.data_ptr()
        m.precision, m.mu_0, m.precision_0 = (self._precision.data_ptr(), self._mu_0.data_ptr(),
                                              self._precision_0.data_ptr())
        m.alpha, m.lms = self._alpha, self._lms
        return m

    def _add_many(self, ids, ks):
        if len(ids) == 0:
            return
        m = self.struct()
        ids_d, ks_d = _lib.dev(np.asarray(ids, dtype=np.int32)), _lib.dev(np.asarray(ks, dtype=np.int32))
        _lib.check(_lib.lib().segb_fixedvar_add_items(m, _lib.ptr(ids_d), _lib.ptr(ks_d), len(ids), _lib.stream_ptr()))

    # ---- mirrors
    @property
    def K(self):
        return int(self._K.item())

    @property
    def counts(self):
        return self._counts.cpu().numpy().astype(np.int64)

    @property
    def assignments(self):
        return self._assign.cpu().numpy().astype(np.int64)

    @property
    def mu_N_numerators(self):
        return self._mu_N_numT.t().contiguous().cpu().numpy()

    @property
    def precision_Ns(self):
        return self._prec_NT.t().contiguous().cpu().numpy()

    @property
    def precision_preds(self):
        return self._prec_predT.t().contiguous().cpu().numpy()

    @property
    def log_prod_precision_preds(self):
        return self._log_prod.cpu().numpy()

    # ---- reference API
    def add_item(self, i, k):
        """:153-170."""
        assert not i == -1
        assert 0 <= k <= self.K, "add_item: component index beyond K"
        self._add_many([i], [k])

    def del_item(self, i):
        """:172-188 (deletes the component when it empties, :190-221)."""
        assert not i == -1
        m = self.struct()
        ids_d = _lib.dev(np.asarray([i], dtype=np.int32))
        rel = self._relabel
        if self.lm is not None:
            _lib.check(_lib.lib().segb_fixedvar_del_items_lm(
                m, self.lm.struct(), _lib.ptr(ids_d), 1, _lib.ptr(rel), 0 if rel is None else rel.numel(),
                _lib.stream_ptr()))
            return
        _lib.check(_lib.lib().segb_fixedvar_del_items(
            m, _lib.ptr(ids_d), 1, _lib.ptr(rel), 0 if rel is None else rel.numel(), _lib.stream_ptr()))

    def _pred_row(self, i):
        m = self.struct()
        _lib.check(_lib.lib().segb_fixedvar_log_pred_row(m, int(i), _lib.ptr(self._scratch_row), _lib.stream_ptr()))
        return self._scratch_row.cpu().numpy()

    def log_post_pred(self, i):
        """:242-253 -- K-vector of posterior predictive log-probabilities of X[i]."""
        return self._pred_row(i)[:self.K].copy()

    def log_post_pred_k(self, i, k):
        """:233-239."""
        assert 0 <= k < self.K
        return float(self._pred_row(i)[k])

    def log_prior(self, i):
        """:224-231 -- probability of X[i] under the prior alone (precision_0, not 1/(var_0+var))."""
        row = self._pred_row(i)
        if self.K < self.K_max:
            return float(row[self.K_max - 1])
        # no empty slot on the device row: closed form, sequential order as in _cython_utils
        delta = self.X[i, :] - self.mu_0
        sq = 0.0
        for a, b in zip(delta, self.precision_0):
            sq += a * a * b
        return self._cached_neg_half_D_log_2pi + 0.5 * _sum_log_sequential(self.precision_0) - 0.5 * sq

    def get_assignments(self, list_of_i):
        return self.assignments[np.asarray(list_of_i)]

    def _log_marg_per_component(self):
        """log_marg_k for every component on the device (csrc/diagnostics.cu): the members' column sums in X's
        dtype and NumPy's order, then the closed form of :270-283 in float64."""
        work = torch.empty(_lib.lib().segb_fixedvar_log_marg_k_work_bytes(self.K_max, self.D) // 8,
                           dtype=torch.float64, device="cuda")
        out = torch.empty(self.K_max, dtype=torch.float64, device="cuda")
        order, seg_off = _lib.members_by_component(self._assign, self.K_max)
        _lib.check(_lib.lib().segb_fixedvar_log_marg_k(self.struct(), _lib.ptr(order), _lib.ptr(seg_off), _lib.ptr(work),
                                                       _lib.ptr(out), _lib.stream_ptr()))
        return out.cpu().numpy()

    def log_marg_k(self, k, _cache=None):
        """:261-283 -- diagnostic, evaluated on the device."""
        per_k = self._log_marg_per_component() if _cache is None else _cache
        return float(per_k[k])

    def log_marg(self):
        """:285-296: the components' log marginals added in component order."""
        per_k = self._log_marg_per_component()
        total = 0.
        for k in range(self.K):
            total += per_k[k]
        return total


def log_norm_pdf(x, mean, var):
    """gaussian_components_fixedvar.py:363-365."""
    return -0.5 * (np.log(2 * np.pi) + np.log(var)) - 1. / (2 * var) * (x - mean) ** 2


def log_post_pred_unvectorized(gmm, i):
    """gaussian_components_fixedvar.py:368-376."""
    return np.array([gmm.log_post_pred_k(i, k) for k in range(gmm.K)])
