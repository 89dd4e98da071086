"""
Diagonal-covariance Gaussian components on the device.

Same constructor, attributes and methods as the reference's `GaussianComponentsDiag`
(segmentalist/gaussian_components_diag.py:23-345): D independent normal-inverse-chi-squared
posteriors per component, posterior predictive = product of univariate Student's t densities.
The statistics live in HBM in the layout of `segb_fixedvar` with `model = SEGB_MODEL_DIAG`
(include/segb200.h) and go through the same kernels as the fixed-variance components: batched
log_marg_i, the cooperative Gibbs sweep and the whole-model resampling sweep.
"""
import math

import numpy as np
import torch
from scipy.special import gammaln

from . import _lib
from .gaussian_components_fixedvar import GaussianComponentsFixedVar


class GaussianComponentsDiag(GaussianComponentsFixedVar):

    def __init__(self, X, prior, assignments=None, K_max=None, alpha=1.0, lms=1.0):
        X = np.ascontiguousarray(X)
        if X.dtype not in (np.float32, np.float64):
            X = X.astype(np.float64)
        self.X = X
        self.prior = prior
        self.N, self.D = X.shape
        if K_max is None:                                                # :90-91
            K_max = self.N
        self.K_max = int(K_max)
        assert len(np.shape(prior.S_0)) == 1, "For diagonal covariance, S_0 needs to be vector."   # :94
        assert int(prior.v_0) == prior.v_0, "v_0 indexes the reference's gammaln table: integer"
        self.lm = None
        dv, z = _lib.dev, lambda *s: torch.zeros(*s, dtype=torch.float64, device="cuda")
        _lib.lib()
        self._X = dv(X)
        # same tables as the fixed-variance model, reinterpreted (include/segb200.h)
        self._mu_N_numT, self._prec_NT = z(self.D, self.K_max), z(self.D, self.K_max)     # m_N_numerators, S_N_partials
        self._prec_predT, self._mu_NT = z(self.D, self.K_max), z(self.D, self.K_max)       # inv_vars, m_N
        self._log_prod = z(self.K_max)                                                     # log_prod_vars
        self._counts = torch.zeros(self.K_max, dtype=torch.int32, device="cuda")
        self._assign = torch.full((self.N,), -1, dtype=torch.int32, device="cuda")
        self._K = torch.zeros(1, dtype=torch.int32, device="cuda")
        self._n_total = torch.zeros(1, dtype=torch.int64, device="cuda")
        self.m_0 = np.asarray(prior.m_0, dtype=np.float64) * np.ones(self.D)
        self.S_0 = np.asarray(prior.S_0, dtype=np.float64) * np.ones(self.D)
        self._mu_0, self._precision_0 = dv(self.m_0), dv(self.S_0)
        self._precision = self._precision_0              # unused by the diagonal model
        self._alpha, self._lms = float(alpha), float(lms)
        self._relabel = None
        self._scratch_row = z(self.K_max)
        self._cached_log_pi = math.log(np.pi)
        if assignments is not None:
            assignments = np.asarray(assignments, dtype=np.int64)        # :108-117
            assert (self.N,) == assignments.shape
            assert set(assignments).difference([-1]) == set(range(assignments.max() + 1))
            if assignments.max() >= 0:
                a_dev = _lib.dev(assignments.astype(np.int32))
                order, seg_off = _lib.members_by_component(a_dev, self.K_max)
                _lib.check(_lib.lib().segb_fixedvar_build(self.struct(), _lib.ptr(order), _lib.ptr(seg_off),
                                                          int(assignments.max()) + 1, _lib.stream_ptr()))

    def struct(self):
        m = super(GaussianComponentsDiag, self).struct_base()
        m.model, m.v_0, m.k_0 = 1, int(self.prior.v_0), float(self.prior.k_0)
        return m

    # ---- mirrors under the reference's names
    @property
    def m_N_numerators(self):
        return self._mu_N_numT.t().contiguous().cpu().numpy()

    @property
    def S_N_partials(self):
        return self._prec_NT.t().contiguous().cpu().numpy()

    @property
    def inv_vars(self):
        return self._prec_predT.t().contiguous().cpu().numpy()

    @property
    def log_prod_vars(self):
        return self._log_prod.cpu().numpy()

    def log_prior(self, i):
        """:216-223."""
        row = self._pred_row(i)
        if self.K < self.K_max:
            return float(row[self.K_max - 1])
        pr = self.prior
        var = (pr.k_0 + 1.) / (pr.k_0 * pr.v_0) * self.S_0
        delta = self.X[i, :].astype(np.float64) - self.m_0
        v = pr.v_0
        return float(self.D * (gammaln((v + 1) / 2.) - gammaln(v / 2.) - 0.5 * math.log(v) - 0.5 * self._cached_log_pi)
                     - 0.5 * np.log(var).sum() - (v + 1.) / 2. * np.log(1. + 1. / v * np.square(delta) / var).sum())

    # ---- diagnostics: closed form over the device-resident sufficient statistics (segb_diag_log_marg_k)
    def _log_marg_all_k(self):
        out = torch.empty(self.K_max, dtype=torch.float64, device="cuda")
        _lib.check(_lib.lib().segb_diag_log_marg_k(self.struct(), _lib.ptr(out), _lib.stream_ptr()))
        return out.cpu().numpy()

    def log_marg_k(self, k):
        """:270-288."""
        return float(self._log_marg_all_k()[k])

    def log_marg(self):
        """:290-301: the per-component values are added in component order like the reference's loop."""
        K = self.K
        if K == 0:
            return 0.
        return float(np.cumsum(self._log_marg_all_k()[:K])[-1])
