"""
Bigram-based finite Bayesian Gaussian mixture model: mirror of the reference's `BigramFBGMM`
(segmentalist/bigram_fbgmm.py:20-100).  Fixed-variance components whose counts are tied to a
`BigramSmoothLM`; the unigram smoothing parameter `a` of the LM plays the role FBGMM's alpha plays in
the segment scores (bigram_acoustic_wordseg.py:314-330).
"""
import numpy as np

from .fbgmm import make_consecutive
from .gaussian_components_fixedvar import GaussianComponentsFixedVar


class BigramFBGMM(object):

    def __init__(self, X, prior, K, assignments="rand", covariance_type="fixed", lms=1.0, lm=None):
        self.prior = prior
        self.covariance_type = covariance_type
        self.lms = lms
        self.setup_components(K, assignments, X, lm)

    def setup_components(self, K, assignments="rand", X=None, lm=None):
        """:45-92."""
        if X is None:
            assert hasattr(self, "components")
            X = self.components.X
        N, D = X.shape
        if isinstance(assignments, str) and assignments == "rand":
            assignments = np.random.randint(0, K, N)
        elif isinstance(assignments, str) and assignments == "each-in-own":
            assignments = np.arange(N)
        assignments = make_consecutive(assignments)
        assert self.covariance_type == "fixed", "bigram sampling on the device: fixed-variance components"
        alpha = 1.0 if lm is None else float(lm.a)
        self.components = GaussianComponentsFixedVar(X, self.prior, assignments, K_max=K, lm=lm, alpha=alpha,
                                                     lms=self.lms)

    def log_prob_X_given_z(self):
        return self.components.log_marg()

    def get_n_assigned(self):
        # counted on the device: `components.assignments` would mirror the whole vector to the host first
        return int((self.components._assign != -1).sum().item())
