"""
Finite Bayesian GMM over device-resident fixed-variance components.

Mirror of the reference's `FBGMM` (segmentalist/fbgmm.py:27-494) for
covariance_type == "fixed" -- the only covariance on the accelerated path
(SURVEY.md section 8).  Uniform draws are taken from Python's `random.random()`
on the host, exactly where the reference takes them, and handed to the kernels,
so a seeded run consumes the same stream as the reference.
"""
import random

import numpy as np
import torch
from scipy.special import gammaln

from . import _lib
from .gaussian_components_fixedvar import GaussianComponentsFixedVar


def make_consecutive(assignments):
    """Relabel so that the used labels are 0..max without gaps (fbgmm.py:124-128).  The reference shifts every
    label above an unused one down, one unused label at a time (a full-array scan per label); the result is the
    order-preserving renumbering of the used labels, computed here in one pass.  In place, like the reference."""
    live = assignments >= 0
    if not live.any():
        return assignments
    used = np.bincount(assignments[live]) > 0
    if used.all():
        return assignments
    new_label = np.cumsum(used) - 1
    assignments[live] = new_label[assignments[live]]
    return assignments


class FBGMM(object):

    def __init__(self, X, prior, alpha, K, assignments="rand", covariance_type="fixed", lms=1.0):
        self.alpha = alpha
        self.prior = prior
        self.covariance_type = covariance_type
        self.lms = lms
        self.setup_components(K, assignments, X)

    def setup_components(self, K, assignments="rand", X=None):
        """fbgmm.py:96-137."""
        if X is None:
            assert hasattr(self, "components")
            X = self.components.X
        N, D = X.shape
        if isinstance(assignments, str) and assignments == "rand":
            assignments = np.random.randint(0, K, N)
        elif isinstance(assignments, str) and assignments == "each-in-own":
            assignments = np.arange(N)
        assignments = make_consecutive(np.asarray(assignments))
        if self.covariance_type == "diag":                                   # fbgmm.py:130-137
            from .gaussian_components_diag import GaussianComponentsDiag
            self.components = GaussianComponentsDiag(X, self.prior, assignments, K_max=K,
                                                     alpha=self.alpha, lms=self.lms)
            return
        assert self.covariance_type == "fixed", (
            "covariance_type 'fixed' and 'diag' are implemented on the B200 path (full covariance: DESIGN.md 7)")
        self.components = GaussianComponentsFixedVar(X, self.prior, assignments, K_max=K,
                                                     alpha=self.alpha, lms=self.lms)

    # ---- hot path
    def log_marg_i(self, i):
        """fbgmm.py:256-285: log p(x_i) marginalised over all K_max slots."""
        assert i != -1
        return float(self.log_marg_items(np.asarray([i]))[0])

    def log_marg_items(self, ids, durations=None, time_power_term=1.0, wip=0.0):
        """Batched log_marg_i (one launch); with `durations` also applies the scaling of
        get_vec_embed_log_probs (unigram_acoustic_wordseg.py:474-511)."""
        c = self.components
        ids_d = _lib.dev(np.asarray(ids, dtype=np.int32))
        durs_d = None if durations is None else _lib.dev(np.asarray(durations, dtype=np.float64))
        out = torch.empty(len(ids), dtype=torch.float64, device="cuda")
        _lib.check(_lib.lib().segb_fixedvar_log_marg(
            c.struct(), _lib.ptr(ids_d), _lib.ptr(durs_d), len(ids), float(time_power_term), float(wip),
            _lib.ptr(out), _lib.stream_ptr()))
        return out.cpu().numpy()

    def log_marg_all(self, tensor_cores=True, method="filter"):
        """log_marg_i of EVERY embedding against the current (frozen) model in one pass.
        tensor_cores=True, method="filter" (default): ONE fp16 tcgen05 pass that keeps, per embedding, the
        components within 25 nats (+ a rigorous rounding bound) of the best one, exact float64 re-scoring of
        those and the logsumexp over the exact scores (segb_fvf_*; isotropic or anisotropic variances,
        float32 embeddings; ~1e-7 absolute).  method="split3": the FP32-accurate three-pass GEMM with the
        logsumexp fused into its epilogue (segb_fvmma_log_marg; isotropic variances only; ~1e-6 relative).
        tensor_cores=False: the exact float64 kernel."""
        c = self.components
        if not tensor_cores:
            return self.log_marg_items(np.arange(c.N))
        assert c._X.dtype == torch.float32, "tensor-core log_marg needs float32 embeddings"
        if method == "filter":
            from .batch import FvScorer
            if getattr(self, "_fv", None) is None or self._fv.c is not c:
                self._fv = FvScorer(c)
            self._fv.score()
            return self._fv.log_marg.cpu().numpy()
        assert method == "split3"
        iso = (np.all(c.precision == c.precision[0]) and np.all(c.precision_0 == c.precision_0[0]))
        assert iso, "the three-pass tensor-core log_marg needs isotropic variances"
        lib = _lib.lib()
        if getattr(self, "_tc", None) is None:
            x_tiles = torch.empty(lib.segb_fvmma_x_tiles_bytes(c.N, c.D), dtype=torch.uint8, device="cuda")
            w_tiles = torch.empty(lib.segb_fvmma_w_tiles_bytes(c.K_max, c.D), dtype=torch.uint8, device="cuda")
            _lib.check(lib.segb_fvmma_pack_x(_lib.ptr(c._X), c.N, c.D, _lib.ptr(x_tiles), _lib.stream_ptr()))
            self._tc = (x_tiles, w_tiles, torch.empty(c.N, dtype=torch.float32, device="cuda"))
        x_tiles, w_tiles, out = self._tc
        _lib.check(lib.segb_fvmma_log_marg(c.struct(), _lib.ptr(x_tiles), _lib.ptr(w_tiles), c.N, _lib.ptr(out),
                                           _lib.stream_ptr()))
        return out.cpu().numpy().astype(np.float64)

    def gibbs_sample(self, n_iter, consider_unassigned=True, anneal_schedule=None, anneal_start_temp_inv=0.1,
                     anneal_end_temp_inv=1, n_anneal_steps=-1):
        """fbgmm.py:288-420: `n_iter` sweeps of collapsed Gibbs sampling over the data items, each
        sweep one cooperative launch (segb_fbgmm_gibbs_items_coop); `random.random()` is consumed
        exactly as the reference consumes it (one draw per considered item, in item order)."""
        import time
        from .unigram_acoustic_wordseg import UniformFeed, _anneal_iter
        c, lib = self.components, _lib.lib()
        record_dict = {k: [] for k in ("sample_time", "log_marg", "log_prob_z", "log_prob_X_given_z",
                                       "anneal_temp", "components")}
        get_anneal_temp = _anneal_iter(n_iter, anneal_schedule, anneal_start_temp_inv, anneal_end_temp_inv,
                                       n_anneal_steps)
        work = torch.empty(lib.segb_gibbs_work_bytes(c.K_max, 1, 1), dtype=torch.uint8, device="cuda")
        start_time = time.time()
        for _ in range(n_iter):
            anneal_temp = next(get_anneal_temp, anneal_end_temp_inv)
            assign = c._assign
            items = (torch.arange(c.N, device="cuda", dtype=torch.int32) if consider_unassigned
                     else (assign != -1).nonzero().flatten().to(torch.int32))
            n = int(items.numel())
            feed = UniformFeed(n)
            _lib.check(lib.segb_fbgmm_gibbs_items_coop(c.struct(), _lib.ptr(items), n, float(anneal_temp),
                                                       _lib.ptr(feed.dev), _lib.ptr(feed.counter), _lib.ptr(work),
                                                       _lib.stream_ptr()))
            used = feed.finish()
            assert used == n
            record_dict["sample_time"].append(time.time() - start_time)
            start_time = time.time()
            record_dict["log_marg"].append(self.log_marg())
            record_dict["log_prob_z"].append(self.log_prob_z())
            record_dict["log_prob_X_given_z"].append(self.log_prob_X_given_z())
            record_dict["anneal_temp"].append(anneal_temp)
            record_dict["components"].append(c.K)
        return record_dict

    def _assign(self, ids, mode, anneal_temp, uniforms):
        c = self.components
        ids_d = _lib.dev(np.asarray(ids, dtype=np.int32))
        ks = torch.empty(len(ids), dtype=torch.int32, device="cuda")
        u_d = None if uniforms is None else _lib.dev(np.asarray(uniforms, dtype=np.float64))
        cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
        _lib.check(_lib.lib().segb_fixedvar_assign_items(
            c.struct(), _lib.ptr(ids_d), len(ids), mode, float(anneal_temp), _lib.ptr(u_d), _lib.ptr(cnt),
            _lib.ptr(ks), _lib.stream_ptr()))
        return ks.cpu().numpy()

    def gibbs_sample_inside_loop_i(self, i, anneal_temp=1):
        """fbgmm.py:422-463: draw a component for X[i] and add it."""
        u = random.random()
        return int(self._assign([i], 0, anneal_temp, [u])[0])

    def map_assign_i(self, i):
        """fbgmm.py:465-494."""
        return int(self._assign([i], 1, 1.0, None)[0])

    # ---- diagnostics (host, from mirrored statistics)
    def log_prob_z(self):
        """fbgmm.py:208-225."""
        counts = self.components.counts
        K_max = self.components.K_max
        return (gammaln(self.alpha) - gammaln(self.alpha + np.sum(counts))
                + np.sum(gammaln(counts + float(self.alpha) / K_max) - gammaln(self.alpha / K_max)))

    def log_prob_X_given_z(self):
        return self.components.log_marg()

    def log_marg(self):
        return self.log_prob_z() + self.log_prob_X_given_z()

    def get_n_assigned(self):
        # counted on the device: `components.assignments` would mirror the whole vector to the host first
        return int((self.components._assign != -1).sum().item())
