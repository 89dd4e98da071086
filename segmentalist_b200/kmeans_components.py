"""
K-means components on the device.

Mirror of the reference's `KMeansComponents`
(segmentalist/kmeans_components.py:18-266).  `means` keeps X's dtype (float32
in practice) and distances are accumulated in that dtype in NumPy's pairwise
order, so `neg_sqrd_norm`, `max_` and `argmax_neg_sqrd_norm_i` return the very
bits the reference returns.
"""
import numpy as np
import torch

from . import _lib


class KMeansComponents(object):

    def __init__(self, X, assignments, K_max):
        X = np.ascontiguousarray(X)
        if X.dtype not in (np.float32, np.float64):
            X = X.astype(np.float64)
        self.X = X
        self.N, self.D = X.shape
        self.K_max = int(K_max)
        assignments = np.asarray(assignments, dtype=np.int64)
        assert (self.N,) == assignments.shape
        n_lab = int(assignments.max()) + 1
        assert np.all(np.bincount(assignments[assignments >= 0], minlength=n_lab) > 0)   # labels 0..max, no gaps (:69-70)
        self.setup_random_means()                                           # :75-76
        _lib.lib()
        tdt = torch.float64 if X.dtype == np.float64 else torch.float32
        self._X = _lib.dev(X)
        self._mean_num = torch.zeros(self.K_max, self.D, dtype=torch.float64, device="cuda")
        self._rnd = _lib.dev(self.random_means)
        self._means = self._rnd.clone()
        self._meansT = self._rnd.t().contiguous()
        self._counts = torch.zeros(self.K_max, dtype=torch.int32, device="cuda")
        self._assign = torch.full((self.N,), -1, dtype=torch.int32, device="cuda")
        self._K = torch.zeros(1, dtype=torch.int32, device="cuda")
        self._row = torch.empty(self.K_max, dtype=tdt, device="cuda")
        self._relabel = None
        if n_lab > 0:                                                       # :79-81, all components in parallel
            a_dev = _lib.dev(assignments.astype(np.int32))
            order, seg_off = _lib.members_by_component(a_dev, self.K_max)
            _lib.check(_lib.lib().segb_kmeans_build(self.struct(), _lib.ptr(order), _lib.ptr(seg_off), n_lab,
                                                    _lib.stream_ptr()))

    @classmethod
    def from_device(cls, X_dev, K_max, random_means_dev):
        """Bench-scale constructor: embeddings already resident in HBM (float32 [N, D] CUDA
        tensor), no items assigned yet.  `X` (host mirror) is not kept."""
        self = cls.__new__(cls)
        assert X_dev.is_cuda and X_dev.dtype == torch.float32 and X_dev.is_contiguous()
        _lib.lib()
        self.X = None
        self.N, self.D = int(X_dev.shape[0]), int(X_dev.shape[1])
        self.K_max = int(K_max)
        self.random_means = None
        self._X = X_dev
        self._x_is_f64 = False
        self._mean_num = torch.zeros(self.K_max, self.D, dtype=torch.float64, device="cuda")
        self._rnd = random_means_dev.contiguous()
        self._means = self._rnd.clone()
        self._meansT = self._rnd.t().contiguous()
        self._counts = torch.zeros(self.K_max, dtype=torch.int32, device="cuda")
        self._assign = torch.full((self.N,), -1, dtype=torch.int32, device="cuda")
        self._K = torch.zeros(1, dtype=torch.int32, device="cuda")
        self._row = torch.empty(self.K_max, dtype=torch.float32, device="cuda")
        self._relabel = None
        return self

    def setup_random_means(self):
        """:90-91 (consumes np.random exactly like the reference)."""
        self.random_means = self.X[np.random.choice(range(self.N), self.K_max, replace=True), :]

    def struct(self):
        m = _lib.KMeansM()
        m.D, m.K_max, m.x_is_f64, m.n_emb = self.D, self.K_max, int(self._X.dtype == torch.float64), self.N
        m.X, m.mean_num = self._X.data_ptr(), self._mean_num.data_ptr()
        m.means, m.meansT, m.random_means = self._means.data_ptr(), self._meansT.data_ptr(), self._rnd.data_ptr()
        m.counts, m.assignments, m.K = self._counts.data_ptr(), self._assign.data_ptr(), self._K.data_ptr()
        return m

    def _add_many(self, ids, ks):
        if len(ids) == 0:
            return
        ids_d, ks_d = _lib.dev(np.asarray(ids, dtype=np.int32)), _lib.dev(np.asarray(ks, dtype=np.int32))
        _lib.check(_lib.lib().segb_kmeans_add_items(self.struct(), _lib.ptr(ids_d), _lib.ptr(ks_d), len(ids),
                                                    _lib.stream_ptr()))

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
    def mean_numerators(self):
        return self._mean_num.cpu().numpy()

    @property
    def means(self):
        return self._means.cpu().numpy()

    # ---- reference API
    def add_item(self, i, k):
        """:93-111 (k > K is clamped to K)."""
        assert not i == -1
        assert int(self._assign[i].item()) == -1
        self._add_many([i], [k])

    def del_item(self, i):
        """:113-132."""
        assert not i == -1
        ids_d = _lib.dev(np.asarray([i], dtype=np.int32))
        _lib.check(_lib.lib().segb_kmeans_del_items(self.struct(), _lib.ptr(ids_d), 1, _lib.stream_ptr()))

    def clean_components(self):
        """:263-266."""
        rel = self._relabel
        _lib.check(_lib.lib().segb_kmeans_clean(self.struct(), _lib.ptr(rel), 0 if rel is None else rel.numel(),
                                                _lib.stream_ptr()))

    def neg_sqrd_norm(self, i):
        """:225-226 over all K_max rows of `means`."""
        _lib.check(_lib.lib().segb_kmeans_neg_sqrd_norm_row(self.struct(), int(i), _lib.ptr(self._row),
                                                            _lib.stream_ptr()))
        return self._row.cpu().numpy()

    def best(self, ids=None):
        """max / first argmax of neg_sqrd_norm for many items in one launch
        (ids=None: every row of X)."""
        n = self.N if ids is None else len(ids)
        ids_d = None if ids is None else _lib.dev(np.asarray(ids, dtype=np.int32))
        val = torch.empty(n, dtype=self._row.dtype, device="cuda")
        arg = torch.empty(n, dtype=torch.int32, device="cuda")
        _lib.check(_lib.lib().segb_kmeans_best(self.struct(), _lib.ptr(ids_d), n, _lib.ptr(val), _lib.ptr(arg),
                                               _lib.stream_ptr()))
        return val, arg

    def max_neg_sqrd_norm_i(self, i):
        return self.best([i])[0].cpu().numpy()[0]                           # :228-229

    def argmax_neg_sqrd_norm_i(self, i):
        return int(self.best([i])[1].item())                                # :231-232

    def get_assignments(self, list_of_i):
        return self.assignments[np.asarray(list_of_i)]

    def get_max_assignments(self, list_of_i):
        """:256-261."""
        return [int(k) for k in self.best(list(list_of_i))[1].cpu().numpy()]

    def sum_neg_sqrd_norm(self):
        """:234-247 -- diagnostic, evaluated on the device (csrc/diagnostics.cu): per-component sums of
        -|mean_numerators[k]/counts[k] - x|^2 over the assigned items, added in component order."""
        out = torch.empty(self.K_max, dtype=torch.float64, device="cuda")
        order, seg_off = _lib.members_by_component(self._assign, self.K_max)
        _lib.check(_lib.lib().segb_kmeans_sum_neg_sqrd_norm_k(self.struct(), _lib.ptr(order), _lib.ptr(seg_off),
                                                              _lib.ptr(out), _lib.stream_ptr()))
        per_k = out.cpu().numpy()
        objective = 0
        for k in range(self.K):
            objective += per_k[k]
        return objective
