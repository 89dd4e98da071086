"""
K-means model over device-resident components.

Mirror of the reference's `KMeans` (segmentalist/kmeans.py:24-176).  `fit` is
the reference's only frozen-state step: every assigned item is scored against
the same means in one batched launch, then the changed items are moved.
"""
import random
import time

import numpy as np

from . import _lib
from .fbgmm import make_consecutive
from .kmeans_components import KMeansComponents


class KMeans(object):

    def __init__(self, X, K, assignments="rand"):
        self.setup_components(K, assignments, X)

    def setup_components(self, K, assignments="rand", X=None):
        """kmeans.py:52-94."""
        if X is None:
            assert hasattr(self, "components")
            X = self.components.X
        N, D = X.shape
        if isinstance(assignments, str) and assignments == "rand":
            assignments = np.random.randint(0, K, N)
        elif isinstance(assignments, str) and assignments == "each-in-own":
            assignments = np.arange(N)
        elif isinstance(assignments, str) and assignments == "spread":
            assignment_list = (list(range(K)) * int(np.ceil(float(N) / K)))[:N]
            random.shuffle(assignment_list)
            assignments = np.array(assignment_list)
        assignments = make_consecutive(np.asarray(assignments))
        self.components = KMeansComponents(X, assignments, K)

    def fit(self, n_iter, consider_unassigned=True, no_empty=True):
        """kmeans.py:97-173.  E-step: one batched exact scoring launch over the items;
        M-step: del_item/add_item for the changed items in index order (:149-151),
        then clean_components."""
        c = self.components
        record_dict = {"sum_neg_sqrd_norm": [], "components": [], "n_mean_updates": [], "sample_time": []}
        start_time = time.time()
        for _ in range(n_iter):
            k_old = c.assignments
            items = np.arange(c.N) if consider_unassigned else np.where(k_old != -1)[0]
            _, arg = c.best(items)
            k_new = arg.cpu().numpy().astype(np.int64)
            changed = k_new != k_old[items]
            upd_i, upd_k = items[changed], k_new[changed]
            if len(upd_i):
                lib = _lib.lib()
                ids_d = _lib.dev(upd_i.astype(np.int32))
                ks_d = _lib.dev(upd_k.astype(np.int32))
                # (del_item(i); add_item(i, k)) pairs in list order
                _lib.check(lib.segb_kmeans_move_items(c.struct(), _lib.ptr(ids_d), _lib.ptr(ks_d), len(upd_i),
                                                      _lib.stream_ptr()))
            c.clean_components()
            record_dict["sum_neg_sqrd_norm"].append(c.sum_neg_sqrd_norm())
            record_dict["components"].append(c.K)
            record_dict["n_mean_updates"].append(int(len(upd_i)))
            record_dict["sample_time"].append(time.time() - start_time)
            start_time = time.time()
            if len(upd_i) == 0:
                break
        return record_dict

    def get_n_assigned(self):
        # counted on the device: `components.assignments` would mirror the whole vector to the host first
        return int((self.components._assign != -1).sum().item())
