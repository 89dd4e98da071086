"""
Smoothed, interpolated maximum-likelihood bigram language model on the device.

Mirror of the reference's `BigramSmoothLM` (segmentalist/bigram_lms.py:18-113).  The counts live in
HBM (K x K int32 table: 100 MB at K = 5000) because the sampling kernels read one row per token and
the acoustic model's `del_component` moves rows and columns when a component dies
(gaussian_components_fixedvar.py:205-221); the attributes `unigram_counts` / `bigram_counts` are
host mirrors.
"""
import numpy as np
import torch

from . import _lib


class BigramSmoothLM(object):

    def __init__(self, intrp_lambda, a, b, K):
        self.intrp_lambda = intrp_lambda
        self.a = a
        self.b = b
        self.K = int(K)
        self._uni = torch.zeros(self.K, dtype=torch.int32, device="cuda")
        self._bi = torch.zeros((self.K, self.K), dtype=torch.int32, device="cuda")
        self._row = torch.empty(self.K, dtype=torch.float64, device="cuda")

    def struct(self):
        lm = _lib.BigramLM()
        lm.K = self.K
        lm.intrp_lambda, lm.a, lm.b = float(self.intrp_lambda), float(self.a), float(self.b)
        lm.unigram_counts, lm.bigram_counts = self._uni.data_ptr(), self._bi.data_ptr()
        return lm

    # ---- mirrors
    @property
    def unigram_counts(self):
        return self._uni.cpu().numpy().astype(np.int64)

    @property
    def bigram_counts(self):
        return self._bi.cpu().numpy().astype(np.int64)

    # ---- probabilities
    def _log_row(self, j):
        _lib.check(_lib.lib().segb_bigram_lm_log_prob_row(self.struct(), -1 if j is None else int(j),
                                                          _lib.ptr(self._row), _lib.stream_ptr()))
        return self._row.cpu().numpy()

    def log_prob_vec_i(self):
        """:64-69."""
        return self._log_row(None)

    def log_prob_vec_given_j(self, j):
        """:78-82."""
        return self._log_row(j)

    def prob_vec_i(self):
        """:71-76 (host, from the mirrored counts: diagnostics only)."""
        uni = self.unigram_counts
        return (uni + float(self.a) / self.K) / (int(uni.sum()) + self.a)

    def prob_vec_given_j(self, j):
        """:84-91 (host)."""
        uni = self.unigram_counts
        row = self._bi[j].cpu().numpy().astype(np.int64)
        return (self.intrp_lambda * self.prob_vec_i() + (1 - self.intrp_lambda) *
                (row + float(self.b) / self.K) / (uni[j] + float(self.b)))

    def prob_i(self, i):
        """:49-54."""
        return self.prob_vec_i()[i]

    def prob_i_given_j(self, i, j):
        """:56-62."""
        uni = self.unigram_counts
        p = (int(self._bi[j, i].item()) + float(self.b) / self.K) / (uni[j] + float(self.b))
        return self.intrp_lambda * self.prob_i(i) + (1 - self.intrp_lambda) * p

    # ---- counts
    def _update(self, utterance, sign):
        tr = np.asarray(list(utterance), dtype=np.int32)
        if len(tr) == 0:
            return
        assert tr.min() >= 0 and tr.max() < self.K
        tr_d = _lib.dev(tr)
        _lib.check(_lib.lib().segb_bigram_lm_update(self.struct(), _lib.ptr(tr_d), len(tr), sign, _lib.stream_ptr()))

    def counts_from_data(self, data):
        """:93-96."""
        for utterance in data:
            self.counts_from_utterance(utterance)

    def counts_from_utterance(self, utterance):
        """:98-105."""
        self._update(utterance, 1)

    def remove_counts_from_utterance(self, utterance):
        """:107-113."""
        self._update(utterance, -1)
