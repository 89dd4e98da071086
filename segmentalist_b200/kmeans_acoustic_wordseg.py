"""
Segmental k-means word segmentation on the device.

Mirror of the reference's `SegmentalKMeansWordseg` and
`forward_backward_kmeans_viterbi`
(segmentalist/kmeans_acoustic_wordseg.py:27-555); `KMeansAcousticWordseg` is an
alias (the name BASELINE.json uses).  `segment()` keeps the reference's
sequential semantics (means updated after every utterance).  `segment_frozen()`
is the new frozen-state batch mode: all utterances are scored and segmented
against the same means, then the means are rebuilt -- this is the mode that
shards over GPUs (see batch.py).
"""
import logging
import random
import time

import numpy as np
import torch

from . import _lib, kmeans
from .fbgmm import make_consecutive
from .batch import FrozenKMeansSweep
from .unigram_acoustic_wordseg import _dp_single
from .utterances import DeviceCorpus, Utterances, process_embeddings

logger = logging.getLogger(__name__)
i_debug_monitor = 0
segment_debug_only = False


class SegmentalKMeansWordseg(object):

    def __init__(self, am_K, embedding_mats, vec_ids_dict, durations_dict, landmarks_dict,
                 seed_boundaries_dict=None, seed_assignments_dict=None, n_slices_min=0, n_slices_max=20,
                 min_duration=0, p_boundary_init=0.5, init_am_assignments="rand", wip=0):
        assert seed_assignments_dict is None or seed_boundaries_dict is not None
        self.n_slices_min = n_slices_min
        self.n_slices_max = n_slices_max
        self.wip = wip
        embeddings, vec_ids, labels = process_embeddings(embedding_mats, vec_ids_dict)
        self.ids_to_utterance_labels = labels
        N = embeddings.shape[0]
        seeds = None if seed_boundaries_dict is None else [seed_boundaries_dict[i] for i in labels]
        self.utterances = Utterances(
            [len(landmarks_dict[i]) for i in labels], vec_ids, [durations_dict[i] for i in labels],
            [landmarks_dict[i] for i in labels], seed_boundaries=seeds, p_boundary_init=p_boundary_init,
            n_slices_min=n_slices_min, n_slices_max=n_slices_max, min_duration=min_duration)
        init_embeds = self.utterances.all_segmented_embeds()     # get_segmented_embeds_i of every utterance, vectorised
        init_embeds = init_embeds[np.where(init_embeds != -1)]
        assignments = -1 * np.ones(N, dtype=int)
        if seed_assignments_dict is not None:
            assert False, "to-do"                                               # :148-149
        elif init_am_assignments == "rand":                                     # :181-196
            a = np.random.randint(0, am_K, len(init_embeds))
            a = make_consecutive(a)
            assignments[init_embeds] = a
        elif init_am_assignments == "spread":                                   # :198-207
            n = len(init_embeds)
            lst = (list(range(am_K)) * int(np.ceil(float(n) / am_K)))[:n]
            random.shuffle(lst)
            assignments[init_embeds] = np.array(lst)
        elif init_am_assignments == "one-by-one":
            assert False, "to-do"                                               # :208
        else:
            assert False, "invalid value for `init_am_assignments`: " + init_am_assignments
        self.acoustic_model = kmeans.KMeans(embeddings, am_K, assignments)
        self._corpus = DeviceCorpus.from_utterances(self.utterances, n_slices_min, n_slices_max)
        comps = self.acoustic_model.components
        comps._relabel = self._corpus.tok_id
        n_slots = self._corpus.N_max * self._corpus.S
        self._scratch = torch.empty(n_slots, dtype=torch.float64, device="cuda")
        self._scratch_arg = torch.empty(n_slots, dtype=torch.int32, device="cuda")
        self._frozen = None

    # ---- sequential (reference semantics)
    def _sweep(self, order):
        corpus, comps = self._corpus, self.acoustic_model.components
        n = len(order)
        order_h = np.ascontiguousarray(order, dtype=np.int32)
        totals = torch.zeros(n, dtype=torch.float64, device="cuda")
        status = torch.zeros(n, dtype=torch.int32, device="cuda")
        _lib.check(_lib.lib().segb_kmeans_segment_sweep(
            comps.struct(), corpus.struct(), order_h.ctypes.data, n, float(self.wip), _lib.ptr(self._scratch),
            None, _lib.ptr(self._scratch_arg), _lib.ptr(totals), _lib.ptr(status), _lib.stream_ptr()))
        st = status.cpu().numpy()
        self.utterances._bflat[:] = corpus.boundaries_flat()
        assert np.all(st == _lib.DP_OK), "segmentation failed for utterances %s (status %s)" % (
            list(order_h[st != 0]), list(st[st != 0]))
        return totals.cpu().numpy()

    def segment_i(self, i):
        """Segment utterance `i` and update the means (:225-332)."""
        return float(self._sweep([i])[0])

    def get_vec_embed_neg_len_sqrd_norms(self, vec_ids, durations):
        """:334-351."""
        comps = self.acoustic_model.components
        vec_ids = np.asarray(vec_ids)
        durations = np.asarray(durations, dtype=np.float64)
        val, _ = comps.best(vec_ids)
        out = val.cpu().numpy().astype(np.float64)
        live = vec_ids != -1
        out[live & np.isnan(durations)] = -np.inf
        ok = live & ~np.isnan(durations)
        out[ok] = out[ok] * durations[ok]
        out[~live] = -np.inf
        return out + self.wip

    def segment(self, n_iter, n_iter_inbetween_kmeans=0):
        """:353-425."""
        record_dict = {k: [] for k in ("sum_neg_sqrd_norm", "sum_neg_len_sqrd_norm", "components",
                                       "sample_time", "n_tokens")}
        for i_iter in range(n_iter):
            start_time = time.time()
            utt_order = list(range(self.utterances.D))
            random.shuffle(utt_order)
            if segment_debug_only:
                utt_order = [i_debug_monitor]
            total = 0
            for v in self._sweep(utt_order):
                total += v
            record_dict["sample_time"].append(time.time() - start_time)
            record_dict["sum_neg_sqrd_norm"].append(self.acoustic_model.components.sum_neg_sqrd_norm())
            record_dict["sum_neg_len_sqrd_norm"].append(total)
            record_dict["components"].append(self.acoustic_model.components.K)
            record_dict["n_tokens"].append(self.acoustic_model.get_n_assigned())
            info = "iteration: " + str(i_iter)
            for key in sorted(record_dict):
                info += ", " + key + ": " + str(record_dict[key][-1])
            logger.info(info)
            if n_iter_inbetween_kmeans > 0:
                self.acoustic_model.fit(n_iter_inbetween_kmeans, consider_unassigned=False)
        return record_dict

    # ---- frozen-state batch mode (new)
    def segment_frozen(self, n_iter, n_iter_inbetween_kmeans=0, scorer="auto", precision="auto"):
        """Frozen-means sweeps: score + Viterbi for every utterance against the same
        means, then rebuild the means from the new tokens (KMeans.fit semantics,
        kmeans.py:124-171, applied to segmentation); n_iter_inbetween_kmeans > 0 runs that many
        frozen hard-assignment steps over the current tokens after every sweep
        (kmeans_acoustic_wordseg.py:414-417), sharded like the sweep.  precision: first-level filter of the tensor-core
        scorer, "auto" / "fp16" / "fp8" (e4m3 cascade; bit-identical results, see batch.FrozenKMeansSweep).  Returns a
        record dict."""
        if self._frozen is None:
            self._frozen = FrozenKMeansSweep(self.acoustic_model.components, self._corpus, wip=self.wip,
                                             scorer=scorer, precision=precision)
        self._frozen.K_host = None      # sequential sweeps / fit() in between may have changed K
        record = {"sum_neg_len_sqrd_norm": [], "components": [], "n_tokens": [], "sample_time": []}
        for _ in range(n_iter):
            t0 = time.time()
            total = self._frozen.sweep()
            self.utterances._bflat[:] = self._corpus.boundaries_flat()
            record["sum_neg_len_sqrd_norm"].append(total)
            record["components"].append(self.acoustic_model.components.K)
            record["n_tokens"].append(self.acoustic_model.get_n_assigned())
            record["sample_time"].append(time.time() - t0)
            if n_iter_inbetween_kmeans > 0:
                record.setdefault("kmeans_fit", []).append(self._frozen.fit(n_iter_inbetween_kmeans))
        return record

    def get_unsup_transcript_i(self, i):
        return list(self.acoustic_model.components.get_assignments(self.utterances.get_segmented_embeds_i(i)))

    def get_max_unsup_transcript_i(self, i):
        return self.acoustic_model.components.get_max_assignments(self.utterances.get_segmented_embeds_i(i))


KMeansAcousticWordseg = SegmentalKMeansWordseg


def forward_backward_kmeans_viterbi(vec_embed_neg_len_sqrd_norms, N, n_slices_min=0, n_slices_max=0, i_utt=None):
    """Segmental k-means Viterbi segmentation of one utterance (:449-555)."""
    return _dp_single(vec_embed_neg_len_sqrd_norms, N, n_slices_min, n_slices_max, _lib.DP_VITERBI_KMEANS, 1.0)
