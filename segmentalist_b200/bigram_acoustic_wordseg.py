"""
Bigram acoustic word segmentation on the device: mirror of the reference's `BigramAcousticWordseg`
(segmentalist/bigram_acoustic_wordseg.py:30-749) for fb_type="unigram" -- segmentation as in the unigram
model (segment scores from `log_marg_i_embed_unigram`, :314-330), component assignments sampled under the
smoothed bigram LM (`gibbs_sample_inside_loop_i_embed`, :333-384).  fb_type="bigram" is a stub in the
reference (its forward_backward is `pass`, :756-789) and asserts here.

A whole sweep is queued on one stream without host synchronisation (segb_gibbs_sweep_bigram): per
utterance the transcript leaves the LM, the tokens leave the components (LM rows/columns follow a
component that moves), segments are scored, FFBS draws the boundaries, the new tokens are sampled left
to right with the bigram prior row of the previous label, and the new transcript enters the LM.
Randomness as in unigram_acoustic_wordseg.py: `random.random()` values are drawn ahead, consumed in the
reference's order and the host generator is rewound to the number used.
"""
import logging
import os
import random
import time

import numpy as np
import torch

from . import _lib
from .fbgmm import make_consecutive
from .bigram_fbgmm import BigramFBGMM
from .bigram_lms import BigramSmoothLM
from .unigram_acoustic_wordseg import UniformFeed, _anneal_iter
from .utterances import DeviceCorpus, Utterances, process_embeddings

logger = logging.getLogger(__name__)
i_debug_monitor = 0
debug_gibbs_only = False


class BigramAcousticWordseg(object):

    def __init__(self, am_K, am_param_prior, lm_params, embedding_mats, vec_ids_dict, durations_dict,
                 landmarks_dict, seed_boundaries_dict=None, seed_assignments_dict=None, covariance_type="fixed",
                 n_slices_min=0, n_slices_max=20, min_duration=0, p_boundary_init=0.5, beta_sent_boundary=2.0,
                 lms=1., wip=0., fb_type="bigram", init_am_assignments="rand", time_power_term=1.):
        assert seed_assignments_dict is None or seed_boundaries_dict is not None
        self.n_slices_min = n_slices_min
        self.n_slices_max = n_slices_max
        self.beta_sent_boundary = beta_sent_boundary
        self.wip = wip
        self.lms = lms
        self.time_power_term = time_power_term
        self.set_fb_type(fb_type)

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

        assert lm_params["type"] == "smooth", "only the smoothed ML bigram LM exists (:184-189)"
        self.lm = BigramSmoothLM(lm_params["intrp_lambda"], lm_params["a"], lm_params["b"], am_K)

        assignments = -1 * np.ones(N, dtype=int)
        assert seed_assignments_dict is None, "seed assignments: not supported on the device path"
        if init_am_assignments == "rand":                                       # :228-243
            a = np.random.randint(0, am_K, len(init_embeds))
            a = make_consecutive(a)
            assignments[init_embeds] = a
        elif init_am_assignments == "one-by-one":
            assert False                                                        # :245-246
        else:
            assert False, "invalid value for `init_am_assignments`: " + init_am_assignments
        self.acoustic_model = BigramFBGMM(embeddings, am_param_prior, am_K, assignments,
                                          covariance_type=covariance_type, lms=lms, lm=self.lm)
        self._corpus = DeviceCorpus.from_utterances(self.utterances, n_slices_min, n_slices_max)
        self.acoustic_model.components._relabel = self._corpus.tok_id
        self._scratch = torch.empty(self._corpus.N_max * self._corpus.S, dtype=torch.float64, device="cuda")
        self.set_lm_counts()

    def set_fb_type(self, fb_type):
        """:253-263."""
        self.fb_type = fb_type
        assert fb_type in ("unigram", "bigram"), "invalid `fb_type`: " + fb_type
        assert fb_type == "unigram", "to-do: the reference's bigram forward-backward is a stub (:756-789)"

    def set_lm_counts(self):
        """:265-267."""
        for i_utt in range(self.utterances.D):
            self.lm.counts_from_utterance(self.get_unsup_transcript_i(i_utt))

    def get_unsup_transcript_i(self, i):
        """:743-747."""
        return list(self.acoustic_model.components.get_assignments(self.utterances.get_segmented_embeds_i(i)))

    # ---- diagnostics (host, over the mirrored transcripts and an LM rebuilt from scratch, as the reference does)
    def log_prob_z(self):
        """:269-305."""
        K, lam, a, b = self.lm.K, self.lm.intrp_lambda, self.lm.a, self.lm.b
        uni = np.zeros(K, np.int64)
        bi = {}
        total = 0
        assign = self.acoustic_model.components.assignments
        log_prob_z = 0.
        for i_utt in range(self.utterances.D):
            j_prev = None
            for e in self.utterances.get_segmented_embeds_i(i_utt):
                i_cur = int(assign[e])
                prob_i = (uni[i_cur] + float(a) / K) / (total + a)
                if j_prev is not None:
                    p = (bi.get((j_prev, i_cur), 0) + float(b) / K) / (uni[j_prev] + float(b))
                    log_prob_z += np.log(lam * prob_i + (1 - lam) * p)
                    bi[(j_prev, i_cur)] = bi.get((j_prev, i_cur), 0) + 1
                else:
                    log_prob_z += np.log(prob_i)
                uni[i_cur] += 1
                total += 1
                # (the reference never advances j_prev in this loop, :283-298, so every term is a unigram
                # term; reproduced as is -- the record values are compared with the reference's)
        return log_prob_z

    def log_marg(self):
        """:307-311."""
        return self.log_prob_z() + self.acoustic_model.log_prob_X_given_z()

    def log_marg_i_embed_unigram(self, i_embed):
        """:314-330: FBGMM.log_marg_i with the LM's unigram smoothing (the counts are tied)."""
        assert i_embed != -1
        c = self.acoustic_model.components
        out = torch.empty(1, dtype=torch.float64, device="cuda")
        ids = _lib.dev(np.asarray([i_embed], dtype=np.int32))
        _lib.check(_lib.lib().segb_fixedvar_log_marg(c.struct(), _lib.ptr(ids), None, 1, 1.0, 0.0, _lib.ptr(out),
                                                     _lib.stream_ptr()))
        return float(out.item())

    def get_vec_embed_log_probs_unigram(self, vec_ids, durations):
        """:696-714."""
        c = self.acoustic_model.components
        ids = _lib.dev(np.asarray(vec_ids, dtype=np.int32))
        durs = _lib.dev(np.asarray(durations, dtype=np.float64))
        out = torch.empty(len(vec_ids), dtype=torch.float64, device="cuda")
        _lib.check(_lib.lib().segb_fixedvar_log_marg(c.struct(), _lib.ptr(ids), _lib.ptr(durs), len(vec_ids),
                                                     float(self.time_power_term), float(self.wip), _lib.ptr(out),
                                                     _lib.stream_ptr()))
        return out.cpu().numpy()

    get_vec_embed_log_probs = get_vec_embed_log_probs_unigram

    def calc_p_continue(self):
        """:719-741."""
        if self.beta_sent_boundary != -1:
            assert False, "to check"
        return 1.0

    # ---- device sweep
    def _sweep(self, order, anneal_temp, anneal_gibbs_am, assignments_only):
        corpus, comps = self._corpus, self.acoustic_model.components
        n = len(order)
        order_h = np.ascontiguousarray(order, dtype=np.int32)
        feed = UniformFeed(int(2 * corpus.lengths[order_h].sum() + 2))
        log_probs = torch.zeros(n, dtype=torch.float64, device="cuda")
        status = torch.zeros(n, dtype=torch.int32, device="cuda")
        assert self.calc_p_continue() == 1.0
        lib = _lib.lib()
        rc = _lib.E_UNSUPPORTED
        if not assignments_only and os.environ.get("SEGB_GIBBS", "coop") != "steps":
            # one cooperative launch for the whole sweep (components sharded over the SMs, CTA 0 keeps the LM)
            if getattr(self, "_gibbs_work", None) is None:
                self._gibbs_work = torch.empty(lib.segb_gibbs_work_bytes(comps.K_max, corpus.N_max, corpus.S),
                                               dtype=torch.uint8, device="cuda")
            rc = lib.segb_gibbs_sweep_bigram_coop(
                comps.struct(), self.lm.struct(), corpus.struct(), _lib.ptr(_lib.dev(order_h)), n,
                float(self.time_power_term), float(self.wip), float(anneal_temp), int(bool(anneal_gibbs_am)),
                _lib.ptr(feed.dev), _lib.ptr(feed.counter), _lib.ptr(self._gibbs_work), _lib.ptr(log_probs),
                _lib.ptr(status), _lib.stream_ptr())
        if rc == _lib.E_UNSUPPORTED:
            # assignments-only sweeps, or a model too large for per-CTA shared memory: four launches per utterance
            rc = lib.segb_gibbs_sweep_bigram(
                comps.struct(), self.lm.struct(), corpus.struct(), order_h.ctypes.data, n, int(bool(assignments_only)),
                float(self.time_power_term), float(self.wip), float(anneal_temp), int(bool(anneal_gibbs_am)),
                _lib.ptr(feed.dev), _lib.ptr(feed.counter), _lib.ptr(self._scratch), _lib.ptr(log_probs), _lib.ptr(status),
                _lib.stream_ptr())
        _lib.check(rc)
        st = status.cpu().numpy()
        feed.finish()
        self.utterances._bflat[:] = corpus.boundaries_flat()
        assert np.all(st == _lib.DP_OK), "segmentation DP failed for %d utterances, first %s (status %s)" % (
            int((st != 0).sum()), list(order_h[st != 0][:8]), list(st[st != 0][:8]))
        lp = log_probs.cpu().numpy()
        assert not np.any(lp == -np.inf)
        return lp

    def gibbs_sample_i(self, i, anneal_temp=1, anneal_gibbs_am=False, assignments_only=False):
        """Block Gibbs sample boundaries and assignments of utterance `i` (:386-543)."""
        return float(self._sweep([i], anneal_temp, anneal_gibbs_am, assignments_only)[0])

    def gibbs_sample(self, n_iter, am_n_iter=0, anneal_schedule=None, anneal_start_temp_inv=0.1,
                     anneal_end_temp_inv=1, n_anneal_steps=-1, anneal_gibbs_am=False, assignments_only=False):
        """Blocked Gibbs sampling over all utterances (:545-694)."""
        get_anneal_temp = _anneal_iter(n_iter, anneal_schedule, anneal_start_temp_inv, anneal_end_temp_inv,
                                       n_anneal_steps)
        record_dict = {k: [] for k in ("sample_time", "log_marg", "log_marg*length", "log_prob_z",
                                       "log_prob_X_given_z", "anneal_temp", "components", "n_tokens")}
        for i_iter in range(n_iter):
            start_time = time.time()
            if am_n_iter > 0:
                assert False, "to-do"                                           # :646-650
            anneal_temp = next(get_anneal_temp, anneal_end_temp_inv)
            utt_order = list(range(self.utterances.D))
            random.shuffle(utt_order)
            if debug_gibbs_only:
                utt_order = [i_debug_monitor]
            log_prob = 0
            for lp in self._sweep(utt_order, anneal_temp, anneal_gibbs_am, assignments_only):
                log_prob += lp
            record_dict["sample_time"].append(time.time() - start_time)
            record_dict["log_marg"].append(self.log_marg())
            record_dict["log_marg*length"].append(log_prob)
            record_dict["log_prob_z"].append(self.log_prob_z())
            record_dict["log_prob_X_given_z"].append(self.acoustic_model.log_prob_X_given_z())
            record_dict["anneal_temp"].append(anneal_temp)
            record_dict["components"].append(self.acoustic_model.components.K)
            record_dict["n_tokens"].append(self.acoustic_model.get_n_assigned())
            info = "iteration: " + str(i_iter)
            for key in sorted(record_dict):
                info += ", " + key + ": " + str(record_dict[key][-1])
            logger.info(info)
        return record_dict
