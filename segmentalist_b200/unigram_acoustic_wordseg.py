"""
Unigram acoustic word segmentation (blocked Gibbs / Viterbi) on the device.

Mirror of the reference's `UnigramAcousticWordseg` and its module-level
`forward_backward` / `forward_backward_viterbi`
(segmentalist/unigram_acoustic_wordseg.py:27-864).  The per-utterance step --
remove the utterance's tokens, score every candidate segment against every
component, run the segmentation DP, assign the new tokens one by one -- runs
entirely on the GPU (segb_gibbs_sweep_fixedvar); a whole sweep is queued on one
stream without host synchronisation, preserving the sequential collapsed-Gibbs
order.  Randomness: the host draws `random.random()` values in advance, the
kernels consume them in the reference's order, and the host RNG is then rewound
to exactly the number consumed -- a seeded run stays in lock-step with the
reference.
"""
import logging
import math
import os
import random
import time

import numpy as np
import torch

from . import _lib
from .fbgmm import make_consecutive
from .utterances import DeviceCorpus, Utterances, band_to_packed, packed_to_band, process_embeddings, tri

logger = logging.getLogger(__name__)
i_debug_monitor = 0
debug_gibbs_only = False


class UniformFeed(object):
    """Hands Python's `random.random()` stream to the device and keeps the host
    generator in step with what the kernels consumed."""

    def __init__(self, n_max, rng=None):
        self.rng = random if rng is None else rng        # the global generator (reference behaviour) or a chain's own
        self.state = self.rng.getstate()
        self.host = np.array([self.rng.random() for _ in range(n_max)], dtype=np.float64)
        self.dev = _lib.dev(self.host) if n_max else torch.zeros(1, dtype=torch.float64, device="cuda")
        self.counter = torch.zeros(1, dtype=torch.int64, device="cuda")

    def finish(self):
        used = int(self.counter.item())
        assert used <= len(self.host), "device consumed more uniforms than were provisioned"
        self.rng.setstate(self.state)
        for _ in range(used):
            self.rng.random()
        return used


def _anneal_iter(n_iter, anneal_schedule, anneal_start_temp_inv, anneal_end_temp_inv, n_anneal_steps):
    """Annealing schedules of unigram_acoustic_wordseg.py:404-421."""
    if anneal_schedule is None:
        return iter([])
    if anneal_schedule == "linear":
        if n_anneal_steps == -1:
            n_anneal_steps = n_iter
        return iter(1. / np.linspace(anneal_start_temp_inv, anneal_end_temp_inv, n_anneal_steps))
    if anneal_schedule == "step":
        assert not n_anneal_steps == -1, "`n_anneal_steps` of -1 not allowed for step annealing schedule"
        per_step = int(round(float(n_iter) / n_anneal_steps))
        temps = 1. / np.linspace(anneal_start_temp_inv, anneal_end_temp_inv, n_anneal_steps)
        return iter(np.repeat(temps, per_step))
    assert False, "invalid anneal_schedule"


class UnigramAcousticWordseg(object):

    def __init__(self, am_class, am_alpha, am_K, am_param_prior, embedding_mats, vec_ids_dict,
                 durations_dict, landmarks_dict, seed_boundaries_dict=None, seed_assignments_dict=None,
                 covariance_type="fixed", n_slices_min=0, n_slices_max=20, min_duration=0,
                 p_boundary_init=0.5, beta_sent_boundary=2.0, lms=1., wip=0., fb_type="standard",
                 init_am_assignments="rand", time_power_term=1.):
        assert seed_assignments_dict is None or seed_boundaries_dict is not None
        self.n_slices_min = n_slices_min
        self.n_slices_max = n_slices_max
        self.beta_sent_boundary = beta_sent_boundary
        self.wip = wip
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

        assignments = -1 * np.ones(N, dtype=int)
        one_by_one = False
        if seed_assignments_dict is not None:                                   # :176-204
            self.seed_to_cluster = {}
            i_cluster = 0
            for i_utt, utt in enumerate(labels):
                utt_embeds = np.array(self.utterances.get_segmented_embeds_i(i_utt), dtype=int)
                utt_assign = np.array(seed_assignments_dict[utt][:])
                utt_assign = utt_assign[np.where(utt_embeds != -1)]
                utt_embeds = utt_embeds[np.where(utt_embeds != -1)]
                for seed in utt_assign:
                    if seed not in self.seed_to_cluster:
                        if isinstance(seed, (int, np.integer)):
                            self.seed_to_cluster[seed] = seed
                        else:
                            self.seed_to_cluster[seed] = i_cluster
                            i_cluster += 1
                assignments[utt_embeds] = [self.seed_to_cluster[i] for i in utt_assign]
            if am_K is None:
                am_K = max(self.seed_to_cluster.values()) + 1
            else:
                assert am_K >= max(self.seed_to_cluster.values()) + 1
        elif init_am_assignments == "rand":                                     # :206-223
            a = np.random.randint(0, am_K, len(init_embeds))
            a = make_consecutive(a)
            assignments[init_embeds] = a
        elif init_am_assignments == "one-by-one":                               # :225-236
            one_by_one = True
        else:
            assert False, "invalid value for `init_am_assignments`: " + init_am_assignments

        self.acoustic_model = am_class(embeddings, am_param_prior, am_alpha, am_K, assignments,
                                       covariance_type=covariance_type, lms=lms)
        self._corpus = DeviceCorpus.from_utterances(self.utterances, n_slices_min, n_slices_max)
        self.acoustic_model.components._relabel = self._corpus.tok_id
        self._scratch = torch.empty(self._corpus.N_max * self._corpus.S, dtype=torch.float64, device="cuda")
        if one_by_one and len(init_embeds):
            us = [random.random() for _ in init_embeds]
            self.acoustic_model._assign(init_embeds, 0, 1.0, us)

    def set_fb_type(self, fb_type):
        """:241-250."""
        self.fb_type = fb_type
        if fb_type == "standard":
            self.fb_func = forward_backward
        elif fb_type == "viterbi":
            self.fb_func = forward_backward_viterbi
        else:
            assert False, "invalid `fb_type`: " + fb_type

    # ---- device sweep
    def _sweep(self, order, anneal_temp, anneal_gibbs_am):
        return self._sweep_finish(self._sweep_launch(order, anneal_temp, anneal_gibbs_am))

    def _sweep_launch(self, order, anneal_temp, anneal_gibbs_am, rng=None, max_ctas=0):
        """Queue one sweep over `order` on the current stream without waiting for it.  rng: a random.Random
        the uniforms come from instead of the global generator; max_ctas: CTAs of the cooperative launch
        (0 = one per SM) -- both for running several independent chains side by side (run_replica_sweeps)."""
        corpus, am = self._corpus, self.acoustic_model
        n = len(order)
        order_h = np.ascontiguousarray(order, dtype=np.int32)
        ffbs = self.fb_type == "standard"
        n_draws = int(2 * corpus.lengths[order_h].sum() + 2) if ffbs else 0
        feed = UniformFeed(n_draws, rng)
        log_probs = torch.zeros(n, dtype=torch.float64, device="cuda")
        status = torch.zeros(n, dtype=torch.int32, device="cuda")
        assert self.calc_p_continue() == 1.0
        lib, comps = _lib.lib(), am.components
        mode = _lib.DP_FFBS if ffbs else _lib.DP_VITERBI_GMM
        rc = _lib.E_UNSUPPORTED
        if os.environ.get("SEGB_GIBBS", "coop") != "steps":
            # one cooperative launch for the whole sweep (components sharded over the SMs)
            if getattr(self, "_gibbs_work", None) is None:
                self._gibbs_work = torch.empty(lib.segb_gibbs_work_bytes(comps.K_max, corpus.N_max, corpus.S),
                                               dtype=torch.uint8, device="cuda")
            lib.segb_gibbs_set_max_ctas(int(max_ctas))
            try:
                rc = lib.segb_gibbs_sweep_fixedvar_coop(
                    comps.struct(), corpus.struct(), _lib.ptr(_lib.dev(order_h)), n, mode, float(self.time_power_term),
                    float(self.wip), float(anneal_temp), int(bool(anneal_gibbs_am)), _lib.ptr(feed.dev),
                    _lib.ptr(feed.counter), _lib.ptr(self._gibbs_work), _lib.ptr(log_probs), _lib.ptr(status),
                    _lib.stream_ptr())
            finally:
                lib.segb_gibbs_set_max_ctas(0)
        if rc == _lib.E_UNSUPPORTED:
            # model too large for per-CTA shared memory: four launches per utterance
            rc = lib.segb_gibbs_sweep_fixedvar(
                comps.struct(), corpus.struct(), order_h.ctypes.data, n, mode, float(self.time_power_term),
                float(self.wip), float(anneal_temp), int(bool(anneal_gibbs_am)), _lib.ptr(feed.dev),
                _lib.ptr(feed.counter), _lib.ptr(self._scratch), _lib.ptr(log_probs), _lib.ptr(status),
                _lib.stream_ptr())
        _lib.check(rc)
        return feed, log_probs, status, order_h

    def _sweep_finish(self, handle):
        feed, log_probs, status, order_h = handle
        corpus = self._corpus
        st = status.cpu().numpy()
        feed.finish()
        self.utterances._bflat[:] = corpus.boundaries_flat()
        assert np.all(st == _lib.DP_OK), "segmentation DP failed for %d utterances, first %s (status %s)" % (
            int((st != 0).sum()), list(order_h[st != 0][:8]), list(st[st != 0][:8]))
        lp = log_probs.cpu().numpy()
        assert not np.any(lp == -np.inf)                                        # :753-754
        return lp

    def gibbs_sample_i(self, i, anneal_temp=1, anneal_gibbs_am=False):
        """Block Gibbs sample boundaries and assignments of utterance `i` (:252-360)."""
        return float(self._sweep([i], anneal_temp, anneal_gibbs_am)[0])

    def gibbs_sample(self, n_iter, am_n_iter=0, anneal_schedule=None, anneal_start_temp_inv=0.1,
                     anneal_end_temp_inv=1, n_anneal_steps=-1, anneal_gibbs_am=False):
        """Blocked Gibbs sampling over all utterances (:362-472)."""
        get_anneal_temp = _anneal_iter(n_iter, anneal_schedule, anneal_start_temp_inv, anneal_end_temp_inv,
                                       n_anneal_steps)
        record_dict = {k: [] for k in ("sample_time", "log_marg", "log_marg*length", "log_prob_z",
                                       "log_prob_X_given_z", "anneal_temp", "components", "n_tokens")}
        for i_iter in range(n_iter):
            start_time = time.time()
            if am_n_iter > 0:                                                   # :440-443
                self.acoustic_model.gibbs_sample(am_n_iter, consider_unassigned=False)
            anneal_temp = next(get_anneal_temp, anneal_end_temp_inv)
            utt_order = list(range(self.utterances.D))
            random.shuffle(utt_order)
            if debug_gibbs_only:
                utt_order = [i_debug_monitor]
            log_prob = 0
            for lp in self._sweep(utt_order, anneal_temp, anneal_gibbs_am):     # summed in visiting order
                log_prob += lp
            record_dict["sample_time"].append(time.time() - start_time)
            record_dict["log_marg"].append(self.acoustic_model.log_marg())
            record_dict["log_marg*length"].append(log_prob)
            record_dict["log_prob_z"].append(self.acoustic_model.log_prob_z())
            record_dict["log_prob_X_given_z"].append(self.acoustic_model.log_prob_X_given_z())
            record_dict["anneal_temp"].append(anneal_temp)
            record_dict["components"].append(self.acoustic_model.components.K)
            record_dict["n_tokens"].append(self.acoustic_model.get_n_assigned())
            info = "iteration: " + str(i_iter)
            for key in sorted(record_dict):
                info += ", " + key + ": " + str(record_dict[key][-1])
            logger.info(info)
        return record_dict

    # ---- frozen-state batch mode (new)
    def segment_frozen(self, n_iter, uniforms=None, precision="auto"):
        """Frozen-model sweeps: every utterance is scored (tensor-core log_marg_i) and segmented (FFBS for
        fb_type "standard", Viterbi for "viterbi") against the same model, every new token picks its
        component from that model (sampled / MAP), then the model is rebuilt from the new assignments
        (batch.FrozenFBGMMSweep) -- the mode that shards over GPUs.  Fixed-variance components only.
        uniforms: optional list (one entry per iteration) of (u_fb, u_assign) float64 arrays [sum of
        utterance lengths]; default: drawn from np.random.  precision: first-level filter of the scorer, "auto" /
        "fp16" / "fp8" (e4m3; same results, see batch.FrozenFBGMMSweep).  Returns a record dict."""
        from .batch import FrozenFBGMMSweep
        comps = self.acoustic_model.components
        assert getattr(self.acoustic_model, "covariance_type", "fixed") == "fixed", \
            "the frozen FBGMM sweep covers fixed-variance components"
        assert self.calc_p_continue() == 1.0
        if getattr(self, "_frozen", None) is None:
            self._frozen = FrozenFBGMMSweep(comps, self._corpus, fb_type=self.fb_type,
                                            time_power_term=self.time_power_term, wip=self.wip, precision=precision)
        self._frozen.K_host = None          # sequential sweeps in between may have changed K
        n_pos = self._corpus.n_pos
        record = {k: [] for k in ("sample_time", "log_marg*length", "components", "n_tokens", "fallback_rows")}
        for it in range(n_iter):
            t0 = time.time()
            u_fb = u_assign = None
            if self.fb_type == "standard":
                if uniforms is not None:
                    u_fb, u_assign = (_lib.dev(np.asarray(a, dtype=np.float64)) for a in uniforms[it])
                else:
                    u_fb, u_assign = _lib.dev(np.random.rand(n_pos)), _lib.dev(np.random.rand(n_pos))
            total = self._frozen.sweep(u_fb, u_assign)
            self.utterances._bflat[:] = self._corpus.boundaries_flat()
            record["sample_time"].append(time.time() - t0)
            record["log_marg*length"].append(total)
            record["components"].append(self._frozen.K_host)
            record["n_tokens"].append(int((self._corpus.tok_id >= 0).sum().item()))
            record["fallback_rows"].append(self._frozen.last_fallback)
        return record

    def get_vec_embed_log_probs(self, vec_ids, durations):
        """log marginals of the `vec_ids` embeddings scaled by `durations` (:474-511)."""
        return self.acoustic_model.log_marg_items(np.asarray(vec_ids), np.asarray(durations, dtype=np.float64),
                                                  self.time_power_term, self.wip)

    def calc_p_continue(self):
        """:513-531."""
        if self.beta_sent_boundary != -1:
            assert False, "to check"
        return 1.0

    def get_unsup_transcript_i(self, i):
        return list(self.acoustic_model.components.get_assignments(self.utterances.get_segmented_embeds_i(i)))

    def get_log_margs_i(self, i):
        """:539-564 -- remove utterance i, evaluate its tokens' log marginals, add them back."""
        comps = self.acoustic_model.components
        embeds = self.utterances.get_segmented_embeds_i(i)
        assign = comps.get_assignments(embeds)
        for e in embeds:
            if e != -1:
                comps.del_item(e)
        out = [self.acoustic_model.log_marg_i(e) for e in embeds if e != -1]
        for e, k in zip(embeds, assign):
            comps.add_item(e, k)
        return out


def run_replica_sweeps(segmenters, orders, rngs, anneal_temp=1, anneal_gibbs_am=False):
    """One sweep of R INDEPENDENT chains side by side (sequential collapsed Gibbs does not shard over
    utterances; replicas do).  Every chain is one cooperative launch on its own stream with n_sm / R CTAs and
    its own random.Random; within a chain the reference's sequential order is untouched, so each chain
    produces exactly the samples it would produce alone.  Returns the per-chain log_prob arrays."""
    n_sm = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    per = max(1, n_sm // len(segmenters))
    if getattr(run_replica_sweeps, "_streams", None) is None or len(run_replica_sweeps._streams) < len(segmenters):
        run_replica_sweeps._streams = [torch.cuda.Stream() for _ in segmenters]
    handles = []
    for seg, order, rng, stream in zip(segmenters, orders, rngs, run_replica_sweeps._streams):
        stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(stream):
            handles.append(seg._sweep_launch(order, anneal_temp, anneal_gibbs_am, rng=rng, max_ctas=per))
    out = []
    for seg, h, stream in zip(segmenters, handles, run_replica_sweeps._streams):
        with torch.cuda.stream(stream):
            out.append(seg._sweep_finish(h))
    return out


# ---------------------------------------------------------------------------
# function-level API on packed vectors (the reference's fb_func seam)
# ---------------------------------------------------------------------------

def _dp_single(vec, N, n_slices_min, n_slices_max, mode, anneal_temp, log_p_continue=0.0):
    """Run one packed-triangular score vector through the batched DP kernel."""
    vec = np.asarray(vec, dtype=np.float64)
    S = N if (n_slices_max == 0 or n_slices_max > N) else n_slices_max
    band = packed_to_band(vec[:N * (N + 1) // 2], N, S, -np.inf)
    corpus = DeviceCorpus([N], np.full((N, S), -1, np.int32), np.full((N, S), np.nan),
                          np.zeros(N, np.uint8), n_slices_min, n_slices_max, S)
    scores = _lib.dev(band)
    ffbs = mode == _lib.DP_FFBS
    feed = UniformFeed(N + 1 if ffbs else 0)
    log_prob = torch.zeros(1, dtype=torch.float64, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.check(_lib.lib().segb_dp_banded(
        corpus.struct(), 0, 1, _lib.ptr(scores), mode, float(log_p_continue),
        1.0 if anneal_temp is None else float(anneal_temp), _lib.ptr(feed.dev), _lib.ptr(feed.counter),
        _lib.ptr(corpus.bounds), _lib.ptr(log_prob), None, None, _lib.ptr(status), _lib.stream_ptr()))
    st = int(status.item())
    feed.finish()
    assert st == _lib.DP_OK, "segmentation DP failed (status %d)" % st
    return float(log_prob.item()), corpus.bounds.cpu().numpy().astype(bool)


def forward_backward(vec_embed_log_probs, log_p_continue, N, n_slices_min=0, n_slices_max=0, i_utt=None,
                     anneal_temp=1):
    """Forward filtering, backward sampling (:653-756)."""
    lp, b = _dp_single(vec_embed_log_probs, N, n_slices_min, n_slices_max, _lib.DP_FFBS, anneal_temp,
                       log_p_continue)
    assert lp != -np.inf
    return lp, b


def forward_backward_viterbi(vec_embed_log_probs, log_p_continue, N, n_slices_min=0, n_slices_max=0,
                             i_utt=None, anneal_temp=None):
    """Viterbi segmentation under the GMM scores (:759-864)."""
    return _dp_single(vec_embed_log_probs, N, n_slices_min, n_slices_max, _lib.DP_VITERBI_GMM, 1.0)
