"""
Frozen-state batch sweep for segmental k-means (new mode).

Given frozen means, every utterance is independent: score all candidate
segments (max / argmax over K_max components), build the banded duration-
weighted scores, run the Viterbi DP, read off the chosen segments and their
components.  The model update is then a pure reduction -- per-component sum of
embeddings and counts -- which is what shards over GPUs: each rank owns a
contiguous range of utterances (and their embeddings) plus a replica of the
means, and one NCCL all-reduce of (sum_x [K_max, D] float64, cnt [K_max] int64)
per sweep rebuilds identical means everywhere.

Semantics (pinned by the test oracle's frozen_kmeans_sweep and
tests/golden/kmeans_wordseg.npz): phase 1 = the reference's pure functions
get_vec_embed_neg_len_sqrd_norms -> forward_backward_kmeans_viterbi ->
get_max_assignments (kmeans_acoustic_wordseg.py:334-351,449-555,
kmeans_components.py:256-261) per utterance; phase 2 = del_item for every old
token, add_item for every new one in utterance order, clean_components()
(kmeans_acoustic_wordseg.py:312-320).

Two scorers produce the per-embedding (max, argmax):
  "exact" -- SIMT kernel, float32 in NumPy order (segb_kmeans_best);
  "mma"   -- tcgen05 filter GEMM + exact refine (segb_mma_filter / segb_mma_refine),
             same bits out, tensor-core speed.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .sharding import dist_on as _dist_on, reduce_packed


def fused_supported(D):
    """The fused score kernel (csrc/score_fused.cu) double-buffers two 128-row operand tiles of
    roundup(D + 6, 16) fp16 columns next to two model tiles: D <= 138; its k-means refine needs an even D."""
    return D % 2 == 0 and D <= 138


class MmaScorer(object):
    """Buffers and call sequence of the tensor-core k-means scorer.

    fused=False (default): the two-kernel path -- segb_mma_filter over a pre-packed fp16 tile image of X,
    then segb_mma_refine -- with its 32-byte per-row records and per-row rounding-error norms.
    fused=True: ONE kernel per sweep reads the fp32 embeddings once, converts them to fp16 operand tiles in
    shared memory, runs the filter GEMM and re-scores the surviving candidates exactly
    (segb_fused_kmeans_best): no fp16 image of X (-0.55 x the size of X in HBM), no pack pass, no per-row
    filter records, HBM traffic = X once.  Same bits out.  Measured at 21M rows x K = 5000 the fused kernel
    takes 23.8 ms against 19.5 + 3.7 ms for filter + refine: it trades ~3 % of sweep time for
    6.7 GB of HBM and the packing pass, which pays when X streams in from the host or memory is tight."""

    def __init__(self, components, fused=None, precision="fp16"):
        lib, c, dev = _lib.lib(), components, "cuda"
        assert c._X.dtype == torch.float32, "tensor-core scorer needs float32 embeddings"
        assert precision in ("fp16", "fp8")
        self.c = c
        self.fused = False if fused is None else bool(fused)
        assert not self.fused or fused_supported(c.D)
        # precision="fp8": the first-level filter runs in e4m3 (segb_mma8_*: twice the MMA rate, a 2.6x smaller image
        # of X); rows it cannot decide take the fp16 second-level pass.  Needs an even D (8-lane refine).
        # K_max <= 65536: the pass keeps its top-3 as packed keys with 12-bit chunk ids (mma_common.cuh, KEY_ID_BITS).
        self.fp8 = precision == "fp8" and not self.fused and c.D % 2 == 0 and c.K_max <= 65536
        self.scale = 1.0
        self.w_tiles = torch.empty(lib.segb_mma_w_tiles_bytes(c.K_max, c.D), dtype=torch.uint8, device=dev)
        # refine scratch: list of undecided rows + (two-kernel path) their compact fp16 image, thresholds and
        # candidate bitmaps for the second-level tensor pass
        self.work = torch.empty(lib.segb_mma_refine_work_bytes(c.N, c.K_max) if self.fused else
                                lib.segb_mma_refine2_work_bytes(c.N, c.K_max, c.D), dtype=torch.uint8, device=dev)
        self.w_err = torch.empty(2 * (c.K_max + 128), dtype=torch.float32, device=dev)  # (|dmu|, |mu^|)
        self.w_max = torch.zeros(2, dtype=torch.float32, device=dev)
        self.n_fallback = torch.zeros(1, dtype=torch.int64, device=dev)
        self.x_tiles = self.cand = self.x_err = self.x_max = None
        # second-level rounds to launch per refine call: 0 = as many as all rows may need (first use); a sweep that knows
        # the previous count of undecided rows sets it to what twice that count needs (rows beyond take the exhaustive scan)
        self.max_rounds = 0
        self.timing = None      # a list: score() appends (start, after filter/fused kernel, after refine) CUDA events
        if self.fp8:
            self.scale = self.pick_scale(c._X)
            self.x_tiles = torch.empty(lib.segb_mma8_x_tiles_bytes(c.N, c.D), dtype=torch.uint8, device=dev)
            self.w_tiles8 = torch.empty(lib.segb_mma8_w_tiles_bytes(c.K_max, c.D), dtype=torch.uint8, device=dev)
            self.w_err8 = torch.empty(4 * (c.K_max + 128), dtype=torch.float32, device=dev)
            self.w_max8 = torch.zeros(4, dtype=torch.float32, device=dev)
        elif not self.fused:
            self.x_tiles = torch.empty(lib.segb_mma_x_tiles_bytes(c.N, c.D), dtype=torch.uint8, device=dev)
        if not self.fused:
            self.cand = torch.empty(lib.segb_mma_cand_bytes(c.N), dtype=torch.uint8, device=dev)
            self.x_err = torch.empty(2 * c.N, dtype=torch.float32, device=dev)              # (|dx|, |x|) per row
            self.x_max = torch.zeros(2, dtype=torch.float32, device=dev)
            self.pack_x()

    def rounds_for(self, n_rows):
        """Second-level rounds that n_rows undecided rows need with this scorer's work buffer (a round takes n_emb / 8
        rows, in whole 256-row work items)."""
        cap = max(16 * 256, (self.c.N // 8 + 255) // 256 * 256)
        return int(max(1, -(-int(n_rows) // cap)))

    @staticmethod
    def pick_scale(X, chunk=1 << 20):
        """The largest power of two s with s * max|x_d| <= 448 (e4m3's largest finite value) and s * max|x| <= 448
        (so that s^2 |mu|^2 / 2 fits three e4m3 terms times the 256.0 constant column; means are averages of rows)."""
        lo, hi = torch.aminmax(X)
        max_elem = max(abs(float(lo)), abs(float(hi)))
        max_norm = 0.0
        for i in range(0, X.shape[0], chunk):
            max_norm = max(max_norm, float(torch.linalg.vector_norm(X[i:i + chunk], dim=1).max()))
        lim = max(max_elem, max_norm, 1e-30)
        if not np.isfinite(lim):
            return 1.0
        return float(2.0 ** min(40.0, np.floor(np.log2(448.0 / lim))))      # s^2 * scores must stay inside float32

    def pack_x(self):
        c = self.c
        if self.fp8:
            _lib.check(_lib.lib().segb_mma8_pack_x(_lib.ptr(c._X), c.N, c.D, self.scale, _lib.ptr(self.x_tiles),
                                                   _lib.ptr(self.x_err), _lib.ptr(self.x_max), _lib.stream_ptr()))
            return
        _lib.check(_lib.lib().segb_mma_pack_x(_lib.ptr(c._X), c.N, c.D, _lib.ptr(self.x_tiles), _lib.ptr(self.x_err),
                                              _lib.ptr(self.x_max), _lib.stream_ptr()))

    def pack_means(self):
        c = self.c
        _lib.check(_lib.lib().segb_mma_pack_means(_lib.ptr(c._means), c.K_max, c.D, _lib.ptr(self.w_tiles),
                                                  _lib.ptr(self.w_err), _lib.ptr(self.w_max), _lib.stream_ptr()))
        if self.fp8:
            _lib.check(_lib.lib().segb_mma8_pack_means(_lib.ptr(c._means), c.K_max, c.D, self.scale, _lib.ptr(self.w_tiles8),
                                                       _lib.ptr(self.w_err8), _lib.ptr(self.w_max8), _lib.stream_ptr()))

    def filter(self):
        c = self.c
        if self.fp8:
            _lib.check(_lib.lib().segb_mma8_filter(_lib.ptr(self.x_tiles), _lib.ptr(self.w_tiles8), c.N, c.K_max, c.D,
                                                   _lib.ptr(self.x_max), _lib.ptr(self.w_max8), _lib.ptr(self.cand),
                                                   _lib.stream_ptr()))
            return
        _lib.check(_lib.lib().segb_mma_filter(_lib.ptr(self.x_tiles), _lib.ptr(self.w_tiles), c.N, c.K_max, c.D,
                                              _lib.ptr(self.x_max), _lib.ptr(self.w_max), _lib.ptr(self.cand),
                                              _lib.stream_ptr()))

    def refine(self, best_val, best_k):
        c = self.c
        if self.fp8:
            _lib.check(_lib.lib().segb_mma8_refine(c.struct(), _lib.ptr(self.cand), _lib.ptr(self.x_err), _lib.ptr(self.w_max8),
                                                   self.scale, _lib.ptr(self.w_tiles), _lib.ptr(self.w_max), c.N,
                                                   _lib.ptr(self.work), self.work.numel(), self.max_rounds, _lib.ptr(best_val),
                                                   _lib.ptr(best_k), _lib.ptr(self.n_fallback), _lib.stream_ptr()))
            return
        _lib.check(_lib.lib().segb_mma_refine2(c.struct(), _lib.ptr(self.x_tiles), _lib.ptr(self.w_tiles),
                                               _lib.ptr(self.cand), _lib.ptr(self.x_err), _lib.ptr(self.w_max), c.N,
                                               _lib.ptr(self.work), self.work.numel(), self.max_rounds, _lib.ptr(best_val),
                                               _lib.ptr(best_k), _lib.ptr(self.n_fallback), _lib.stream_ptr()))

    def fused_score(self, best_val, best_k):
        """The model image must be current (pack_means)."""
        c = self.c
        _lib.check(_lib.lib().segb_fused_kmeans_best(c.struct(), _lib.ptr(self.w_tiles), _lib.ptr(self.w_max), c.N,
                                                     _lib.ptr(self.work), _lib.ptr(best_val), _lib.ptr(best_k),
                                                     _lib.ptr(self.n_fallback), _lib.stream_ptr()))

    def score(self, best_val, best_k):
        self.pack_means()
        ev = None
        if self.timing is not None:          # device time of the kernels INSIDE a running sweep (bench.py roofline)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
        if self.fused:
            self.fused_score(best_val, best_k)
            if ev:
                ev[1].record()
        else:
            self.filter()
            if ev:
                ev[1].record()
            self.refine(best_val, best_k)
        if ev:
            ev[2].record()
            self.timing.append(ev)

    def score_streamed(self, X_host, best_val, best_k, chunk_rows=1 << 20):
        """Score embeddings that still live in (pinned) HOST memory: the rows are uploaded in
        chunks on a copy stream while the previous chunk is scored on the compute stream, so the
        sweep costs max(PCIe, compute) instead of their sum.  Chunks start on 256-row boundaries
        (whole work items); every kernel is the same C-ABI call as in score(), handed pointers
        offset to the chunk."""
        import ctypes
        lib, c, sp = _lib.lib(), self.c, _lib.stream_ptr()
        assert X_host.shape == c._X.shape and X_host.dtype == torch.float32 and X_host.is_pinned()
        assert chunk_rows % 256 == 0
        if getattr(self, "copy_stream", None) is None:
            self.copy_stream = torch.cuda.Stream()
            self.fb_total = torch.zeros(1, dtype=torch.int64, device="cuda")
        main = torch.cuda.current_stream()
        self.copy_stream.wait_stream(main)             # earlier kernels may still read X
        self.pack_means()
        self.fb_total.zero_()
        m = c.struct()
        x_base = c._X.data_ptr()
        vp = ctypes.c_void_p
        if not self.fused:
            kp2 = (lib.segb_mma8_x_tiles_bytes(256, c.D) if self.fp8 else lib.segb_mma_x_tiles_bytes(256, c.D)) // 256   # bytes of tile image per row
            rec = lib.segb_mma_cand_bytes(1)
        for lo in range(0, c.N, chunk_rows):
            hi = min(c.N, lo + chunk_rows)
            n = hi - lo
            with torch.cuda.stream(self.copy_stream):
                c._X[lo:hi].copy_(X_host[lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
            main.wait_event(ev)
            m.X, m.n_emb = x_base + 4 * c.D * lo, n
            bv, bk = vp(best_val.data_ptr() + 4 * lo), vp(best_k.data_ptr() + 4 * lo)
            if self.fused:
                _lib.check(lib.segb_fused_kmeans_best(m, _lib.ptr(self.w_tiles), _lib.ptr(self.w_max), n,
                                                      _lib.ptr(self.work), bv, bk, _lib.ptr(self.n_fallback), sp))
            elif self.fp8:
                xt = vp(self.x_tiles.data_ptr() + lo * kp2)
                xe = vp(self.x_err.data_ptr() + 8 * lo)
                cd = vp(self.cand.data_ptr() + rec * lo)
                _lib.check(lib.segb_mma8_pack_x(vp(x_base + 4 * c.D * lo), n, c.D, self.scale, xt, xe, _lib.ptr(self.x_max), sp))
                _lib.check(lib.segb_mma8_filter(xt, _lib.ptr(self.w_tiles8), n, c.K_max, c.D, _lib.ptr(self.x_max),
                                                _lib.ptr(self.w_max8), cd, sp))
                _lib.check(lib.segb_mma8_refine(m, cd, xe, _lib.ptr(self.w_max8), self.scale, _lib.ptr(self.w_tiles),
                                                _lib.ptr(self.w_max), n, _lib.ptr(self.work), self.work.numel(), self.max_rounds, bv, bk,
                                                _lib.ptr(self.n_fallback), sp))
            else:
                xt = vp(self.x_tiles.data_ptr() + lo * kp2)
                xe = vp(self.x_err.data_ptr() + 8 * lo)
                cd = vp(self.cand.data_ptr() + rec * lo)
                _lib.check(lib.segb_mma_pack_x(vp(x_base + 4 * c.D * lo), n, c.D, xt, xe, _lib.ptr(self.x_max), sp))
                _lib.check(lib.segb_mma_filter(xt, _lib.ptr(self.w_tiles), n, c.K_max, c.D, _lib.ptr(self.x_max),
                                               _lib.ptr(self.w_max), cd, sp))
                _lib.check(lib.segb_mma_refine2(m, xt, _lib.ptr(self.w_tiles), cd, xe, _lib.ptr(self.w_max), n,
                                                _lib.ptr(self.work), self.work.numel(), self.max_rounds, bv, bk,
                                                _lib.ptr(self.n_fallback), sp))
            self.fb_total += self.n_fallback
        self.n_fallback.copy_(self.fb_total)


class FrozenKMeansSweep(object):

    # precision of the tensor-core scorer's first-level filter: "fp16", "fp8" (e4m3 first level, fp16 second level) or
    # "auto" = start in e4m3 and fall back to fp16 once a sweep leaves more than AUTO_FP16_FRACTION of the rows to the
    # second level (a diffuse model: the e4m3 pass would only add work); the e4m3 level is tried again after
    # AUTO_RETRY_SWEEPS sweeps (a model that has converged to separated components is served better by it), with the
    # interval doubling after every failed retry.  Results are bit-identical in every mode.
    AUTO_FP16_FRACTION = 0.35
    AUTO_RETRY_SWEEPS = 8

    def __init__(self, components, corpus, wip=0.0, scorer="auto", fused=None, precision="auto"):
        self.c, self.corpus, self.wip = components, corpus, float(wip)
        assert precision in ("auto", "fp16", "fp8")
        self.precision_mode = precision
        self._auto_interval = self.AUTO_RETRY_SWEEPS        # sweeps to stay in fp16 before e4m3 is tried again
        self._auto_countdown = 0
        lib = _lib.lib()
        c = components
        if scorer == "auto":
            scorer = "mma" if (c._X.dtype == torch.float32 and c.D <= 157) else "exact"
        assert scorer in ("exact", "mma")
        self.scorer = scorer
        dev = "cuda"
        self.best_val = torch.empty(c.N, dtype=c._row.dtype, device=dev)
        self.best_k = torch.empty(c.N, dtype=torch.int32, device=dev)
        self.scores = torch.empty(corpus.n_pos * corpus.S, dtype=torch.float64, device=dev)
        self.log_prob = torch.zeros(corpus.n_utt, dtype=torch.float64, device=dev)
        self.status = torch.zeros(corpus.n_utt, dtype=torch.int32, device=dev)
        # one flat float64 buffer = what a sweep all-reduces: [sum_x (K_max*D) | counts (K_max)]
        self.red = torch.zeros(c.K_max * c.D + c.K_max + 1, dtype=torch.float64, device=dev)
        self.sum_x = self.red[:c.K_max * c.D].view(c.K_max, c.D)
        self.cnt_f = self.red[c.K_max * c.D:c.K_max * c.D + c.K_max]
        self.obj = self.red[c.K_max * c.D + c.K_max:]         # several ranks: the sweep objective rides in the same all-reduce
        self.cnt = torch.zeros(c.K_max, dtype=torch.int64, device=dev)
        # end-of-sweep scalars read with ONE device->host copy: [bad DP statuses, fallback rows, emptied components]
        self.flags = torch.zeros(4, dtype=torch.int64, device=dev)            # [3]: bits of the all-reduced objective
        self.flags_h = torch.zeros(4, dtype=torch.int64).pin_memory()
        self.log_prob_h = torch.zeros(corpus.n_utt, dtype=torch.float64).pin_memory()
        self.side = torch.cuda.Stream()
        self.log_prob_ready = torch.cuda.Event()             # the per-utterance objectives have reached pinned memory
        self.last_fallback = 0
        # finite embeddings (a finite sum has no NaN / inf term) give finite or -inf band scores, so the DP may skip
        # its NaN compares (SEGB_DP_SCORES_FINITE); embeddings streamed from the host are not vouched for
        self.scores_finite = bool(torch.isfinite(c._X.sum(dtype=torch.float64)).item()) and np.isfinite(self.wip)
        self._streamed = False
        self._fused = fused
        self.mma = MmaScorer(c, fused=fused, precision="fp16" if precision == "fp16" else "fp8") if scorer == "mma" else None
        self.K_host = None                     # host copy of the active-component count (no .item() per sweep)
        # add_item's clamp and clean_components as device kernels (csrc/frozen.cu): no host logic per sweep
        self.clamp = NewComponentClamp(corpus, c.K_max)
        self.clean_work = torch.empty(lib.segb_kmeans_frozen_clean_work_bytes(c.K_max, c.D), dtype=torch.uint8,
                                      device=dev)

    # ---- phases (each is one or two launches; no host sync inside)
    def score(self, X_host=None):
        lib, c, sp = _lib.lib(), self.c, _lib.stream_ptr()
        m = c.struct()
        self._streamed = X_host is not None
        if X_host is not None:
            assert self.scorer == "mma", "streaming from host memory uses the tensor-core scorer"
            self.mma.score_streamed(X_host, self.best_val, self.best_k)
        elif self.scorer == "exact":
            _lib.check(lib.segb_kmeans_best(m, None, c.N, _lib.ptr(self.best_val), _lib.ptr(self.best_k), sp))
        else:
            self.mma.score(self.best_val, self.best_k)

    def segment(self):
        lib, c, cp, sp = _lib.lib(), self.c, self.corpus, _lib.stream_ptr()
        cs = cp.struct()
        _lib.check(lib.segb_kmeans_band_scores(c.struct(), cs, 0, cp.n_pos, _lib.ptr(self.best_val), self.wip,
                                               _lib.ptr(self.scores), sp))
        mode = _lib.DP_VITERBI_KMEANS | (_lib.DP_SCORES_FINITE if self.scores_finite and not self._streamed else 0)
        _lib.check(lib.segb_dp_banded(cs, 0, cp.n_utt, _lib.ptr(self.scores), mode, 0.0, 1.0,
                                      None, None, _lib.ptr(cp.bounds), _lib.ptr(self.log_prob), None, None,
                                      _lib.ptr(self.status), sp))

    def collect(self, assign_from=None):
        """Tokens of the current boundaries -> (sum_x, cnt); assignments[id] = k."""
        lib, c, cp, sp = _lib.lib(), self.c, self.corpus, _lib.stream_ptr()
        src = self.best_k if assign_from is None else assign_from
        if assign_from is None:
            c._assign.fill_(-1)
        self.sum_x.zero_()
        self.cnt.zero_()
        _lib.check(lib.segb_kmeans_collect(c.struct(), cp.struct(), 0, cp.n_utt, _lib.ptr(src), _lib.ptr(self.sum_x),
                                           _lib.ptr(self.cnt), sp))

    def summarize(self):
        """Start the end-of-sweep summary on a side stream, overlapping the token collection:
        the per-utterance objectives go to pinned host memory (their sum must be formed in
        utterance order -- the reference accumulates it one utterance at a time,
        kmeans_acoustic_wordseg.py:398-406 -- a serial float64 chain that a host core runs
        faster than one GPU thread: dependent DADDs cost ~24 cycles each on B200) and the DP statuses
        are counted on the device into self.flags[0]."""
        self.side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.side):
            self.flags[0:1].copy_((self.status != _lib.DP_OK).sum())
            if not _dist_on():
                self.log_prob_h.copy_(self.log_prob, non_blocking=True)
                self.log_prob_ready.record(self.side)

    def reduce_and_update(self):
        """All-reduce the sufficient statistics over ranks (NCCL over NVLink) -- ONE collective
        over the flat buffer [sum_x | counts] -- and rebuild the means."""
        lib, c, sp = _lib.lib(), self.c, _lib.stream_ptr()
        reduce_packed(self.red, self.cnt_f, self.cnt)        # counts ride along as float64 (exact below 2^53)
        _lib.check(lib.segb_kmeans_set_means(c.struct(), _lib.ptr(self.sum_x), _lib.ptr(self.cnt), sp))

    def init_means_from_assignments(self):
        """Build replicated means from the current (sharded) assignments: used once before
        the first distributed sweep."""
        c = self.c
        self.collect(assign_from=c._assign.clone())
        self.reduce_and_update()
        K = int((self.cnt > 0).sum().item())
        assert bool((self.cnt[:K] > 0).all().item()), "initial assignments must use labels 0..K-1"
        c._K.fill_(K)
        self.K_host = K

    def profile_phases(self):
        """Device time of each phase of one sweep (CUDA events on the launching stream); the
        model is advanced exactly as by sweep().  Diagnostic for bench.py / profiles/."""
        names, evs = [], [torch.cuda.Event(enable_timing=True)]
        evs[0].record()

        def mark(name):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            names.append(name)
            evs.append(e)
        c = self.c
        if self.scorer == "mma" and self.mma.fused:
            self.mma.pack_means()
            mark("pack_means")
            self.mma.fused_score(self.best_val, self.best_k)
            mark("score_fused(convert+filter_gemm+refine)")
        elif self.scorer == "mma":
            self.mma.pack_means()
            mark("pack_means")
            self.mma.filter()
            mark("filter_gemm")
            self.mma.refine(self.best_val, self.best_k)
            mark("refine_exact")
        else:
            self.score()
            mark("score_exact")
        self.segment()
        mark("band_scores+viterbi_dp")
        self.collect()
        mark("collect_tokens")
        self.reduce_and_update()
        mark("allreduce+set_means")
        torch.cuda.synchronize()
        return {n: evs[i].elapsed_time(evs[i + 1]) for i, n in enumerate(names)}

    def sweep(self, X_host=None):
        """One frozen sweep; returns sum_neg_len_sqrd_norm (summed over all ranks).  No host
        round trip until the end: one all-reduce, one small device->host copy, one sync.
        X_host: pinned float32 host copy of the embeddings to (re)upload while scoring
        (out-of-core / end-to-end use); None = the embeddings are already resident in HBM.
        Precondition of the bit-identical means: per-component sums of the embeddings are exact in
        float64 (true for float32 data of ordinary dynamic range; the sums are formed with atomics)."""
        c, cp = self.c, self.corpus
        if self.K_host is None:
            self.K_host = c.K
        K_before = self.K_host
        self.score(X_host)
        self.segment()
        self.summarize()
        if K_before < c.K_max:
            self._clamp_inactive_winners(K_before)
        self.collect()
        if _dist_on():
            # with several ranks the reference's single serial sum cannot be formed anyway: each rank contributes a
            # device-side sum of its utterances' objectives to the sweep's one all-reduce (no second collective, no
            # copy of the per-utterance values to the host)
            self.obj.copy_(self.log_prob.sum().reshape(1))
        self.reduce_and_update()
        self._clean_components()
        if self.scorer == "mma":
            self.flags[1:2].copy_(self.mma.n_fallback)
        self.flags[2:3].copy_(c._K)
        self.flags[3:4].copy_(self.obj.view(torch.int64))
        torch.cuda.current_stream().wait_stream(self.side)
        self.flags_h.copy_(self.flags, non_blocking=True)
        objective = None
        if not _dist_on():
            # objective: utterance-order float64 sum (the reference accumulates it one utterance at a time) -- a serial
            # chain of n_utt additions on the host, formed while the device still collects the tokens and rebuilds the means
            self.log_prob_ready.synchronize()
            objective = float(np.cumsum(self.log_prob_h.numpy())[-1]) if cp.n_utt else 0.0
        torch.cuda.current_stream().synchronize()
        n_bad, n_fb, K_now, obj_bits = (int(v) for v in self.flags_h.tolist())
        assert n_bad == 0, "segmentation failed for %d utterances (status %s)" % (
            n_bad, np.unique(self.status.cpu().numpy()))
        self.last_fallback = n_fb
        self.K_host = K_now
        if self.mma is not None and not self.mma.fused:
            self.mma.max_rounds = self.mma.rounds_for(2 * n_fb + 4096)
        if self.precision_mode == "auto" and self.mma is not None and not self.mma.fused and c.D % 2 == 0:
            switch_to = None
            if self.mma.fp8 and n_fb > self.AUTO_FP16_FRACTION * c.N:
                # the e4m3 pass decided too little: this model is served better by the fp16 first level -- for a while
                switch_to, self._auto_countdown = "fp16", self._auto_interval
                self._auto_interval = min(2 * self._auto_interval, 1 << 20)
            elif not self.mma.fp8:
                self._auto_countdown -= 1
                if self._auto_countdown <= 0:
                    switch_to = "fp8"                        # try the e4m3 level again
            elif self.mma.fp8 and n_fb <= self.AUTO_FP16_FRACTION * c.N:
                self._auto_interval = self.AUTO_RETRY_SWEEPS  # e4m3 works: a later fallback starts with a short interval
            if switch_to is not None:
                timing = self.mma.timing
                self.mma = None                             # release one image of X before the other one is built
                self.mma = MmaScorer(c, fused=self._fused, precision=switch_to)
                self.mma.timing = timing
        if _dist_on():
            return float(np.array([obj_bits], dtype=np.int64).view(np.float64)[0])
        return objective

    def fit(self, n_iter):
        """Frozen hard-assignment E-step + M-step over the CURRENT tokens -- KMeans.fit(n_iter,
        consider_unassigned=False) (kmeans.py:97-173) in sharded form: every rank re-assigns its own
        tokens against the same means (tensor-core scorer), one all-reduce of (sum_x, counts)
        rebuilds identical means everywhere, clean_components is replayed identically.  Stops when
        no token changed component on any rank.  Returns the reference's record keys."""
        c, cp = self.c, self.corpus
        if self.K_host is None:
            self.K_host = c.K
        record = {"components": [], "n_mean_updates": []}
        tok = cp.tok_id[cp.tok_id >= 0].long()
        for _ in range(n_iter):
            K_before = self.K_host
            self.score()
            changed = (self.best_k[tok] != c._assign[tok]).sum().to(torch.int64).reshape(1)
            if K_before < c.K_max:
                self._clamp_inactive_winners(K_before)
            self.collect()                                  # tokens of the unchanged boundaries, k = argmax
            self.reduce_and_update()
            self._clean_components()
            if _dist_on():
                dist.all_reduce(changed, op=dist.ReduceOp.SUM)
            n_changed, self.K_host = int(changed.item()), int(c._K.item())
            record["components"].append(self.K_host)
            record["n_mean_updates"].append(n_changed)
            if n_changed == 0:
                break
        return record

    def _clean_components(self):
        """clean_components() (kmeans_components.py:263-266) on the device: the swap-with-last deletions of
        the emptied components are replayed on the (global) counts by one thread, rows gathered and live
        tokens relabelled in parallel (segb_kmeans_frozen_clean); K stays on the device."""
        c, cp = self.c, self.corpus
        _lib.check(_lib.lib().segb_kmeans_frozen_clean(c.struct(), cp.struct(), 0, cp.n_pos, _lib.ptr(self.cnt),
                                                       _lib.ptr(self.clean_work), _lib.stream_ptr()))

    def _clamp_inactive_winners(self, K_before):
        """add_item's `k > K -> K` clamp (kmeans_components.py:103-106) for tokens won by an inactive slot:
        sequential in token order by nature.  Device version: the affected tokens are compacted in order,
        ranks exchange their lists with a fixed-size all-gather, ONE serial pass by a single warp resolves
        them, the winners are rewritten in best_k before the statistics are collected (NewComponentClamp)."""
        c, cp = self.c, self.corpus
        _lib.check(_lib.lib().segb_tokens_from_bounds(cp.struct(), 0, cp.n_utt, _lib.stream_ptr()))
        self.clamp.run(self.best_k, K_before)
        c._K.copy_(self.clamp.K_out)


# ---------------------------------------------------------------------------
# Frozen-state FBGMM sweep (fixed-variance components)
# ---------------------------------------------------------------------------

LSE_T = 20.0          # nats below the best component at which a component is dropped from log_marg_i's sum
                      # (what is dropped is at most K_max * exp(-20) = 1e-5 of the sum for K_max = 5000)


def _is_aniso(c):
    return not (np.all(c.precision == c.precision[0]) and np.all(c.precision_0 == c.precision_0[0]))


class FvScorer(object):
    """Buffers and call sequence of the tensor-core log_marg_i (segb_fvf_* / segb_fused_fv_log_marg): fp16
    model image + exact float64 row tables (packed per model state), float64 log marginals and MAP slots
    out.  fused=False (default): pre-packed fp16 tile image of X, filter GEMM, refine kernel (isotropic and
    anisotropic variances); fused=True (isotropic, even D <= 138): one kernel reads the fp32 embeddings once
    (see MmaScorer for the trade-off).
    keep_records: also keep the 16-byte thresholded row records (needed to DRAW components afterwards)."""

    def __init__(self, components, T=LSE_T, fused=None, keep_records=False, precision="fp16"):
        lib, c, dev = _lib.lib(), components, "cuda"
        assert c._X.dtype == torch.float32, "tensor-core log_marg needs float32 embeddings"
        assert precision in ("fp16", "fp8")
        self.c, self.T = c, float(T)
        self.aniso = int(_is_aniso(c))
        can_fuse = (not self.aniso) and fused_supported(c.D)
        self.fused = False if fused is None else (bool(fused) and can_fuse)
        u8 = torch.uint8
        self.w_tiles = torch.empty(lib.segb_fvf_w_tiles_bytes(c.K_max, c.D, self.aniso), dtype=u8, device=dev)
        self.model = torch.empty(lib.segb_fvf_model_bytes(c.K_max, c.D, self.aniso), dtype=u8, device=dev)
        self.work = torch.empty(lib.segb_fvf_work_bytes(c.N), dtype=u8, device=dev)
        self.w_max = torch.zeros(4, dtype=torch.float32, device=dev)
        self.n_fallback = torch.zeros(1, dtype=torch.int64, device=dev)
        self.log_marg = torch.empty(c.N, dtype=torch.float64, device=dev)
        self.map_k = torch.empty(c.N, dtype=torch.int32, device=dev)
        self.recs = torch.empty(16 * c.N, dtype=u8, device=dev) if keep_records else None
        self.x_tiles = self.cand = self.x_err = self.x_max = None
        # precision="fp8": e4m3 first level (segb_fvf8_*; isotropic variances, two-kernel path).  Rows it cannot decide
        # take the exhaustive scan, so callers watch n_fallback (FrozenFBGMMSweep falls back to fp16 by itself).
        self.fp8 = precision == "fp8" and not self.fused and not self.aniso and c.K_max <= 65536 - 16   # 12-bit chunk ids
        if self.fp8:
            self.sx = MmaScorer.pick_scale(c._X)
            n2_max = 0.0
            for i in range(0, c.N, 1 << 20):
                n2_max = max(n2_max, float((c._X[i:i + (1 << 20)].double() ** 2).sum(dim=1).max()))
            self.alpha = float(2.0 ** min(40.0, np.floor(np.log2(448.0 / max(n2_max, 1e-30)))))
            self.x_tiles = torch.empty(lib.segb_fvf8_x_tiles_bytes(c.N, c.D), dtype=u8, device=dev)
            self.w_tiles8 = torch.empty(lib.segb_fvf8_w_tiles_bytes(c.K_max, c.D), dtype=u8, device=dev)
            self.w_err8 = torch.empty(lib.segb_fvf8_w_err_bytes(c.K_max), dtype=u8, device=dev)
            self.w_max8 = torch.zeros(8, dtype=torch.float32, device=dev)
        elif not self.fused:
            self.x_tiles = torch.empty(lib.segb_fvf_x_tiles_bytes(c.N, c.D, self.aniso), dtype=u8, device=dev)
        if not self.fused:
            self.cand = torch.empty(lib.segb_mma_cand_bytes(c.N), dtype=u8, device=dev)
            self.x_err = torch.empty(2 * c.N, dtype=torch.float32, device=dev)
            self.x_max = torch.zeros(2, dtype=torch.float32, device=dev)
            self.pack_x()

    def pack_x(self):
        c = self.c
        if self.fp8:
            _lib.check(_lib.lib().segb_fvf8_pack_x(_lib.ptr(c._X), c.N, c.D, self.sx, self.alpha, _lib.ptr(self.x_tiles),
                                                   _lib.ptr(self.x_err), _lib.ptr(self.x_max), _lib.stream_ptr()))
            return
        _lib.check(_lib.lib().segb_fvf_pack_x(_lib.ptr(c._X), c.N, c.D, self.aniso, _lib.ptr(self.x_tiles),
                                              _lib.ptr(self.x_err), _lib.ptr(self.x_max), _lib.stream_ptr()))

    def pack_model(self):
        _lib.check(_lib.lib().segb_fvf_pack_model(self.c.struct(), self.aniso, _lib.ptr(self.w_tiles),
                                                  _lib.ptr(self.model), _lib.ptr(self.w_max), _lib.stream_ptr()))
        if self.fp8:
            c = self.c
            _lib.check(_lib.lib().segb_fvf8_pack_model(c.K_max, c.D, _lib.ptr(self.model), _lib.ptr(self.w_max), self.sx,
                                                       self.alpha, _lib.ptr(self.w_tiles8), _lib.ptr(self.w_err8),
                                                       _lib.ptr(self.w_max8), _lib.stream_ptr()))

    def filter(self):
        c = self.c
        if self.fp8:
            _lib.check(_lib.lib().segb_fvf8_filter(_lib.ptr(self.x_tiles), _lib.ptr(self.w_tiles8), c.N, c.K_max, c.D,
                                                   _lib.ptr(self.x_max), _lib.ptr(self.w_max8), self.sx, self.alpha, self.T,
                                                   _lib.ptr(self.cand), _lib.stream_ptr()))
            return
        _lib.check(_lib.lib().segb_fvf_filter(_lib.ptr(self.x_tiles), _lib.ptr(self.w_tiles), c.N, c.K_max, c.D,
                                              self.aniso, _lib.ptr(self.x_max), _lib.ptr(self.w_max), self.T,
                                              _lib.ptr(self.cand), _lib.stream_ptr()))

    def refine(self):
        c = self.c
        if self.fp8:
            _lib.check(_lib.lib().segb_fvf8_refine(_lib.ptr(c._X), c.N, c.D, c.K_max, _lib.ptr(self.model), _lib.ptr(self.cand),
                                                   _lib.ptr(self.x_err), _lib.ptr(self.w_max8), self.sx, self.alpha, self.T,
                                                   _lib.ptr(self.work), _lib.ptr(self.log_marg), _lib.ptr(self.map_k),
                                                   _lib.ptr(self.recs), _lib.ptr(self.n_fallback), _lib.stream_ptr()))
            return
        _lib.check(_lib.lib().segb_fvf_refine(_lib.ptr(c._X), c.N, c.D, c.K_max, self.aniso, _lib.ptr(self.model),
                                              _lib.ptr(self.cand), _lib.ptr(self.x_err), _lib.ptr(self.w_max), self.T,
                                              _lib.ptr(self.work), _lib.ptr(self.log_marg), _lib.ptr(self.map_k),
                                              _lib.ptr(self.recs), _lib.ptr(self.n_fallback), _lib.stream_ptr()))

    def fused_score(self):
        """The model image must be current (pack_model)."""
        c = self.c
        _lib.check(_lib.lib().segb_fused_fv_log_marg(_lib.ptr(c._X), c.N, c.D, c.K_max, _lib.ptr(self.w_tiles),
                                                     _lib.ptr(self.model), _lib.ptr(self.w_max), self.T,
                                                     _lib.ptr(self.work), _lib.ptr(self.log_marg), _lib.ptr(self.map_k),
                                                     _lib.ptr(self.recs), _lib.ptr(self.n_fallback), _lib.stream_ptr()))

    def score(self):
        """log_marg_i of every embedding (self.log_marg) and its MAP slot (self.map_k)."""
        self.pack_model()
        if self.fused:
            self.fused_score()
        else:
            self.filter()
            self.refine()


class FrozenFBGMMSweep(object):
    """One frozen-model sweep of the unigram FBGMM segmenter: every utterance is scored
    (get_vec_embed_log_probs, unigram_acoustic_wordseg.py:474-511) and segmented (forward_backward /
    forward_backward_viterbi, :653-864) against the SAME model, every new token picks its component from
    that model (gibbs_sample_inside_loop_i / map_assign_i without the add, fbgmm.py:422-494), then the
    model is rebuilt from the new assignments (FBGMM.setup_components, fbgmm.py:96-137).  Utterances (with
    their embeddings) shard over ranks; ONE all-reduce of [sum_x | counts] per sweep, like the k-means
    sweep.  Semantics pinned by the test oracle's frozen_fbgmm_sweep."""

    # precision of the scorer's first-level filter: "fp16", "fp8" (e4m3; isotropic variances) or "auto" = e4m3 until a
    # sweep leaves more than AUTO_FP16_FRACTION of the rows undecided (they take the exhaustive scan, which is slow: the
    # e4m3 pass only serves trained models), then fp16 for good.  Same results in every mode.
    AUTO_FP16_FRACTION = 0.002

    def __init__(self, components, corpus, fb_type="standard", time_power_term=1.0, wip=0.0, T=LSE_T, fused=None,
                 precision="auto"):
        self.c, self.corpus = components, corpus
        assert fb_type in ("standard", "viterbi")
        assert precision in ("auto", "fp16", "fp8")
        self.precision_mode = precision
        self.fb_type, self.tpt, self.wip = fb_type, float(time_power_term), float(wip)
        lib, c, cp, dev = _lib.lib(), components, corpus, "cuda"
        self.lms_range_ok = True
        if fb_type == "viterbi" and c._lms != 1.0:
            # map_assign_i ranks the slots WITHOUT lms (fbgmm.py:475-479) while the filter ranks them with it:
            # widen the threshold by the largest possible difference between the two rankings
            n_tok = max(1, int(cp.n_pos))
            T = T + abs(1.0 - c._lms) * float(np.log((c._alpha / c.K_max + n_tok) / (c._alpha / c.K_max)))
        self._fv_args = (T, fused, fb_type == "standard")
        self.fv = FvScorer(c, T, fused=fused, keep_records=(fb_type == "standard"),
                           precision="fp16" if precision == "fp16" else "fp8")
        self.scores = torch.empty(cp.n_pos * cp.S, dtype=torch.float64, device=dev)
        self.log_prob = torch.zeros(cp.n_utt, dtype=torch.float64, device=dev)
        self.status = torch.zeros(cp.n_utt, dtype=torch.int32, device=dev)
        self.choice = torch.full((c.N,), -1, dtype=torch.int32, device=dev)
        self.red = torch.zeros(c.K_max * c.D + c.K_max, dtype=torch.float64, device=dev)
        self.sum_x = self.red[:c.K_max * c.D].view(c.K_max, c.D)
        self.cnt_f = self.red[c.K_max * c.D:]
        self.cnt = torch.zeros(c.K_max, dtype=torch.int64, device=dev)
        self.new_label = torch.empty(c.K_max, dtype=torch.int32, device=dev)
        self.clamp = NewComponentClamp(cp, c.K_max)
        self.K_host = None
        self.last_fallback = 0

    def score(self):
        self.fv.score()

    def segment(self, u_fb):
        lib, cp, sp = _lib.lib(), self.corpus, _lib.stream_ptr()
        cs = cp.struct()
        _lib.check(lib.segb_fixedvar_band_scores(cs, 0, cp.n_pos, _lib.ptr(self.fv.log_marg), self.tpt, self.wip,
                                                 _lib.ptr(self.scores), sp))
        mode = _lib.DP_FFBS if self.fb_type == "standard" else _lib.DP_VITERBI_GMM
        _lib.check(lib.segb_dp_banded(cs, 0, cp.n_utt, _lib.ptr(self.scores), mode, 0.0, 1.0, _lib.ptr(u_fb), None,
                                      _lib.ptr(cp.bounds), _lib.ptr(self.log_prob), None, None,
                                      _lib.ptr(self.status), sp))
        _lib.check(lib.segb_tokens_from_bounds(cs, 0, cp.n_utt, sp))

    def choose(self, u_assign, K_before):
        lib, c, cp, fv, sp = _lib.lib(), self.c, self.corpus, self.fv, _lib.stream_ptr()
        mode = 0 if self.fb_type == "standard" else 1
        _lib.check(lib.segb_fvf_choose_tokens(_lib.ptr(c._X), c.D, c.K_max, int(K_before), fv.aniso, _lib.ptr(fv.model),
                                              _lib.ptr(fv.recs), cp.struct(), 0, cp.n_pos, mode, _lib.ptr(fv.map_k),
                                              _lib.ptr(u_assign), _lib.ptr(self.choice), sp))

    def collect(self):
        lib, c, cp, sp = _lib.lib(), self.c, self.corpus, _lib.stream_ptr()
        self.red.zero_()
        self.cnt.zero_()
        _lib.check(lib.segb_fixedvar_frozen_collect(c.struct(), cp.struct(), 0, cp.n_pos, _lib.ptr(self.choice),
                                                    _lib.ptr(self.sum_x), _lib.ptr(self.cnt), sp))

    def reduce_and_update(self):
        lib, c, cp, sp = _lib.lib(), self.c, self.corpus, _lib.stream_ptr()
        reduce_packed(self.red, self.cnt_f, self.cnt)
        c._assign.fill_(-1)
        _lib.check(lib.segb_fixedvar_frozen_update(c.struct(), cp.struct(), 0, cp.n_pos, _lib.ptr(self.choice),
                                                   _lib.ptr(self.sum_x), _lib.ptr(self.cnt), _lib.ptr(self.new_label), sp))

    def init_from_assignments(self):
        """Build the replicated model from the current (sharded) assignments of the current tokens
        (labels 0..K-1, all in use): used once before the first distributed sweep."""
        c, cp = self.c, self.corpus
        self.choice.copy_(c._assign)
        self.collect()
        self.reduce_and_update()
        self.K_host = int(c._K.item())

    def profile_phases(self, u_fb=None, u_assign=None):
        """Device time of each phase of one sweep (CUDA events on the launching stream); the model is
        advanced exactly as by sweep().  Diagnostic for bench.py / profiles/."""
        names, evs = [], [torch.cuda.Event(enable_timing=True)]
        evs[0].record()

        def mark(name):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            names.append(name)
            evs.append(e)
        c = self.c
        K_before = self.K_host if self.K_host is not None else c.K
        self.fv.pack_model()
        mark("pack_model")
        if self.fv.fused:
            self.fv.fused_score()
            mark("score_fused(convert+filter_gemm+refine)")
        else:
            self.fv.filter()
            mark("filter_gemm")
            self.fv.refine()
            mark("refine_exact")
        self.segment(u_fb)
        mark("band_scores+dp+tokens")
        self.choose(u_assign, K_before)
        if K_before < c.K_max:
            self.clamp.run(self.choice, K_before)
        mark("choose_components")
        self.collect()
        mark("collect_tokens")
        self.reduce_and_update()
        mark("allreduce+rebuild")
        torch.cuda.synchronize()
        self.K_host = int(c._K.item())
        return {n: evs[i].elapsed_time(evs[i + 1]) for i, n in enumerate(names)}

    def sweep(self, u_fb=None, u_assign=None):
        """One sweep.  u_fb / u_assign: float64 device arrays [n_pos] of uniforms (utterance u's i-th
        back-sampled segment reads u_fb[pos_off[u] + i]; the token ending at position p reads u_assign[p]);
        unused for fb_type "viterbi".  Returns the sum of the utterances' log_prob over all ranks."""
        c, cp = self.c, self.corpus
        if self.K_host is None:
            self.K_host = c.K
        K_before = self.K_host
        if self.fb_type == "standard":
            assert u_fb is not None and u_assign is not None and u_fb.numel() == cp.n_pos == u_assign.numel()
        self.score()
        self.segment(u_fb)
        self.choose(u_assign, K_before)
        if K_before < c.K_max:
            self.clamp.run(self.choice, K_before)
        self.collect()
        self.reduce_and_update()
        flags = torch.stack([(self.status != _lib.DP_OK).sum(), self.fv.n_fallback[0], c._K[0].to(torch.int64)])
        n_bad, n_fb, K_now = (int(v) for v in flags.tolist())            # the sweep's one host sync
        assert n_bad == 0, "segmentation failed for %d utterances (status %s)" % (
            n_bad, np.unique(self.status.cpu().numpy()))
        self.last_fallback, self.K_host = n_fb, K_now
        if self.precision_mode == "auto" and self.fv.fp8 and n_fb > self.AUTO_FP16_FRACTION * c.N:
            T_, fused_, keep_ = self._fv_args
            self.fv = None                                   # release the e4m3 image before the fp16 one is built
            self.fv = FvScorer(c, T_, fused=fused_, keep_records=keep_, precision="fp16")
        lp = self.log_prob.cpu().numpy()
        total = float(np.cumsum(lp)[-1]) if cp.n_utt else 0.0          # utterance-order float64 sum
        if _dist_on():
            t = torch.tensor([total], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            total = float(t.item())
        return total


class NewComponentClamp(object):
    """add_item's `k > K -> K` / "slot K opens a component" rule for a whole sweep on the device
    (segb_frozen_new_list + segb_frozen_clamp): ordered compaction of the affected tokens, a fixed-size
    all-gather across ranks (rank order = global token order), one serial pass, parallel write-back."""

    def __init__(self, corpus, K_max, cap=None):
        lib, dev = _lib.lib(), "cuda"
        self.corpus, self.K_max = corpus, int(K_max)
        self.cap = int(cap if cap is not None else max(1024, corpus.n_pos))
        i32 = torch.int32
        self.work = torch.empty(lib.segb_frozen_new_work_bytes(corpus.n_pos), dtype=torch.uint8, device=dev)
        self.list_j = torch.zeros(self.cap, dtype=i32, device=dev)
        self.list_id = torch.zeros(self.cap, dtype=i32, device=dev)
        self.n_list = torch.zeros(1, dtype=i32, device=dev)
        self.out_k = torch.zeros(self.cap, dtype=i32, device=dev)
        self.K_out = torch.zeros(1, dtype=i32, device=dev)
        self.overflow = torch.zeros(1, dtype=i32, device=dev)

    def run(self, choice, K_before):
        """Rewrites choice[id] of this rank's tokens in place; the new K is left in self.K_out (device)."""
        lib, cp, sp = _lib.lib(), self.corpus, _lib.stream_ptr()
        _lib.check(lib.segb_frozen_new_list(cp.struct(), 0, cp.n_pos, _lib.ptr(choice), int(K_before),
                                            _lib.ptr(self.work), self.cap, _lib.ptr(self.list_j),
                                            _lib.ptr(self.list_id), _lib.ptr(self.n_list), sp))
        if _dist_on():
            world, rank = dist.get_world_size(), dist.get_rank()
            if getattr(self, "_cap_all", None) is None:
                caps = torch.tensor([self.cap], dtype=torch.int64, device="cuda")
                dist.all_reduce(caps, op=dist.ReduceOp.MAX)              # ranks may own different numbers of positions
                self._cap_all = int(caps.item())                         # once: the corpus split is fixed
            cap_all = self._cap_all
            mine = torch.zeros(cap_all, dtype=torch.int32, device="cuda")
            mine[:self.cap] = self.list_j
            lists = torch.empty(world * cap_all, dtype=torch.int32, device="cuda")
            counts = torch.empty(world, dtype=torch.int32, device="cuda")
            dist.all_gather_into_tensor(lists, mine)
            dist.all_gather_into_tensor(counts, self.n_list)
        else:
            world, rank, cap_all, lists, counts = 1, 0, self.cap, self.list_j, self.n_list
        _lib.check(lib.segb_frozen_clamp(_lib.ptr(lists), _lib.ptr(counts), world, cap_all, rank, int(K_before),
                                         self.K_max, _lib.ptr(self.list_id), _lib.ptr(self.out_k), _lib.ptr(choice),
                                         _lib.ptr(self.K_out), _lib.ptr(self.overflow), sp))
