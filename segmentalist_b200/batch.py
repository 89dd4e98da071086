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
from .sharding import clamp_plan, compaction_plan, dist_on as _dist_on, reduce_packed


class MmaScorer(object):
    """Buffers and call sequence of the tensor-core scorer (segb_mma_*): fp16 tile images of
    the embeddings (packed once) and of the means (packed per sweep), the 32-byte per-row
    filter records, and the per-row / per-component rounding-error norms behind the rigorous
    candidate threshold."""

    def __init__(self, components):
        lib, c, dev = _lib.lib(), components, "cuda"
        assert c._X.dtype == torch.float32, "tensor-core scorer needs float32 embeddings"
        self.c = c
        self.x_tiles = torch.empty(lib.segb_mma_x_tiles_bytes(c.N, c.D), dtype=torch.uint8, device=dev)
        self.w_tiles = torch.empty(lib.segb_mma_w_tiles_bytes(c.K_max, c.D), dtype=torch.uint8, device=dev)
        self.cand = torch.empty(lib.segb_mma_cand_bytes(c.N), dtype=torch.uint8, device=dev)
        self.work = torch.empty(lib.segb_mma_refine_work_bytes(c.N, c.K_max), dtype=torch.uint8, device=dev)
        self.x_err = torch.empty(2 * c.N, dtype=torch.float32, device=dev)              # (|dx|, |x|) per row
        self.w_err = torch.empty(2 * (c.K_max + 128), dtype=torch.float32, device=dev)  # (|dmu|, |mu^|)
        self.x_max = torch.zeros(2, dtype=torch.float32, device=dev)
        self.w_max = torch.zeros(2, dtype=torch.float32, device=dev)
        self.n_fallback = torch.zeros(1, dtype=torch.int64, device=dev)
        self.pack_x()

    def pack_x(self):
        c = self.c
        _lib.check(_lib.lib().segb_mma_pack_x(_lib.ptr(c._X), c.N, c.D, _lib.ptr(self.x_tiles), _lib.ptr(self.x_err),
                                              _lib.ptr(self.x_max), _lib.stream_ptr()))

    def pack_means(self):
        c = self.c
        _lib.check(_lib.lib().segb_mma_pack_means(_lib.ptr(c._means), c.K_max, c.D, _lib.ptr(self.w_tiles),
                                                  _lib.ptr(self.w_err), _lib.ptr(self.w_max), _lib.stream_ptr()))

    def filter(self):
        c = self.c
        _lib.check(_lib.lib().segb_mma_filter(_lib.ptr(self.x_tiles), _lib.ptr(self.w_tiles), c.N, c.K_max, c.D,
                                              _lib.ptr(self.x_max), _lib.ptr(self.w_max), _lib.ptr(self.cand),
                                              _lib.stream_ptr()))

    def refine(self, best_val, best_k):
        c = self.c
        _lib.check(_lib.lib().segb_mma_refine(c.struct(), _lib.ptr(self.cand), _lib.ptr(self.x_err),
                                              _lib.ptr(self.w_max), c.N, _lib.ptr(self.work), _lib.ptr(best_val),
                                              _lib.ptr(best_k), _lib.ptr(self.n_fallback), _lib.stream_ptr()))

    def score(self, best_val, best_k):
        self.pack_means()
        self.filter()
        self.refine(best_val, best_k)

    def score_streamed(self, X_host, best_val, best_k, chunk_rows=1 << 20):
        """Score embeddings that still live in (pinned) HOST memory: the rows are uploaded in
        chunks on a copy stream while the previous chunk is packed to fp16 tiles, filtered and
        refined on the compute stream, so the sweep costs max(PCIe, compute) instead of their
        sum.  Chunks start on 256-row boundaries (whole operand tiles); every kernel is the
        same C-ABI call as in score(), handed pointers offset to the chunk."""
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
        kp2 = lib.segb_mma_x_tiles_bytes(256, c.D) // 256        # bytes of tile image per row
        rec = lib.segb_mma_cand_bytes(1)
        m = c.struct()
        x_base = c._X.data_ptr()
        vp = ctypes.c_void_p
        for lo in range(0, c.N, chunk_rows):
            hi = min(c.N, lo + chunk_rows)
            n = hi - lo
            with torch.cuda.stream(self.copy_stream):
                c._X[lo:hi].copy_(X_host[lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
            main.wait_event(ev)
            xt = vp(self.x_tiles.data_ptr() + lo * kp2)
            xe = vp(self.x_err.data_ptr() + 8 * lo)
            cd = vp(self.cand.data_ptr() + rec * lo)
            _lib.check(lib.segb_mma_pack_x(vp(x_base + 4 * c.D * lo), n, c.D, xt, xe, _lib.ptr(self.x_max), sp))
            _lib.check(lib.segb_mma_filter(xt, _lib.ptr(self.w_tiles), n, c.K_max, c.D, _lib.ptr(self.x_max),
                                           _lib.ptr(self.w_max), cd, sp))
            m.X, m.n_emb = x_base + 4 * c.D * lo, n
            _lib.check(lib.segb_mma_refine(m, cd, xe, _lib.ptr(self.w_max), n, _lib.ptr(self.work),
                                           vp(best_val.data_ptr() + 4 * lo), vp(best_k.data_ptr() + 4 * lo),
                                           _lib.ptr(self.n_fallback), sp))
            self.fb_total += self.n_fallback
        self.n_fallback.copy_(self.fb_total)


class FrozenKMeansSweep(object):

    def __init__(self, components, corpus, wip=0.0, scorer="auto"):
        self.c, self.corpus, self.wip = components, corpus, float(wip)
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
        self.red = torch.zeros(c.K_max * c.D + c.K_max, dtype=torch.float64, device=dev)
        self.sum_x = self.red[:c.K_max * c.D].view(c.K_max, c.D)
        self.cnt_f = self.red[c.K_max * c.D:]
        self.cnt = torch.zeros(c.K_max, dtype=torch.int64, device=dev)
        # end-of-sweep scalars read with ONE device->host copy: [bad DP statuses, fallback rows, emptied components]
        self.flags = torch.zeros(3, dtype=torch.int64, device=dev)
        self.flags_h = torch.zeros(3, dtype=torch.int64).pin_memory()
        self.log_prob_h = torch.zeros(corpus.n_utt, dtype=torch.float64).pin_memory()
        self.side = torch.cuda.Stream()
        self.last_fallback = 0
        self.mma = MmaScorer(c) if scorer == "mma" else None
        self.K_host = None                     # host copy of the active-component count (no .item() per sweep)

    # ---- phases (each is one or two launches; no host sync inside)
    def score(self, X_host=None):
        lib, c, sp = _lib.lib(), self.c, _lib.stream_ptr()
        m = c.struct()
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
        _lib.check(lib.segb_dp_banded(cs, 0, cp.n_utt, _lib.ptr(self.scores), _lib.DP_VITERBI_KMEANS, 0.0, 1.0,
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
            self.log_prob_h.copy_(self.log_prob, non_blocking=True)

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
        if self.scorer == "mma":
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
        (out-of-core / end-to-end use); None = the embeddings are already resident in HBM."""
        c, cp = self.c, self.corpus
        if self.K_host is None:
            self.K_host = c.K
        K_before = self.K_host
        self.score(X_host)
        self.segment()
        self.summarize()
        self.collect()
        if K_before < c.K_max:
            self._clamp_inactive_winners(K_before)
        self.reduce_and_update()
        K_now = self.K_host
        if self.scorer == "mma":
            self.flags[1:2].copy_(self.mma.n_fallback)
        self.flags[2:3].copy_((self.cnt[:K_now] == 0).sum())
        torch.cuda.current_stream().wait_stream(self.side)
        self.flags_h.copy_(self.flags, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        n_bad, n_fb, n_empty = (int(v) for v in self.flags_h.tolist())
        assert n_bad == 0, "segmentation failed for %d utterances (status %s)" % (
            n_bad, np.unique(self.status.cpu().numpy()))
        self.last_fallback = n_fb
        # objective: utterance-order float64 sum (the reference accumulates it one utterance at a time)
        total = float(np.cumsum(self.log_prob_h.numpy())[-1]) if cp.n_utt else 0.0
        if _dist_on():
            t = torch.tensor([total], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            total = float(t.item())
        if n_empty:
            self._clean_components(K_now)
        return total

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
            self.collect()                                  # tokens of the unchanged boundaries, k = argmax
            if K_before < c.K_max:
                self._clamp_inactive_winners(K_before)
            self.reduce_and_update()
            if _dist_on():
                dist.all_reduce(changed, op=dist.ReduceOp.SUM)
            K_now = self.K_host
            self.flags[2:3].copy_((self.cnt[:K_now] == 0).sum())
            n_changed, n_empty = int(changed.item()), int(self.flags[2].item())
            if n_empty:
                self._clean_components(K_now)
            record["components"].append(self.K_host)
            record["n_mean_updates"].append(n_changed)
            if n_changed == 0:
                break
        return record

    def _clean_components(self, K_old):
        """clean_components() (kmeans_components.py:263-266) without the per-deletion token
        scan: the swap-with-last sequence is replayed on the host over the K counts (cheap),
        giving for every surviving slot the component that ends up there; rows are then
        gathered and tokens relabelled in one pass each."""
        c = self.c
        cnt = self.cnt[:K_old].cpu().numpy()
        if not np.any(cnt == 0):
            return
        K, dst, src = compaction_plan(cnt, K_old)
        if len(dst):
            dst_d, src_d = _lib.dev(dst), _lib.dev(src)
            c._mean_num[dst_d] = c._mean_num[src_d]
            c._counts[dst_d] = c._counts[src_d]
            c._means[dst_d] = c._means[src_d]
            inv = torch.arange(c.K_max, dtype=torch.int32, device="cuda")
            inv[src_d] = dst_d.to(torch.int32)
            tok = self.corpus.tok_id[self.corpus.tok_id >= 0].long()
            c._assign[tok] = inv[c._assign[tok].long()]
        c._mean_num[K:K_old] = 0
        c._counts[K:K_old] = 0
        c._means[K:K_old] = c._rnd[K:K_old]
        c._meansT.copy_(c._means.t())
        c._K.fill_(int(K))
        self.K_host = int(K)

    def _clamp_inactive_winners(self, K_before):
        """add_item's `k > K -> K` clamp (kmeans_components.py:103-106) for tokens won by an
        inactive slot.  Sequential by nature; resolved on the host over the (few) affected
        tokens in utterance order, then the statistics of exactly those tokens are moved (sums of
        float32 embeddings are exact in float64, so subtracting and re-adding a row leaves the
        same bits as collecting again)."""
        c, cp = self.c, self.corpus
        flag = (self.cnt[K_before:] > 0).any().to(torch.int32).reshape(1)
        if _dist_on():
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if int(flag.item()) == 0:
            return
        tid = cp.tok_id
        k_at = torch.where(tid >= 0, c._assign[tid.clamp(min=0).long()], torch.full_like(tid, -1))
        pos = (k_at >= K_before).nonzero().flatten()            # landmark order == token order
        ids_d = tid[pos].long()
        ks_old_d = k_at[pos].long()
        ks = ks_old_d.cpu().numpy().astype(np.int64)
        if _dist_on():
            # ranks hold consecutive utterance ranges: rank order IS the global token order
            parts = [None] * dist.get_world_size()
            dist.all_gather_object(parts, ks.tolist())
            all_ks, K = clamp_plan([k for part in parts for k in part], K_before)
            lo = sum(len(part) for part in parts[:dist.get_rank()])
            ks_new = np.asarray(all_ks[lo:lo + len(ks)], dtype=np.int64)
        else:
            ks_new, K = clamp_plan(ks, K_before)
        c._K.fill_(int(K))
        self.K_host = int(K)
        if len(ks):
            ks_new_d = _lib.dev(np.asarray(ks_new, dtype=np.int64))
            rows = c._X[ids_d].to(torch.float64)
            self.sum_x.index_add_(0, ks_old_d, -rows)
            self.sum_x.index_add_(0, ks_new_d, rows)
            ones = torch.ones_like(ks_old_d)
            self.cnt.index_add_(0, ks_old_d, -ones)
            self.cnt.index_add_(0, ks_new_d, ones)
            c._assign[ids_d] = ks_new_d.to(torch.int32)
            self.best_k[ids_d] = ks_new_d.to(torch.int32)
