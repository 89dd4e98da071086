"""
Corpus / segmentation state: host-side mirror of the reference's `Utterances`
(segmentalist/utterances.py:14-229) plus the banded device layout the kernels
read.

The reference keeps `vec_ids` and `durations` packed-triangular per utterance
(O(N^2) slots, utterances.py:91-102).  On the device the same information is
banded: row = landmark position, column = span-1, S columns (longest span that
carries an embedding), so an utterance costs N*S slots and a warp walks it with
unit stride.  See include/segb200.h (segb_corpus).
"""
import numpy as np
import torch

from . import _lib


def tri(t):
    """Offset of the block of segments ending at landmark t in the packed layout."""
    return t * (t - 1) // 2


class Utterances(object):
    """Same constructor, attributes and accessors as the reference class (utterances.py:74-229).

    Storage is ragged-flat: the packed-triangular `vec_ids` / `durations` of all utterances back to back
    (O(sum N^2) but without padding every utterance to N_max(N_max+1)/2 slots, and built with a handful of
    NumPy calls instead of a Python loop per utterance); the reference's padded [D, N_max(N_max+1)/2]
    matrices are materialised only if somebody reads the `vec_ids` / `durations` attributes.  The device
    layout (DeviceCorpus) is built straight from the flat arrays."""

    def __init__(self, lengths, vec_ids, durations, landmarks, seed_boundaries=None,
                 p_boundary_init=0.5, n_slices_min=0, n_slices_max=6, min_duration=0):
        assert lengths == [len(i) for i in landmarks]
        self.lengths = lengths
        self.D = len(lengths)
        assert self.D == len(vec_ids)
        self.N_max = max(lengths)
        self.landmarks = landmarks
        self._len = np.asarray(lengths, dtype=np.int64)
        self._pos_off = np.concatenate([[0], np.cumsum(self._len)]).astype(np.int64)
        n_packed = np.fromiter((len(v) for v in vec_ids), dtype=np.int64, count=self.D)
        assert np.all(n_packed <= self._len * (self._len + 1) // 2)
        self._poff = np.concatenate([[0], np.cumsum(n_packed)]).astype(np.int64)
        self._ids = (np.concatenate([np.asarray(v, dtype=np.int64) for v in vec_ids])
                     if self.D else np.zeros(0, np.int64))
        assert len(durations) == self.D
        dur = np.concatenate([np.asarray(d, dtype=np.float64) for d in durations]) if self.D else np.zeros(0)
        assert dur.shape == self._ids.shape, "vec_ids and durations must have the same packed lengths"
        if min_duration != 0:                                     # utterances.py:96-101
            raw = dur.copy()
            multi = np.repeat(n_packed != 1, n_packed)            # single-slot utterances are left alone
            dur = np.where(multi & (raw < min_duration), np.nan, raw)
            all_nan = np.logical_and.reduceat(np.isnan(dur), self._poff[:-1][n_packed > 0])
            for u in np.where(n_packed > 0)[0][all_nan]:          # rare: keep the longest candidate
                lo, hi = self._poff[u], self._poff[u + 1]
                k = int(np.argmax(raw[lo:hi]))
                dur[lo + k] = raw[lo + k]
        self._dur = dur
        self._vec_ids_mat = self._dur_mat = None
        self._bflat = np.zeros(int(self._pos_off[-1]), dtype=bool)
        self.boundaries = _BoundaryMatrix(self)
        if seed_boundaries is not None:                          # utterances.py:106-115
            for u, seeds in enumerate(seed_boundaries):
                marks = np.asarray(landmarks[u])
                hits = [int(np.argmin(np.abs(s - marks))) for s in seeds]
                self._bflat[self._pos_off[u] + np.asarray(hits, dtype=np.int64)] = True
        elif p_boundary_init == 0:                               # utterances.py:128-135
            self._bflat[self._pos_off[1:] - 1] = True
        else:                                                    # utterances.py:136-157
            self._init_random_boundaries(p_boundary_init, n_slices_min, n_slices_max)

    def _init_random_boundaries(self, p, n_slices_min, n_slices_max):
        """The reference redraws `np.random.rand(N) < p` per utterance until the segmentation is usable; the
        loop runs in C over a block of uniforms drawn in advance, and the global NumPy generator is then
        advanced by exactly the number of draws the reference would have made."""
        lib = _lib.load()
        n_pos = int(self._pos_off[-1])
        state = np.random.get_state()
        block = max(1024, 2 * n_pos)
        out = np.zeros(n_pos, dtype=np.uint8)
        ptr = lambda a: a.ctypes.data
        while True:
            np.random.set_state(state)
            uni = np.random.rand(block)
            used = lib.segb_host_init_boundaries(ptr(self._len), self.D, ptr(self._poff), ptr(self._ids),
                                                 ptr(self._pos_off), ptr(uni), block, float(p), int(n_slices_min),
                                                 int(n_slices_max), ptr(out))
            if used >= 0:
                break
            block *= 2
        np.random.set_state(state)
        if used:
            np.random.rand(int(used))
        self._bflat[:] = out.astype(bool)

    # ---- the reference's padded matrices, on demand
    def _padded(self, flat, fill, dtype):
        width = self.N_max * (self.N_max + 1) // 2
        out = np.full((self.D, width), fill, dtype=dtype)
        n_packed = np.diff(self._poff)
        rows = np.repeat(np.arange(self.D), n_packed)
        cols = np.arange(len(flat)) - np.repeat(self._poff[:-1], n_packed)
        out[rows, cols] = flat
        return out

    @property
    def vec_ids(self):
        if self._vec_ids_mat is None:
            self._vec_ids_mat = self._padded(self._ids, -1, np.int64)
        return self._vec_ids_mat

    @property
    def durations(self):
        if self._dur_mat is None:
            self._dur_mat = self._padded(self._dur, np.nan, np.float64)
        return self._dur_mat

    def _packed_at(self, flat, u, k, fill):
        k = np.asarray(k, dtype=np.int64)
        n = self._poff[u + 1] - self._poff[u]
        out = np.full(k.shape, fill, dtype=flat.dtype)
        ok = k < n
        out[ok] = flat[self._poff[u] + k[ok]]
        return out

    def all_segmented_embeds(self):
        """get_segmented_embeds_i for every utterance in order, concatenated (vectorised)."""
        idx = np.where(self._bflat)[0]
        if len(idx) == 0:
            return np.zeros(0, np.int64)
        utt = np.searchsorted(self._pos_off, idx, "right") - 1
        first = np.concatenate([[True], utt[1:] != utt[:-1]])
        start = np.where(first, self._pos_off[utt], np.concatenate([[0], idx[:-1] + 1])) - self._pos_off[utt]
        end = idx - self._pos_off[utt] + 1
        k = end * (end - 1) // 2 + start
        n = (self._poff[1:] - self._poff[:-1])[utt]
        out = np.full(len(idx), -1, dtype=np.int64)
        ok = k < n
        out[ok] = self._ids[self._poff[utt[ok]] + k[ok]]
        return out

    def get_segmented_landmark_indices(self, i):
        """(start, end) landmark index of every hypothesised word (utterances.py:199-208)."""
        out, start = [], 0
        for j in np.where(self.boundaries[i][:self.lengths[i]])[0]:
            out.append((start, int(j) + 1))
            start = int(j) + 1
        return out

    def get_segmented_embeds_i(self, i):
        """Embedding ids of the current segmentation (utterances.py:159-174)."""
        seg = self.get_segmented_landmark_indices(i)
        return list(self._packed_at(self._ids, i, [tri(e) + s for s, e in seg], -1))

    def get_segmented_durations_i(self, i):
        """utterances.py:176-190."""
        seg = self.get_segmented_landmark_indices(i)
        return list(self._packed_at(self._dur, i, [tri(e) + s for s, e in seg], np.nan))

    def get_original_segmented_embeds_i(self, i):
        """utterances.py:192-204."""
        ids = self._ids[self._poff[i]:self._poff[i + 1]]
        lo = np.min(ids[np.where(ids != -1)])
        return list(np.asarray(self.get_segmented_embeds_i(i)) - lo)

    def get_segmented_landmarks(self, i):
        """utterances.py:210-221."""
        assert self.landmarks is not None
        out, prev = [], 0
        for _, e in self.get_segmented_landmark_indices(i):
            out.append((prev, self.landmarks[i][e - 1]))
            prev = self.landmarks[i][e - 1]
        return out

    def __getstate__(self):
        d = dict(self.__dict__)
        d["boundaries"] = None
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self.boundaries = _BoundaryMatrix(self)


class _BoundaryMatrix(object):
    """`Utterances.boundaries` with the reference's [D, N_max] indexing over the flat per-landmark storage:
    `boundaries[i]`, `boundaries[i, j]`, `boundaries[i, :N]`, `boundaries[:, :] = matrix`, np.asarray(.)."""

    def __init__(self, utts):
        self._u = utts

    @property
    def shape(self):
        return (self._u.D, self._u.N_max)

    def _dense(self):
        u = self._u
        out = np.zeros((u.D, u.N_max), dtype=bool)
        rows = np.repeat(np.arange(u.D), u._len)
        cols = np.arange(len(u._bflat)) - np.repeat(u._pos_off[:-1], u._len)
        out[rows, cols] = u._bflat
        return out

    def __array__(self, dtype=None, copy=None):
        d = self._dense()
        return d if dtype is None else d.astype(dtype)

    def copy(self):
        return self._dense()

    def __eq__(self, other):
        return self._dense() == np.asarray(other)

    def __getitem__(self, key):
        u = self._u
        if isinstance(key, (int, np.integer)):
            row = np.zeros(u.N_max, dtype=bool)
            row[:u._len[key]] = u._bflat[u._pos_off[key]:u._pos_off[key + 1]]
            return row
        if isinstance(key, tuple) and isinstance(key[0], (int, np.integer)):
            return self[key[0]][key[1]]
        return self._dense()[key]

    def __setitem__(self, key, value):
        u = self._u
        if isinstance(key, tuple) and isinstance(key[0], (int, np.integer)):
            i = int(key[0])
            row = self[i]
            row[key[1]] = value
            u._bflat[u._pos_off[i]:u._pos_off[i + 1]] = row[:u._len[i]]
            return
        if isinstance(key, (int, np.integer)):
            row = np.zeros(u.N_max, dtype=bool)
            row[:] = value
            u._bflat[u._pos_off[key]:u._pos_off[key + 1]] = row[:u._len[key]]
            return
        d = self._dense()
        d[key] = value
        rows = np.repeat(np.arange(u.D), u._len)
        cols = np.arange(len(u._bflat)) - np.repeat(u._pos_off[:-1], u._len)
        u._bflat[:] = d[rows, cols]


# ---------------------------------------------------------------------------
# banded device layout
# ---------------------------------------------------------------------------

_BAND_CACHE = {}


def _band_index(N, S):
    """(rows, cols, packed) index triplets mapping packed slots of an N-landmark
    utterance into its [N, S] band."""
    key = (N, S)
    if key not in _BAND_CACHE:
        t = np.repeat(np.arange(1, N + 1), S)
        l = np.tile(np.arange(1, S + 1), N)
        ok = l <= t
        t, l = t[ok], l[ok]
        _BAND_CACHE[key] = (t - 1, l - 1, tri(t) + (t - l))
    return _BAND_CACHE[key]


def band_width(lengths, vec_ids_rows, n_slices_max):
    """Longest span that carries an embedding, limited by n_slices_max (0 = no limit)."""
    S_data = 1
    for N, ids in zip(lengths, vec_ids_rows):
        t = np.repeat(np.arange(1, N + 1), np.arange(1, N + 1))      # end landmark of every packed slot
        j = np.arange(len(t)) - tri(t)                               # start landmark
        live = np.asarray(ids[:len(t)]) != -1
        if live.any():
            S_data = max(S_data, int((t - j)[live].max()))
    N_max = max(lengths)
    S = min(S_data, N_max)
    if n_slices_max and n_slices_max > 0:
        S = min(S, n_slices_max)
    return max(S, 1)


def band_width_flat(utts, n_slices_max):
    """band_width over the flat packed storage of an Utterances object (vectorised)."""
    n_packed = np.diff(utts._poff)
    k = np.arange(len(utts._ids)) - np.repeat(utts._poff[:-1], n_packed)       # packed slot within the utterance
    t = ((np.sqrt(8.0 * k + 1.0) + 1.0) // 2).astype(np.int64)                 # end landmark: tri(t) <= k < tri(t+1)
    t = np.where(tri(t) > k, t - 1, t)
    t = np.where(tri(t + 1) <= k, t + 1, t)
    span = t - (k - tri(t))
    live = utts._ids != -1
    S_data = int(span[live].max()) if live.any() else 1
    S = min(max(S_data, 1), utts.N_max)
    if n_slices_max and n_slices_max > 0:
        S = min(S, n_slices_max)
    return max(S, 1)


def band_arrays(utts, S):
    """(seg_id [n_pos, S] int32, seg_dur [n_pos, S] float64) of an Utterances object: band slot (landmark
    position p of utterance u, span l) <- packed slot tri(t) + (t - l) of that utterance, t = p - pos_off[u] + 1."""
    n_pos = int(utts._pos_off[-1])
    utt = np.repeat(np.arange(utts.D), utts._len)
    t = (np.arange(n_pos) - utts._pos_off[utt] + 1)[:, None]              # [n_pos, 1]
    l = np.arange(1, S + 1)[None, :]                                      # [1, S]
    k = tri(t) + (t - l)
    n_packed = (utts._poff[1:] - utts._poff[:-1])[utt][:, None]
    ok = (l <= t) & (k < n_packed)
    src = np.where(ok, utts._poff[utt][:, None] + k, 0)
    seg_id = np.where(ok, utts._ids[src], -1).astype(np.int32)
    seg_dur = np.where(ok, utts._dur[src], np.nan)
    return seg_id, seg_dur


def packed_to_band(vec, N, S, fill):
    """Packed-triangular vector of one utterance -> [N, S] band."""
    rows, cols, packed = _band_index(N, S)
    out = np.full((N, S), fill, dtype=np.asarray(vec).dtype)
    out[rows, cols] = np.asarray(vec)[packed]
    return out


def band_to_packed(band, N, S, fill):
    rows, cols, packed = _band_index(N, S)
    out = np.full(N * (N + 1) // 2, fill, dtype=band.dtype)
    out[packed] = band[rows, cols]
    return out


class DeviceCorpus(object):
    """Banded device arrays for a set of utterances (segb_corpus)."""

    def __init__(self, lengths, seg_id, seg_dur, boundaries_flat, n_slices_min, n_slices_max, S):
        self.lengths = np.asarray(lengths, dtype=np.int64)
        self.n_utt = len(self.lengths)
        self.S = int(S)
        self.N_max = int(self.lengths.max())
        self.n_slices_min, self.n_slices_max = int(n_slices_min), int(n_slices_max)
        self.pos_off_h = np.concatenate([[0], np.cumsum(self.lengths)]).astype(np.int64)
        self.n_pos = int(self.pos_off_h[-1])
        assert seg_id.shape == (self.n_pos, self.S) and seg_dur.shape == (self.n_pos, self.S)
        self.pos_off = _lib.dev(self.pos_off_h)
        self.seg_id = _lib.dev(seg_id.astype(np.int32))
        self.seg_dur = _lib.dev(seg_dur.astype(np.float64))
        self.bounds = _lib.dev(np.asarray(boundaries_flat, dtype=np.uint8))
        self.tok_id = torch.full((self.n_pos,), -1, dtype=torch.int32, device="cuda")
        self.refresh_tokens_from_bounds()

    @classmethod
    def from_utterances(cls, utts, n_slices_min, n_slices_max):
        """Banded arrays straight from the flat packed storage (no padded-triangular intermediate, no
        per-utterance loop): band slot (landmark position p of utterance u, span l) <- packed slot
        tri(t) + (t - l) of that utterance, t = p - pos_off[u] + 1."""
        S = band_width_flat(utts, n_slices_max)
        seg_id, seg_dur = band_arrays(utts, S)
        return cls(utts.lengths, seg_id, seg_dur, utts._bflat, n_slices_min, n_slices_max, S)

    def struct(self):
        c = _lib.Corpus()
        c.n_utt, c.S, c.N_max = self.n_utt, self.S, self.N_max
        c.n_slices_min, c.n_slices_max, c.n_pos = self.n_slices_min, self.n_slices_max, self.n_pos
        c.pos_off = self.pos_off.data_ptr()
        c.seg_id = self.seg_id.data_ptr()
        c.seg_dur = self.seg_dur.data_ptr()
        c.bounds = self.bounds.data_ptr()
        c.tok_id = self.tok_id.data_ptr()
        return c

    def refresh_tokens_from_bounds(self):
        """tok_id[p] = embedding id of the token ending at landmark p (vectorised host pass;
        used at initialisation and after the host edits `boundaries`)."""
        b = self.bounds.cpu().numpy().astype(bool)
        seg = self.seg_id.cpu().numpy()
        tok = np.full(self.n_pos, -1, dtype=np.int32)
        idx = np.where(b)[0]
        if len(idx):
            utt = np.searchsorted(self.pos_off_h, idx, "right") - 1
            first = np.concatenate([[True], utt[1:] != utt[:-1]])
            start = np.where(first, self.pos_off_h[utt], np.concatenate([[0], idx[:-1] + 1]))
            span = idx - start + 1
            ok = span <= self.S
            tok[idx[ok]] = seg[idx[ok], span[ok] - 1]
        self.tok_id.copy_(torch.from_numpy(tok))

    def boundaries_matrix(self):
        """[n_utt, N_max] bool matrix in the reference's `Utterances.boundaries` format."""
        b = self.bounds.cpu().numpy().astype(bool)
        out = np.zeros((self.n_utt, self.N_max), dtype=bool)
        rows = np.repeat(np.arange(self.n_utt), self.lengths)
        cols = np.arange(self.n_pos) - np.repeat(self.pos_off_h[:-1], self.lengths)
        out[rows, cols] = b
        return out

    def boundaries_flat(self):
        """The current boundaries, one bool per landmark (utterances back to back)."""
        return self.bounds.cpu().numpy().astype(bool)

    def set_boundaries_matrix(self, boundaries):
        boundaries = np.asarray(boundaries)
        rows = np.repeat(np.arange(self.n_utt), self.lengths)
        cols = np.arange(self.n_pos) - np.repeat(self.pos_off_h[:-1], self.lengths)
        self.bounds.copy_(torch.from_numpy(boundaries[rows, cols].astype(np.uint8)))
        self.refresh_tokens_from_bounds()


def process_embeddings(embedding_mats, vec_ids_dict):
    """Stack the per-utterance matrices in sorted-label order and rewrite the per-utterance row indices
    into global embedding ids (unigram_acoustic_wordseg.py:571-646).  Vectorised over the whole corpus: one
    concatenation and one offset addition instead of an `np.where` per matrix row."""
    labels = sorted(embedding_mats)
    mats = [np.asarray(embedding_mats[u]) for u in labels]
    srcs = [np.asarray(vec_ids_dict[u]) for u in labels]
    n_rows = np.fromiter((m.shape[0] for m in mats), dtype=np.int64, count=len(mats))
    n_slots = np.fromiter((len(v) for v in srcs), dtype=np.int64, count=len(srcs))
    base = np.concatenate([[0], np.cumsum(n_rows)[:-1]]) if len(mats) else np.zeros(0, np.int64)
    src = np.concatenate(srcs) if srcs else np.zeros(0, np.int64)
    live = (src >= 0) & (src < np.repeat(n_rows, n_slots))
    cur = np.where(live, src + np.repeat(base, n_slots), src)
    vec_ids = np.split(cur, np.cumsum(n_slots)[:-1]) if len(srcs) else []
    return np.concatenate(mats, axis=0), vec_ids, labels
