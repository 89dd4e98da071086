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
    """Same constructor, attributes and accessors as the reference class
    (utterances.py:74-229).  Pure host logic (O(N) per utterance)."""

    def __init__(self, lengths, vec_ids, durations, landmarks, seed_boundaries=None,
                 p_boundary_init=0.5, n_slices_min=0, n_slices_max=6, min_duration=0):
        assert lengths == [len(i) for i in landmarks]
        self.lengths = lengths
        self.D = len(lengths)
        assert self.D == len(vec_ids)
        self.N_max = max(lengths)
        self.landmarks = landmarks
        width = self.N_max * (self.N_max + 1) // 2
        self.vec_ids = np.full((self.D, width), -1, dtype=np.int64)
        for u, ids in enumerate(vec_ids):
            self.vec_ids[u, :len(ids)] = ids
        self.durations = np.full((self.D, width), np.nan)
        for u, dur in enumerate(durations):
            if not (min_duration == 0 or len(dur) == 1):          # utterances.py:96-101
                cur = np.array(dur, dtype=np.float64)
                cur[cur < min_duration] = np.nan
                if np.all(np.isnan(cur)):
                    cur[np.argmax(dur)] = np.max(dur)
                dur = cur
            self.durations[u, :len(dur)] = dur
        self.boundaries = np.zeros((self.D, self.N_max), dtype=bool)
        if seed_boundaries is not None:                          # utterances.py:106-115
            for u, seeds in enumerate(seed_boundaries):
                marks = landmarks[u]
                hits = [int(np.argmin([abs(s - lm) for lm in marks])) for s in seeds]
                self.boundaries[u, hits] = True
        elif p_boundary_init == 0:                               # utterances.py:128-135
            for u in range(self.D):
                self.boundaries[u, self.lengths[u] - 1] = True
        else:                                                    # utterances.py:136-157
            for u in range(self.D):
                N = self.lengths[u]
                while True:
                    self.boundaries[u, 0:N] = (np.random.rand(N) < p_boundary_init)
                    self.boundaries[u, N - 1] = True
                    if np.all(np.asarray(self.get_segmented_embeds_i(u)) == -1):
                        continue
                    spans = [e - s for s, e in self.get_segmented_landmark_indices(u)]
                    if ((np.max(spans) <= n_slices_max and np.min(spans) >= n_slices_min)
                            or N <= n_slices_min):
                        break

    def get_segmented_landmark_indices(self, i):
        """(start, end) landmark index of every hypothesised word (utterances.py:199-208)."""
        out, start = [], 0
        for j in np.where(self.boundaries[i][:self.lengths[i]])[0]:
            out.append((start, int(j) + 1))
            start = int(j) + 1
        return out

    def get_segmented_embeds_i(self, i):
        """Embedding ids of the current segmentation (utterances.py:159-174)."""
        return [self.vec_ids[i, tri(e) + s] for s, e in self.get_segmented_landmark_indices(i)]

    def get_segmented_durations_i(self, i):
        """utterances.py:176-190."""
        return [self.durations[i, tri(e) + s] for s, e in self.get_segmented_landmark_indices(i)]

    def get_original_segmented_embeds_i(self, i):
        """utterances.py:192-204."""
        ids = self.vec_ids[i]
        lo = np.min(ids[np.where(ids != -1)])
        return list(self.get_segmented_embeds_i(i) - lo)

    def get_segmented_landmarks(self, i):
        """utterances.py:210-221."""
        assert self.landmarks is not None
        out, prev = [], 0
        for _, e in self.get_segmented_landmark_indices(i):
            out.append((prev, self.landmarks[i][e - 1]))
            prev = self.landmarks[i][e - 1]
        return out


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
        S = band_width(utts.lengths, utts.vec_ids, n_slices_max)
        ids, durs, bflat = [], [], []
        for u in range(utts.D):
            N = utts.lengths[u]
            n_packed = N * (N + 1) // 2
            ids.append(packed_to_band(utts.vec_ids[u, :n_packed], N, S, -1))
            durs.append(packed_to_band(utts.durations[u, :n_packed], N, S, np.nan))
            bflat.append(utts.boundaries[u, :N])
        return cls(utts.lengths, np.concatenate(ids), np.concatenate(durs), np.concatenate(bflat),
                   n_slices_min, n_slices_max, S)

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
        for u in range(self.n_utt):
            out[u, :self.lengths[u]] = b[self.pos_off_h[u]:self.pos_off_h[u + 1]]
        return out

    def set_boundaries_matrix(self, boundaries):
        flat = np.concatenate([boundaries[u, :self.lengths[u]] for u in range(self.n_utt)])
        self.bounds.copy_(torch.from_numpy(flat.astype(np.uint8)))
        self.refresh_tokens_from_bounds()


def process_embeddings(embedding_mats, vec_ids_dict):
    """Stack the per-utterance matrices in sorted-label order and rewrite the
    per-utterance row indices into global embedding ids
    (unigram_acoustic_wordseg.py:571-646).  Vectorised: one lookup table per
    utterance instead of one `np.where` per row."""
    labels = sorted(embedding_mats)
    mats, vec_ids, base = [], [], 0
    for utt in labels:
        mat = np.asarray(embedding_mats[utt])
        src = np.asarray(vec_ids_dict[utt])
        cur = src.copy()
        live = (src >= 0) & (src < mat.shape[0])
        cur[live] = src[live] + base
        mats.append(mat)
        vec_ids.append(cur)
        base += mat.shape[0]
    return np.concatenate(mats, axis=0), vec_ids, labels
