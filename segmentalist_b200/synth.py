"""
Synthetic corpora in the reference's input format.

The reference segmenters take four dicts keyed by utterance label
(unigram_acoustic_wordseg.py:118-125): `embedding_mats` (one [n_seg, D] matrix
per utterance), `vec_ids_dict` (packed-triangular map slot -> matrix row, -1 =
no embedding), `durations_dict` (same shape, frames) and `landmarks_dict`.
The packed layout is the one built in
segmentalist/tests/test_unigram_acoustic_wordseg.py:35-46: slot t(t-1)/2 + j
holds the segment covering landmarks j..t-1, rows numbered start-major.

Two generators:
  make_corpus_dicts  -- reference-format dicts (small/medium cases, tests, CPU baseline)
  make_corpus_flat   -- the same corpus, vectorised, directly as flat arrays
                        (bench sizes: 200k utterances / 21M embeddings)
"""
import numpy as np


def _unit_rows(a):
    n = np.sqrt((a.astype(np.float64) ** 2).sum(axis=1, keepdims=True))
    return (a / n).astype(np.float32)


def cluster_centres(K_true, D, rng):
    return _unit_rows(rng.standard_normal((K_true, D)))


def utterance_layout(N, n_slices_max):
    """Packed slot, start, end for every candidate segment of an N-landmark
    utterance, in the reference's row order (start-major)."""
    starts, ends = [], []
    for s in range(N):
        for e in range(s + 1, min(N, s + n_slices_max) + 1):
            starts.append(s)
            ends.append(e)
    starts = np.asarray(starts)
    ends = np.asarray(ends)
    slots = ends * (ends - 1) // 2 + starts
    return slots, starts, ends


def make_corpus_dicts(n_utt, D=130, K_true=200, n_min=15, n_max=25, n_slices_max=6,
                      noise=0.05, seed=0, gap_lo=3, gap_hi=14):
    """Reference-format dicts.  x = normalise(c_z + noise*N(0,I)), one latent
    word z per candidate segment."""
    rng = np.random.RandomState(seed)
    centres = cluster_centres(K_true, D, rng)
    embedding_mats, vec_ids_dict, durations_dict, landmarks_dict = {}, {}, {}, {}
    for u in range(n_utt):
        label = "utt%07d" % u
        N = int(rng.randint(n_min, n_max + 1))
        gaps = rng.randint(gap_lo, gap_hi + 1, size=N)
        landmarks = np.cumsum(gaps)
        slots, starts, ends = utterance_layout(N, n_slices_max)
        n_seg = len(slots)
        z = rng.randint(0, K_true, size=n_seg)
        X = _unit_rows(centres[z] + noise * rng.standard_normal((n_seg, D)).astype(np.float32))
        vec_ids = -1 * np.ones(N * (N + 1) // 2, dtype=np.int64)
        vec_ids[slots] = np.arange(n_seg)
        bounds = np.concatenate([[0], landmarks])
        durations = -1 * np.ones(N * (N + 1) // 2, dtype=np.int64)
        durations[slots] = bounds[ends] - bounds[starts]
        embedding_mats[label] = X
        vec_ids_dict[label] = vec_ids
        durations_dict[label] = durations
        landmarks_dict[label] = [int(v) for v in landmarks]
    return embedding_mats, vec_ids_dict, durations_dict, landmarks_dict


def make_corpus_flat(n_utt, D=130, K_true=200, n_min=15, n_max=25, n_slices_max=6,
                     noise=0.05, seed=0, gap_lo=3, gap_hi=14, chunk=4096):
    """Banded/flat arrays for bench-size corpora, generated chunk-wise.

    Returns dict with
      X          float32 [n_emb, D]   embeddings, rows in (utterance, start-major) order
      lengths    int32   [n_utt]      landmarks per utterance
      seg_id     int32   [sum N, S]   banded map: row (off[u]+t-1), col l-1 -> embedding id of
                                      segment [t-l, t), -1 if absent
      seg_dur    float64 [sum N, S]   duration in frames (NaN where absent)
    """
    rng = np.random.RandomState(seed)
    S = n_slices_max
    centres = cluster_centres(K_true, D, rng)
    lengths = rng.randint(n_min, n_max + 1, size=n_utt).astype(np.int32)
    pos_off = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    n_pos = int(pos_off[-1])
    seg_id = -1 * np.ones((n_pos, S), dtype=np.int32)
    seg_dur = np.full((n_pos, S), np.nan)
    # embeddings per utterance: sum_{s} min(S, N-s)
    n_seg_u = np.array([sum(min(S, int(N) - s) for s in range(int(N))) for N in range(n_max + 1)])
    emb_off = np.concatenate([[0], np.cumsum(n_seg_u[lengths])]).astype(np.int64)
    n_emb = int(emb_off[-1])
    X = np.empty((n_emb, D), dtype=np.float32)
    layouts = {}
    for u in range(n_utt):
        N = int(lengths[u])
        if N not in layouts:
            layouts[N] = utterance_layout(N, S)[1:]
        starts, ends = layouts[N]
        gaps = rng.randint(gap_lo, gap_hi + 1, size=N)
        b = np.concatenate([[0], np.cumsum(gaps)])
        rows = pos_off[u] + ends - 1
        cols = ends - starts - 1
        seg_id[rows, cols] = emb_off[u] + np.arange(len(starts))
        seg_dur[rows, cols] = b[ends] - b[starts]
    for lo in range(0, n_emb, chunk * 128):
        hi = min(n_emb, lo + chunk * 128)
        z = rng.randint(0, K_true, size=hi - lo)
        X[lo:hi] = _unit_rows(centres[z] + noise * rng.standard_normal((hi - lo, D)).astype(np.float32))
    return {"X": X, "lengths": lengths, "seg_id": seg_id, "seg_dur": seg_dur,
            "pos_off": pos_off, "emb_off": emb_off, "centres": centres}
