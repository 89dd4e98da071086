"""
Host-side ingestion (no GPU): segmentalist_b200.utterances.Utterances / process_embeddings against the oracle's
restatement of the reference's Utterances.__init__ and process_embeddings (utterances.py:74-157,
unigram_acoustic_wordseg.py:571-646) -- identical matrices, identical random initial boundaries from the same
np.random state (the C loop consumes the generator exactly as the reference's per-utterance redraws do),
identical accessors; and the banded index arithmetic DeviceCorpus.from_utterances uses.
"""
import numpy as np
import numpy.testing as npt
import pytest

from oracle import seg_oracle as so
from segmentalist_b200 import synth, utterances as ut


def _corpus(n_utt, seed, n_min=3, n_max=14, S=4, D=6, holes=True):
    mats, vids, durs, lms = synth.make_corpus_dicts(n_utt, D=D, K_true=5, n_min=n_min, n_max=n_max, n_slices_max=S,
                                                    noise=0.1, seed=seed)
    if holes:                                             # knock out some embeddings (-1 slots inside the band)
        rng = np.random.RandomState(seed + 1)
        for k in vids:
            v = vids[k]
            live = np.where(v >= 0)[0]
            v[live[rng.rand(len(live)) < 0.15]] = -1
    return mats, vids, durs, lms


@pytest.mark.parametrize("p_init,n_min,n_max,min_dur", [(0.5, 0, 4, 0), (0.3, 2, 4, 0), (0.5, 0, 3, 12), (0, 0, 4, 0),
                                                        (0.8, 0, 2, 0)])
def test_utterances_match_the_reference_restatement(p_init, n_min, n_max, min_dur):
    mats, vids, durs, lms = _corpus(120, seed=5)
    emb, vec_ids, labels = ut.process_embeddings(mats, vids)
    oemb, ovec_ids, olabels = so.process_embeddings(mats, vids)
    assert labels == olabels
    npt.assert_array_equal(emb, oemb)
    for a, b in zip(vec_ids, ovec_ids):
        npt.assert_array_equal(a, b)
    lengths = [len(lms[l]) for l in labels]
    args = (lengths, vec_ids, [durs[l] for l in labels], [lms[l] for l in labels])
    kw = dict(p_boundary_init=p_init, n_slices_min=n_min, n_slices_max=n_max, min_duration=min_dur)
    np.random.seed(7)
    u = ut.Utterances(*args, **kw)
    after = np.random.rand(3)
    np.random.seed(7)
    o = so.Utterances(*args, **kw)
    oafter = np.random.rand(3)
    npt.assert_array_equal(after, oafter)                 # the generator was advanced exactly as the reference advances it
    npt.assert_array_equal(np.asarray(u.boundaries), o.boundaries)
    npt.assert_array_equal(u.vec_ids, o.vec_ids)
    npt.assert_array_equal(np.isnan(u.durations), np.isnan(o.durations))
    npt.assert_array_equal(np.nan_to_num(u.durations, nan=-1.), np.nan_to_num(o.durations, nan=-1.))
    for i in range(0, u.D, 7):
        assert list(u.get_segmented_embeds_i(i)) == list(o.get_segmented_embeds_i(i))
        assert u.get_segmented_landmark_indices(i) == o.get_segmented_landmark_indices(i)
        npt.assert_array_equal(u.boundaries[i], o.boundaries[i])
        npt.assert_array_equal(u.boundaries[i, :lengths[i]], o.boundaries[i, :lengths[i]])
    allemb = u.all_segmented_embeds()
    ref = [e for i in range(u.D) for e in o.get_segmented_embeds_i(i)]
    npt.assert_array_equal(allemb, np.asarray(ref))
    # boundary matrix writes (what the sweeps do with the device result)
    new = np.asarray(u.boundaries).copy()
    new[:, 0] = True
    u.boundaries[:, :] = new
    for i in range(u.D):
        new[i, lengths[i]:] = False
    npt.assert_array_equal(np.asarray(u.boundaries), new)
    u.boundaries[3, :lengths[3]] = o.boundaries[3, :lengths[3]]
    npt.assert_array_equal(u.boundaries[3], o.boundaries[3])


def test_seed_boundaries_and_band_width():
    mats, vids, durs, lms = _corpus(40, seed=9, holes=False)
    emb, vec_ids, labels = ut.process_embeddings(mats, vids)
    lengths = [len(lms[l]) for l in labels]
    seeds = [[lms[l][0] + 1, lms[l][-1]] for l in labels]
    args = (lengths, vec_ids, [durs[l] for l in labels], [lms[l] for l in labels])
    u = ut.Utterances(*args, seed_boundaries=seeds, n_slices_max=4)
    o = so.Utterances(*args, seed_boundaries=seeds, n_slices_max=4)
    npt.assert_array_equal(np.asarray(u.boundaries), o.boundaries)
    assert ut.band_width_flat(u, 4) == ut.band_width(u.lengths, u.vec_ids, 4)
    assert ut.band_width_flat(u, 0) == ut.band_width(u.lengths, u.vec_ids, 0)
    assert ut.band_width_flat(u, 2) == 2


def test_band_arrays_equal_the_per_utterance_conversion():
    """The vectorised band construction of DeviceCorpus.from_utterances == packed_to_band per utterance."""
    mats, vids, durs, lms = _corpus(60, seed=11)
    emb, vec_ids, labels = ut.process_embeddings(mats, vids)
    lengths = [len(lms[l]) for l in labels]
    np.random.seed(3)
    u = ut.Utterances(lengths, vec_ids, [durs[l] for l in labels], [lms[l] for l in labels], n_slices_max=4)
    S = ut.band_width_flat(u, 4)
    seg_id, seg_dur = ut.band_arrays(u, S)
    ids = np.concatenate([ut.packed_to_band(u.vec_ids[i, :n * (n + 1) // 2], n, S, -1) for i, n in enumerate(lengths)])
    dur = np.concatenate([ut.packed_to_band(u.durations[i, :n * (n + 1) // 2], n, S, np.nan) for i, n in enumerate(lengths)])
    npt.assert_array_equal(seg_id, ids)
    npt.assert_array_equal(np.nan_to_num(seg_dur, nan=-1.), np.nan_to_num(dur, nan=-1.))
