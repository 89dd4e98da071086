"""Host-side check of the arithmetic behind the e4m3 filter's packed top-3 keys (csrc/mma_common.cuh:
top3_insert_key, key_unpack, C_KEY8): a NumPy restatement of the bit manipulation, checked for the properties the
rigorous bound relies on.  The device code itself is exercised by the -m gpu parity tests."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = open(os.path.join(ROOT, "segmentalist_b200", "csrc", "mma_common.cuh")).read()


def _const(name):
    m = re.search(r"constexpr\s+\w+\s+%s\s*=\s*([^;]+);" % name, HDR)
    assert m, name
    return m.group(1)


KEY_ID_BITS = int(_const("KEY_ID_BITS"))
MASK = np.uint32((1 << KEY_ID_BITS) - 1)
KEY_FLOOR = np.float32(-3.0e38)
KEY_FLOOR_TEST = np.float32(-2.9e38)


def pack(cm, cid):
    bits = np.asarray(cm, dtype=np.float32).view(np.uint32)
    return ((bits & ~MASK) | np.asarray(cid, dtype=np.uint32)).view(np.float32)


def insert(a, key):
    """the five-instruction network: a = [a1, a2, a3] sorted descending"""
    t1 = np.minimum(a[0], key)
    a1 = np.maximum(a[0], key)
    t2 = np.minimum(a[1], t1)
    a2 = np.maximum(a[1], t1)
    a3 = np.maximum(a[2], t2)
    return [a1, a2, a3]


def test_header_constants():
    assert KEY_ID_BITS == 12
    assert "1.05f / (float)(1u << (23 - KEY_ID_BITS))" in _const("C_KEY8")
    assert float(_const("KEY_FLOOR").rstrip("f")) == float(KEY_FLOOR) or abs(float(_const("KEY_FLOOR").rstrip("f")) + 3.0e38) < 1e31
    # both thresholds and the bound term are used by the two e4m3 taus
    assert HDR.count("+ C_KEY8") == 2


def test_key_is_within_the_relative_term_of_its_maximum():
    rng = np.random.RandomState(0)
    cm = np.concatenate([rng.standard_normal(200000) * 10.0 ** rng.uniform(-6, 7, 200000),
                         [0.0, -0.0, 1e-38, -1e-38, 3.2e7, -3.2e6]]).astype(np.float32)
    cid = rng.randint(0, 1 << KEY_ID_BITS, cm.shape[0])
    key = pack(cm, cid)
    assert np.all(np.isfinite(key))
    rel = 2.0 ** (KEY_ID_BITS - 23)
    err = np.abs(key.astype(np.float64) - cm.astype(np.float64))
    assert np.all(err <= rel * np.abs(cm.astype(np.float64)) + 1e-30)          # what C_KEY8 = 1.05 * rel covers
    # the id travels with the value, and a key orders like its maximum whenever the maxima differ by more than the term
    assert np.array_equal(key.view(np.uint32) & MASK, cid.astype(np.uint32))
    i, j = rng.randint(0, cm.shape[0], (2, 100000))
    sep = np.abs(cm[i].astype(np.float64) - cm[j]) > rel * (np.abs(cm[i]) + np.abs(cm[j])) + 1e-30
    assert np.array_equal((key[i] > key[j])[sep], (cm[i] > cm[j])[sep])


def test_network_keeps_the_three_largest_keys():
    rng = np.random.RandomState(1)
    n_rows, n_chunks = 4000, 320
    cm = (rng.standard_normal((n_rows, n_chunks)) * 100).astype(np.float32)
    cm[::7, ::5] = np.nan                                     # NaN chunk maxima enter as the floor: fmaxf drops NaN operands
    cmf = np.where(np.isnan(cm), KEY_FLOOR, np.maximum(cm, KEY_FLOOR))
    a = [np.full(n_rows, -np.inf, np.float32)] * 3
    for c in range(n_chunks):
        a = insert(a, pack(cmf[:, c], np.full(n_rows, c)))
    keys = pack(cmf, np.broadcast_to(np.arange(n_chunks), cm.shape))
    want = -np.sort(-keys, axis=1)[:, :3]
    for r in range(3):
        assert np.array_equal(a[r], want[:, r])
    # unpacking: chunk id from the low bits; the floor / -inf mean "no chunk"
    real = a[0] > KEY_FLOOR_TEST
    assert real.all()
    i1 = a[0].view(np.uint32) & MASK
    assert np.array_equal(i1, np.argmax(keys, axis=1).astype(np.uint32))
    # the best key's chunk is within the relative term of the true best chunk maximum
    best = np.nanmax(cm, axis=1).astype(np.float64)
    got = cm[np.arange(n_rows), i1].astype(np.float64)
    rel = 2.0 ** (KEY_ID_BITS - 23)
    assert np.all(best - got <= 2 * rel * np.abs(best) + 1e-30)


def test_all_nan_row_unpacks_as_no_chunk():
    a = [np.full(1, -np.inf, np.float32)] * 3
    for c in range(40):
        a = insert(a, pack(np.full(1, KEY_FLOOR, np.float32), np.full(1, c)))
    assert not (a[0] > KEY_FLOOR_TEST).any()                  # key_unpack: (value, id) = (-inf, -1) -> exhaustive scan
