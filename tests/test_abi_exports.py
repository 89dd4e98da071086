"""CPU tests: the C-ABI shared library loads and exports every symbol include/segb200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "segb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(segb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from segmentalist_b200 import _lib
    assert os.path.exists(_lib.SO_PATH)
    lib = ctypes.CDLL(_lib.SO_PATH)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "missing export " + n
    # the ctypes prototypes cover exactly the declared functions
    assert sorted(_lib.EXPORTS) == names
    assert _lib.load().segb_version() >= 100


def test_product_path_fails_loudly_without_gpu():
    import torch
    from segmentalist_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        _lib.lib()


def test_no_oracle_import_in_product():
    pkg = os.path.join(ROOT, "segmentalist_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                # neither an import of the package nor any mention of its modules / shared library
                for needle in ("from oracle", "import oracle", "seg_oracle", "libseg_oracle", "oracle/"):
                    assert needle not in txt, "%s mentions %r" % (os.path.join(dp, fn), needle)


def test_members_by_component_groups_like_np_where():
    """Host logic of the device diagnostics: the stable sort of the assignments lists every component's
    members in index order -- the order of np.where(assignments == k) in the reference
    (gaussian_components_fixedvar.py:270, kmeans_components.py:242)."""
    import numpy as np
    import torch
    from segmentalist_b200 import _lib
    rng = np.random.RandomState(0)
    K_max = 9
    assign = rng.randint(-1, 7, size=200).astype(np.int32)
    order, seg_off = _lib.members_by_component(torch.from_numpy(assign), K_max)
    order, seg_off = order.numpy(), seg_off.numpy()
    assert seg_off.shape == (K_max + 1,) and seg_off[0] == (assign == -1).sum() and seg_off[-1] == len(assign)
    for k in range(K_max):
        np.testing.assert_array_equal(order[seg_off[k]:seg_off[k + 1]], np.where(assign == k)[0])


def test_e4m3_scale_choice():
    """Host logic of the e4m3 first-level filter: the operand scale is the largest power of two that keeps every
    scaled element and every scaled row norm inside e4m3's finite range (448), so that s^2 |mu|^2 / 2 fits three e4m3
    terms times the 256.0 constant column (means are averages of rows)."""
    import numpy as np
    import torch
    from segmentalist_b200.batch import MmaScorer
    rng = np.random.RandomState(0)
    for mag in (1e-4, 0.03, 1.0, 37.0, 5e3):
        X = torch.from_numpy((rng.standard_normal((1000, 130)) * mag).astype(np.float32))
        s = MmaScorer.pick_scale(X, chunk=256)
        assert s > 0 and np.log2(s) == int(np.log2(s))
        lim = max(float(X.abs().max()), float(torch.linalg.vector_norm(X, dim=1).max()))
        assert s * lim <= 448.0 and 2 * s * lim > 448.0
    assert MmaScorer.pick_scale(torch.zeros(10, 4)) == 2.0 ** 40          # degenerate data: capped, finite
    assert MmaScorer.pick_scale(torch.full((3, 4), float("inf"))) == 1.0
