"""Helpers to load tests/golden/*.npz (written by oracle/make_golden.py)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def unpack_dicts(z, prefix="in_"):
    labels = [str(l) for l in z[prefix + "labels"]]
    mats, vids, durs, lms = {}, {}, {}, {}
    for i, l in enumerate(labels):
        mats[l] = z["%smat_%d" % (prefix, i)]
        vids[l] = z["%svid_%d" % (prefix, i)]
        durs[l] = z["%sdur_%d" % (prefix, i)]
        lms[l] = [int(v) for v in z["%slm_%d" % (prefix, i)]]
    return mats, vids, durs, lms


def dp_cases():
    z = load("dp_cases.npz")
    for i in range(int(z["n_cases"])):
        N, S, mode, temp, used, ok = z["c%d_meta" % i]
        yield dict(N=int(N), S=int(S), mode=int(mode), temp=float(temp), used=int(used), ok=int(ok),
                   vec=z["c%d_vec" % i], u=z["c%d_u" % i], lp=float(z["c%d_lp" % i]),
                   b=z["c%d_b" % i])
