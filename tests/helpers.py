"""Shared test helpers: oracle-vs-engine array comparison, golden IO."""

import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

EXACT_FIELDS = ["call", "direction", "lifts", "clade1", "clade2", "lca", "best1", "best2",
                "synteny", "locus_flags", "ann_winner", "member_off", "n_members_a", "members",
                "call_counts", "call_index"]


def compare_results(ref, got, score_rtol=0.0):
    """All integer/byte outputs bit-exact; crit/rank bit-exact (score_rtol=0) or within rtol."""
    diffs = []
    for k in EXACT_FIELDS:
        a, b = np.asarray(ref[k]), np.asarray(got[k])
        if a.shape != b.shape:
            diffs.append((k, "shape", a.shape, b.shape))
        elif not np.array_equal(a, b):
            bad = np.nonzero(a.reshape(len(a), -1) != b.reshape(len(b), -1))[0]
            diffs.append((k, "first_bad", int(bad[0]), a[bad[0]].tolist(), b[bad[0]].tolist(),
                          "n_bad", len(np.unique(bad))))
    for k in ("crit", "rank"):
        a, b = np.asarray(ref[k]), np.asarray(got[k])
        if score_rtol == 0.0:
            ok = a.view(np.int64) == b.view(np.int64)
        else:
            ok = np.abs(a - b) <= score_rtol * np.maximum(np.abs(a), np.abs(b))
        if not ok.all():
            i = int(np.nonzero(~ok)[0][0])
            diffs.append((k, "first_bad", i, float(a[i]).hex(), float(b[i]).hex(),
                          "n_bad", int((~ok).sum())))
    return diffs


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as fh:
        return json.load(fh)
