"""Shared helpers for the parity tests."""
from __future__ import annotations

import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def golden_module(name="make_golden"):
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, os.path.join(GOLDEN, name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def assert_topk_matches(ids, scores, ref_scores_full, k, rel_tol, tie_ulps=8, what=""):
    """``ids``/``scores``: the k results of the path under test for ONE query.
    ``ref_scores_full``: the oracle's score of every row (fp64 [N]).

    North-star bar: top-k ids bit-exact except at documented score ties; scores within
    ``rel_tol`` relative.  A "documented tie" is a pair of rows whose ORACLE scores differ
    by at most ``tie_ulps`` fp32 ulps: the oracle's own fp32/BLAS rounding cannot order
    them reliably, so membership/ordering may differ inside such a group."""
    n = ref_scores_full.shape[0]
    k_eff = min(k, n)
    ids = np.asarray(ids)[:k_eff]
    scores = np.asarray(scores, dtype=np.float64)[:k_eff]
    assert (ids >= 0).all() and len(set(ids.tolist())) == k_eff, f"{what}: bad/duplicate ids {ids}"
    ref_at = ref_scores_full[ids]
    denom = np.maximum(np.abs(ref_at), 1e-30)
    assert np.all(np.abs(scores - ref_at) <= rel_tol * denom + 1e-12), \
        f"{what}: scores off: {scores} vs {ref_at}"
    order = np.argsort(ref_scores_full, kind="stable")[::-1]
    ref_ids = order[:k_eff]
    if np.array_equal(ids, ref_ids):
        return 0
    # differences must be explained by near-ties
    tol = tie_ulps * np.spacing(np.float32(np.max(np.abs(ref_scores_full[ref_ids])))).astype(np.float64)
    kth = ref_scores_full[ref_ids[-1]]
    for r in ids:
        assert ref_scores_full[r] >= kth - tol, f"{what}: row {r} does not belong to the top-{k_eff}"
    for a, b in zip(ids, ref_ids):
        if a != b:
            assert abs(ref_scores_full[a] - ref_scores_full[b]) <= tol, \
                f"{what}: order differs outside a tie group ({a} vs {b})"
    return 1


def dicts_to_lists(results, key="score"):
    return [[r["chunk_id"], r["metadata"]["row"], r[key]] for r in results]
