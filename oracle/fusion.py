"""Score-fusion restatement (oracle; test infrastructure only).

Follows ``HybridIndexer._reciprocal_rank_fusion`` (reference
``src/indexing/hybrid_indexer.py:112-156``) and ``_weighted_fusion`` (``:158-218``)
statement by statement, on result dicts keyed by ``chunk_id``.  Pinned against the
reference's own functions by ``tests/golden/fusion_golden.json`` and against the
RRF values stored in the reference's ``results/*.json``.
"""
from __future__ import annotations


def rrf_fuse(vector_results, bm25_results, top_k=10, k=60):
    """hybrid_indexer.py:127-156.  1-based ranks, fp64, duplicates accumulate,
    first dict seen wins (vector list first), stable sort => ties keep insertion
    order (vector ids in rank order, then BM25-only ids)."""
    rrf = {}
    for rank, r in enumerate(vector_results, 1):          # :129-131
        cid = r["chunk_id"]
        rrf[cid] = rrf.get(cid, 0) + 1 / (k + rank)
    for rank, r in enumerate(bm25_results, 1):            # :134-136
        cid = r["chunk_id"]
        rrf[cid] = rrf.get(cid, 0) + 1 / (k + rank)
    first = {}
    for r in list(vector_results) + list(bm25_results):   # :139-143
        first.setdefault(r["chunk_id"], r)
    ranked = sorted(rrf.items(), key=lambda kv: kv[1], reverse=True)   # :146
    out = []
    for cid, score in ranked[:top_k]:                     # :150-154
        d = first[cid].copy()
        d["score"] = float(score)
        d["fusion_method"] = "rrf"
        out.append(d)
    return out


def _minmax_inplace(results):
    """hybrid_indexer.py:175-185: writes ``normalized_score`` INTO the input dicts;
    leaves them untouched when the list is empty or max == min."""
    if not results:
        return results
    scores = [r["score"] for r in results]
    lo, hi = min(scores), max(scores)
    if hi == lo:
        return results
    for r in results:
        r["normalized_score"] = (r["score"] - lo) / (hi - lo)
    return results


def weighted_fuse(vector_results, bm25_results, vector_weight=0.5, bm25_weight=0.5, top_k=10):
    """hybrid_indexer.py:187-218.  The representative dict is the LAST vector
    occurrence of the id if any (``result_map[cid] = result`` per occurrence,
    :198), else the first BM25 occurrence (:204-205)."""
    vector_results = _minmax_inplace(vector_results)
    bm25_results = _minmax_inplace(bm25_results)
    combined, rep = {}, {}
    for r in vector_results:                              # :194-198
        cid = r["chunk_id"]
        s = r.get("normalized_score", r["score"])
        combined[cid] = combined.get(cid, 0) + vector_weight * s
        rep[cid] = r
    for r in bm25_results:                                # :200-205
        cid = r["chunk_id"]
        s = r.get("normalized_score", r["score"])
        combined[cid] = combined.get(cid, 0) + bm25_weight * s
        if cid not in rep:
            rep[cid] = r
    ranked = sorted(combined.items(), key=lambda kv: kv[1], reverse=True)
    out = []
    for cid, score in ranked[:top_k]:
        d = rep[cid].copy()
        d["score"] = float(score)
        d["fusion_method"] = "weighted"
        out.append(d)
    return out


def as_dicts(rows, scores, keys=None):
    """Minimal result dicts for array-level comparisons: ``chunk_id`` is the
    fusion key (``keys[row]`` or ``chunk_<row>``), ``row`` rides along."""
    out = []
    for r, s in zip(rows, scores):
        r = int(r)
        cid = f"chunk_{r}" if keys is None else keys[r]
        out.append({"chunk_id": cid, "row": r, "score": float(s)})
    return out
