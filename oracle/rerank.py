"""Simple cosine rerank restatement (oracle; test infrastructure only).

Follows ``BaseReranker.rerank`` (reference ``src/reranker/rerank.py:29-43``) and
``SimpleSimilarityReranker._cosine`` / ``score_pairs`` (``:53-68``).  Pinned by
``tests/golden/rerank_golden.json`` (outputs of the reference's own ``_cosine`` and
``rerank`` run in the build container under NumPy 2.x).
"""
from __future__ import annotations

import math

import numpy as np


def cosine_ref(a, b):
    """rerank.py:53-58, verbatim semantics: three sequential Python ``sum``s over
    the elements.  With ``np.float32`` inputs under NumPy >= 2 the sums run in
    fp32 and the result is ``np.float32``."""
    num = sum(x * y for x, y in zip(a, b))
    da = math.sqrt(sum(x * x for x in a))
    db = math.sqrt(sum(y * y for y in b))
    return 0.0 if (da == 0 or db == 0) else (num / (da * db))


def cosine_f32_seq(a: np.ndarray, b: np.ndarray) -> np.float32:
    """Same value as ``cosine_ref`` on fp32 arrays, ~50x faster: sequential fp32
    multiply-then-add (``np.cumsum`` on fp32 is a left-to-right running sum),
    fp64 sqrt of the fp32 sums, the fp64 product rounded to fp32, fp32 divide."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    num = np.cumsum(a * b, dtype=np.float32)[-1]
    da = math.sqrt(np.cumsum(a * a, dtype=np.float32)[-1])
    db = math.sqrt(np.cumsum(b * b, dtype=np.float32)[-1])
    if da == 0 or db == 0:
        return np.float32(0.0)
    return np.float32(num) / np.float32(da * db)


def rerank(results, scores, top_n=50, top_k=10):
    """rerank.py:29-43 given precomputed pool scores: pool = results[:top_n];
    each dict copied with ``rerank_score`` and ``fusion_method + '+rerank'``;
    STABLE sort by rerank_score desc (ties keep fused order); [:top_k]."""
    if not results:
        return results
    pool = results[: min(top_n, len(results))]
    out = []
    for c, s in zip(pool, scores):
        r = c.copy()
        r["rerank_score"] = float(s)
        r["fusion_method"] = (r.get("fusion_method") or "pre") + "+rerank"
        out.append(r)
    out.sort(key=lambda x: x["rerank_score"], reverse=True)
    return out[:top_k]
