"""Dense cosine top-k restatement (oracle; test infrastructure only).

Restates what ``VectorIndexer.search`` (reference ``src/indexing/vector_indexer.py:165-230``)
gets from ``QdrantClient.query_points`` in LOCAL mode with ``Distance.COSINE``
(third-party ``qdrant-client``, un-vendored, pin ``>=1.7.0`` ``requirements.txt:6``):

* upsert: each vector is L2-normalised (fp32) and appended to an fp32 matrix;
* query: the query arrives as a Python list (``vector_indexer.py:204`` ``.tolist()``)
  => float64 array, normalised (zero norm guarded by float32 eps), scores =
  ``matrix @ q`` (fp32 rows widened), ``order = argsort(scores)[::-1]``, walk the
  order skipping rows that fail the payload filter, break when below
  ``score_threshold``, stop at ``limit``.

"parity unpinned": the library source is not in this image; the dtype of the dot
(fp32 vs fp64) differs from this restatement by <= 1e-6 relative, inside the 1e-3
tolerance the north star states for cosine scores.
"""
from __future__ import annotations

import numpy as np

EPS = float(np.finfo(np.float32).eps)  # 1.19e-7, the library's zero-norm guard


def normalize_rows(x: np.ndarray) -> np.ndarray:
    """Upsert-time normalisation: fp32 rows / fp32 norm (zero rows stay zero)."""
    x = np.ascontiguousarray(x, dtype=np.float32).copy()
    nrm = np.linalg.norm(x, axis=1).astype(np.float32)
    x /= np.where(nrm != 0.0, nrm, np.float32(1.0))[:, None]
    return x


def cosine_scores(rows_normed: np.ndarray, query, exact_rows: bool = False,
                  block: int = 1 << 16) -> np.ndarray:
    """fp64 [N] scores of one query against upsert-normalised fp32 rows.

    ``exact_rows=True`` uses a per-row reduction whose rounding does not depend
    on the row's position, so byte-identical rows get byte-identical scores
    (needed when a test plants duplicates); the default uses BLAS.
    """
    q = np.array(list(np.asarray(query).tolist()), dtype=np.float64)
    qn = np.linalg.norm(q)
    q = q / (qn if qn != 0.0 else EPS)
    n = rows_normed.shape[0]
    out = np.empty(n, dtype=np.float64)
    for a in range(0, n, block):
        blk = rows_normed[a:a + block].astype(np.float64)
        out[a:a + block] = (blk * q).sum(axis=1) if exact_rows else blk @ q
    return out


def dot_scores(rows: np.ndarray, query) -> np.ndarray:
    """``Distance.DOT`` (``vector_indexer.py:57-62``): nothing is normalised, score = row . query.
    fp64 [N]; ``rows`` are the stored vectors as the index holds them (here: the bf16 values widened)."""
    q = np.array(list(np.asarray(query).tolist()), dtype=np.float64)
    return np.ascontiguousarray(rows, dtype=np.float32).astype(np.float64) @ q


def euclid_scores(rows: np.ndarray, query) -> np.ndarray:
    """``Distance.EUCLID`` (``vector_indexer.py:57-62``): nothing is normalised, score = ||query - row||_2 of the
    stored fp32 rows against the fp64 query; SMALLER is better (``dense_search(..., smaller_is_better=True)``)."""
    q = np.array(list(np.asarray(query).tolist()), dtype=np.float64)
    return np.linalg.norm(q[None, :] - np.ascontiguousarray(rows, dtype=np.float32).astype(np.float64), axis=1)


def cosine_scores_batch(rows_normed: np.ndarray, queries: np.ndarray, block: int = 1 << 15) -> np.ndarray:
    """fp64 [Q, N] scores for a batch (BLAS dgemm per block)."""
    q = np.asarray(queries, dtype=np.float64)
    qn = np.linalg.norm(q, axis=1)
    q = q / np.where(qn != 0.0, qn, EPS)[:, None]
    n = rows_normed.shape[0]
    out = np.empty((q.shape[0], n), dtype=np.float64)
    for a in range(0, n, block):
        out[:, a:a + block] = q @ rows_normed[a:a + block].astype(np.float64).T
    return out


def dense_search(scores: np.ndarray, limit: int, row_passes=None, score_threshold=None, smaller_is_better=False):
    """Walk of the sorted order (local-mode ``search``): filter BEFORE counting,
    threshold break.  Tie rule: larger row first (``argsort(kind='stable')[::-1]``); for distances
    (``smaller_is_better``) the order is ascending and the threshold cuts what lies ABOVE it.
    Returns (rows int64, scores float64 holding Python-float values)."""
    order = np.argsort(-scores if smaller_is_better else scores, kind="stable")[::-1]
    rows, out = [], []
    for idx in order:
        if len(rows) >= limit:
            break
        if row_passes is not None and not row_passes(int(idx)):
            continue
        s = float(scores[idx])
        if score_threshold is not None and (s > score_threshold if smaller_is_better else s < score_threshold):
            break
        rows.append(int(idx))
        out.append(s)
    return np.array(rows, dtype=np.int64), np.array(out, dtype=np.float64)
