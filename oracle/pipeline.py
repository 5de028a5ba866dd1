"""End-to-end restatement of the reference's retrieval path on arrays
(oracle; test infrastructure only).

``HybridOracle.search`` follows ``HybridIndexer.search`` (reference
``src/indexing/hybrid_indexer.py:60-110``): vector search with ``top_k*2``, BM25
search with ``top_k*2`` (which itself over-fetches ``*2`` again,
``bm25_indexer.py:128``), then RRF or weighted fusion; ``rerank`` follows
``scripts/query.py:153-160`` + ``src/reranker/rerank.py:29-68``.  Results are dicts
with ``row`` (corpus row), ``chunk_id`` (fusion key) and ``score``.
"""
from __future__ import annotations

import numpy as np

from . import bm25 as _bm25
from . import dense as _dense
from . import fusion as _fusion
from . import rerank as _rerank


class HybridOracle:
    def __init__(self, emb_f32: np.ndarray | None, bm25_index, chunk_ids=None, exact_rows=True):
        """``emb_f32``: [N, D] fp32 values of the stored (bf16-rounded) embeddings;
        ``bm25_index``: ``Bm25Csr`` or ``BM25OkapiRestated`` (``get_scores``);
        ``chunk_ids``: optional list of fusion keys per row (default ``chunk_<row>``)."""
        self.raw = None if emb_f32 is None else np.ascontiguousarray(emb_f32, dtype=np.float32)
        self.rows = None if emb_f32 is None else _dense.normalize_rows(self.raw)
        self.bm25 = bm25_index
        self.chunk_ids = chunk_ids
        self.exact_rows = exact_rows

    # -- the two retrievers ------------------------------------------------
    def vector_search(self, qvec, limit, row_passes=None, score_threshold=None):
        sc = _dense.cosine_scores(self.rows, qvec, exact_rows=self.exact_rows)
        rows, s = _dense.dense_search(sc, limit, row_passes, score_threshold)
        return _fusion.as_dicts(rows, s, self.chunk_ids)

    def bm25_search(self, qterms, top_k, row_passes=None):
        sc = self.bm25.get_scores(qterms)
        rows, s = _bm25.bm25_search(sc, top_k, row_passes)
        return _fusion.as_dicts(rows, s, self.chunk_ids)

    # -- HybridIndexer.search ----------------------------------------------
    def search(self, qvec, qterms, top_k=10, vector_weight=0.5, bm25_weight=0.5,
               fusion_method="rrf", row_passes=None):
        v = self.vector_search(qvec, top_k * 2, row_passes) if self.rows is not None else []
        b = self.bm25_search(qterms, top_k * 2, row_passes) if self.bm25 is not None else []
        if fusion_method == "rrf":
            return _fusion.rrf_fuse(v, b, top_k=top_k)
        return _fusion.weighted_fuse(v, b, vector_weight, bm25_weight, top_k=top_k)

    # -- simple rerank -------------------------------------------------------
    def rerank(self, qvec_raw, results, top_n=50, top_k=10):
        """The pool texts re-embed to the stored rows (same model, same text), so
        ``d_vec`` = ``self.raw[row]`` (fp32 values of the bf16 rows)."""
        if not results:
            return results
        pool = results[: min(top_n, len(results))]
        q = np.asarray(qvec_raw, dtype=np.float32)
        sc = [_rerank.cosine_f32_seq(q, self.raw[r["row"]]) for r in pool]
        return _rerank.rerank(results, sc, top_n=top_n, top_k=top_k)
