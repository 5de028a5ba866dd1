"""Drop-in rerankers (reference ``src/reranker/rerank.py``) on the CUDA path.

``SimpleSimilarityReranker`` keeps the reference's contract -- cosine between the embedding of
the (raw) query and the embedding of each pool text, stable sort, top_k -- but the pool texts
are the stored chunks, so their embeddings are the rows already in HBM: the work becomes
"gather rows + cosine + stable sort" in ``csrc/rerank.cu`` (bit-identical to the reference's
``_cosine`` under NumPy >= 2).  Texts that are not stored chunks are embedded with the supplied
embedder and scored by the same kernel.  ``CrossEncoderReranker`` is transformer inference and
out of scope; ``build_reranker('crossencoder')`` degrades to the simple reranker exactly as the
reference does when the cross-encoder cannot be constructed (rerank.py:96-100).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Literal, Optional, Tuple

import numpy as np
import torch

from .. import _lib
from ..indexing.vector_indexer import indexer_for_embedder


class BaseReranker:
    """Minimal rerank interface (rerank.py:19-43)."""

    def __init__(self, top_n: int = 50, name: str = "base"):
        self.top_n = top_n
        self.name = name

    def score_pairs(self, pairs: Iterable[Tuple[str, str]]) -> List[float]:
        raise NotImplementedError

    def rerank(self, query: str, results: List[Dict], top_k: int = 10) -> List[Dict]:
        if not results:
            return results
        pool = results[: min(self.top_n, len(results))]
        pairs = [(query, c.get("text", "") or "") for c in pool]
        scores = self.score_pairs(pairs)
        rescored = []
        for c, s in zip(pool, scores):
            r = c.copy()
            r["rerank_score"] = float(s)
            r["fusion_method"] = (r.get("fusion_method") or "pre") + "+rerank"
            rescored.append(r)
        rescored.sort(key=lambda x: x["rerank_score"], reverse=True)   # stable, as in the reference
        return rescored[:top_k]


class SimpleSimilarityReranker(BaseReranker):
    """Cosine rerank on the device (rerank.py:46-68)."""

    def __init__(self, embedder=None, top_n: int = 50):
        super().__init__(top_n=top_n, name="simple-sim")
        if embedder is None:
            raise ValueError("an embedder object is required (model inference is outside this package)")
        _lib.load()
        self.embedder = embedder
        self.dim = self.embedder.embedding_dim

    def score_pairs(self, pairs: Iterable[Tuple[str, str]]) -> List[float]:
        pairs = list(pairs)
        if not pairs:
            return []
        q_vec = np.asarray(self.embedder.encode(pairs[0][0]), dtype=np.float32)[:1]
        texts = [p for _, p in pairs]
        vi = indexer_for_embedder(self.embedder)
        rows = vi.rows_of_texts(texts) if vi is not None and vi.rows is not None else [-1] * len(texts)
        missing = [i for i, r in enumerate(rows) if r < 0]
        if vi is not None and vi.rows is not None and not missing:
            corpus, dev = vi.rows, vi.device
            pool = torch.tensor([rows], dtype=torch.int64, device=dev)
        else:  # texts that are not stored chunks: embed them, score against an ad-hoc matrix
            dev = vi.device if vi is not None else torch.device("cuda")
            d_vecs = np.asarray(self.embedder.encode(texts), dtype=np.float32)
            corpus = torch.from_numpy(d_vecs).to(dev).to(torch.bfloat16).contiguous()
            pool = torch.arange(len(texts), dtype=torch.int64, device=dev).view(1, -1)
        n = pool.shape[1]
        if n > _lib.MAX_POOL:
            raise NotImplementedError(f"rerank pool larger than {_lib.MAX_POOL}")
        qd = torch.from_numpy(np.ascontiguousarray(q_vec)).to(dev)
        sc, _, pos = _lib.rerank_cosine(corpus, qd, pool, n)
        out = np.empty(n, dtype=np.float32)
        out[pos[0].cpu().numpy()] = sc[0].cpu().numpy()
        return [np.float32(v) for v in out]


def build_reranker(kind: Literal["none", "simple", "crossencoder"] = "simple", top_n: int = 50,
                   cross_model: str = "cross-encoder/ms-marco-MiniLM-L-6-v2", embedder=None) -> Optional[BaseReranker]:
    """Factory (rerank.py:87-101)."""
    if kind == "none":
        return None
    return SimpleSimilarityReranker(embedder=embedder, top_n=top_n)
