"""Drop-in ``HybridIndexer`` (reference ``src/indexing/hybrid_indexer.py``) on the CUDA path.

``search`` keeps the reference's orchestration (vector top-2K, BM25 top-2K, RRF or weighted
fusion, top-K; hybrid_indexer.py:60-110); the fusion arithmetic runs in ``csrc/fuse.cu``.
``search_batch`` is the additive batched surface: tensors in, tensors out,
no Python per query.
"""
from __future__ import annotations

import math
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .. import _lib
from .bm25_indexer import BM25Indexer
from .vector_indexer import VectorIndexer


class HybridIndexer:
    def __init__(self, vector_indexer: VectorIndexer = None, bm25_indexer: BM25Indexer = None,
                 storage_path: str = "./data/indices"):
        self.storage_path = Path(storage_path)
        self.storage_path.mkdir(parents=True, exist_ok=True)
        self.vector_indexer = vector_indexer
        self.bm25_indexer = bm25_indexer
        self._fusion_key = None      # int64 [N] device: rows sharing a chunk_id share a key
        self._fusion_key_src = None

    # -- hybrid_indexer.py:30-58 ----------------------------------------------------------------
    def index_chunks(self, chunks: List):
        print("\n" + "=" * 60)
        print("Building Hybrid Index")
        print("=" * 60)
        if self.vector_indexer:
            print("\n[1/2] Building vector index...")
            self.vector_indexer.create_collection(force_recreate=True)
            self.vector_indexer.index_chunks(chunks)
        else:
            print("\nWarning: No vector indexer provided")
        if self.bm25_indexer:
            print("\n[2/2] Building BM25 index...")
            self.bm25_indexer.index_chunks(chunks)
        else:
            print("\nWarning: No BM25 indexer provided")
        self._fusion_key = None
        print("\n" + "=" * 60)
        print("Hybrid index built successfully!")
        print("=" * 60)

    # -- hybrid_indexer.py:60-110 ---------------------------------------------------------------
    def search(self, query: str, top_k: int = 10, vector_weight: float = 0.5, bm25_weight: float = 0.5,
               filters: Dict = None, fusion_method: str = "rrf") -> List[Dict]:
        vector_results, bm25_results = [], []
        if self.vector_indexer:
            vector_results = self.vector_indexer.search(query, top_k=top_k * 2, filters=filters)
        if self.bm25_indexer:
            bm25_results = self.bm25_indexer.search(query, top_k=top_k * 2, filters=filters)
        if fusion_method == "rrf":
            return self._reciprocal_rank_fusion(vector_results, bm25_results, top_k=top_k)
        return self._weighted_fusion(vector_results, bm25_results, vector_weight=vector_weight,
                                     bm25_weight=bm25_weight, top_k=top_k)

    # -- fusion of result-dict lists: encode chunk_id strings as keys, fuse on the GPU, rebuild dicts --
    def _device(self):
        if self.vector_indexer is not None:
            return self.vector_indexer.device
        if self.bm25_indexer is not None:
            return self.bm25_indexer.device
        return torch.device("cuda")

    def _fuse_dicts(self, mode, vector_results, bm25_results, top_k, k=60, wv=0.5, wb=0.5):
        if top_k <= 0 or (not vector_results and not bm25_results):
            return [], None, None
        dev = self._device()
        entries = list(vector_results) + list(bm25_results)
        keys: Dict[str, int] = {}
        ids = [keys.setdefault(r["chunk_id"], len(keys)) for r in entries]
        nv, nb = len(vector_results), len(bm25_results)
        # entry index doubles as the "row": the kernel returns the representative entry index
        fusion_key = torch.tensor(ids, dtype=torch.int64, device=dev)
        v_ids = torch.arange(nv, dtype=torch.int64, device=dev).view(1, nv) if nv else None
        b_ids = torch.arange(nv, nv + nb, dtype=torch.int64, device=dev).view(1, nb) if nb else None
        v_sc = torch.tensor([[r["score"] for r in vector_results]], dtype=torch.float64, device=dev) if nv else None
        b_sc = torch.tensor([[r["score"] for r in bm25_results]], dtype=torch.float64, device=dev) if nb else None
        out_ids, out_sc, out_norm = _lib.fuse_topk(mode, v_sc, v_ids, b_sc, b_ids, top_k, fusion_key=fusion_key,
                                                   rrf_k=k, w_vector=wv, w_bm25=wb, want_norm=True)
        return entries, out_ids[0].cpu().numpy(), (out_sc[0].cpu().numpy(), out_norm[0].cpu().numpy())

    # -- hybrid_indexer.py:112-156 ----------------------------------------------------------------
    def _reciprocal_rank_fusion(self, vector_results: List[Dict], bm25_results: List[Dict], top_k: int = 10,
                                k: int = 60) -> List[Dict]:
        entries, ids, aux = self._fuse_dicts("rrf", vector_results, bm25_results, top_k, k=k)
        fused = []
        if ids is None:
            return fused
        for i, s in zip(ids, aux[0]):
            if i < 0:
                break
            r = entries[i].copy()
            r["score"] = float(s)
            r["fusion_method"] = "rrf"
            fused.append(r)
        return fused

    # -- hybrid_indexer.py:158-218 ----------------------------------------------------------------
    def _weighted_fusion(self, vector_results: List[Dict], bm25_results: List[Dict], vector_weight: float = 0.5,
                         bm25_weight: float = 0.5, top_k: int = 10) -> List[Dict]:
        # the reference writes `normalized_score` INTO the input dicts (:175-185); reproduce that
        for lst in (vector_results, bm25_results):
            if lst:
                sc = [r["score"] for r in lst]
                lo, hi = min(sc), max(sc)
                if hi != lo:
                    for r in lst:
                        r["normalized_score"] = (r["score"] - lo) / (hi - lo)
        entries, ids, aux = self._fuse_dicts("weighted", vector_results, bm25_results, top_k, wv=vector_weight,
                                             wb=bm25_weight)
        fused = []
        if ids is None:
            return fused
        for i, s in zip(ids, aux[0]):
            if i < 0:
                break
            r = entries[i].copy()
            r["score"] = float(s)
            r["fusion_method"] = "weighted"
            fused.append(r)
        return fused

    # -- batched surface ----------------------------------------------------------------------------
    def fusion_keys(self) -> Optional[torch.Tensor]:
        """int64 [N] on the device: key[row] = first row carrying the same chunk_id, or None when
        every chunk_id is unique (then the row id is the key)."""
        vi = self.vector_indexer
        payloads = vi.payloads if vi is not None and vi.payloads else None
        if payloads is None and self.bm25_indexer is not None and self.bm25_indexer.chunks:
            from .bm25_indexer import chunk_as_dict
            payloads = [{"chunk_id": chunk_as_dict(c).get("chunk_id", f"chunk_{i}")}
                        for i, c in enumerate(self.bm25_indexer.chunks)]
        if payloads is None:
            return None
        if self._fusion_key_src is not payloads:
            first: Dict[str, int] = {}
            key = np.fromiter((first.setdefault(p["chunk_id"], i) for i, p in enumerate(payloads)),
                              dtype=np.int64, count=len(payloads))
            self._fusion_key = None if len(first) == len(payloads) else torch.from_numpy(key).to(self._device())
            self._fusion_key_src = payloads
        return self._fusion_key

    def search_batch(self, query_vecs: torch.Tensor, q_terms: torch.Tensor, q_offsets: torch.Tensor, top_k: int = 10,
                     vector_weight: float = 0.5, bm25_weight: float = 0.5, fusion_method: str = "rrf",
                     rerank_top_n: int = 0, rerank_vecs: Optional[torch.Tensor] = None, final_k: Optional[int] = None,
                     fusion_key: Optional[torch.Tensor] = "auto", dense_kernel: int = _lib.DENSE_AUTO):
        """Whole hybrid path for Q queries, tensors only.

        ``query_vecs`` fp32 [Q, D] and ``q_terms``/``q_offsets`` (``Bm25DeviceIndex.encode_queries``)
        are DEVICE tensors.  Returns a dict of device tensors: ``ids`` int64 [Q, K], ``scores`` fp64
        [Q, K] (fused), and when ``rerank_top_n`` > 0 also ``rerank_ids`` / ``rerank_scores`` fp32
        [Q, final_k] (the pool is the fused top-K, K = top_k; scripts/query.py:139-160)."""
        nq = query_vecs.shape[0]
        kprime = top_k * 2
        v_sc = v_ids = b_sc = b_ids = None
        if self.vector_indexer is not None:
            kp = min(kprime, self.vector_indexer.rows.shape[0])
            v_sc, v_ids, _ = self.vector_indexer.topk_vectors(query_vecs, kp, kernel=dense_kernel)
        if self.bm25_indexer is not None:
            kp = min(kprime, self.bm25_indexer.bm25.n_docs)
            b_sc, b_ids = self.bm25_indexer.topk_terms(q_terms, q_offsets, nq, kp)
        fk = self.fusion_keys() if isinstance(fusion_key, str) else fusion_key
        ids, scores, _ = _lib.fuse_topk(fusion_method, v_sc, v_ids, b_sc, b_ids, top_k, fusion_key=fk,
                                        w_vector=vector_weight, w_bm25=bm25_weight)
        out = {"ids": ids, "scores": scores, "vector": (v_sc, v_ids), "bm25": (b_sc, b_ids)}
        if rerank_top_n and self.vector_indexer is not None:
            pool = ids[:, : min(rerank_top_n, ids.shape[1])].contiguous()
            rq = query_vecs if rerank_vecs is None else rerank_vecs
            fk_ = final_k or top_k
            r_sc, r_ids, r_pos = _lib.rerank_cosine(self.vector_indexer.rows, rq, pool, min(fk_, pool.shape[1]))
            out.update(rerank_scores=r_sc, rerank_ids=r_ids, rerank_pos=r_pos)
        return out

    def search_many(self, queries: Sequence[str], top_k: int = 10, vector_weight: float = 0.5,
                    bm25_weight: float = 0.5, fusion_method: str = "rrf", rerank_top_n: int = 0,
                    rerank_queries: Optional[Sequence[str]] = None, final_k: Optional[int] = None) -> List[List[Dict]]:
        """The batched caller (the evaluation loop of evaluation/comprehensive_evaluation.py:298-328 in one call):
        for every query the list ``search(query, top_k, ...)`` would return -- and, with ``rerank_top_n`` > 0, what
        ``build_reranker('simple', top_n=rerank_top_n).rerank(rerank_query, that_list, top_k=final_k)`` makes of it
        -- with ONE embedding call, one query encoding and one pass of the CUDA path for the whole batch.
        ``rerank_queries``: the raw queries the reranker embeds (scripts/query.py:160); default: ``queries``."""
        if not queries:
            return []
        vi, bi = self.vector_indexer, self.bm25_indexer
        dev = self._device()
        qv = None
        if vi is not None:
            qv = torch.from_numpy(np.ascontiguousarray(np.asarray(vi.embedder.encode(list(queries)), dtype=np.float32))).to(dev)
        if bi is not None:
            if bi.bm25.vocab is not None:
                qt, qo = bi.encode_texts(queries)            # tokeniser + vocabulary lookup on the device
            else:
                qt, qo = bi.bm25.encode_queries([bi.tokenize(q) for q in queries])
                qt, qo = qt.to(dev), qo.to(dev)
        else:
            qt = qo = None
        nq = len(queries)
        kprime = top_k * 2
        v_sc = v_ids = b_sc = b_ids = None
        if vi is not None and vi.rows is not None and vi.rows.shape[0]:
            v_sc, v_ids, _ = vi.topk_vectors(qv, min(kprime, vi.rows.shape[0]))
        if bi is not None:
            b_sc, b_ids = bi.topk_terms(qt, qo, nq, min(kprime, bi.bm25.n_docs))
        if v_ids is None and b_ids is None:
            return [[] for _ in queries]
        ids, scores, norm = _lib.fuse_topk(fusion_method, v_sc, v_ids, b_sc, b_ids, top_k, fusion_key=self.fusion_keys(),
                                           w_vector=vector_weight, w_bm25=bm25_weight, want_norm=True)
        r_ids = r_sc = r_pos = None
        if rerank_top_n and vi is not None:
            rq = qv
            if rerank_queries is not None:
                rq = torch.from_numpy(np.ascontiguousarray(np.asarray(vi.embedder.encode(list(rerank_queries)),
                                                                      dtype=np.float32))).to(dev)
            pool = ids[:, : min(rerank_top_n, ids.shape[1])].contiguous()
            r_sc, r_ids, r_pos = _lib.rerank_cosine(vi.rows, rq, pool, min(final_k or top_k, pool.shape[1]))
            r_sc, r_pos = r_sc.cpu().numpy(), r_pos.cpu().numpy()
        ids_h, sc_h, nm_h = ids.cpu().numpy(), scores.cpu().numpy(), norm.cpu().numpy()
        out = []
        for i in range(nq):
            fused = []
            for row, s, nm in zip(ids_h[i], sc_h[i], nm_h[i]):
                if row < 0:
                    break
                r = self._payload_dict(int(row))
                if not math.isnan(nm):
                    r["normalized_score"] = float(nm)       # the key the weighted mode leaks (:184)
                r["score"] = float(s)
                r["fusion_method"] = fusion_method if fusion_method == "rrf" else "weighted"
                fused.append(r)
            if r_pos is None:
                out.append(fused)
                continue
            rer = []
            for s, pos in zip(r_sc[i], r_pos[i]):
                if pos < 0:
                    break
                r = fused[int(pos)].copy()
                r["rerank_score"] = float(s)
                r["fusion_method"] = (r.get("fusion_method") or "pre") + "+rerank"
                rer.append(r)
            out.append(rer)
        return out

    def _payload_dict(self, row: int) -> Dict:
        """The result dict of a row, as VectorIndexer.search / BM25Indexer.search build it (without a score)."""
        vi = self.vector_indexer
        if vi is not None and vi.payloads:
            p = vi.payloads[row]
            return {"chunk_id": p["chunk_id"], "doc_id": p["doc_id"], "source": p["source"],
                    "chunk_type": p["chunk_type"], "text": p["text"], "metadata": p["metadata"]}
        from .bm25_indexer import format_result
        r = format_result(self.bm25_indexer.chunks[row], row, 0.0)
        r.pop("score", None)
        return r

    # -- hybrid_indexer.py:220-246 ----------------------------------------------------------------
    def save_bm25(self, filepath: str = None):
        if filepath is None:
            filepath = self.storage_path / "bm25_index.pkl"
        if self.bm25_indexer:
            self.bm25_indexer.save(filepath)

    def load_bm25(self, filepath: str = None):
        if filepath is None:
            filepath = self.storage_path / "bm25_index.pkl"
        if self.bm25_indexer:
            self.bm25_indexer.load(filepath)

    def get_stats(self) -> Dict:
        stats = {}
        if self.vector_indexer:
            stats["vector"] = self.vector_indexer.get_stats()
        if self.bm25_indexer:
            stats["bm25"] = self.bm25_indexer.get_stats()
        return stats
