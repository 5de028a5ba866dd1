"""Drop-in ``VectorIndexer`` (reference ``src/indexing/vector_indexer.py``) on the CUDA path.

The embedded Qdrant collection becomes a bf16 row-major matrix in HBM plus fp32 inverse
row norms; ``query_points`` becomes ``drs_dense_topk``.  Payload dicts stay on the host,
indexed by row (the reference's point id == row index, vector_indexer.py:133-135).
"""
from __future__ import annotations

import json
import pickle
import weakref
from pathlib import Path
from typing import Dict, List

import numpy as np
import torch

from .. import _lib
from .bm25_indexer import chunk_as_dict

# embedder object -> VectorIndexer holding the rows that embedder produced.  Lets the simple
# reranker (which only receives the embedder, scripts/query.py:153-158) find stored rows
# instead of re-embedding texts.
_BY_EMBEDDER: "weakref.WeakValueDictionary[int, VectorIndexer]" = weakref.WeakValueDictionary()


def _hashable(v):
    try:
        hash(v)
        return v
    except TypeError:
        return repr(v)


def indexer_for_embedder(embedder):
    return _BY_EMBEDDER.get(id(embedder))


class VectorIndexer:
    def __init__(self, collection_name: str = "medical_docs", embedder=None, storage_path: str = "./data/vector_db",
                 distance_metric: str = "cosine", use_in_memory: bool = False, device=None,
                 import_qdrant: bool = False):
        """Same arguments as the reference's class; ``device`` and ``import_qdrant`` are additions.
        ``import_qdrant=True``: when ``storage_path`` holds a collection written by the reference's qdrant-client
        local mode (and none in this package's format), import its points (``import_qdrant_local``).  Off by
        default: the points are pickles, and reading pickles should be something the caller asked for."""
        _lib.load()  # ImportError when the CUDA library is missing (reference: qdrant-client missing, :44-45)
        if embedder is None:
            raise ValueError("an embedder object is required (model inference is outside this package); "
                             "pass the reference's MedicalEmbedder or a TableEmbedder")
        self.collection_name = collection_name
        self.embedder = embedder
        self.use_in_memory = use_in_memory
        if not use_in_memory:
            self.storage_path = Path(storage_path)
            self.storage_path.mkdir(parents=True, exist_ok=True)
        else:
            self.storage_path = None
        metric = distance_metric.lower()
        self.distance_metric = metric if metric in ("cosine", "euclidean", "dot") else "cosine"   # (:57-62)
        self._row_scale = 1.0                        # 'dot' / 'euclidean': largest row norm R (ranked by q.c / R)
        self.device = torch.device(device) if device is not None else torch.device("cuda")
        self.rows: torch.Tensor | None = None        # bf16 [N, D] in HBM: the stored vectors
        self._search_rows: torch.Tensor | None = None  # what the kernels scan (== rows except for 'euclidean')
        self.inv_norm: torch.Tensor | None = None    # fp32 [N]
        self.payloads: List[Dict] = []
        self._exists = False
        self._ws = _lib.Workspace()
        self._mask_cache: Dict = {}
        self._codes: Dict = {}                       # payload key -> (value -> code, int32 codes per row on device)
        self._text_rows: Dict[str, int] | None = None
        _BY_EMBEDDER[id(embedder)] = self
        if use_in_memory:
            print("Initialized in-memory vector store")
        else:
            print(f"Initialized persistent vector store at: {self.storage_path}")
            self._try_attach()
            if import_qdrant and not self._exists:
                self.import_qdrant_local()
        print(f"Collection: {collection_name}")
        print(f"Embedding model: {self.embedder.model_name}")

    # -- persistence: <storage>/collection/<name>/{rows.pt,payload.pkl} ----------------------
    def _dir(self) -> Path:
        return self.storage_path / "collection" / self.collection_name

    def _try_attach(self):
        d = self._dir()
        if (d / "rows.pt").exists():
            self.set_rows(torch.load(d / "rows.pt", map_location="cpu"))
            with open(d / "payload.pkl", "rb") as f:
                self.payloads = pickle.load(f)
            self._exists = True

    def import_qdrant_local(self, storage_path=None) -> int:
        """Import the points of a collection written by the reference's qdrant-client local mode
        (vector_indexer.py:51-53,68) from ``storage_path`` (default: this indexer's) and persist them in this
        package's format next to it.  Returns the number of points (0: no such collection)."""
        from .qdrant_import import read_qdrant_local
        root = Path(storage_path) if storage_path is not None else self.storage_path
        if root is None or not (root / "collection" / self.collection_name / "storage.sqlite").exists():
            return 0
        vecs, payloads, _cfg = read_qdrant_local(root, self.collection_name)
        print(f"Importing {len(payloads)} points from the Qdrant local collection at {root}")
        if len(payloads):
            self.payloads = payloads
            self.set_rows(torch.from_numpy(vecs))
            self._persist()
        self._exists = True
        return len(payloads)

    def _persist(self):
        if self.storage_path is None or self.rows is None:
            return
        d = self._dir()
        d.mkdir(parents=True, exist_ok=True)
        torch.save(self.rows.cpu(), d / "rows.pt")
        with open(d / "payload.pkl", "wb") as f:
            pickle.dump(self.payloads, f)
        # (not "meta.json": that name belongs to qdrant-client's local mode, whose directory we may be sharing)
        with open(self.storage_path / "drs_b200_meta.json", "w") as f:
            json.dump({"collections": {self.collection_name: {"size": int(self.rows.shape[1]),
                                                              "distance": self.distance_metric}}}, f)

    # -- vector_indexer.py:74-103 ------------------------------------------------------------
    def create_collection(self, force_recreate: bool = True):
        if self._exists and force_recreate:
            print(f"Deleting existing collection: {self.collection_name}")
            self.delete_collection(_quiet=True)
        if not self._exists:
            print(f"Creating collection: {self.collection_name}")
            self.rows, self.inv_norm, self.payloads = None, None, []
            self._exists = True
            print(f"Collection created with dimension: {self.embedder.embedding_dim}")
        else:
            print(f"Collection {self.collection_name} already exists")

    def set_rows(self, rows):
        """Install the corpus matrix: any float tensor/array [N, D] -> bf16 in HBM + inverse norms."""
        t = torch.as_tensor(rows)
        t = t.to(self.device, non_blocking=True)
        if t.dtype != torch.bfloat16:
            t = t.to(torch.bfloat16)
        self.rows = t.contiguous()
        self._search_rows = self.rows
        self.inv_norm = _lib.row_inv_norms(self.rows)
        if self.distance_metric in ("dot", "euclidean"):
            # Distance.DOT / EUCLID (vector_indexer.py:57-62): rows and query are NOT normalised.
            # EUCLID: ||q - c||^2 = ||q||^2 - 2 (q.c - ||c||^2 / 2), so the nearest rows are the largest values of
            # q'.c' with q' = [q, 1, 1, 1, 0..] and c' = [c, b_hi, b_mid, b_lo, 0..], b = -||c||^2 / 2 split into three
            # bf16 pieces that add up to its fp32 value exactly: a dot product over a 64-column wider matrix.
            if self.distance_metric == "euclidean":
                n, d = self.rows.shape
                half = -0.5 * (self.rows.double() ** 2).sum(dim=1)
                b32 = half.float()
                hi = b32.to(torch.bfloat16)
                mid = (b32 - hi.float()).to(torch.bfloat16)
                lo = (b32 - hi.float() - mid.float()).to(torch.bfloat16)
                wide = torch.zeros(n, d + 64, dtype=torch.bfloat16, device=self.device)
                wide[:, :d] = self.rows
                wide[:, d], wide[:, d + 1], wide[:, d + 2] = hi, mid, lo
                self._search_rows = wide
                self.inv_norm = _lib.row_inv_norms(wide)
            # The kernels scale every row by inv_norm[row] and need |c| * inv_norm <= 1 for their exactness bound,
            # so all rows get the same factor 1/R (R = largest row norm): the ranking is the dot product's, and
            # topk_vectors() multiplies the scores back by R * |q|.
            nz = self.inv_norm > 0
            r = float((1.0 / self.inv_norm[nz]).max()) if bool(nz.any()) else 1.0
            self._row_scale = r
            self.inv_norm = torch.full_like(self.inv_norm, 1.0 / r)
        self._mask_cache.clear()
        self._codes.clear()
        self._text_rows = None

    # -- vector_indexer.py:105-163 -----------------------------------------------------------
    def index_chunks(self, chunks: List, batch_size: int = 100):
        print(f"\nIndexing {len(chunks)} chunks...")
        if not self._exists:
            raise RuntimeError(f"Collection {self.collection_name} does not exist")  # qdrant raises on upsert
        texts = [c.text if hasattr(c, "text") else c["text"] for c in chunks]
        print("Generating embeddings...")
        emb = np.asarray(self.embedder.encode(texts), dtype=np.float32)
        # PointStruct(id=idx) with idx counted from 0 on EVERY call (:126-135): an upsert OVERWRITES the points
        # 0 .. len(chunks)-1 of an existing collection and only extends it beyond its current size.
        new_payloads = []
        for idx, chunk in enumerate(chunks):
            d = chunk_as_dict(chunk)
            new_payloads.append({"chunk_id": d.get("chunk_id", f"chunk_{idx}"), "doc_id": d.get("doc_id", ""),
                                 "source": d.get("source", ""), "chunk_type": d.get("chunk_type", ""),
                                 "text": d.get("text", ""), "metadata": d.get("metadata", {})})
        new = torch.from_numpy(emb).to(self.device).to(torch.bfloat16)
        n_old = 0 if self.rows is None else self.rows.shape[0]
        if n_old > len(chunks):
            new = torch.cat([new, self.rows[len(chunks):]])          # on the device: the old tail stays
            new_payloads = new_payloads + self.payloads[len(chunks):]
        self.payloads = new_payloads
        self.set_rows(new)
        self._persist()
        print(f"Successfully indexed {len(chunks)} chunks")
        print(f"Collection points count: {self.rows.shape[0]}")

    # -- filters: equality on top-level payload keys, applied BEFORE top-k (Qdrant semantics) ------
    def _key_codes(self, key):
        """Dictionary encoding of one payload key: (value -> code, int32 codes per row on the device).  Built once
        per key (one pass over the payloads), after which a filter on any value of it is a device comparison."""
        got = self._codes.get(key)
        if got is None:
            table: Dict = {}
            codes = np.fromiter((table.setdefault(_hashable(p.get(key)), len(table)) for p in self.payloads),
                                dtype=np.int32, count=len(self.payloads))
            got = (table, torch.from_numpy(codes).to(self.device))
            self._codes[key] = got
        return got

    def _row_mask(self, filters: Dict | None):
        if not filters:
            return None
        key = tuple(sorted((k, repr(v)) for k, v in filters.items()))
        m = self._mask_cache.get(key)
        if m is None:
            m = torch.ones(len(self.payloads), dtype=torch.bool, device=self.device)
            for k, v in filters.items():
                table, codes = self._key_codes(k)
                code = table.get(_hashable(v), -1)                   # a value no row carries: nothing passes
                m &= codes == code
            m = m.to(torch.uint8)
            self._mask_cache[key] = m
        return m

    # -- batched device API (additive) ---------------------------------------------------------
    def topk_vectors(self, q_f32: torch.Tensor, kprime: int, row_mask=None, kernel=_lib.DENSE_AUTO):
        """fp32 device queries [Q, D] -> (scores fp32 [Q,k'], ids int64 [Q,k'], flags int32 [Q]).  Scores are
        cosines, dot products or (euclidean) distances -- ascending for distances."""
        if kprime > _lib.MAX_KPRIME:
            raise ValueError(f"top_k: {kprime} results per retriever requested, the CUDA path serves at most "
                             f"{_lib.MAX_KPRIME} (HybridIndexer.search asks for 2 * top_k)")
        q = q_f32
        if self.distance_metric == "euclidean":
            q = torch.zeros(q_f32.shape[0], self._search_rows.shape[1], dtype=torch.float32, device=q_f32.device)
            d = q_f32.shape[1]
            q[:, :d] = q_f32
            q[:, d:d + 3] = 1.0
        scores, ids, flags = _lib.dense_topk(q, self._search_rows, self.inv_norm, kprime, row_mask=row_mask,
                                             kernel=kernel, ws=self._ws)
        if self.distance_metric != "cosine":
            scores = scores.double() * (torch.linalg.vector_norm(q.double(), dim=1, keepdim=True) * self._row_scale)
            if self.distance_metric == "euclidean":
                qq = (q_f32.double() ** 2).sum(dim=1, keepdim=True)
                scores = torch.sqrt(torch.clamp(qq - 2.0 * scores, min=0.0))
                scores = torch.where(ids >= 0, scores, torch.full_like(scores, float("inf")))
            scores = scores.float()
        return scores, ids, flags

    # -- vector_indexer.py:165-230 ---------------------------------------------------------------
    def search(self, query: str, top_k: int = 10, filters: Dict = None, score_threshold: float = None) -> List[Dict]:
        if self.rows is None or self.rows.shape[0] == 0:
            return []
        q = np.asarray(self.embedder.encode(query), dtype=np.float32)[:1]
        qd = torch.from_numpy(np.ascontiguousarray(q)).to(self.device)
        kprime = max(1, min(top_k, self.rows.shape[0]))
        scores, ids, _ = self.topk_vectors(qd, kprime, self._row_mask(filters))
        scores, ids = scores[0].cpu().numpy(), ids[0].cpu().numpy()
        results = []
        smaller_is_better = self.distance_metric == "euclidean"
        for idx, s in zip(ids, scores):
            if idx < 0:
                break
            s = float(s)
            if score_threshold is not None and (s > score_threshold if smaller_is_better else s < score_threshold):
                break
            p = self.payloads[idx]
            results.append({"chunk_id": p["chunk_id"], "doc_id": p["doc_id"], "source": p["source"],
                            "chunk_type": p["chunk_type"], "text": p["text"], "metadata": p["metadata"],
                            "score": s})
        return results

    def rows_of_texts(self, texts: List[str]) -> List[int]:
        """Row of each text (-1 if it is not a stored chunk text); used by the simple reranker."""
        if self._text_rows is None:
            self._text_rows = {}
            for i, p in enumerate(self.payloads):
                self._text_rows.setdefault(p["text"], i)
        return [self._text_rows.get(t, -1) for t in texts]

    # -- vector_indexer.py:232-250 ---------------------------------------------------------------
    def get_stats(self) -> Dict:
        if not self._exists:
            return {"error": f"Collection {self.collection_name} not found"}
        return {"collection_name": self.collection_name,
                "points_count": 0 if self.rows is None else int(self.rows.shape[0]),
                "embedding_dim": self.embedder.embedding_dim, "model_name": self.embedder.model_name,
                "storage_path": str(self.storage_path) if self.storage_path else "in-memory"}

    def delete_collection(self, _quiet: bool = False):
        self.rows, self._search_rows, self.inv_norm, self.payloads = None, None, None, []
        self._exists = False
        self._mask_cache.clear()
        self._codes.clear()
        self._text_rows = None
        if self.storage_path is not None:
            d = self._dir()
            for f in ("rows.pt", "payload.pkl"):
                if (d / f).exists():
                    (d / f).unlink()
        if not _quiet:
            print(f"Deleted collection: {self.collection_name}")
