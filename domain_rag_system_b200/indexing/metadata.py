"""``index_metadata.json``: what ``scripts/build_index.py:78-90`` writes next to ``bm25_index.pkl`` and
``scripts/query.py:25-92`` / ``scripts/rag.py:14-70`` read to re-attach to a built index.

The reference's file lacks the vector store's location (``query.py:58`` falls back to ``data/vector_db``);
``write_index_metadata`` records it as ``vector_storage_path``, the key ``query.py`` already looks for.
``load_hybrid_indexer`` is ``query.py:load_hybrid_indexer`` over this package's classes; the embedder is passed in
(model inference is outside this package)."""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict, Optional

from .bm25_indexer import BM25Indexer
from .hybrid_indexer import HybridIndexer
from .vector_indexer import VectorIndexer

FILE_NAME = "index_metadata.json"


def write_index_metadata(output_dir, hybrid: HybridIndexer, num_chunks: Optional[int] = None) -> Path:
    vi, bi = hybrid.vector_indexer, hybrid.bm25_indexer
    meta = {"num_chunks": int(num_chunks if num_chunks is not None else (len(vi.payloads) if vi else len(bi.chunks))),
            "embedding_model": getattr(vi.embedder, "model_name", None) if vi else None,
            "embedding_dim": getattr(vi.embedder, "embedding_dim", None) if vi else None,
            "collection_name": vi.collection_name if vi else None,
            "bm25_params": {"k1": bi.k1, "b": bi.b} if bi else {"k1": 1.5, "b": 0.75},
            "distance_metric": vi.distance_metric if vi else "cosine",
            "vector_storage_path": str(vi.storage_path) if vi and vi.storage_path else None}
    path = Path(output_dir) / FILE_NAME
    path.parent.mkdir(parents=True, exist_ok=True)
    with open(path, "w") as f:
        json.dump(meta, f, indent=2)
    return path


def load_index_metadata(metadata_path) -> Dict:
    with open(metadata_path, "r", encoding="utf-8") as f:
        return json.load(f)


def load_hybrid_indexer(indices_dir, embedder, metadata: Optional[Dict] = None, medical_normalizer=None, device=None,
                        import_qdrant: bool = False) -> HybridIndexer:
    """Attach to a built index: ``indices_dir`` holds index_metadata.json and bm25_index.pkl (query.py:31-92)."""
    indices = Path(indices_dir)
    meta = metadata if metadata is not None else load_index_metadata(indices / FILE_NAME)
    vi = VectorIndexer(collection_name=meta["collection_name"], embedder=embedder,
                       storage_path=meta.get("vector_storage_path") or "data/vector_db",
                       distance_metric=meta.get("distance_metric", "cosine"), device=device, import_qdrant=import_qdrant)
    params = meta.get("bm25_params", {"k1": 1.5, "b": 0.75})
    bi = BM25Indexer(k1=params.get("k1", 1.5), b=params.get("b", 0.75), medical_normalizer=medical_normalizer,
                     device=device)
    bm25_path = indices / "bm25_index.pkl"
    if not bm25_path.exists():
        raise FileNotFoundError(f"BM25 index not found: {bm25_path}")
    bi.load(str(bm25_path))
    return HybridIndexer(vector_indexer=vi, bm25_indexer=bi, storage_path=str(indices))
