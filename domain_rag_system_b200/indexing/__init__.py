from .embedder import TableEmbedder
from .vector_indexer import VectorIndexer
from .bm25_indexer import BM25Indexer
from .bm25_index import Bm25DeviceIndex
from .hybrid_indexer import HybridIndexer

__all__ = ["TableEmbedder", "VectorIndexer", "BM25Indexer", "Bm25DeviceIndex", "HybridIndexer"]
