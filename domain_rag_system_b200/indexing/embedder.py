"""Embedder contract of the retrieval path.

Model inference is out of scope; the retrievers only need the
contract of ``MedicalEmbedder`` (reference ``src/indexing/embedder.py:8-98``):
``encode(str | List[str]) -> np.float32 [n, dim]`` (unit-norm), ``embedding_dim``,
``model_name``.  Any object with that surface works, including the reference's own
``MedicalEmbedder``.  ``TableEmbedder`` serves precomputed vectors (benchmarks, tests,
and deployments that embed offline).
"""
from __future__ import annotations

import hashlib
from typing import Dict, Iterable, List, Union

import numpy as np


class TableEmbedder:
    def __init__(self, dim: int = 768, model_name: str = "precomputed-table"):
        self.embedding_dim = dim
        self.model_name = model_name
        self.table: Dict[str, np.ndarray] = {}

    def add(self, texts: Iterable[str], vectors) -> None:
        for t, v in zip(texts, vectors):
            self.table[t] = np.asarray(v, dtype=np.float32)

    def _one(self, text: str) -> np.ndarray:
        v = self.table.get(text)
        if v is None:  # unseen text: reproducible pseudo-random unit vector
            seed = int.from_bytes(hashlib.sha256(text.encode("utf-8")).digest()[:8], "little")
            v = np.random.default_rng(seed).standard_normal(self.embedding_dim).astype(np.float32)
            v /= np.linalg.norm(v)
        return v

    def encode(self, texts: Union[str, List[str]], normalize: bool = True) -> np.ndarray:
        if isinstance(texts, str):
            texts = [texts]
        if not len(texts):
            return np.zeros((0, self.embedding_dim), dtype=np.float32)
        return np.stack([self._one(t) for t in texts]).astype(np.float32)

    def encode_chunks(self, chunks, text_key: str = "text") -> np.ndarray:
        return self.encode([c[text_key] if isinstance(c, dict) else c.text for c in chunks])

    def get_model_info(self) -> Dict:
        return {"model_name": self.model_name, "embedding_dim": self.embedding_dim, "max_length": None,
                "device": "n/a"}
