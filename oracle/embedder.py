"""Deterministic stand-in for ``MedicalEmbedder`` (reference
``src/indexing/embedder.py:8-98``; model inference is out of scope): looks texts
up in a table of precomputed vectors.  Oracle-side twin of
``domain_rag_system_b200.indexing.embedder.TableEmbedder``; test infrastructure only."""
from __future__ import annotations

import hashlib

import numpy as np


class TableEmbedder:
    model_name = "table-embedder"

    def __init__(self, table=None, dim=768):
        self.table = dict(table or {})
        self.embedding_dim = dim

    def add(self, texts, vectors):
        for t, v in zip(texts, vectors):
            self.table[t] = np.asarray(v, dtype=np.float32)

    def _one(self, text):
        v = self.table.get(text)
        if v is None:  # unseen text: a reproducible pseudo-random unit vector
            seed = int.from_bytes(hashlib.sha256(text.encode("utf-8")).digest()[:8], "little")
            v = np.random.default_rng(seed).standard_normal(self.embedding_dim).astype(np.float32)
            v /= np.linalg.norm(v)
        return v

    def encode(self, texts, normalize=True):
        if isinstance(texts, str):
            texts = [texts]
        if not texts:
            return np.zeros((0, self.embedding_dim), dtype=np.float32)
        return np.stack([self._one(t) for t in texts]).astype(np.float32)
