"""Drop-in ``BM25Indexer`` (reference ``src/indexing/bm25_indexer.py``) on the CUDA path.

Same constructor, methods, attributes, result-dict schema and error behaviour; the
arithmetic ``rank_bm25`` did on the host (``BM25Okapi.__init__`` / ``get_scores`` +
``np.argsort``) runs in ``csrc/bm25.cu`` through ``drs_bm25_topk``.
"""
from __future__ import annotations

import pickle
import re
from pathlib import Path
from typing import Dict, List

import numpy as np
import torch

from .. import _lib
from .bm25_index import Bm25DeviceIndex

_TOKEN = re.compile(r"\b\w+\b")


class _NoAbbreviations:
    """Stand-in for ``MedicalTermNormalizer`` when none is supplied.  The tokenizer only
    reads ``.abbreviations`` to keep 1-character tokens that are abbreviation keys
    (bm25_indexer.py:64-76); every key of the reference's default table has >= 2
    characters (medical_term_normalizer.py:170-218), so an empty table tokenises
    identically."""
    abbreviations: Dict[str, str] = {}


class _PickledObject:
    """Attribute bag standing in for a class whose module is not installed here (e.g. the
    ``rank_bm25.BM25Okapi`` object inside a ``bm25_index.pkl`` written by the reference,
    ``src/indexing/bm25_indexer.py:165-181``).  Nothing is computed with it: the device index is rebuilt
    from the ``tokenized_corpus`` the same pickle carries."""

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {"state": state})


def _load_pickle_with_shims(f):
    """``pickle.load`` that tolerates classes from modules missing in this environment."""
    class _Unpickler(pickle.Unpickler):
        def find_class(self, module, name):
            try:
                return super().find_class(module, name)
            except (ImportError, AttributeError):
                if module.split(".")[0] in ("rank_bm25",):
                    return type(name, (_PickledObject,), {"__module__": module})
                raise
    return _Unpickler(f).load()


def chunk_as_dict(chunk) -> Dict:
    return chunk.to_dict() if hasattr(chunk, "to_dict") else chunk


def format_result(chunk, idx: int, score) -> Dict:
    """The result-dict schema of bm25_indexer.py:149-157 / vector_indexer.py:217-226."""
    d = chunk_as_dict(chunk)
    return {"chunk_id": d.get("chunk_id", f"chunk_{idx}"), "doc_id": d.get("doc_id", ""),
            "source": d.get("source", ""), "chunk_type": d.get("chunk_type", ""), "text": d.get("text", ""),
            "metadata": d.get("metadata", {}), "score": score}


class BM25Indexer:
    def __init__(self, k1: float = 1.5, b: float = 0.75, medical_normalizer=None, device=None):
        _lib.load()  # ImportError if the CUDA library is missing (reference: rank-bm25 missing, :36-37)
        self.k1 = k1
        self.b = b
        self.normalizer = medical_normalizer or _NoAbbreviations()
        self.bm25: Bm25DeviceIndex | None = None
        self.chunks: List = []
        self.tokenized_corpus: List[List[str]] = []
        self.device = torch.device(device) if device is not None else torch.device("cuda")
        self._ws = _lib.Workspace()
        print(f"Initialized BM25Indexer (k1={k1}, b={b})")

    # -- bm25_indexer.py:50-78 ---------------------------------------------------------
    def tokenize(self, text: str) -> List[str]:
        text = text.lower()
        preserved = [a.lower() for a in self.normalizer.abbreviations.keys() if a.lower() in text]
        return [t for t in _TOKEN.findall(text) if len(t) >= 2 or t in preserved]

    def encode_texts(self, texts):
        """Batch of query strings -> (q_terms, q_offsets) device tensors: ``tokenize`` + the vocabulary lookup on the
        GPU for ASCII strings (exactly the host tokeniser's tokens), the host tokeniser for the others."""
        keys = {a.lower() for a in self.normalizer.abbreviations.keys()}
        return self.bm25.encode_texts(list(texts), keep_single=[k for k in keys if len(k) == 1],
                                      host_tokenize=self.tokenize)

    # -- bm25_indexer.py:80-111 --------------------------------------------------------
    def index_chunks(self, chunks: List):
        print(f"\nIndexing {len(chunks)} chunks with BM25...")
        self.chunks = chunks
        for chunk in chunks:
            text = chunk.text if hasattr(chunk, "text") else chunk["text"]
            # NB the reference APPENDS and never clears (:92-99); kept for behavioural parity
            self.tokenized_corpus.append(self.tokenize(text))
        print("Building BM25 index...")
        self.bm25 = Bm25DeviceIndex.from_token_lists(self.tokenized_corpus, k1=self.k1, b=self.b,
                                                     device=self.device)
        print(f"Successfully indexed {len(chunks)} chunks")
        print(f"Average document length: {np.mean([len(d) for d in self.tokenized_corpus]):.2f} tokens")

    # -- batched device API (additive) ----------------------------------------------------
    def topk_terms(self, q_terms: torch.Tensor, q_offsets: torch.Tensor, n_queries: int, kprime: int):
        """Device tensors in, (scores fp64 [Q,k'], ids int64 [Q,k']) out."""
        return _lib.bm25_topk(self.bm25.struct(), q_terms, q_offsets, n_queries, kprime, ws=self._ws)

    # -- bm25_indexer.py:113-163 -------------------------------------------------------
    def search(self, query: str, top_k: int = 10, filters: Dict = None) -> List[Dict]:
        if self.bm25 is None:
            raise ValueError("Index not built. Call index_chunks() first.")
        qt, qo = self.bm25.encode_queries([self.tokenize(query)])
        fetch = top_k * 2 if filters else top_k        # the reference takes argsort[::-1][:top_k*2] (:128)
        fetch = max(1, min(fetch, self.bm25.n_docs))
        if fetch <= _lib.MAX_KPRIME:
            scores, ids = self.topk_terms(qt.to(self.device), qo.to(self.device), 1, fetch)
            scores, ids = scores[0].cpu().numpy(), ids[0].cpu().numpy()
        else:
            # a window wider than the selection kernel's lists (e.g. top_k = 100 with filters => 200): exact fp64
            # scores of every document on the device, same arithmetic, then the reference's argsort order
            scores, ids = self.bm25.wide_topk(qt.tolist(), fetch)
            scores, ids = scores.cpu().numpy(), ids.cpu().numpy()
        results = []
        for idx, s in zip(ids, scores):
            if idx < 0:
                break
            chunk = self.chunks[idx]
            if filters:
                d = chunk_as_dict(chunk)
                if any(d.get(k) != v for k, v in filters.items()):
                    continue
            results.append(format_result(chunk, int(idx), float(s)))
            if len(results) >= top_k:
                break
        return results

    # -- bm25_indexer.py:165-200 (own on-disk format, same keys) ---------------------------
    def save(self, filepath: str):
        filepath = Path(filepath)
        filepath.parent.mkdir(parents=True, exist_ok=True)
        data = {"bm25": None if self.bm25 is None else self.bm25.state_dict(), "chunks": self.chunks,
                "tokenized_corpus": self.tokenized_corpus, "k1": self.k1, "b": self.b, "format": "drs_b200/1"}
        with open(filepath, "wb") as f:
            pickle.dump(data, f)
        print(f"Saved BM25 index to {filepath}")

    def load(self, filepath: str):
        filepath = Path(filepath)
        if not filepath.exists():
            raise FileNotFoundError(f"Index file not found: {filepath}")
        with open(filepath, "rb") as f:
            data = _load_pickle_with_shims(f)
        self.chunks = data["chunks"]
        self.tokenized_corpus = data["tokenized_corpus"]
        self.k1, self.b = data["k1"], data["b"]
        if data.get("format") == "drs_b200/1" and data["bm25"] is not None:
            self.bm25 = Bm25DeviceIndex.from_state_dict(data["bm25"], device=self.device)
        else:  # a foreign pickle (e.g. the reference's): rebuild from the tokenised corpus it carries
            self.bm25 = Bm25DeviceIndex.from_token_lists(self.tokenized_corpus, k1=self.k1, b=self.b,
                                                         device=self.device)
        print(f"Loaded BM25 index from {filepath}")
        print(f"Index contains {len(self.chunks)} chunks")

    # -- bm25_indexer.py:202-213 -------------------------------------------------------
    def get_stats(self) -> Dict:
        if self.bm25 is None:
            return {"error": "Index not built"}
        return {"num_documents": len(self.chunks),
                "avg_doc_length": float(np.mean([len(doc) for doc in self.tokenized_corpus])),
                "vocab_size": len(set(tok for doc in self.tokenized_corpus for tok in doc)),
                "k1": self.k1, "b": self.b}
