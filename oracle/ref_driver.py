"""Import the reference's OWN retrieval classes over the stand-in modules
(oracle; test infrastructure only; build container only -- ``/root/reference`` does
not exist on the GPU box, so nothing under ``-m gpu``, ``smoke()`` or ``bench.py``
may call this).

Used by ``tests/golden/make_golden.py`` to produce the committed golden fixtures
and by ``tests/test_reference_crosscheck.py`` (skipped when the mount is absent).
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys

REFERENCE_ROOT = os.environ.get("DRS_REFERENCE_ROOT", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "indexing"))


def load():
    """Returns a namespace with the reference's HybridIndexer, VectorIndexer,
    BM25Indexer, SimpleSimilarityReranker, BaseReranker, build_reranker."""
    if not available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    repo_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (repo_root, _STUBS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    with contextlib.redirect_stdout(io.StringIO()):
        idx = importlib.import_module("src.indexing")
        rr = importlib.import_module("src.reranker.rerank")
    from types import SimpleNamespace
    return SimpleNamespace(HybridIndexer=idx.HybridIndexer, VectorIndexer=idx.VectorIndexer,
                           BM25Indexer=idx.BM25Indexer, BaseReranker=rr.BaseReranker,
                           SimpleSimilarityReranker=rr.SimpleSimilarityReranker,
                           build_reranker=rr.build_reranker)


@contextlib.contextmanager
def quiet():
    """The reference prints progress; swallow it."""
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        yield
