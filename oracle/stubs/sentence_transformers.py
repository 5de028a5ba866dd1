"""Empty shell: the reference imports these names at module import time
(src/indexing/embedder.py:3, src/reranker/rerank.py:7); nothing on the path calls them."""


class SentenceTransformer:  # pragma: no cover
    def __init__(self, *a, **k):
        raise RuntimeError("sentence_transformers is not available; pass an embedder object")


class CrossEncoder(SentenceTransformer):  # pragma: no cover
    pass
