from .rerank import BaseReranker, SimpleSimilarityReranker, build_reranker

CrossEncoderReranker = None  # transformer inference: out of scope (see rerank.py docstring)

__all__ = ["build_reranker", "SimpleSimilarityReranker", "BaseReranker", "CrossEncoderReranker"]
