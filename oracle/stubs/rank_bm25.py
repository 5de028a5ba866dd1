"""Stand-in for the un-vendored ``rank_bm25`` package (oracle; tests only)."""
from oracle.bm25 import BM25OkapiRestated as BM25Okapi  # noqa: F401
