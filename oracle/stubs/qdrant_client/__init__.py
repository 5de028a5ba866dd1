"""Stand-in for ``qdrant-client`` LOCAL mode (oracle; tests only).

Only the calls the reference makes (src/indexing/vector_indexer.py:65-68,81-97,
151-159,202-209,235,248) are provided; the arithmetic is ``oracle.dense``.
"""
from types import SimpleNamespace

import numpy as np

from oracle import dense as _dense
from . import models


class QdrantClient:
    def __init__(self, location=None, path=None, **_):
        self._c = {}

    def get_collections(self):
        return SimpleNamespace(collections=[SimpleNamespace(name=n) for n in self._c])

    def create_collection(self, collection_name, vectors_config, **_):
        self._c[collection_name] = {"cfg": vectors_config, "ids": [], "vecs": [], "payload": [],
                                    "mat": None}

    def delete_collection(self, collection_name, **_):
        self._c.pop(collection_name, None)

    def upsert(self, collection_name, points, **_):
        """Insert, or OVERWRITE the point that already carries the id (that is what an upsert is)."""
        c = self._c[collection_name]
        pos = {pid: i for i, pid in enumerate(c["ids"])}
        for p in points:
            v = np.asarray(p.vector, dtype=np.float32)[None, :]
            if c["cfg"].distance == models.Distance.COSINE:
                v = _dense.normalize_rows(v)
            if p.id in pos:
                i = pos[p.id]
                c["vecs"][i], c["payload"][i] = v[0], p.payload
            else:
                pos[p.id] = len(c["ids"])
                c["ids"].append(p.id)
                c["vecs"].append(v[0])
                c["payload"].append(p.payload)
        c["mat"] = None

    def get_collection(self, collection_name):
        return SimpleNamespace(points_count=len(self._c[collection_name]["ids"]))

    def query_points(self, collection_name, query, limit=10, query_filter=None,
                     with_payload=True, score_threshold=None, **_):
        c = self._c[collection_name]
        if c["mat"] is None:
            c["mat"] = (np.stack(c["vecs"]) if c["vecs"]
                        else np.zeros((0, c["cfg"].size), dtype=np.float32))
        smaller = False
        if c["cfg"].distance == models.Distance.COSINE:
            scores = _dense.cosine_scores(c["mat"], query, exact_rows=True)
        elif c["cfg"].distance == models.Distance.DOT:
            scores = _dense.dot_scores(c["mat"], query)
        else:
            scores, smaller = _dense.euclid_scores(c["mat"], query), True
        passes = None
        if query_filter is not None and query_filter.must:
            conds = query_filter.must

            def passes(i):
                pl = c["payload"][i]
                return all(pl.get(fc.key) == fc.match.value for fc in conds)
        rows, sc = _dense.dense_search(scores, limit, passes, score_threshold, smaller_is_better=smaller)
        pts = [SimpleNamespace(id=c["ids"][r], score=float(s),
                               payload=c["payload"][r] if with_payload else None)
               for r, s in zip(rows, sc)]
        return SimpleNamespace(points=pts)
