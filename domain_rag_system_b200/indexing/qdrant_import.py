"""Reader for a Qdrant LOCAL-mode directory (what the reference's ``VectorIndexer`` leaves under
``storage_path`` when ``use_in_memory=False``; reference ``src/indexing/vector_indexer.py:51-53,68``).

qdrant-client is not installed here and no such directory ships with the reference, so the format below is
restated from qdrant-client >= 1.7's local mode (``qdrant_client/local/persistence.py``) and is **unpinned**:

* ``<storage>/meta.json``: ``{"collections": {<name>: {"vectors": {"size": D, "distance": "Cosine"}, ...}}, ...}``
* ``<storage>/collection/<name>/storage.sqlite``: table ``points(id TEXT PRIMARY KEY, point BLOB)`` where ``point``
  is ``pickle.dumps(PointStruct)`` -- a pydantic model with ``id``, ``vector`` (list of floats, or a dict of named
  vectors) and ``payload`` (the reference stores ``chunk_id, doc_id, source, chunk_type, text, metadata``,
  ``vector_indexer.py:132-143``) -- and ``id`` an encoding of the point id.

The reader is deliberately permissive: it takes the first table that has a BLOB column, unpickles with stand-in
classes for everything that lives in ``qdrant_client`` / ``pydantic`` (pydantic v1 and v2 pickle states are both
understood) and refuses every other non-builtin class.  Point ids written by the reference are the chunk
positions (``vector_indexer.py:128``), so rows are ordered by id when all ids are integers, else by rowid.
"""
from __future__ import annotations

import io
import json
import pickle
import sqlite3
from pathlib import Path
from typing import Dict, List, Tuple

import numpy as np

_SHIM_ROOTS = ("qdrant_client", "pydantic", "pydantic_core")
_SAFE = {
    ("builtins", n) for n in ("dict", "list", "tuple", "set", "frozenset", "str", "int", "float", "bool", "bytes",
                              "bytearray", "complex", "object", "NoneType")
} | {("collections", "OrderedDict"), ("datetime", "datetime"), ("datetime", "date"), ("datetime", "timezone"),
     ("datetime", "timedelta"), ("uuid", "UUID"), ("copyreg", "_reconstructor"), ("copyreg", "__newobj__")}


class _Shim:
    """Stand-in for a pickled qdrant/pydantic object: keeps whatever state the pickle carries."""

    def __new__(cls, *args, **kwargs):
        self = object.__new__(cls)
        if args:
            self.__dict__["_args"] = args          # enum members are rebuilt as cls(value)
        return self

    def __init__(self, *args, **kwargs):
        pass

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2 and isinstance(state[0], dict):   # (dict, slots)
            state = state[0]
        if isinstance(state, dict):
            inner = state.get("__dict__")
            self.__dict__.update(inner if isinstance(inner, dict) else state)


_NUMPY_OK = {("numpy", "ndarray"), ("numpy", "dtype"), ("numpy.core.multiarray", "_reconstruct"),
             ("numpy._core.multiarray", "_reconstruct"), ("numpy.core.multiarray", "scalar"),
             ("numpy._core.multiarray", "scalar")}


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] in _SHIM_ROOTS:
            return type(name, (_Shim,), {"__module__": module})
        if (module, name) in _SAFE:
            return super().find_class(module, name)
        if (module, name) in _NUMPY_OK:                # vectors may have been stored as arrays: only what rebuilds one
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"refusing to load {module}.{name} from a Qdrant point blob")


def _field(obj, name):
    if isinstance(obj, dict):
        return obj.get(name)
    return getattr(obj, name, None)


def _as_vector(v):
    if isinstance(v, dict):                             # named vectors: the reference uses the unnamed one
        v = v.get("") if "" in v else next(iter(v.values()))
    if v is None:
        raise ValueError("point without a vector")
    return np.asarray(v, dtype=np.float32).reshape(-1)


def read_qdrant_local(storage_path, collection_name: str) -> Tuple[np.ndarray, List[Dict], Dict]:
    """-> (vectors fp32 [N, D] in point-id order, payload dicts, collection config from meta.json or {})."""
    root = Path(storage_path)
    db = root / "collection" / collection_name / "storage.sqlite"
    if not db.exists():
        raise FileNotFoundError(f"no Qdrant local collection at {db}")
    cfg: Dict = {}
    meta = root / "meta.json"
    if meta.exists():
        try:
            cfg = json.loads(meta.read_text()).get("collections", {}).get(collection_name, {}) or {}
        except (ValueError, OSError):
            cfg = {}
    con = sqlite3.connect(f"file:{db}?mode=ro", uri=True)
    try:
        tables = [r[0] for r in con.execute("SELECT name FROM sqlite_master WHERE type='table'")]
        table, blob_col = None, None
        for t in (["points"] if "points" in tables else []) + [t for t in tables if t != "points"]:
            cols = con.execute(f'PRAGMA table_info("{t}")').fetchall()      # (cid, name, type, ...)
            blobs = [c[1] for c in cols if "BLOB" in (c[2] or "").upper()] or [c[1] for c in cols if c[1] == "point"]
            if blobs:
                table, blob_col = t, blobs[0]
                break
        if table is None:
            raise ValueError(f"{db}: no table with a BLOB column")
        rows = con.execute(f'SELECT rowid, "{blob_col}" FROM "{table}" ORDER BY rowid').fetchall()
    finally:
        con.close()
    pts = []
    for rowid, blob in rows:
        p = _Unpickler(io.BytesIO(blob)).load()
        pts.append((rowid, _field(p, "id"), _as_vector(_field(p, "vector")), _field(p, "payload") or {}))
    if pts and all(isinstance(p[1], (int, np.integer)) and not isinstance(p[1], bool) for p in pts):
        pts.sort(key=lambda p: int(p[1]))               # reference: point id == chunk position
    if not pts:
        return np.zeros((0, int(_field(_field(cfg, "vectors") or {}, "size") or 0)), dtype=np.float32), [], cfg
    dim = pts[0][2].shape[0]
    if any(p[2].shape[0] != dim for p in pts):
        raise ValueError("points with different vector sizes")
    return np.stack([p[2] for p in pts]), [dict(p[3]) for p in pts], cfg
