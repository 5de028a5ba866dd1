"""Model classes the reference constructs (src/indexing/vector_indexer.py:9-12)."""
from dataclasses import dataclass, field
from enum import Enum
from typing import Any, List, Optional


class Distance(str, Enum):
    COSINE = "Cosine"
    EUCLID = "Euclid"
    DOT = "Dot"


@dataclass
class VectorParams:
    size: int
    distance: Distance


@dataclass
class PointStruct:
    id: Any
    vector: Any
    payload: Optional[dict] = None


@dataclass
class MatchValue:
    value: Any


@dataclass
class FieldCondition:
    key: str
    match: MatchValue


@dataclass
class Filter:
    must: Optional[List[FieldCondition]] = field(default=None)
