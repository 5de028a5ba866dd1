"""Model of bm25_select_kernel's "crossing mode" (csrc/bm25.cu): with non-negative impacts a document's running
sum only grows, so it reaches a fixed positive threshold at exactly one addition; noting documents at that addition
(old < tau <= new), in ANY order of additions, yields exactly the set a scan of the finished tile would find, each
document once.  With a negative impact the property fails -- which is why indexes that carry negative impacts
(the epsilon floor of a negative average idf) take the scan path (`impacts_nonneg == 0`)."""
import numpy as np
import pytest


def _notes(n_docs, adds, tau):
    score = np.zeros(n_docs, dtype=np.int64)
    noted = []
    for doc, v in adds:
        old = score[doc]
        score[doc] = old + v
        if old < tau <= old + v:
            noted.append(doc)
    return score, noted


@pytest.mark.parametrize("seed", range(10))
def test_crossings_equal_scan_for_nonnegative_impacts(seed):
    rng = np.random.default_rng(seed)
    n_docs, n_adds = 500, 4000
    adds = [(int(rng.integers(0, n_docs)), int(rng.integers(0, 50))) for _ in range(n_adds)]   # zeros included
    tau = int(rng.integers(1, 600))
    want = None
    for _ in range(3):                                   # any schedule of the same additions
        rng.shuffle(adds)
        score, noted = _notes(n_docs, adds, tau)
        assert len(noted) == len(set(noted))             # each document noted once
        got = sorted(noted)
        assert got == sorted(np.flatnonzero(score >= tau).tolist())
        assert want is None or got == want
        want = got


def test_a_negative_impact_breaks_the_property():
    adds = [(0, 10), (0, -8), (0, 10), (1, 3)]           # doc 0: 10 -> 2 -> 12 crosses tau = 5 twice
    _, noted = _notes(2, adds, 5)
    assert noted.count(0) == 2
    _, noted = _notes(2, [(0, 10), (0, -8)], 5)          # ... or is noted although it ends below the threshold
    assert noted == [0]
