"""The lower bounds the merge kernels pre-filter with (csrc/select.cu: merge_keys_kernel) are valid:
every bound is <= the kk-th largest key of the union of the partial lists, so no member of the global top-kk
is dropped.  Each partial list holds the kk best keys of its own share of the rows (what a selection CTA /
epilogue heap hands in)."""
import numpy as np
import pytest


def _lists(rng, n_lists, rows_per_list, kk, ties):
    allk = []
    lists = []
    for p in range(n_lists):
        n = int(rng.integers(max(1, rows_per_list // 3), rows_per_list + 1))
        s = rng.integers(0, 50 if ties else 1 << 30, size=n).astype(np.uint64)
        keys = (s << np.uint64(32)) | (np.arange(n, dtype=np.uint64) + np.uint64(p * 100000))   # distinct keys
        allk.append(keys)
        top = np.sort(keys)[::-1][:kk]
        lists.append(np.concatenate([top, np.zeros(kk - len(top), dtype=np.uint64)]))           # 0 = empty slot
    return lists, np.sort(np.concatenate(allk))[::-1]


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("ties", [False, True])
def test_bound_from_list_maxima(seed, ties):
    """Latency mode: at least kk lists come in; bound = kk-th largest of the lists' maxima."""
    rng = np.random.default_rng(seed)
    kk = int(rng.integers(1, 40))
    lists, union = _lists(rng, int(rng.integers(kk, 4 * kk + 1)), 60, kk, ties)
    maxima = np.sort(np.array([l.max() for l in lists]))[::-1]
    bound = maxima[kk - 1]
    kth = union[kk - 1] if len(union) >= kk else np.uint64(0)
    assert bound <= kth
    survivors = np.concatenate([l[l >= bound] for l in lists]) if bound else np.concatenate(lists)
    assert set(union[:kk].tolist()) <= set(survivors.tolist())


@pytest.mark.parametrize("seed", range(8))
def test_tensor_kernel_thresholds(seed):
    """gtau = max over heaps of their kk-th best (when full); pooled = min over heaps of their m-th best with
    n_heaps * m >= kk.  Both are <= the global kk-th best."""
    rng = np.random.default_rng(100 + seed)
    kk = int(rng.integers(4, 40))
    n_heaps = int(rng.integers(2, 40))
    lists, union = _lists(rng, n_heaps, 300, kk, False)
    kth = union[kk - 1]
    roots = [l[kk - 1] for l in lists if l[kk - 1] != 0]                  # full heaps only
    if roots:
        assert max(roots) <= kth
    m = -(-kk // n_heaps)
    if all(l[m - 1] != 0 for l in lists):
        assert min(l[m - 1] for l in lists) <= kth
