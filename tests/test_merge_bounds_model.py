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


@pytest.mark.parametrize("seed", range(10))
def test_pair_kernel_pool_part_bound_and_thresholded_append(seed):
    """dense_pair.cu: the corpus is cut into pool parts (slab x column group), every part publishes its running
    MAXIMUM, and the threshold is the kk-th largest of the published maxima -- with dead parts (no valid row: -inf) and
    parts that have not reported yet left out.  Replayed tile by tile with stale, partial information:
      * the threshold never exceeds the kk-th best score of the whole corpus (kk different rows reach it), so
      * appending every score >= the threshold of the moment keeps all of the final top-kk, however late the
        thresholds arrive, and the append count only shrinks as they tighten."""
    rng = np.random.default_rng(500 + seed)
    kk = int(rng.integers(4, 70))
    n_slabs, n_tiles_per_slab, groups, cols = int(rng.integers(2, 12)), int(rng.integers(2, 9)), int(rng.choice([1, 2, 4, 8])), 64
    n_parts = n_slabs * groups
    if n_parts < kk:                       # make_plan picks G so that kk <= n_slabs * G
        kk = n_parts
    scores = rng.standard_normal((n_slabs, n_tiles_per_slab, cols)).astype(np.float32)
    if seed % 3 == 0:
        scores = np.round(scores, 1)                                   # mass ties
    valid = rng.random(scores.shape) > (0.3 if seed % 2 else 0.0)      # masked rows / rows beyond N
    if seed % 4 == 1:
        valid[0] = False                                               # a whole slab without a valid row: dead parts
    flat = scores[valid]
    kth = np.sort(flat)[::-1][kk - 1] if flat.size >= kk else -np.inf
    part_max = np.full((n_slabs, groups), np.nan)                      # nan = not reported yet
    appended = []
    threshold_seen = []
    gtau = -np.inf
    for t in range(n_tiles_per_slab):                                  # all slabs advance tile by tile (co-resident)
        order = rng.permutation(n_slabs)
        for s in order:
            # thresholds are refreshed from whatever the other slabs have published so far (stale is fine)
            rep = part_max[~np.isnan(part_max)]
            live = np.sort(rep[np.isfinite(rep)])[::-1]
            if live.size >= kk:
                gtau = max(gtau, live[kk - 1])                         # atomicMax: thresholds only rise
            threshold_seen.append(gtau)
            tile, ok = scores[s, t], valid[s, t]
            keep = ok & (tile >= gtau)
            appended.extend(tile[keep].tolist())
            for g in range(groups):                                    # publish this tile's part maxima
                sl = slice(g * cols // groups, (g + 1) * cols // groups)
                m = tile[sl][ok[sl]].max() if ok[sl].any() else -np.inf
                prev = part_max[s, g]
                part_max[s, g] = m if np.isnan(prev) else max(prev, m)
    assert all(th <= kth for th in threshold_seen)                     # a valid lower bound at every moment
    assert threshold_seen == sorted(threshold_seen)
    top = np.sort(flat)[::-1][:kk]
    app = np.sort(np.array(appended, dtype=np.float32))[::-1]
    assert app.size >= min(kk, flat.size)
    # multiset inclusion of the final top-kk in what was appended
    vals, cnt = np.unique(top, return_counts=True)
    for v, c in zip(vals, cnt):
        assert (app == v).sum() >= c
