"""CPU model of the BM25 exactness proof (csrc/bm25.cu: bm25_select_kernel + bm25_rescore_kernel, DESIGN.md
"Exactness of the BM25 path"): the integer pre-selection is replayed with NumPy from the SAME index tensors the
kernel reads (fixed-point postings, packed 14-bit head-term rows, per-query shift), and checked against the fp64
oracle:

  * the error bound E really bounds |exact - integer score| for every document;
  * whenever the proof declares a query "proven" (flag 0), the exact top-k' is inside the candidate list --
    i.e. the fast path returned the reference's ids; unproven queries are the ones the kernel escalates."""
import math

import numpy as np
import pytest

from domain_rag_system_b200 import synth
from domain_rag_system_b200.indexing.bm25_index import Bm25DeviceIndex
from oracle import bm25 as obm25


def _query_shift(nt):
    sh = 0
    while (16 << sh) < nt:
        sh += 1
    return sh


def _integer_scores(ix, q):
    """What bm25_select_kernel accumulates for one query: fx[doc] and the error budget in units of 2^-s."""
    n, t = ix.n_docs, ix.t
    to, pd, pi = t["term_offsets"].numpy(), t["post_doc"].numpy(), t["post_impact"].numpy().astype(np.int64)
    slot = t["dense_slot"].numpy()
    t8 = ix.tile_docs // 8
    dense = t["dense_impact"].view(-1, ix.n_tiles, t8, 2, 4).permute(0, 1, 3, 2, 4).reshape(t["dense_impact"].shape[0], -1)
    dense = dense.numpy().astype(np.int64)[:, :n]
    nt = len(q)
    sh = _query_shift(nt)
    dk = ix.dense_shift - sh
    fx = np.zeros(n, dtype=np.int64)
    rows, units = [], 0.0
    for term in q:
        if term < 0 or term >= ix.n_terms:
            continue
        if slot[term] >= 0:
            rows.append(int(slot[term]))
            units += 2.0 ** sh + (2.0 ** (ix.dense_shift - 1) if ix.dense_shift > 0 else 0.0) + 1e-6
        else:
            a, e = to[term], to[term + 1]
            np.add.at(fx, pd[a:e], pi[a:e] >> sh)
            units += 2.0 ** sh - 0.5 + 1e-6
    for g0 in range(0, len(rows), 4):                      # head-term rows: groups of four, summed, then shifted
        grp = dense[rows[g0:g0 + 4]].sum(axis=0)
        fx += (grp << dk) if dk >= 0 else (grp >> -dk)
    return fx, sh, units


@pytest.mark.parametrize("n,vocab,qlen,seed", [(3000, 60, 8, 1), (5000, 900, 8, 2), (4000, 200, 40, 3), (2500, 30, 150, 4),
                                               (6000, 3000, 5, 5)])
def test_error_bound_and_proof_are_sound(n, vocab, qlen, seed):
    offs, toks = synth.zipf_docs(n, vocab=vocab, seed=seed, len_min=2, len_max=50)
    ix = Bm25DeviceIndex.from_token_ids(offs, toks, vocab, device="cpu", tile_docs=256)
    orc = obm25.Bm25Csr(offs, toks, vocab)
    if not ix.impacts_nonneg:
        assert ix.n_dense == 0          # negative impacts keep every term on the postings path
    rng = np.random.default_rng(seed)
    s = ix.impact_scale_log2
    kprime = 10
    kk = kprime + 8 + kprime // 4
    proven = 0
    for qi in range(25):
        q = synth.zipf_queries(1, offs, toks, vocab=vocab, seed=100 * seed + qi, ntok=qlen)[0]
        if qi % 5 == 4:
            q = rng.integers(0, vocab, size=qlen).astype(np.int32)      # uniform terms: rarer, fewer matches
        fx, sh, units = _integer_scores(ix, q)
        exact = orc.get_scores(q)
        # (1) the bound holds for every document
        err_units = np.abs(exact * 2.0 ** s - fx.astype(np.float64) * 2.0 ** sh)
        assert err_units.max() <= units + 1e-6 * (1 + np.abs(exact).max() * 2.0 ** s * 1e-6), (err_units.max(), units)
        # (2) the proof, as bm25_rescore_kernel computes it
        order = np.lexsort((np.arange(n), fx))[::-1][:kk]             # top-kk by (fx desc, doc desc)
        want = obm25.stable_desc_order(exact)[:kprime]
        if len(order) < kk or n <= kk:
            continue
        cand_exact = exact[order]
        rank = np.lexsort((order, cand_exact))[::-1]                  # candidates by (exact desc, doc desc)
        L = cand_exact[rank][kprime - 1]
        E = units * 2.0 ** -s
        thr = (L - E - abs(L) * 1e-12) * 2.0 ** (s - sh)
        fx_low = math.floor(thr) - 1
        if fx[order[-1]] < fx_low:                                    # flag 0: proven
            proven += 1
            assert np.array_equal(order[rank][:kprime], want)
        else:                                                         # flag 1: every doc that can still qualify has fx >= fx_low
            assert (fx[want] >= fx_low).all()
    assert proven >= 1 or vocab <= 60                                 # the bound is not vacuous on ordinary corpora
