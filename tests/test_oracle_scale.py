"""The benchmark-scale oracle (oracle/scale.py) is the plain tier-B oracle, rearranged: bit-identical results."""
import numpy as np

from domain_rag_system_b200 import synth
from oracle import bm25 as obm25
from oracle import dense as odense
from oracle import scale as oscale


def test_topk_desc_rows_equals_stable_argsort_with_ties():
    rng = np.random.default_rng(3)
    for n, k in ((1, 1), (5, 10), (1000, 7), (1000, 1000), (4096, 60)):
        s = np.round(rng.normal(size=n), 1)            # many exact ties
        rows, sc = oscale.topk_desc_rows(s, k)
        ref = np.argsort(s, kind="stable")[::-1][:k]
        assert np.array_equal(rows, ref) and np.array_equal(sc, s[ref])
    rows, sc = oscale.topk_desc_rows(np.zeros(50), 8)  # all zero: largest rows first
    assert rows.tolist() == list(range(49, 41, -1))


def test_sampled_bm25_equals_tier_b_on_blocks():
    n, vocab = 6000, 800
    offs, toks = synth.zipf_docs(n, vocab=vocab, seed=11)
    queries = synth.zipf_queries(40, offs, toks, vocab=vocab, seed=5)
    full = obm25.Bm25Csr(offs, toks, vocab)
    terms = np.concatenate(queries)
    sm = oscale.SampledBm25(vocab, terms)
    cuts = [0, 1000, 1001, 2500, 6000]                 # uneven blocks, one of a single document
    blocks = []
    for a, e in zip(cuts[:-1], cuts[1:]):
        blocks.append((offs[a:e + 1] - offs[a], toks[offs[a]:offs[e]]))
    sm.add_blocks(blocks, threads=3)
    sm.finalize()
    assert sm.corpus_size == full.corpus_size and sm.avgdl == full.avgdl
    assert np.array_equal(sm.idf, full.idf) and sm.average_idf == full.average_idf
    for q in queries:
        a, b = sm.get_scores(q), full.get_scores(q)
        assert np.array_equal(a, b)
        r1, s1 = oscale.bm25_topk(a, 20)
        r2, s2 = obm25.bm25_search(b, 20)
        assert np.array_equal(r1, r2) and np.array_equal(s1, s2)


def test_blocked_dense_equals_batch_oracle():
    n, q, k = 5000, 9, 12
    emb = synth.unit_embeddings(n, seed=2, dup_frac=0.01).float().numpy()
    qv = synth.query_embeddings(q, seed=8).numpy().astype(np.float64)
    qv[3] = 0.0                                        # zero query: pure tie order
    normed = odense.normalize_rows(emb)
    ref = odense.cosine_scores_batch(normed, qv, block=1 << 15)
    bd = oscale.BlockedDense(qv, k)
    for a, e in ((0, 700), (700, 701), (701, 3000), (3000, 5000)):
        bd.add_block(a, odense.normalize_rows(emb[a:e]))
    for i in range(q):
        rows, sc = bd.result(i)
        r2, s2 = odense.dense_search(ref[i], k)
        assert np.array_equal(rows, r2)
        assert np.allclose(sc, s2, rtol=0, atol=1e-15)
