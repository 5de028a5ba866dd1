"""Parity at BASELINE.json's full sizes, through size-independent properties and independent code paths.

* configs[1] (1 M chunks x 1000 queries, k'=10): always run (about a minute).
* configs[2]/[4] (10 M chunks, k'=60 weighted + rerank top_n=30; single query k'=100): about a minute and
  ~40 GB of HBM; set DRS_FULLSIZE=0 to skip.

The CPU oracle cannot hold 10 M documents, so at that size the checks are: two independent CUDA paths
must agree bit for bit (tensor-core vs fp32 streaming kernel for dense; dense-row vs postings-only
accumulation for BM25), outputs are sorted / unique / deterministic, a batch equals its halves, and the
fusion + rerank stages are checked against the oracle fed with the GPU candidate lists.
"""
import os

import numpy as np
import pytest
import torch

from domain_rag_system_b200 import _lib, synth
from domain_rag_system_b200.indexing.bm25_index import Bm25DeviceIndex
from oracle import bm25 as obm25
from oracle import dense as odense
from oracle import fusion as ofusion
from oracle import rerank as orerank

pytestmark = pytest.mark.gpu
DEV = "cuda"
FULL = os.environ.get("DRS_FULLSIZE", "1") != "0"


def _sorted_unique(scores, ids):
    s, i = scores.cpu().numpy(), ids.cpu().numpy()
    assert (i >= 0).all()
    assert (np.diff(s.astype(np.float64), axis=1) <= 0).all(), "scores not non-increasing"
    tie = np.diff(s.astype(np.float64), axis=1) == 0
    assert (np.diff(i, axis=1)[tie] < 0).all(), "ties not ordered by larger row first"
    assert all(len(set(r.tolist())) == len(r) for r in i), "duplicate ids"


def _dense_properties(rows, inv, qv, kprime, sample):
    sc, ids, flags = _lib.dense_topk(qv, rows, inv, kprime, kernel=_lib.DENSE_UMMA)
    _sorted_unique(sc, ids)
    sc2, ids2, _ = _lib.dense_topk(qv, rows, inv, kprime, kernel=_lib.DENSE_UMMA)
    assert torch.equal(ids, ids2) and torch.equal(sc, sc2), "not deterministic"
    h = qv.shape[0] // 2
    sa, ia, _ = _lib.dense_topk(qv[:h].contiguous(), rows, inv, kprime, kernel=_lib.DENSE_UMMA)
    assert torch.equal(ia, ids[:h]) and torch.equal(sa, sc[:h]), "batch != its first half"
    # independent kernel: fp32 streaming path on a sample of the queries
    pick = torch.arange(0, qv.shape[0], max(1, qv.shape[0] // sample), device=qv.device)[:sample]
    ss, is_, _ = _lib.dense_topk(qv[pick].contiguous(), rows, inv, kprime, kernel=_lib.DENSE_SIMT)
    assert torch.equal(is_, ids[pick]), "tensor-core and fp32 kernels disagree on ids"
    assert torch.equal(ss, sc[pick]), "tensor-core and fp32 kernels disagree on scores"
    assert int((flags == 1).sum()) == 0
    return sc, ids, pick


def test_config2_one_million_chunks_1000_queries():
    n, q, vocab, top_k = 1_000_000, 1000, 50_000, 5
    kprime = 2 * top_k
    rows = synth.unit_embeddings(n, seed=1234, device=DEV, dup_frac=0.0001)
    inv = _lib.row_inv_norms(rows)
    qv = synth.query_embeddings(q, seed=4321, near_frac=0.0).to(DEV)
    v_sc, v_ids, pick = _dense_properties(rows, inv, qv, kprime, sample=16)
    # oracle on 8 of the queries (full corpus on the host)
    normed = odense.normalize_rows(rows.float().cpu().numpy())
    for i in pick[:8].tolist():
        ref = normed @ (qv[i].cpu().numpy() / np.linalg.norm(qv[i].cpu().numpy())).astype(np.float32)
        order = np.argsort(ref.astype(np.float64), kind="stable")[::-1][:kprime]
        got = v_ids[i].cpu().numpy()
        if not np.array_equal(got, order):  # only near-ties of the fp32 oracle may differ
            assert set(got) == set(order) or np.allclose(ref[got], ref[order], rtol=0, atol=4e-7)
        assert np.allclose(v_sc[i].cpu().numpy(), ref[got], rtol=1e-3)
    del normed
    offs, toks = synth.zipf_docs(n, vocab=vocab, seed=2025)
    qt = synth.zipf_queries(q, offs, toks, vocab=vocab, seed=777)
    ix = Bm25DeviceIndex.from_token_ids(offs, toks, vocab, device=DEV, build_device=DEV)
    ix0 = Bm25DeviceIndex.from_token_ids(offs, toks, vocab, device=DEV, build_device=DEV, dense_frac=0.0)
    assert ix.n_dense > 5 and ix0.n_dense == 0
    q_terms, q_offs = ix.encode_queries(qt)
    q_terms, q_offs = q_terms.to(DEV), q_offs.to(DEV)
    b_sc, b_ids = _lib.bm25_topk(ix.struct(), q_terms, q_offs, q, kprime)
    b_sc0, b_ids0 = _lib.bm25_topk(ix0.struct(), q_terms, q_offs, q, kprime)
    assert torch.equal(b_ids, b_ids0) and torch.equal(b_sc, b_sc0), "dense-row and postings-only paths disagree"
    _sorted_unique(b_sc, b_ids)
    orc = obm25.Bm25Csr(offs, toks, vocab)
    for i in range(0, q, 125):
        full = orc.get_scores(qt[i])
        order = obm25.stable_desc_order(full)[:kprime]
        assert np.array_equal(b_ids[i].cpu().numpy(), order)
        assert np.array_equal(b_sc[i].cpu().numpy(), full[order])           # fp64, bit-identical
    ids, sc, _ = _lib.fuse_topk("rrf", v_sc, v_ids, b_sc, b_ids, top_k)
    for i in range(0, q, 50):
        v = ofusion.as_dicts(v_ids[i].cpu().numpy(), v_sc[i].cpu().numpy().astype(np.float64))
        b = ofusion.as_dicts(b_ids[i].cpu().numpy(), b_sc[i].cpu().numpy())
        ref = ofusion.rrf_fuse(v, b, top_k)
        assert ids[i].cpu().tolist() == [r["row"] for r in ref]
        assert sc[i].cpu().tolist() == [r["score"] for r in ref]


@pytest.mark.skipif(not FULL, reason="DRS_FULLSIZE=0")
def test_config3_and_5_ten_million_chunks():
    n, q, vocab = 10_000_000, 1000, 50_000
    rows = synth.unit_embeddings(n, seed=1234, device=DEV)
    inv = _lib.row_inv_norms(rows)
    qv = synth.query_embeddings(q, seed=4321, near_frac=0.0).to(DEV)
    # config 3: weighted fusion + simple rerank top_n=30 => fetch_k = 30, k' = 60
    v_sc, v_ids, _ = _dense_properties(rows, inv, qv, 60, sample=8)
    offs, toks = synth.zipf_docs(n, vocab=vocab, seed=2025)
    qt = synth.zipf_queries(q, offs, toks, vocab=vocab, seed=777)
    ix = Bm25DeviceIndex.from_token_ids(offs, toks, vocab, device=DEV, build_device=DEV)
    q_terms, q_offs = ix.encode_queries(qt)
    q_terms, q_offs = q_terms.to(DEV), q_offs.to(DEV)
    b_sc, b_ids = _lib.bm25_topk(ix.struct(), q_terms, q_offs, q, 60)
    _sorted_unique(b_sc, b_ids)
    del ix
    torch.cuda.empty_cache()
    ix0 = Bm25DeviceIndex.from_token_ids(offs, toks, vocab, device=DEV, build_device=DEV, dense_frac=0.0)
    b_sc0, b_ids0 = _lib.bm25_topk(ix0.struct(), q_terms[: int(q_offs[64])].contiguous(), q_offs[:65].contiguous(), 64, 60)
    assert torch.equal(b_ids[:64], b_ids0) and torch.equal(b_sc[:64], b_sc0)
    for wv, wb in ((0.5, 0.5), (0.7, 0.3)):
        ids, sc, _ = _lib.fuse_topk("weighted", v_sc, v_ids, b_sc, b_ids, 30, w_vector=wv, w_bm25=wb)
        r_sc, r_ids, _ = _lib.rerank_cosine(rows, qv, ids, 5)
        for i in range(0, q, 100):
            v = ofusion.as_dicts(v_ids[i].cpu().numpy(), v_sc[i].cpu().numpy().astype(np.float64))
            b = ofusion.as_dicts(b_ids[i].cpu().numpy(), b_sc[i].cpu().numpy())
            ref = ofusion.weighted_fuse(v, b, wv, wb, 30)
            assert ids[i].cpu().tolist() == [r["row"] for r in ref]
            assert sc[i].cpu().tolist() == [r["score"] for r in ref]
            pool = rows[ids[i]].float().cpu().numpy()
            cs = np.array([orerank.cosine_f32_seq(qv[i].cpu().numpy(), p) for p in pool], dtype=np.float32)
            order = np.argsort(-cs.astype(np.float64), kind="stable")[:5]
            assert r_ids[i].cpu().tolist() == ids[i][order].cpu().tolist()
            assert np.array_equal(r_sc[i].cpu().numpy(), cs[order])
    # config 5: single query, rerank_top_n = 50 => fetch_k = 50, k' = 100 (fp32 streaming kernel)
    for i in (0, 7):
        s1, i1, _ = _lib.dense_topk(qv[i:i + 1].contiguous(), rows, inv, 100)
        assert _lib.load().drs_dense_last_kernel() == _lib.DENSE_SIMT
        s2, i2, _ = _lib.dense_topk(qv[i:i + 1].contiguous(), rows, inv, 100, kernel=_lib.DENSE_UMMA)
        assert torch.equal(i1, i2) and torch.equal(s1, s2)
        _sorted_unique(s1, i1)
