"""Reduced parity subset for compute-sanitizer (memcheck / racecheck / synccheck), small shapes, every kernel:

    compute-sanitizer --tool memcheck python tests/sanitize_subset.py

Dense: the CTA-pair tensor kernel, the single-CTA tensor kernel (1 and 3 issuing warps, CTA-pair TS variant), the
fp32 streaming kernel; BM25 in its three exactness modes; fusion, rerank, the sharded-path kernels.  Each result is
compared with the CPU oracle, so a sanitizer-clean run is also a parity run.  Summaries: profiles/sanitizer_r2.txt."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from domain_rag_system_b200 import _lib, synth  # noqa: E402
from domain_rag_system_b200.dist import _Layout  # noqa: E402
from domain_rag_system_b200.indexing.bm25_index import Bm25DeviceIndex  # noqa: E402
from oracle import bm25 as obm25  # noqa: E402
from oracle import dense as odense  # noqa: E402
from oracle import fusion as ofusion  # noqa: E402

DEV = "cuda"


def dense_case(n, q, kp, kernel, name):
    emb = synth.unit_embeddings(n, seed=n % 89 + q, dup_frac=0.002)
    qv = synth.query_embeddings(q, corpus=emb, seed=q + 7)
    rows = emb.to(DEV)
    inv = _lib.row_inv_norms(rows)
    sc, ids, fl = _lib.dense_topk(qv.to(DEV).contiguous(), rows, inv, kp, kernel=kernel)
    torch.cuda.synchronize()
    ref = odense.cosine_scores_batch(odense.normalize_rows(emb.float().numpy()), qv.numpy().astype(np.float64))
    bad = 0
    for i in range(q):
        want, _ = odense.dense_search(ref[i], kp)
        got = ids[i].cpu().numpy()
        if not np.array_equal(got, want):
            bad += not np.allclose(ref[i][got], ref[i][want], rtol=0, atol=1e-6)
    print(f"[sanitize] dense {name}: n={n} q={q} k'={kp} mismatches beyond near-ties: {bad}", flush=True)
    assert bad == 0


def main():
    which = os.environ.get("DRS_SANITIZE_PART", "all")
    if which in ("all", "dense"):
        dense_case(6000, 300, 10, _lib.DENSE_UMMA_PAIR, "pair")
        dense_case(9000, 130, 60, _lib.DENSE_UMMA_PAIR, "pair k'=60")
        dense_case(5000, 100, 10, _lib.DENSE_UMMA_TS, "single-CTA (3 issuers)")
        dense_case(5000, 100, 60, _lib.DENSE_UMMA_TS, "single-CTA (1 issuer)")
        dense_case(3000, 3, 20, _lib.DENSE_SIMT, "fp32 streaming")
    if which in ("all", "rest"):
        n, vocab, nq, kp = 6000, 500, 40, 10
        offs, toks = synth.zipf_docs(n, vocab=vocab, seed=3)
        qt = synth.zipf_queries(nq, offs, toks, vocab=vocab, seed=4)
        ix = Bm25DeviceIndex.from_token_ids(offs, toks, vocab, device=DEV, tile_docs=1024)
        orc = obm25.Bm25Csr(offs, toks, vocab)
        q_terms, q_offs = [t.to(DEV) for t in ix.encode_queries(qt)]
        for mode in ("auto", "no_escalation", "escalate_all"):
            b_sc, b_ids, fl = _lib.bm25_topk(ix.struct(), q_terms, q_offs, nq, kp, mode=mode, want_flags=True)
            torch.cuda.synchronize()
            for i in range(nq):
                if mode == "no_escalation" and int(fl[i]) == 1:
                    continue
                r, s = obm25.bm25_search(orc.get_scores(qt[i]), kp)
                assert np.array_equal(b_ids[i].cpu().numpy(), r) and np.array_equal(b_sc[i].cpu().numpy(), s)
            print(f"[sanitize] bm25 mode {mode}: ok", flush=True)
        emb = synth.unit_embeddings(n, seed=5)
        qv = synth.query_embeddings(nq, corpus=emb, seed=6)
        rows = emb.to(DEV)
        inv = _lib.row_inv_norms(rows)
        v_sc, v_ids, _ = _lib.dense_topk(qv.to(DEV).contiguous(), rows, inv, kp)
        for method in ("rrf", "weighted"):
            ids, sc, _ = _lib.fuse_topk(method, v_sc, v_ids, b_sc, b_ids, 5)
            for i in range(0, nq, 7):
                v = ofusion.as_dicts(v_ids[i].cpu().numpy(), v_sc[i].cpu().numpy().astype(np.float64))
                b = ofusion.as_dicts(b_ids[i].cpu().numpy(), b_sc[i].cpu().numpy())
                ref = ofusion.rrf_fuse(v, b, 5) if method == "rrf" else ofusion.weighted_fuse(v, b, 0.5, 0.5, 5)
                assert ids[i].cpu().tolist() == [r["row"] for r in ref]
        r_sc, r_ids, r_pos = _lib.rerank_cosine(rows, qv.to(DEV).contiguous(), ids, 3)
        lay = _Layout(nq, kp)
        gathered = torch.zeros(2 * lay.nbytes, dtype=torch.uint8, device=DEV)
        id_off = torch.tensor([0, n // 2], dtype=torch.int64, device=DEV)
        for p in range(2):
            lo, hi = p * (n // 2), (p + 1) * (n // 2)
            v = lay.views(gathered[p * lay.nbytes:(p + 1) * lay.nbytes])
            s, i_, _ = _lib.dense_topk(qv.to(DEV).contiguous(), rows[lo:hi].contiguous(), inv[lo:hi].contiguous(), kp)
            v["v_sc"].copy_(s), v["v_ids"].copy_(i_)
            _lib.candidate_cosines(rows[lo:hi].contiguous(), qv.to(DEV).contiguous(), v["v_ids"], out=v["v_cos"])
        m_sc, m_ids, m_cos = _lib.merge_gathered(gathered, 2, lay.nbytes, lay.off_vsc, False, lay.off_vid, lay.off_vcos,
                                                 id_off, nq, kp, kp)
        assert torch.equal(m_ids, v_ids)
        _lib.rerank_from_cosines(m_ids[:, :5].contiguous(), 3, m_ids, m_cos, None, None)
        torch.cuda.synchronize()
        print("[sanitize] fuse / rerank / sharded-path kernels: ok", flush=True)
    print("SANITIZE_SUBSET_OK", flush=True)


if __name__ == "__main__":
    main()
