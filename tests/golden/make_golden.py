"""Generate the committed golden fixtures by running the REFERENCE'S OWN classes
(imported unmodified from /root/reference/src over the stand-in modules in
oracle/stubs) in the build container.  Run from the repo root:

    python tests/golden/make_golden.py

Outputs (all small JSON): fusion_golden.json, rerank_golden.json,
tokenizer_golden.json, hybrid_golden.json.  The GPU box has no /root/reference;
tests there read only these files.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from domain_rag_system_b200 import synth  # noqa: E402
from oracle import ref_driver  # noqa: E402
from oracle.embedder import TableEmbedder  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# Parameters of the small hybrid corpus; tests regenerate the inputs from these.
HYB = dict(n=600, vocab=400, emb_seed=1234, doc_seed=2025, q=24, q_seed=4321, qt_seed=777,
           dup_frac=0.02)


def hybrid_inputs(p=HYB):
    emb = synth.unit_embeddings(p["n"], seed=p["emb_seed"], dup_frac=p["dup_frac"]).float().numpy()
    offs, toks = synth.zipf_docs(p["n"], vocab=p["vocab"], seed=p["doc_seed"])
    qv = synth.query_embeddings(p["q"], seed=p["q_seed"], near_frac=0.0).numpy()
    # half of the queries are rounded-to-bf16 copies of corpus rows + noise: realistic near hits
    qt = synth.zipf_queries(p["q"], offs, toks, vocab=p["vocab"], seed=p["qt_seed"])
    return emb, offs, toks, qv, qt


def _digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def _dump(name, obj):
    with open(os.path.join(OUT, name), "w") as f:
        json.dump(obj, f, separators=(",", ":"))
    print("wrote", name)


def fusion_cases():
    rng = np.random.default_rng(5)
    cases = []

    def mk(ids, scores):
        return [{"chunk_id": i, "score": float(s), "text": f"t{i}"} for i, s in zip(ids, scores)]

    # the SURVEY probes
    cases.append(dict(v=mk(list("abcd"), [.9, .8, .7, .6]), b=mk(list("xbya"), [12, 9, 5, 1]),
                      top_k=10, wv=0.7, wb=0.3))
    # duplicates inside one list accumulate; last vector occurrence is the representative
    cases.append(dict(v=mk(["a", "b", "a", "c"], [.9, .8, .7, .6]), b=mk(["c", "c", "d"], [3, 2, 1]),
                      top_k=10, wv=0.5, wb=0.5))
    # empty sides
    cases.append(dict(v=[], b=mk(["p", "q"], [2.0, 1.0]), top_k=5, wv=0.5, wb=0.5))
    cases.append(dict(v=mk(["p", "q"], [0.3, 0.1]), b=[], top_k=5, wv=0.5, wb=0.5))
    cases.append(dict(v=[], b=[], top_k=5, wv=0.5, wb=0.5))
    # max == min: raw scores are used
    cases.append(dict(v=mk(["a", "b"], [.5, .5]), b=mk(["b", "c", "d"], [0.0, 0.0, 0.0]),
                      top_k=4, wv=0.6, wb=0.4))
    cases.append(dict(v=mk(["a"], [.42]), b=mk(["z"], [7.5]), top_k=3, wv=0.5, wb=0.5))
    # random, with overlap and truncation
    for n in (10, 20, 60, 100):
        ids_v = [f"c{j}" for j in rng.choice(3 * n, n, replace=False)]
        ids_b = [f"c{j}" for j in rng.choice(3 * n, n, replace=False)]
        sv = np.sort(rng.random(n).astype(np.float32))[::-1]
        sb = np.sort(rng.random(n) * 20)[::-1]
        sb[n // 2:] = 0.0  # tail of zero-score BM25 hits
        cases.append(dict(v=mk(ids_v, sv), b=mk(ids_b, sb), top_k=n // 2,
                          wv=float(rng.random()), wb=float(rng.random())))
    return cases


def main():
    ref = ref_driver.load()
    with ref_driver.quiet():
        hyb = ref.HybridIndexer(None, None, storage_path="/tmp/drs_golden_idx")

    # ---- fusion ---------------------------------------------------------
    out = []
    for c in fusion_cases():
        import copy
        rrf = hyb._reciprocal_rank_fusion(copy.deepcopy(c["v"]), copy.deepcopy(c["b"]), top_k=c["top_k"])
        wf = hyb._weighted_fusion(copy.deepcopy(c["v"]), copy.deepcopy(c["b"]), vector_weight=c["wv"],
                                  bm25_weight=c["wb"], top_k=c["top_k"])
        out.append(dict(case=c,
                        rrf=[[r["chunk_id"], r["score"], r["text"]] for r in rrf],
                        weighted=[[r["chunk_id"], r["score"], r["text"], r.get("normalized_score")]
                                  for r in wf]))
    # known-answer values stored in the reference's results/*.json
    known = {"1/61": 0.01639344262295082, "1/61+1/64": 0.032018442622950824,
             "dup_accumulation_gt_2/61": 0.0694912486938349}
    _dump("fusion_golden.json", dict(cases=out, known_answer=known))

    # ---- tokenizer ------------------------------------------------------
    with ref_driver.quiet():
        bi = ref.BM25Indexer()
    probes = ["Pt w/ T2DM & HbA1c=7.5% on ACE-inhibitor; a b I.V. 5mg",
              "What are the side effects of aspirin?", "", "a I x", "naïve café Ünïcode_word 12",
              "w00001 w00002 zzzz", "MI-CHF/HTN: BP 120/80 mmHg (n=3)"]
    _dump("tokenizer_golden.json", [[p, bi.tokenize(p)] for p in probes])

    # ---- rerank ---------------------------------------------------------
    cos = ref.SimpleSimilarityReranker._cosine
    rng = np.random.default_rng(11)
    pairs = []
    for i in range(6):
        a = rng.standard_normal(768).astype(np.float32)
        b = rng.standard_normal(768).astype(np.float32)
        if i == 1:
            b = a.copy()
        if i == 2:
            b[:] = 0
        if i >= 3:  # bf16-representable rows, unit-ish, like the stored corpus
            import torch
            a /= np.linalg.norm(a)
            b = torch.from_numpy(b / np.linalg.norm(b)).to(torch.bfloat16).float().numpy()
        v = cos(a, b)
        pairs.append(dict(seed_index=i, a=[float(x) for x in a], b=[float(x) for x in b],
                          cos=float(v), cos_type=type(v).__name__))

    class _Fixed(ref.BaseReranker):
        def __init__(self, scores, top_n):
            super().__init__(top_n=top_n)
            self._s = scores

        def score_pairs(self, pairs_):
            return self._s[: len(list(pairs_))]

    rr_cases = []
    for top_n, top_k, n in ((5, 3, 8), (50, 10, 6), (3, 5, 3), (4, 2, 0)):
        res = [{"chunk_id": f"c{j}", "text": f"t{j}", "score": 1.0 / (j + 1), "fusion_method": "rrf" if j % 2 else None}
               for j in range(n)]
        sc = [0.5, 0.9, 0.5, 0.1, 0.9, 0.3, 0.2, 0.8][:n]
        got = _Fixed(sc, top_n).rerank("q", res, top_k=top_k)
        rr_cases.append(dict(results=res, scores=sc, top_n=top_n, top_k=top_k,
                             out=[[r["chunk_id"], r.get("rerank_score"), r.get("fusion_method")] for r in got]))
    _dump("rerank_golden.json", dict(cosine=pairs, rerank=rr_cases))

    # ---- full hybrid path through the reference classes -----------------
    emb, offs, toks, qv, qt = hybrid_inputs()
    texts = synth.docs_as_text(offs, toks)
    chunks = synth.make_chunks(texts)
    # non-unique chunk ids on a few rows (the real corpus has them)
    for r in (5, 17, 301):
        chunks[r]["chunk_id"] = chunks[r - 1]["chunk_id"]
    te = TableEmbedder(dim=768)
    te.add(texts, emb)
    qtexts = [" ".join(synth.term_string(t) for t in q) for q in qt]
    te.add(qtexts, qv)
    with ref_driver.quiet():
        vi = ref.VectorIndexer(collection_name="g", embedder=te, use_in_memory=True)
        bi = ref.BM25Indexer()
        hi = ref.HybridIndexer(vi, bi, storage_path="/tmp/drs_golden_idx")
        hi.index_chunks(chunks)
    rows_of = {}
    for i, c in enumerate(chunks):
        rows_of.setdefault(c["chunk_id"], i)

    def pack(results, key="score"):
        return [[r["chunk_id"], r["metadata"]["row"], r[key]] for r in results]

    cases = []
    with ref_driver.quiet():
        for qi, qtext in enumerate(qtexts):
            for (top_k, method, wv, wb, top_n, final_k) in ((5, "rrf", .5, .5, 0, 5),
                                                            (30, "weighted", .5, .5, 30, 5),
                                                            (10, "weighted", .7, .3, 0, 10),
                                                            (50, "rrf", .5, .5, 50, 5)):
                vres = vi.search(qtext, top_k=top_k * 2)
                bres = bi.search(qtext, top_k=top_k * 2)
                fused = hi.search(qtext, top_k=top_k, fusion_method=method, vector_weight=wv, bm25_weight=wb)
                entry = dict(q=qi, top_k=top_k, method=method, wv=wv, wb=wb, top_n=top_n, final_k=final_k,
                             vector=pack(vres), bm25=pack(bres), fused=pack(fused))
                if top_n:
                    rr = ref.build_reranker("simple", top_n=top_n, embedder=te)
                    out_rr = rr.rerank(qtext, fused, top_k=final_k)
                    entry["reranked"] = pack(out_rr, "rerank_score")
                cases.append(entry)
        stats = hi.get_stats()
    _dump("hybrid_golden.json", dict(params=HYB, digest=_digest(emb, offs, toks, qv, np.concatenate(qt)),
                                     dup_rows=[5, 17, 301], stats=stats, cases=cases,
                                     idf_probe={t: bi.bm25.idf.get(t) for t in ("w00000", "w00001", "w00050", "w00399")},
                                     avgdl=bi.bm25.avgdl, average_idf=bi.bm25.average_idf))


if __name__ == "__main__":
    main()
