"""Golden fixtures for payload filters, score thresholds, re-indexing (upsert) and the euclidean / dot metrics:
outputs of the REFERENCE'S OWN classes (imported unmodified from /root/reference/src over the stand-in modules in
oracle/stubs), run in the build container:

    python tests/golden/make_filters_golden.py        ->  tests/golden/filters_golden.json

Covers `filters=` on all three indexers (src/indexing/vector_indexer.py:185-200 pre-filter, bm25_indexer.py:140-147
post-filter of the 2*top_k window, hybrid_indexer.py:81-92), `score_threshold` (vector_indexer.py:198-200), a second
`index_chunks` on an existing collection (vector_indexer.py:126-135: ids restart at 0, so it overwrites) and
`distance_metric` (vector_indexer.py:57-62).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from domain_rag_system_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
P = dict(n=900, vocab=300, emb_seed=41, doc_seed=42, q=8, q_seed=43, qt_seed=44)
SOURCES = ("pubmed", "fda", "kaggle")


def inputs(p=P):
    emb = synth.unit_embeddings(p["n"], seed=p["emb_seed"]).float().numpy()
    emb *= (1.0 + 0.5 * np.sin(np.arange(p["n"], dtype=np.float32)))[:, None] ** 2 + 0.25   # non-unit rows (dot / euclid)
    emb = __import__("torch").from_numpy(emb).to(__import__("torch").bfloat16).float().numpy()
    offs, toks = synth.zipf_docs(p["n"], vocab=p["vocab"], seed=p["doc_seed"])
    qv = synth.query_embeddings(p["q"], seed=p["q_seed"], near_frac=0.0).numpy()
    qt = synth.zipf_queries(p["q"], offs, toks, vocab=p["vocab"], seed=p["qt_seed"])
    return emb, offs, toks, qv, qt


def chunks_of(texts):
    chunks = synth.make_chunks(texts)
    for i, c in enumerate(chunks):
        c["source"] = SOURCES[i % 3]
        c["chunk_type"] = "title" if i % 5 == 0 else "body"
    return chunks


def pack(results, key="score"):
    return [[r["chunk_id"], r["metadata"]["row"], r[key], r["source"], r["chunk_type"]] for r in results]


FILTERS = [{"source": "fda"}, {"source": "fda", "chunk_type": "title"}, {"source": "nope"}, {"doc_id": "doc_3"},
           {"chunk_type": "body"}]


def main():
    from oracle import ref_driver          # (needs /root/reference: only the generator imports it)
    from oracle.embedder import TableEmbedder
    ref = ref_driver.load()
    emb, offs, toks, qv, qt = inputs()
    texts = synth.docs_as_text(offs, toks)
    chunks = chunks_of(texts)
    te = TableEmbedder(dim=768)
    te.add(texts, emb)
    qtexts = [" ".join(synth.term_string(t) for t in q) for q in qt]
    te.add(qtexts, qv)
    out = dict(params=P, filters=FILTERS, cases=[], metrics={}, upsert={})
    with ref_driver.quiet():
        vi = ref.VectorIndexer(collection_name="f", embedder=te, use_in_memory=True)
        bi = ref.BM25Indexer()
        hi = ref.HybridIndexer(vi, bi, storage_path="/tmp/drs_golden_idx_f")
        hi.index_chunks(chunks)
        for qi, qtext in enumerate(qtexts):
            for fi, flt in enumerate(FILTERS):
                entry = dict(q=qi, f=fi,
                             vector=pack(vi.search(qtext, top_k=10, filters=flt)),
                             bm25=pack(bi.search(qtext, top_k=10, filters=flt)),
                             hybrid_rrf=pack(hi.search(qtext, top_k=5, filters=flt)),
                             hybrid_weighted=pack(hi.search(qtext, top_k=5, filters=flt, fusion_method="weighted",
                                                            vector_weight=0.7, bm25_weight=0.3)))
                if fi == 0:   # the reference's default rerank pool: top_k = 50 => BM25 window of 200 with filters
                    entry["hybrid_top50"] = pack(hi.search(qtext, top_k=50, filters=flt))
                    entry["bm25_top100"] = pack(bi.search(qtext, top_k=100, filters=flt))
                out["cases"].append(entry)
            top13 = vi.search(qtext, top_k=13)                        # between the 12th and the 13th best score:
            thr = 0.5 * (top13[11]["score"] + top13[12]["score"])     # cuts a top-40 request short after 12 hits
            out["cases"].append(dict(q=qi, f=None, threshold=thr,
                                     vector_thr=pack(vi.search(qtext, top_k=40, score_threshold=thr)),
                                     vector_thr_filtered=pack(vi.search(qtext, top_k=40, filters=FILTERS[0],
                                                                        score_threshold=thr)),
                                     vector_thr_none=pack(vi.search(qtext, top_k=5, score_threshold=2.0))))
        # ---- a second index_chunks on the existing collection: points 0..49 are overwritten, 50.. stay
        again = [dict(c) for c in chunks[100:150]]
        for c in again:
            c["chunk_id"] = "re_" + c["chunk_id"]
        vi.index_chunks(again)
        out["upsert"] = dict(points_count=vi.get_stats()["points_count"],
                             searches=[pack(vi.search(qtext, top_k=10)) for qtext in qtexts[:4]])
        # ---- other metrics
        for metric in ("dot", "euclidean"):
            vm = ref.VectorIndexer(collection_name="m_" + metric, embedder=te, use_in_memory=True, distance_metric=metric)
            vm.create_collection(force_recreate=True)
            vm.index_chunks(chunks)
            res = []
            for qtext in qtexts:
                full = vm.search(qtext, top_k=10)
                thr = 0.5 * (full[5]["score"] + full[6]["score"])
                res.append(dict(top=pack(full), thr=thr, top_thr=pack(vm.search(qtext, top_k=10, score_threshold=thr)),
                                filtered=pack(vm.search(qtext, top_k=10, filters=FILTERS[1]))))
            out["metrics"][metric] = res
    with open(os.path.join(OUT, "filters_golden.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote filters_golden.json", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
