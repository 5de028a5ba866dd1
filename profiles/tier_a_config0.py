"""Tier-A CPU timing of BASELINE configs[0] (SURVEY.md 8d "Config 1"): the REFERENCE'S OWN classes
(HybridIndexer / VectorIndexer / BM25Indexer imported unmodified from /root/reference/src) over the stand-in
rank_bm25 / qdrant_client modules of oracle/stubs, 10 000 synthetic chunks x 768-d, top_k = 5, RRF, one query at a
time: 5 warm-ups, then 50 sequential queries; median ms and 1/median queries/s.

/root/reference exists only in the build container, so this runs THERE (CPU) and the result is committed as
profiles/tier_a_config0.json:      python profiles/tier_a_config0.py
(The GPU box times the tier-B port of the same path at 10 M chunks: bench.py `cpu_baseline` / `--impl reference`.)
"""
import json
import os
import platform
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from domain_rag_system_b200 import synth  # noqa: E402
from oracle import ref_driver  # noqa: E402
from oracle.embedder import TableEmbedder  # noqa: E402


def main():
    n, vocab, nq = 10_000, 50_000, 55
    ref = ref_driver.load()
    emb = synth.unit_embeddings(n, seed=1234).float().numpy()
    offs, toks = synth.zipf_docs(n, vocab=vocab, seed=2025)
    qv = synth.query_embeddings(nq, seed=4321, near_frac=0.0).numpy()
    qt = synth.zipf_queries(nq, offs, toks, vocab=vocab, seed=777)
    texts = synth.docs_as_text(offs, toks)
    chunks = synth.make_chunks(texts)
    te = TableEmbedder(dim=768)
    te.add(texts, emb)
    qtexts = [" ".join(synth.term_string(t) for t in q) for q in qt]
    te.add(qtexts, qv)
    with ref_driver.quiet():
        vi = ref.VectorIndexer(collection_name="c0", embedder=te, use_in_memory=True)
        bi = ref.BM25Indexer()
        hi = ref.HybridIndexer(vi, bi, storage_path="/tmp/drs_tier_a_c0")
        t0 = time.time()
        hi.index_chunks(chunks)
        build_s = time.time() - t0
        for q in qtexts[:5]:
            hi.search(q, top_k=5, fusion_method="rrf")
        ms = []
        for q in qtexts[5:55]:
            t = time.perf_counter()
            out = hi.search(q, top_k=5, fusion_method="rrf")
            ms.append(1e3 * (time.perf_counter() - t))
            assert len(out) == 5
    med = statistics.median(ms)
    try:
        from threadpoolctl import threadpool_info
        blas = [{k: d.get(k) for k in ("internal_api", "num_threads")} for d in threadpool_info()]
    except Exception:
        blas = os.environ.get("OMP_NUM_THREADS")
    cpu = ""
    try:
        cpu = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        pass
    res = {"config": "BASELINE configs[0]: 10k synthetic chunks x 768-d unit embeddings + BM25 corpus (Zipf s=1.07, vocab 50k, "
                     "doc length U[32,96]), top_k=5 (k'=10), RRF, 1 query at a time",
           "implementation": "tier A: the reference's own HybridIndexer / VectorIndexer / BM25Indexer classes over the "
                             "stand-in rank_bm25 (per-document dicts, Python loops) and qdrant_client local mode (NumPy)",
           "queries_timed": len(ms), "warmups": 5, "median_ms": med, "p10_ms": float(np.percentile(ms, 10)),
           "p90_ms": float(np.percentile(ms, 90)), "queries_per_s": 1e3 / med, "index_build_s": build_s,
           "host": {"cpu": cpu, "os_cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0)), "blas": blas,
                    "python": platform.python_version(), "numpy": np.__version__},
           "where": "build container (CPU only); /root/reference is not available on the GPU box"}
    with open(os.path.join(ROOT, "profiles", "tier_a_config0.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
