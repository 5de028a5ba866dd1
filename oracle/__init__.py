"""CPU oracle for the hybrid-retrieval hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithm the reference
(Luyu-H/Domain-RAG-System) runs for ``HybridIndexer.search`` +
``BaseReranker.rerank``.  It exists so the CUDA path can be checked; it is never
the thing shipped or measured.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The
product package (``domain_rag_system_b200``) must not, and does not.

Provenance of the arithmetic:

* The reference's own first-party logic (fusion, over-fetch, post-filter, rerank,
  tokenizer) is restated from ``/root/reference/src`` with file:line citations on
  every function, and is PINNED: ``tests/golden/*.json`` were produced by running
  the reference's unmodified classes in the build container
  (``tests/golden/make_golden.py``), and ``tests/test_oracle_golden.py`` checks the
  oracle against them, together with the RRF known-answer values stored in the
  reference's ``results/*.json``.
* The two third-party providers are absent from ``/root/reference`` and from
  this image, so they are restated from their published algorithm:
  ``rank_bm25`` (requirements.txt:7, ``>=0.2.1``; restated after 0.2.2
  ``BM25Okapi``) and ``qdrant-client`` local mode (requirements.txt:6,
  ``>=1.7.0``; ``query_points`` needs >=1.10).  No test of the reference pins
  numbers at either boundary => **parity unpinned** for raw BM25 scores and raw
  cosine scores; parity for those is defined against this restatement.
"""
