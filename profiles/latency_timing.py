#!/usr/bin/env python
"""Latency-mode timing driver (BASELINE configs[4]): one query against n chunks, k' = 100, simple rerank of 50.

    python profiles/latency_timing.py [n_chunks] [calls]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from domain_rag_system_b200 import _lib, synth                      # noqa: E402
from domain_rag_system_b200.dist import ShardedHybridSearcher       # noqa: E402
from domain_rag_system_b200.indexing import Bm25DeviceIndex         # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    calls = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    dev = "cuda:0"
    rows, offs, toks, _ = synth.shard_corpus(n, 1, 0, device=dev)
    inv = _lib.row_inv_norms(rows)
    ix = Bm25DeviceIndex.from_token_ids(offs, toks, synth.VOCAB, device=dev)
    del offs, toks
    qv, qt = synth.block_queries(1, n)
    q_terms, q_offs = ix.encode_queries(qt)
    qv, q_terms, q_offs = qv.to(dev), q_terms.to(dev), q_offs.to(dev)
    s = ShardedHybridSearcher(rows, inv, ix, 0, world=1)
    kw = dict(top_k=50, fusion_method="weighted", rerank_top_n=50, final_k=10)
    ref = None
    for v in os.environ.get("STREAM_AB", "").split(","):      # STREAM_AB=0,3,4,6: 0 = load-then-score kernel, n = ring stages
        if v == "":
            continue
        os.environ["DRS_DENSE_STREAM"] = "0" if v == "0" else "1"
        os.environ["DRS_STREAM_STAGES"] = v if v != "0" else "3"
        for _ in range(3):
            out = s.search_batch(qv, q_terms, q_offs, **kw)
        torch.cuda.synchronize()
        same = None if ref is None else bool(torch.equal(ref[0], out[0]) and torch.equal(ref[1], out[1]))
        ref = ref or (out[0].clone(), out[1].clone())
        _lib.profile_enable(True)
        for _ in range(50):
            s.search_batch(qv, q_terms, q_offs, **kw)
        pr = _lib.profile_read()
        _lib.profile_enable(False)
        print(json.dumps({"dense_stream": v, "dense_select_ms": round(pr["dense_select"][0] / pr["dense_select"][1], 4),
                          "same_result_as_first": same}))
    os.environ.pop("DRS_DENSE_STREAM", None)
    os.environ.pop("DRS_STREAM_STAGES", None)
    for _ in range(5):
        s.search_batch(qv, q_terms, q_offs, **kw)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(calls)]
    _lib.profile_enable(True)
    for a, b in ev:
        a.record()
        s.search_batch(qv, q_terms, q_offs, **kw)
        b.record()
    torch.cuda.synchronize()
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    ms = np.array([a.elapsed_time(b) for a, b in ev])
    print(json.dumps({"n_chunks": n, "calls": calls, "device_ms_p50": float(np.percentile(ms, 50)),
                      "device_ms_p99": float(np.percentile(ms, 99)),
                      "hbm_floor_ms": 2.0 * n * rows.shape[1] / 6.542e12 * 1e3,
                      "kernel_ms": {k: round(v[0] / max(v[1], 1), 4) for k, v in prof.items() if v[1]}}))


if __name__ == "__main__":
    main()
