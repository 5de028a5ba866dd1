#!/usr/bin/env python
"""Timing driver for the BM25 kernels alone (no dense side): the bench workload's documents and queries,
`drs_bm25_topk` timed with the library's CUDA-event hooks.

    python profiles/bm25_timing.py [n_docs] [n_queries] [k'] [iters]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from domain_rag_system_b200 import _lib, synth                      # noqa: E402
from domain_rag_system_b200.indexing import Bm25DeviceIndex         # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    q = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    kprime = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    iters = int(sys.argv[4]) if len(sys.argv) > 4 else 20
    dev = "cuda:0"
    _, offs, toks, _ = synth.shard_corpus(n, 1, 0, dim=8, device=dev)
    ix = Bm25DeviceIndex.from_token_ids(offs, toks, synth.VOCAB, device=dev)
    _, qt = synth.block_queries(q, n, dim=8)
    q_terms, q_offs = ix.encode_queries(qt)
    q_terms, q_offs = q_terms.to(dev), q_offs.to(dev)
    ws = _lib.Workspace()
    for _ in range(3):
        sc, ids, flags = _lib.bm25_topk(ix.struct(), q_terms, q_offs, q, kprime, ws=ws, want_flags=True)
    torch.cuda.synchronize()
    ab = {}
    for dbg in os.environ.get("BM25_AB", "").split(","):      # BM25_AB=0,1,2,...: DRS_BM25_DEBUG values, one after the other
        if dbg == "":
            continue
        os.environ["DRS_BM25_DEBUG"] = dbg
        _lib.profile_enable(True)
        for _ in range(iters):
            _lib.bm25_topk(ix.struct(), q_terms, q_offs, q, kprime, ws=ws, mode="no_escalation")
        ab.setdefault(dbg, []).append(round(_lib.profile_read()["bm25_select"][0] / iters, 4))
        _lib.profile_enable(False)
    if ab:
        print(json.dumps({"bm25_select_ms_by_debug_flags": ab}))
        os.environ.pop("DRS_BM25_DEBUG", None)
    ab = {}
    for cps in os.environ.get("BM25_CTAS_AB", "").split(","):  # BM25_CTAS_AB=4,5,4,5: resident CTAs per SM, interleaved
        if cps == "":
            continue
        os.environ["DRS_BM25_CTAS"] = cps
        for _ in range(2):
            out = _lib.bm25_topk(ix.struct(), q_terms, q_offs, q, kprime, ws=ws)
        _lib.profile_enable(True)
        for _ in range(iters):
            _lib.bm25_topk(ix.struct(), q_terms, q_offs, q, kprime, ws=ws)
        pr = _lib.profile_read()
        ab.setdefault(cps, []).append((round(pr["bm25_select"][0] / iters, 4), int(out[1].sum().item())))
        _lib.profile_enable(False)
    if ab:
        print(json.dumps({"bm25_select_ms_and_checksum_by_ctas_per_sm": ab}))
        os.environ.pop("DRS_BM25_CTAS", None)
    _lib.profile_enable(True)
    for _ in range(iters):
        _lib.bm25_topk(ix.struct(), q_terms, q_offs, q, kprime, ws=ws)
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    fl = np.bincount(flags.cpu().numpy(), minlength=3).tolist()
    print(json.dumps({"n_docs": n, "queries": q, "kprime": kprime,
                      "ms": {k: round(v[0] / max(v[1], 1), 4) for k, v in prof.items() if v[1]},
                      "flags_proven_unproven_escalated": fl,
                      "checksum": int(ids.sum().item())}))


if __name__ == "__main__":
    main()
