"""Multi-GPU parity check, launched with torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
        tests/dist_gpu_check.py

Every rank builds the SAME small corpus, keeps its row shard, and runs the sharded hybrid search (local
top-k' -> all-gather of candidate lists over NCCL -> in-place merge -> fusion -> rerank from the carried cosines),
eagerly and as a replayed CUDA graph.  Rank 0 also runs the unsharded path on the whole corpus; everything must
agree bit for bit (ids, fused scores, rerank scores)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from domain_rag_system_b200 import _lib, synth  # noqa: E402
from domain_rag_system_b200.dist import ShardedHybridSearcher, shard_bounds  # noqa: E402
from domain_rag_system_b200.indexing.bm25_index import Bm25DeviceIndex  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import datetime
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=90))
    n, vocab, nq, top_k = 60_000, 3000, 200, 10
    emb = synth.unit_embeddings(n, seed=1234, dup_frac=0.001)
    offs, toks = synth.zipf_docs(n, vocab=vocab, seed=2025)
    qv = synth.query_embeddings(nq, corpus=emb, seed=4321)
    qt = synth.zipf_queries(nq, offs, toks, vocab=vocab, seed=777)

    lo, hi = shard_bounds(n, world, rank)
    rows = emb[lo:hi].to(dev)
    inv = _lib.row_inv_norms(rows)
    ix = Bm25DeviceIndex.from_token_ids(offs[lo:hi + 1] - offs[lo], toks[offs[lo]:offs[hi]], vocab, device=dev,
                                        build_device=dev, group=True)
    q_terms, q_offs = ix.encode_queries(qt)
    searcher = ShardedHybridSearcher(rows, inv, ix, row_offset=lo)
    ok = True
    qd, qtd, qod = qv.to(dev), q_terms.to(dev), q_offs.to(dev)
    full = None
    if rank == 0:
        full_rows = emb.to(dev)
        full = ShardedHybridSearcher(full_rows, _lib.row_inv_norms(full_rows),
                                     Bm25DeviceIndex.from_token_ids(offs, toks, vocab, device=dev, build_device=dev),
                                     row_offset=0, world=1)
    for method, top_n in (("rrf", 0), ("weighted", 0), ("weighted", 8), ("rrf", 10)):
        kw = dict(top_k=top_k, fusion_method=method, vector_weight=0.7, bm25_weight=0.3, rerank_top_n=top_n, final_k=5)
        ids, sc = searcher.search_batch(qd, qtd, qod, **kw)
        torch.cuda.synchronize()
        print(f"[dist check r{rank}] eager step done ({method}, rerank {top_n})", file=sys.stderr, flush=True)
        run = searcher.graphed(qd, qtd, qod, **kw)
        ids_g, sc_g = run()
        torch.cuda.synchronize()
        print(f"[dist check r{rank}] graph replay done", file=sys.stderr, flush=True)
        graph_ok = bool(torch.equal(ids, ids_g)) and bool(torch.equal(sc, sc_g))
        if rank == 0:
            ids1, sc1 = full.search_batch(qd, qtd, qod, **kw)
            same_ids = bool(torch.equal(ids, ids1))
            same_sc = bool(torch.equal(sc, sc1))
            print(f"[dist check] world={world} {method} rerank_top_n={top_n}: ids equal to single-GPU = {same_ids}, "
                  f"scores bit-equal = {same_sc}, graph replay equal = {graph_ok}")
            ok = ok and same_ids and same_sc
        ok = ok and graph_ok
    # every rank holds the same answer
    t = torch.tensor([int(ids.sum().item())], device=dev)
    lst = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(lst, t)
    if rank == 0:
        agree = all(int(x) == int(lst[0]) for x in lst)
        print(f"[dist check] all ranks agree: {agree}")
        ok = ok and agree
        print("DIST_CHECK_OK" if ok else "DIST_CHECK_FAILED")
    # (no destroy_process_group: tearing NCCL down while CUDA graphs that captured its kernels are alive can hang)
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0 if (rank != 0 or ok) else 1)


if __name__ == "__main__":
    main()
