"""Row-sharded hybrid retrieval: one process per GPU, one small exchange per query batch.

The reference is single-process; the path shards naturally because both retrievers score every chunk
independently given BM25's GLOBAL statistics (reduced once at build time, ``Bm25DeviceIndex.from_token_ids(group=...)``).
Per batch each rank computes its local top-k' for both retrievers straight into ONE send buffer, the ranks all-gather
it (``Q x k' x 36 bytes`` per rank -- no row data ever crosses NVLink), and every rank merges the gathered lists in
place (``drs_merge_gathered``), fuses, and -- for the "simple" reranker -- orders the fused pool by cosines that were
computed where the rows live and travelled with the candidates (``drs_candidate_cosines`` ->
``drs_rerank_from_cosines``): no second exchange.  Fusion cannot precede the merge (RRF ranks are global ranks).
The whole step is stream-ordered on the current stream and can be captured in a CUDA graph (``graphed``).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib


_SIDE_STREAM = os.environ.get("DRS_LATENCY_SIDE_STREAM", "0") == "1"


def shard_bounds(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row shard [lo, hi) of rank `rank`; sizes differ by at most one."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class _Layout:
    """Byte layout of one rank's send buffer for Q queries and lists of exactly k entries."""

    def __init__(self, q: int, k: int):
        n = q * k
        self.q, self.k = q, k
        self.off_vsc, self.off_vcos, self.off_bcos = 0, 4 * n, 8 * n
        self.off_vid = (12 * n + 7) // 8 * 8
        self.off_bid = self.off_vid + 8 * n
        self.off_bsc = self.off_bid + 8 * n
        self.nbytes = self.off_bsc + 8 * n

    def views(self, buf: torch.Tensor):
        q, k, n = self.q, self.k, self.q * self.k
        cut = lambda off, nb, dt: buf[off:off + nb].view(dt).view(q, k)  # noqa: E731
        return {"v_sc": cut(self.off_vsc, 4 * n, torch.float32), "v_cos": cut(self.off_vcos, 4 * n, torch.float32),
                "b_cos": cut(self.off_bcos, 4 * n, torch.float32), "v_ids": cut(self.off_vid, 8 * n, torch.int64),
                "b_ids": cut(self.off_bid, 8 * n, torch.int64), "b_sc": cut(self.off_bsc, 8 * n, torch.float64)}


    def read_parts(self, gathered: torch.Tensor, world: int):
        """The gathered buffer as [P, Q, k] tensors (what ``drs_merge_gathered`` reads in place); host-side helper
        for tests and debugging."""
        g = gathered.view(world, self.nbytes)
        return {name: torch.stack([self.views(g[p])[name] for p in range(world)]) for name in
                ("v_sc", "v_cos", "b_cos", "v_ids", "b_ids", "b_sc")}


class ShardedHybridSearcher:
    """Batched hybrid search over a row-sharded corpus.  Every rank returns the same result (GLOBAL row ids)."""

    def __init__(self, rows_bf16: torch.Tensor, inv_norm: torch.Tensor, bm25_index, row_offset: int,
                 group=None, fusion_key: Optional[torch.Tensor] = None, world: Optional[int] = None):
        """``world`` = 1 makes a purely local searcher even inside an initialised process group (NO collective is
        issued, here or later); by default the group's size is used and the constructor is itself a collective (one
        all-gather of the shards' row offsets), so every rank of the group must construct its searcher."""
        self.rows, self.inv_norm, self.bm25 = rows_bf16, inv_norm, bm25_index
        self.row_offset = int(row_offset)
        self.group = group
        self.world = int(world) if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.fusion_key = fusion_key            # int64 [N_global] or None
        self._ws_d, self._ws_b = _lib.Workspace(), _lib.Workspace()
        self._bufs = {}
        self._side = None
        self._id_offset = None
        if self.world > 1:
            dev = rows_bf16.device
            mine = torch.tensor([self.row_offset], dtype=torch.int64, device=dev)
            offs = torch.empty(self.world, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(offs, mine, group=group)
            self._id_offset = offs

    # -- buffers of one (Q, k') shape: allocated once so that a step allocates nothing (graph capture) ----------
    def _buffers(self, nq: int, kprime: int):
        key = (nq, kprime)
        b = self._bufs.get(key)
        if b is None:
            dev = self.rows.device
            lay = _Layout(nq, kprime)
            send = torch.empty(lay.nbytes, dtype=torch.uint8, device=dev)
            b = {"layout": lay, "send": send, "views": lay.views(send),
                 "gather": torch.empty(self.world * lay.nbytes, dtype=torch.uint8, device=dev) if self.world > 1 else None,
                 "flags": torch.empty(nq, dtype=torch.int32, device=dev)}
            self._bufs[key] = b
        return b

    def local_topk(self, qvecs, q_terms, q_offsets, kprime, dense_kernel=_lib.DENSE_AUTO, out=None):
        """Local top-k' of both retrievers, lists padded to exactly k' (id -1, score -inf): the exchange size must not
        depend on how many rows a shard happens to hold."""
        nq = qvecs.shape[0]
        v = out if out is not None else self._buffers(nq, kprime)["views"]
        flags = self._buffers(nq, kprime)["flags"]
        kv = min(kprime, self.rows.shape[0])
        kb = min(kprime, self.bm25.n_docs)
        # Experiment hook (DRS_LATENCY_SIDE_STREAM=1): in latency mode run BM25 beside the HBM-bound dense pass on a
        # second stream.  Measured SLOWER on B200 (p50 3.3 -> 4.4 ms: the streaming kernel's statically partitioned
        # CTAs lose their SM slots to the BM25 CTAs and finish in two waves), so it is off by default.
        side = None
        if _SIDE_STREAM and nq <= 8 and kb == kprime:
            if self._side is None:
                self._side = torch.cuda.Stream()
            side = self._side
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                _lib.bm25_topk(self.bm25.struct(), q_terms, q_offsets, nq, kprime, ws=self._ws_b, out=(v["b_sc"], v["b_ids"]))
        if kv == kprime:
            _lib.dense_topk(qvecs, self.rows, self.inv_norm, kprime, kernel=dense_kernel, ws=self._ws_d,
                            out=(v["v_sc"], v["v_ids"], flags))
        else:
            sc, ids, _ = _lib.dense_topk(qvecs, self.rows, self.inv_norm, kv, kernel=dense_kernel, ws=self._ws_d)
            v["v_sc"].fill_(float("-inf"))
            v["v_ids"].fill_(-1)
            v["v_sc"][:, :kv], v["v_ids"][:, :kv] = sc, ids
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
        elif kb == kprime:
            _lib.bm25_topk(self.bm25.struct(), q_terms, q_offsets, nq, kprime, ws=self._ws_b, out=(v["b_sc"], v["b_ids"]))
        else:
            sc, ids = _lib.bm25_topk(self.bm25.struct(), q_terms, q_offsets, nq, kb, ws=self._ws_b)
            v["b_sc"].fill_(float("-inf"))
            v["b_ids"].fill_(-1)
            v["b_sc"][:, :kb], v["b_ids"][:, :kb] = sc, ids
        return v["v_sc"], v["v_ids"], v["b_sc"], v["b_ids"]

    def search_batch(self, qvecs, q_terms, q_offsets, top_k=10, fusion_method="rrf", vector_weight=0.5,
                     bm25_weight=0.5, dense_kernel=_lib.DENSE_AUTO, rerank_top_n=0, final_k=None, rerank_vecs=None):
        """Device tensors in.  Returns (ids int64 [Q, K] GLOBAL rows, scores fp64 [Q, K]) of the fused top-K; with
        ``rerank_top_n`` > 0 the fused top-K is the rerank pool (scripts/query.py:139-160) and the result is
        (ids int64 [Q, final_k], rerank scores fp32 [Q, final_k])."""
        nq, kprime = qvecs.shape[0], 2 * top_k
        b = self._buffers(nq, kprime)
        v = b["views"]
        v_sc, v_ids, b_sc, b_ids = self.local_topk(qvecs, q_terms, q_offsets, kprime, dense_kernel, out=v)
        rq = qvecs if rerank_vecs is None else rerank_vecs
        v_cos = b_cos = None
        if self.world > 1:
            lay = b["layout"]
            if rerank_top_n:
                _lib.candidate_cosines(self.rows, rq, v_ids, out=v["v_cos"])
                _lib.candidate_cosines(self.rows, rq, b_ids, out=v["b_cos"])
            dist.all_gather_into_tensor(b["gather"], b["send"], group=self.group)
            v_sc, v_ids, v_cos = _lib.merge_gathered(b["gather"], self.world, lay.nbytes, lay.off_vsc, False, lay.off_vid,
                                                     lay.off_vcos if rerank_top_n else None, self._id_offset, nq,
                                                     kprime, kprime)
            b_sc, b_ids, b_cos = _lib.merge_gathered(b["gather"], self.world, lay.nbytes, lay.off_bsc, True, lay.off_bid,
                                                     lay.off_bcos if rerank_top_n else None, self._id_offset, nq,
                                                     kprime, kprime)
        ids, scores, _ = _lib.fuse_topk(fusion_method, v_sc, v_ids, b_sc, b_ids, top_k, fusion_key=self.fusion_key,
                                        w_vector=vector_weight, w_bm25=bm25_weight)
        if not rerank_top_n:
            return ids, scores
        pool = ids[:, : min(rerank_top_n, ids.shape[1])].contiguous()
        fk = min(final_k or top_k, pool.shape[1])
        if self.world > 1:
            r_sc, r_ids, _ = _lib.rerank_from_cosines(pool, fk, v_ids, v_cos, b_ids, b_cos)
        else:
            r_sc, r_ids, _ = _lib.rerank_cosine(self.rows, rq, pool, fk)
        return r_ids, r_sc

    def graphed(self, qvecs, q_terms, q_offsets, **kw):
        """Capture one ``search_batch`` step on the given (static) input tensors in a CUDA graph -- the local
        kernels, the all-gather and the merge / fusion / rerank tail become ONE launch.  Returns ``run()``, which
        replays the step and returns the (static) result tensors; refresh the inputs in place between replays."""
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):      # warm-up off the default stream: workspaces grow, attributes get set
            for _ in range(2):
                self.search_batch(qvecs, q_terms, q_offsets, **kw)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = self.search_batch(qvecs, q_terms, q_offsets, **kw)

        def run():
            g.replay()
            return out
        run.graph = g
        return run
