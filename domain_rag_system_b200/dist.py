"""Row-sharded hybrid retrieval: one process per GPU, one small exchange per query batch.

The reference is single-process; the path shards naturally because both
retrievers score every chunk independently given BM25's GLOBAL statistics (reduced once at
build time, ``Bm25DeviceIndex.from_token_ids(group=...)``).  Per batch each rank computes its
local top-k' for both retrievers, the ranks all-gather ONE packed buffer of candidates
(``Q x k' x 28 bytes`` per rank -- no row data ever crosses NVLink), and every rank runs the
merge -> fusion -> top-K kernels on the gathered lists (ranks are global ranks only after the
merge, so fusion cannot happen earlier).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib


def shard_bounds(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row shard [lo, hi) of rank `rank`; sizes differ by at most one."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


# ---- candidate packing: [Q, k'] x {f32 score, i64 id} + [Q, k'] x {f64 score, i64 id} -> bytes ----
def pack_candidates(v_sc: torch.Tensor, v_ids: torch.Tensor, b_sc: torch.Tensor, b_ids: torch.Tensor) -> torch.Tensor:
    parts = [v_sc.contiguous().view(torch.uint8).reshape(-1), v_ids.contiguous().view(torch.uint8).reshape(-1),
             b_sc.contiguous().view(torch.uint8).reshape(-1), b_ids.contiguous().view(torch.uint8).reshape(-1)]
    return torch.cat(parts)


def unpack_candidates(buf: torch.Tensor, world: int, q: int, kv: int, kb: int):
    """``buf``: [world, nbytes] uint8 -> (v_sc [P,Q,kv] f32, v_ids i64, b_sc [P,Q,kb] f64, b_ids i64)."""
    sizes = [q * kv * 4, q * kv * 8, q * kb * 8, q * kb * 8]
    offs = [0]
    for s in sizes:
        offs.append(offs[-1] + s)
    buf = buf.view(world, -1)
    cut = lambda i, dt, k: buf[:, offs[i]:offs[i + 1]].contiguous().view(dt).view(world, q, k)  # noqa: E731
    return cut(0, torch.float32, kv), cut(1, torch.int64, kv), cut(2, torch.float64, kb), cut(3, torch.int64, kb)


def globalize_ids(ids: torch.Tensor, row_offset: int) -> torch.Tensor:
    """Local row ids -> global row ids (padding -1 stays -1)."""
    return torch.where(ids >= 0, ids + row_offset, ids)


class ShardedHybridSearcher:
    """Batched hybrid search over a row-sharded corpus.  Every rank returns the same result."""

    def __init__(self, rows_bf16: torch.Tensor, inv_norm: torch.Tensor, bm25_index, row_offset: int,
                 group=None, fusion_key: Optional[torch.Tensor] = None):
        self.rows, self.inv_norm, self.bm25 = rows_bf16, inv_norm, bm25_index
        self.row_offset = int(row_offset)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.fusion_key = fusion_key            # int64 [N_global] or None
        self._ws_d, self._ws_b = _lib.Workspace(), _lib.Workspace()
        self._gather = None

    def local_topk(self, qvecs, q_terms, q_offsets, kprime, dense_kernel=_lib.DENSE_AUTO):
        nq = qvecs.shape[0]
        kv = min(kprime, self.rows.shape[0])
        kb = min(kprime, self.bm25.n_docs)
        v_sc, v_ids, _ = _lib.dense_topk(qvecs, self.rows, self.inv_norm, kv, kernel=dense_kernel, ws=self._ws_d)
        b_sc, b_ids = _lib.bm25_topk(self.bm25.struct(), q_terms, q_offsets, nq, kb, ws=self._ws_b)
        return v_sc, v_ids, b_sc, b_ids

    def search_batch(self, qvecs, q_terms, q_offsets, top_k=10, fusion_method="rrf", vector_weight=0.5,
                     bm25_weight=0.5, dense_kernel=_lib.DENSE_AUTO):
        """Device tensors in; returns (ids int64 [Q,K] GLOBAL rows, scores fp64 [Q,K])."""
        nq, kprime = qvecs.shape[0], 2 * top_k
        v_sc, v_ids, b_sc, b_ids = self.local_topk(qvecs, q_terms, q_offsets, kprime, dense_kernel)
        if self.world > 1:
            v_ids, b_ids = globalize_ids(v_ids, self.row_offset), globalize_ids(b_ids, self.row_offset)
            kv, kb = v_ids.shape[1], b_ids.shape[1]
            send = pack_candidates(v_sc, v_ids, b_sc, b_ids)
            if self._gather is None or self._gather.numel() != self.world * send.numel():
                self._gather = torch.empty(self.world * send.numel(), dtype=torch.uint8, device=send.device)
            dist.all_gather_into_tensor(self._gather, send, group=self.group)
            gv_sc, gv_ids, gb_sc, gb_ids = unpack_candidates(self._gather, self.world, nq, kv, kb)
            v_sc, v_ids = _lib.merge_candidates(gv_sc, gv_ids, kprime)
            b_sc, b_ids = _lib.merge_candidates(gb_sc, gb_ids, kprime)
        ids, scores, _ = _lib.fuse_topk(fusion_method, v_sc, v_ids, b_sc, b_ids, top_k, fusion_key=self.fusion_key,
                                        w_vector=vector_weight, w_bm25=bm25_weight)
        return ids, scores
