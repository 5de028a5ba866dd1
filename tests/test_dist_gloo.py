"""World-size-2 gloo tests (CPU) of the multi-rank host logic: shard bounds, the global BM25
statistics reduction, and the candidate pack -> all_gather -> unpack -> merge exchange."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from domain_rag_system_b200 import synth
from domain_rag_system_b200.dist import _Layout, shard_bounds
from domain_rag_system_b200.indexing.bm25_index import Bm25DeviceIndex
from oracle import bm25 as obm25
from oracle import dense as odense

N, VOCAB, Q, KP = 700, 120, 6, 10


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _corpus():
    offs, toks = synth.zipf_docs(N, vocab=VOCAB, seed=31, len_min=2, len_max=30)
    emb = synth.unit_embeddings(N, dim=64, seed=32).float().numpy()
    qv = synth.query_embeddings(Q, dim=64, seed=33, near_frac=0).numpy()
    qt = synth.zipf_queries(Q, offs, toks, vocab=VOCAB, seed=34)
    return offs, toks, emb, qv, qt


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        offs, toks, emb, qv, qt = _corpus()
        lo, hi = shard_bounds(N, world, rank)
        s_offs = offs[lo:hi + 1] - offs[lo]
        s_toks = toks[offs[lo]:offs[hi]]
        ix = Bm25DeviceIndex.from_token_ids(s_offs, s_toks, VOCAB, device="cpu", tile_docs=256, group=True)
        assert ix.row_offset == lo and ix.n_docs_global == N
        # local top-k' with the oracle's arithmetic on this shard, using the GLOBAL statistics
        orc = obm25.Bm25Csr(s_offs, s_toks, VOCAB)
        orc.idf = ix.t["idf"].numpy()
        orc.avgdl = ix.avgdl
        normed = odense.normalize_rows(emb[lo:hi])
        v_sc = np.full((Q, KP), -np.inf, np.float32)
        v_id = np.full((Q, KP), -1, np.int64)
        b_sc = np.full((Q, KP), -np.inf, np.float64)
        b_id = np.full((Q, KP), -1, np.int64)
        for i in range(Q):
            r, s = odense.dense_search(odense.cosine_scores(normed, qv[i], exact_rows=True), KP)
            v_id[i, :len(r)], v_sc[i, :len(r)] = r, s.astype(np.float32)
            r, s = obm25.bm25_search(orc.get_scores(qt[i]), KP)
            b_id[i, :len(r)], b_sc[i, :len(r)] = r, s
        # the product's exchange: every rank fills ONE send buffer with its LOCAL lists, one all-gather, and the
        # gathered buffer is read in place (row offsets of the parts turn local ids into global ones)
        t = torch.from_numpy
        lay = _Layout(Q, KP)
        send = torch.zeros(lay.nbytes, dtype=torch.uint8)
        v = lay.views(send)
        v["v_sc"].copy_(t(v_sc)), v["v_ids"].copy_(t(v_id)), v["b_sc"].copy_(t(b_sc)), v["b_ids"].copy_(t(b_id))
        gathered = torch.empty(world * send.numel(), dtype=torch.uint8)
        dist.all_gather_into_tensor(gathered, send)
        offs_t = torch.zeros(world, dtype=torch.int64)
        offs_t[rank] = lo
        dist.all_reduce(offs_t)
        parts = lay.read_parts(gathered, world)
        glob = lambda ids: torch.where(ids >= 0, ids + offs_t.view(-1, 1, 1), ids)  # noqa: E731
        if rank == 0:
            torch.save({"idf": ix.t["idf"], "avgdl": ix.avgdl, "average_idf": ix.average_idf, "gv_sc": parts["v_sc"],
                        "gv_id": glob(parts["v_ids"]), "gb_sc": parts["b_sc"], "gb_id": glob(parts["b_ids"])}, out)
    finally:
        dist.destroy_process_group()


def _merge_np(sc, ids, k):
    s, i = sc.reshape(-1), ids.reshape(-1)
    keep = i >= 0
    s, i = s[keep], i[keep]
    order = np.lexsort((-i, -s.astype(np.float64)))[:k]
    return i[order], s[order]


def test_sharded_statistics_and_candidate_exchange(tmp_path):
    out = str(tmp_path / "rank0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    offs, toks, emb, qv, qt = _corpus()
    full = obm25.Bm25Csr(offs, toks, VOCAB)
    # global statistics reduced across the two shards == statistics of the whole corpus, bit for bit
    assert got["avgdl"] == full.avgdl and got["average_idf"] == full.average_idf
    assert np.array_equal(got["idf"].numpy(), full.idf)
    normed = odense.normalize_rows(emb)
    for i in range(Q):
        ids, sc = _merge_np(got["gb_sc"][:, i].numpy(), got["gb_id"][:, i].numpy(), KP)
        r, s = obm25.bm25_search(full.get_scores(qt[i]), KP)
        assert np.array_equal(ids, r) and np.array_equal(sc, s)
        ids, sc = _merge_np(got["gv_sc"][:, i].numpy(), got["gv_id"][:, i].numpy(), KP)
        r, s = odense.dense_search(odense.cosine_scores(normed, qv[i], exact_rows=True), KP)
        assert np.array_equal(ids, r)


def test_shard_bounds_cover_everything():
    for n in (1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_send_buffer_layout_roundtrip():
    g = torch.Generator().manual_seed(0)
    q, k = 3, 5
    lay = _Layout(q, k)
    assert lay.off_vid % 8 == 0 and lay.off_bid % 8 == 0 and lay.off_bsc % 8 == 0
    bufs = []
    vals = []
    for p in range(2):
        buf = torch.zeros(lay.nbytes, dtype=torch.uint8)
        v = lay.views(buf)
        cur = {"v_sc": torch.rand(q, k, generator=g), "v_cos": torch.rand(q, k, generator=g),
               "b_cos": torch.rand(q, k, generator=g), "v_ids": torch.randint(0, 99, (q, k), generator=g),
               "b_ids": torch.randint(0, 99, (q, k), generator=g), "b_sc": torch.rand(q, k, generator=g, dtype=torch.float64)}
        for name, val in cur.items():
            v[name].copy_(val)
        bufs.append(buf)
        vals.append(cur)
    parts = lay.read_parts(torch.cat(bufs), 2)
    for p in range(2):
        for name, val in vals[p].items():
            assert torch.equal(parts[name][p], val)
