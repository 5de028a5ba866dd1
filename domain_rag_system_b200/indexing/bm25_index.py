"""Device-resident BM25 index: term-major CSR postings + per-tile offsets.

Host/torch-side construction of what ``rank_bm25.BM25Okapi.__init__`` builds (called at
reference ``src/indexing/bm25_indexer.py:102-106``), laid out for the CUDA kernels in
``csrc/bm25.cu`` (see ``include/drs_b200.h: drs_bm25_index``):

* ``idf[t] = ln(N - df + 0.5) - ln(df + 0.5)`` (``math.log``), summed in the library's
  dict-insertion order (first appearance of the term in the token stream) to get
  ``average_idf``; negative idfs are replaced by ``epsilon * average_idf``;
* ``doc_kd[d] = k1 * (1 - b + b * doc_len[d] / avgdl)`` in fp64, same operation order;
* postings sorted by (term, doc): ``post_doc`` u32, ``post_tf`` u16,
  ``post_impact`` = int32 fixed point round(idf * tf*(k1+1) / (tf + doc_kd[doc]) * 2^s) with
  ``max |impact| * 2^s <= 2^26`` (``impact_scale_log2``); head terms also get dense rows of them;
* ``tile_ptr[t, j]`` = first posting of term t whose doc >= j*tile_docs (relative).

The construction uses torch ops so it runs on the GPU for large corpora (sort-based CSR
build) and on the CPU in unit tests; only the V-long idf loop is Python.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .. import _lib

IMPACT_BITS = 26   # |stored impact| <= 2^26: 16 query tokens sum without a shift (csrc/bm25.cu)
DENSE_MAX = (1 << 14) - 1   # dense row values are 14-bit unsigned (see include/drs_b200.h)
POST_PAD = 4       # zeroed entries after the last posting (aligned 4-wide reads)


def _pad_tail(t: torch.Tensor) -> torch.Tensor:
    return torch.cat([t, torch.zeros(POST_PAD, dtype=t.dtype, device=t.device)])


def choose_tile_docs(n_docs: int, sm_count: int = 148) -> int:
    """Documents per tile (power of two, 256..8192).  Every (tile, query) pair pays a fixed ~7 us of
    barriers and dependent loads, so tiles are as large as shared memory allows (8192) unless the shard is
    so small that even a single query could not spread over a handful of CTAs."""
    t = 8192
    while t > 256 and n_docs // t < 8:
        t //= 2
    return t


class Bm25DeviceIndex:
    def __init__(self):
        self.n_docs = 0
        self.n_terms = 0
        self.tile_docs = 0
        self.n_tiles = 0
        self.k1 = 1.5
        self.b = 0.75
        self.epsilon = 0.25
        self.avgdl = 0.0
        self.average_idf = 0.0
        self.vocab: Dict[str, int] | None = None   # term string -> id (None for id-only corpora)
        self._vtab = None                          # device lookup table over the vocabulary (encode_texts)
        self.vocab_present = 0                     # number of distinct terms that occur
        self.row_offset = 0                        # global row of local document 0 (sharded corpora)
        self.n_docs_global = 0
        self.total_tokens = 0
        self.n_dense = 0                           # head terms that also have a dense impact row
        self.impact_scale_log2 = 0                 # stored impacts = round(impact * 2^s)
        self.dense_shift = 0                       # dense row value = round(stored impact / 2^dense_shift)
        self.impacts_nonneg = 0                    # 1: no negative impact anywhere (enables crossing detection)
        self.t: Dict[str, torch.Tensor] = {}       # device tensors
        self._struct = None

    # rank_bm25-style attribute names, used by get_stats / tests
    @property
    def corpus_size(self):
        return self.n_docs

    # -----------------------------------------------------------------------------
    @classmethod
    def from_token_ids(cls, doc_offsets, tokens, vocab_size: int, k1=1.5, b=0.75, epsilon=0.25,
                       device="cuda", tile_docs: int | None = None, build_device=None, group=None,
                       dense_frac: float = 0.25, max_dense_terms: int = 64):
        """``doc_offsets`` int64 [N+1], ``tokens`` int32 [total] (term ids in [0, vocab_size)).

        ``group``: a ``torch.distributed`` process group (or ``True`` for the default group) when
        this rank holds one ROW SHARD of a larger corpus.  BM25's statistics are global (df, N,
        avgdl, and the order in which idf values are summed), so they are reduced across ranks
        once, here; postings keep LOCAL document ids and ``row_offset`` maps them to global rows."""
        self = cls()
        build_device = torch.device(build_device or device)
        device = torch.device(device)
        # (numpy arrays / lists, or torch tensors already on a device: nothing is round-tripped through the host)
        offs = (doc_offsets if torch.is_tensor(doc_offsets) else torch.as_tensor(np.asarray(doc_offsets))).to(torch.int64)
        n = int(offs.numel() - 1)
        if n <= 0:
            raise ValueError("empty corpus")  # rank_bm25 divides by corpus_size
        total = int(offs[-1])
        self.n_docs, self.n_terms, self.k1, self.b, self.epsilon = n, int(vocab_size), k1, b, epsilon
        self.total_tokens = total
        n_global, total_global, tok_offset = n, total, 0
        dist = None
        if group is not None:
            import torch.distributed as dist
            grp = None if group is True else group
            world, rank = dist.get_world_size(grp), dist.get_rank(grp)
            sizes = torch.zeros(world, 2, dtype=torch.int64, device=build_device)
            sizes[rank, 0], sizes[rank, 1] = n, total
            dist.all_reduce(sizes, group=grp)
            sizes = sizes.cpu()
            n_global, total_global = int(sizes[:, 0].sum()), int(sizes[:, 1].sum())
            self.row_offset = int(sizes[:rank, 0].sum())
            tok_offset = int(sizes[:rank, 1].sum())
        self.n_docs_global = n_global
        self.avgdl = total_global / n_global
        doc_len = (offs[1:] - offs[:-1]).to(build_device)
        toks = (tokens if torch.is_tensor(tokens) else torch.as_tensor(np.asarray(tokens))).to(build_device, dtype=torch.int64)
        doc_of_tok = torch.repeat_interleave(torch.arange(n, device=build_device, dtype=torch.int64), doc_len)
        # (term, doc) pairs -> sorted unique with counts == postings with tf
        ukey, tf = torch.unique(toks * n + doc_of_tok, return_counts=True)
        del doc_of_tok
        post_term = torch.div(ukey, n, rounding_mode="floor")
        post_doc = ukey - post_term * n
        del ukey
        df = torch.bincount(post_term, minlength=self.n_terms)
        term_offsets = torch.zeros(self.n_terms + 1, dtype=torch.int64, device=build_device)
        term_offsets[1:] = torch.cumsum(df, 0)
        # idf in dict-insertion order (first appearance in the token stream)
        first_pos = torch.full((self.n_terms,), total, dtype=torch.int64, device=build_device)
        if total:
            first_pos.scatter_reduce_(0, toks, torch.arange(total, device=build_device, dtype=torch.int64),
                                      reduce="amin", include_self=True)
        del toks
        df_stat = df
        if dist is not None:
            first_pos = torch.where(first_pos < total, first_pos + tok_offset,
                                    torch.full_like(first_pos, torch.iinfo(torch.int64).max))
            dist.all_reduce(first_pos, op=dist.ReduceOp.MIN, group=grp)
            df_stat = df.clone()
            dist.all_reduce(df_stat, group=grp)
        df_h = df_stat.cpu().numpy()
        n_local, n = n, n_global          # idf uses the GLOBAL corpus size
        # idf[t] = math.log(N - df + 0.5) - math.log(df + 0.5): Python's math.log (not bit-equal to a vectorised log),
        # evaluated once per DISTINCT df (a few thousand values, not V); the running sum that becomes average_idf
        # follows dict-insertion order = order of first appearance (math.fsum would differ: plain left-to-right adds)
        order = np.argsort(first_pos.cpu().numpy(), kind="stable")
        order = order[df_h[order] > 0]
        if order.size == 0:
            raise ValueError("corpus has no tokens")  # rank_bm25: division by zero in average_idf
        udf, inv = np.unique(df_h[order], return_inverse=True)
        uval = np.array([math.log(n - int(d) + 0.5) - math.log(int(d) + 0.5) for d in udf], dtype=np.float64)
        vals = uval[inv]
        idf = np.zeros(self.n_terms, dtype=np.float64)
        idf[order] = vals
        idf_sum = 0
        for v in vals.tolist():          # sequential fp64 adds in insertion order, as rank_bm25 does
            idf_sum += v
        self.vocab_present = int(order.size)
        self.average_idf = idf_sum / self.vocab_present
        neg = order[vals < 0]
        if neg.size:
            idf[neg] = self.epsilon * self.average_idf
        n = n_local
        # per-document length normaliser, fp64, reference operation order
        dl_h = doc_len.cpu().numpy()
        kd_h = self.k1 * (1 - self.b + self.b * dl_h / self.avgdl)
        kd = torch.from_numpy(kd_h).to(build_device)
        idf_d = torch.from_numpy(idf).to(build_device)
        tf64 = tf.to(torch.float64)
        impact64 = idf_d[post_term] * (tf64 * (self.k1 + 1) / (tf64 + kd[post_doc]))
        # Pre-selection works on INTEGER sums of fixed-point impacts (order-free, so ties are exact):
        # scale 2^s with max |impact| * 2^s <= 2^IMPACT_BITS.  idf is global, so the bound and therefore
        # the scale are the same on every shard of a sharded corpus.
        max_imp = float(np.abs(idf).max()) * (self.k1 + 1)
        self.impact_scale_log2 = int(min(40, math.floor(IMPACT_BITS - math.log2(max_imp)))) if max_imp > 0 else 0
        impact = torch.round(impact64 * float(2.0 ** self.impact_scale_log2)).clamp_(-2 ** IMPACT_BITS, 2 ** IMPACT_BITS)
        impact = impact.to(torch.int32)
        self.impacts_nonneg = int(impact.numel() == 0 or int(impact.min()) >= 0)
        del impact64
        if int(tf.max()) > 65535:
            raise NotImplementedError("term frequency above 65535 in one chunk")
        # per-tile offsets
        self.tile_docs = int(tile_docs or choose_tile_docs(n))
        self.n_tiles = (n + self.tile_docs - 1) // self.tile_docs
        shift = int(math.log2(self.tile_docs))
        cell = post_term * self.n_tiles + (post_doc >> shift)
        cnt = torch.bincount(cell, minlength=self.n_terms * self.n_tiles).view(self.n_terms, self.n_tiles)
        del cell
        tile_ptr = torch.zeros(self.n_terms, self.n_tiles + 1, dtype=torch.int64, device=build_device)
        torch.cumsum(cnt, 1, out=tile_ptr[:, 1:])
        del cnt
        # dense impact rows for the head terms (local df >= dense_frac * n): vector-add path
        n_pad = self.n_tiles * self.tile_docs
        dense_slot = torch.full((self.n_terms,), -1, dtype=torch.int32, device=build_device)
        heavy = torch.nonzero(df >= max(1.0, dense_frac * n)).flatten()
        if heavy.numel() > max_dense_terms:
            heavy = heavy[torch.argsort(df[heavy], descending=True)[:max_dense_terms]]
        dense_impact = torch.zeros(max(int(heavy.numel()), 1), n_pad, dtype=torch.int16, device=build_device)
        self.dense_shift = 0
        if heavy.numel() and dense_frac > 0:
            dense_slot[heavy] = torch.arange(heavy.numel(), dtype=torch.int32, device=build_device)
            slot_of_post = dense_slot[post_term]
            sel = slot_of_post >= 0
            imp_d = impact[sel].to(torch.int64)
            if int(imp_d.min()) < 0:
                # packed 16-bit lanes are unsigned: a corpus whose head terms carry NEGATIVE impacts (the
                # epsilon floor of a negative average idf) keeps every term on the postings path
                dense_slot.fill_(-1)
            else:
                # 14-bit row values: four rows add up inside a 16-bit lane without a carry
                mx = int(imp_d.max())
                while ((mx + ((1 << self.dense_shift) >> 1)) >> self.dense_shift) > DENSE_MAX:
                    self.dense_shift += 1
                v14 = (imp_d + ((1 << self.dense_shift) >> 1)) >> self.dense_shift
                dense_impact[slot_of_post[sel].long(), post_doc[sel]] = v14.to(torch.int16)
                # per tile: group i = documents 4i..4i+3 of the first half, then 4i..4i+3 of the second half
                t8 = self.tile_docs // 8
                dense_impact = (dense_impact.view(-1, self.n_tiles, 2, t8, 4).permute(0, 1, 3, 2, 4).contiguous()
                                .view(-1, n_pad))
            del slot_of_post, sel, imp_d
        else:
            dense_slot.fill_(-1)
        self.n_dense = int((dense_slot >= 0).sum())
        self.t = {
            "dense_slot": dense_slot.to(device),
            "dense_impact": dense_impact.to(device),
            "term_offsets": term_offsets.to(device),
            # postings are read in aligned groups of four: POST_PAD zeroed entries follow the last posting
            "post_doc": _pad_tail(post_doc.to(torch.int32)).to(device),   # bit pattern of u32 (n < 2^31 enforced below)
            "post_impact": _pad_tail(impact).to(device),
            "post_tf": tf.to(torch.int16).to(device),             # bit pattern of u16
            "tile_ptr": tile_ptr.to(torch.int32).contiguous().to(device),
            "idf": idf_d.to(device),
            "doc_kd": kd.to(device),
            "doc_len": doc_len.to(torch.int32).to(device),
        }
        if n >= (1 << 31):
            raise NotImplementedError("more than 2^31-1 documents in one shard")
        self._struct = None
        return self

    @classmethod
    def from_token_lists(cls, tokenized_corpus: Sequence[Sequence[str]], k1=1.5, b=0.75, epsilon=0.25,
                         device="cuda", tile_docs=None, build_device=None):
        """The drop-in path: documents are lists of token strings."""
        lens = np.fromiter((len(d) for d in tokenized_corpus), dtype=np.int64, count=len(tokenized_corpus))
        offs = np.zeros(len(tokenized_corpus) + 1, dtype=np.int64)
        np.cumsum(lens, out=offs[1:])
        # term ids in order of first appearance (= rank_bm25's dict insertion order), without a Python loop per
        # token: one flat pass builds the vocabulary (dict.setdefault runs in C under map), one fromiter the ids
        vocab: Dict[str, int] = {}
        flat = [w for doc in tokenized_corpus for w in doc]
        toks = np.fromiter(map(lambda w: vocab.setdefault(w, len(vocab)), flat), dtype=np.int32, count=len(flat))
        self = cls.from_token_ids(offs, toks, max(len(vocab), 1), k1, b, epsilon, device, tile_docs, build_device)
        self.vocab = vocab
        return self

    # -----------------------------------------------------------------------------
    def struct(self) -> _lib.Bm25IndexStruct:
        if self._struct is None:
            s = _lib.Bm25IndexStruct()
            s.n_docs, s.n_terms, s.tile_docs, s.n_tiles, s.impact_scale_log2 = (self.n_docs, self.n_terms, self.tile_docs,
                                                                       self.n_tiles, self.impact_scale_log2)
            t = self.t
            s.d_term_offsets = t["term_offsets"].data_ptr()
            s.d_post_doc = t["post_doc"].data_ptr()
            s.d_post_impact = t["post_impact"].data_ptr()
            s.d_post_tf = t["post_tf"].data_ptr()
            s.d_tile_ptr = t["tile_ptr"].data_ptr()
            s.d_idf = t["idf"].data_ptr()
            s.d_doc_kd = t["doc_kd"].data_ptr()
            s.k1_plus_1 = self.k1 + 1
            s.d_dense_slot = t["dense_slot"].data_ptr()
            s.d_dense_impact = t["dense_impact"].data_ptr()
            s.dense_stride = self.n_tiles * self.tile_docs
            s.dense_shift = self.dense_shift
            s.impacts_nonneg = self.impacts_nonneg
            self._struct = s
        return self._struct

    @property
    def device(self):
        return self.t["post_doc"].device

    @property
    def nnz(self):
        return int(self.t["post_doc"].numel()) - POST_PAD

    def encode_queries(self, queries: Sequence[Sequence], pinned: bool = False):
        """Token strings (or term ids) -> (q_terms int32, q_offsets int32) HOST tensors.
        Unknown strings become -1 (``idf.get(q) or 0`` in rank_bm25 => no contribution)."""
        offs = np.zeros(len(queries) + 1, dtype=np.int32)
        if len(queries) and all(isinstance(q, np.ndarray) and q.dtype.kind in "iu" for q in queries):
            # term-id queries (benchmarks, pre-encoded callers): no Python loop per token
            np.cumsum([len(q) for q in queries], out=offs[1:])
            ids = np.concatenate(queries).astype(np.int64) if offs[-1] else np.zeros(0, dtype=np.int64)
            flat_a = np.where((ids >= 0) & (ids < self.n_terms), ids, -1).astype(np.int32)
        else:
            flat: List[int] = []
            get = self.vocab.get if self.vocab is not None else (lambda w, d: d)
            for i, q in enumerate(queries):
                for w in q:
                    if isinstance(w, str):
                        flat.append(get(w, -1))
                    else:
                        w = int(w)
                        flat.append(w if 0 <= w < self.n_terms else -1)
                offs[i + 1] = len(flat)
            flat_a = np.asarray(flat, dtype=np.int32).reshape(-1)
        qt = torch.from_numpy(flat_a)
        qo = torch.from_numpy(offs)
        if pinned:
            qt, qo = qt.pin_memory(), qo.pin_memory()
        return qt, qo

    # -- query strings -> term ids on the device --------------------------------
    @staticmethod
    def vocab_table_arrays(words: Sequence[Optional[str]]):
        """Host arrays of the device lookup table over ``words[term_id]`` (None = id without a string): 64-bit
        FNV-1a of every string (vectorised over character positions), entries placed in order of their home slot
        with no wrap-around (``pos_i = max(home_i, pos_{i-1} + 1)``), so a probe walks forward from the home slot
        to the first empty one.  Returns (slot_hash uint64 [n_slots], slot_term int32 [n_slots], str_offsets int64
        [V+1], pool uint8, capacity)."""
        v = len(words)
        enc = [w.encode("utf-8") if w else b"" for w in words]
        lens = np.fromiter((len(b) for b in enc), dtype=np.int64, count=v)
        offs = np.zeros(v + 1, dtype=np.int64)
        np.cumsum(lens, out=offs[1:])
        pool = np.frombuffer(b"".join(enc), dtype=np.uint8).copy() if offs[-1] else np.zeros(1, dtype=np.uint8)
        h = np.full(v, 0xCBF29CE484222325, dtype=np.uint64)
        prime = np.uint64(0x100000001B3)
        for pos in range(int(lens.max()) if v else 0):
            m = lens > pos
            h[m] = (h[m] ^ pool[offs[:-1][m] + pos].astype(np.uint64)) * prime
        ids = np.flatnonzero(lens > 0)
        n = ids.size
        cap = 16
        while cap < 2 * max(n, 1):
            cap *= 2
        home = (h[ids] & np.uint64(cap - 1)).astype(np.int64)
        order = np.argsort(home, kind="stable")
        ar = np.arange(n, dtype=np.int64)
        pos = np.maximum.accumulate(home[order] - ar) + ar if n else ar
        n_slots = cap + n + 1
        slot_hash = np.zeros(n_slots, dtype=np.uint64)
        slot_term = np.full(n_slots, -1, dtype=np.int32)
        slot_hash[pos] = h[ids][order]
        slot_term[pos] = ids[order].astype(np.int32)
        return slot_hash, slot_term, offs, pool, cap

    def _vocab_table(self):
        """The device copy of ``vocab_table_arrays`` for this index's vocabulary (built once, on first use)."""
        if self._vtab is None:
            words: List[Optional[str]] = [None] * self.n_terms
            for w, t in (self.vocab or {}).items():
                if 0 <= t < self.n_terms:
                    words[t] = w
            sh, st, offs, pool, cap = self.vocab_table_arrays(words)
            dev = self.t["post_doc"].device
            keep = {"sh": torch.from_numpy(sh.view(np.int64)).to(dev), "st": torch.from_numpy(st).to(dev),
                    "offs": torch.from_numpy(offs).to(dev), "pool": torch.from_numpy(pool).to(dev)}
            s = _lib.VocabTableStruct()
            s.d_slot_hash, s.d_slot_term = keep["sh"].data_ptr(), keep["st"].data_ptr()
            s.d_str_offsets, s.d_pool = keep["offs"].data_ptr(), keep["pool"].data_ptr()
            s.capacity, s.n_slots = (cap if self.vocab else 0), sh.size
            self._vtab = (s, keep)
        return self._vtab[0]

    def encode_texts(self, texts: Sequence[str], keep_single: Sequence[str] = (), host_tokenize=None):
        """Query STRINGS -> (q_terms int32, q_offsets int32 [Q+1]) DEVICE tensors: the reference's tokeniser
        (``BM25Indexer.tokenize``, src/indexing/bm25_indexer.py:50-78) and vocabulary lookup run on the GPU
        (``drs_tokenize_lookup``), one launch pair for the whole batch, no Python work per token.
        ``keep_single``: the lower-cased abbreviation keys of one character (the only one-character tokens the
        reference keeps).  ASCII strings are exact on the device; a string with non-ASCII characters (Unicode
        ``lower()`` / ``\\w``) goes through ``host_tokenize`` + the vocabulary dict and is spliced in."""
        dev = self.t["post_doc"].device
        ok = [t.isascii() for t in texts]
        enc = [t.encode("ascii") if a else b"" for t, a in zip(texts, ok)]
        offs = np.zeros(len(texts) + 1, dtype=np.int64)
        np.cumsum([len(b) for b in enc], out=offs[1:])
        raw = np.frombuffer(b"".join(enc), dtype=np.uint8).copy() if offs[-1] else np.zeros(1, dtype=np.uint8)
        keep1 = _lib.SingleChars()
        for c in keep_single:
            if len(c) == 1 and c.isascii():
                keep1.bits[ord(c) >> 5] |= 1 << (ord(c) & 31)
        terms, q_offs = _lib.tokenize_lookup(torch.from_numpy(raw).to(dev), torch.from_numpy(offs).to(dev), keep1,
                                             self._vocab_table())
        if all(ok):
            return terms, q_offs
        if host_tokenize is None:
            raise ValueError("encode_texts: non-ASCII query text needs host_tokenize")
        # rare path: splice the host-tokenised strings into the device result
        t_h, o_h = terms.cpu().numpy(), q_offs.cpu().numpy()
        parts = []
        for i, a in enumerate(ok):
            if a:
                parts.append(t_h[o_h[i]:o_h[i + 1]])
            else:
                parts.append(self.encode_queries([host_tokenize(texts[i])])[0].numpy())
        qo = np.zeros(len(texts) + 1, dtype=np.int32)
        np.cumsum([len(x) for x in parts], out=qo[1:])
        qt = np.concatenate(parts).astype(np.int32) if qo[-1] else np.zeros(0, dtype=np.int32)
        return torch.from_numpy(qt).to(dev), torch.from_numpy(qo).to(dev)

    # -- wide windows ---------------------------------------------------------------------------
    def exact_scores(self, term_ids) -> torch.Tensor:
        """fp64 BM25 score of EVERY local document for one query (term ids of its token occurrences, -1 = OOV), on
        the device with rank_bm25's arithmetic and operation order (``score += idf * (tf * (k1 + 1) / (tf + kd))``
        per occurrence, in query order; ``BM25Okapi.get_scores``): bit-identical to the CUDA kernel's re-scored
        values.  O(sum of df) torch work per query: the path for result windows wider than DRS_MAX_KPRIME."""
        t = self.t
        dev = t["post_doc"].device
        score = torch.zeros(self.n_docs, dtype=torch.float64, device=dev)
        to = t["term_offsets"]
        for term in term_ids:
            term = int(term)
            if term < 0 or term >= self.n_terms:
                continue
            a, e = int(to[term]), int(to[term + 1])
            if e == a:
                continue
            docs = t["post_doc"][a:e].long()
            tf = (t["post_tf"][a:e].to(torch.int32) & 0xFFFF).to(torch.float64)
            contrib = t["idf"][term] * (tf * (self.k1 + 1) / (tf + t["doc_kd"][docs]))
            score.index_add_(0, docs, contrib)       # at most one posting per (term, document): no collisions
        return score

    def wide_topk(self, term_ids, k: int):
        """(scores fp64 [k], ids int64 [k]) of one query in the reference's order (score descending, larger row first
        among ties), for any k <= n_docs."""
        sc = self.exact_scores(term_ids)
        order = torch.flip(torch.argsort(sc, stable=True), dims=[0])[:k]     # == np.argsort(kind='stable')[::-1][:k]
        return sc[order], order

    def df_of(self, term_ids) -> np.ndarray:
        to = self.t["term_offsets"]
        ids = torch.as_tensor(np.asarray(term_ids, dtype=np.int64), device=to.device)
        ok = (ids >= 0) & (ids < self.n_terms)
        ids = ids.clamp(0, self.n_terms - 1)
        return ((to[ids + 1] - to[ids]) * ok).cpu().numpy()

    # -- persistence (our own format; tensors moved to the host) ---------------------
    def state_dict(self):
        return {"meta": dict(n_docs=self.n_docs, n_terms=self.n_terms, tile_docs=self.tile_docs, n_tiles=self.n_tiles,
                             k1=self.k1, b=self.b, epsilon=self.epsilon, avgdl=self.avgdl,
                             average_idf=self.average_idf, vocab_present=self.vocab_present,
                             total_tokens=self.total_tokens, row_offset=self.row_offset,
                             impact_scale_log2=self.impact_scale_log2, n_dense=self.n_dense,
                             dense_shift=self.dense_shift, impacts_nonneg=self.impacts_nonneg,
                             n_docs_global=self.n_docs_global),
                "vocab": self.vocab, "tensors": {k: v.cpu() for k, v in self.t.items()}}

    @classmethod
    def from_state_dict(cls, sd, device="cuda"):
        self = cls()
        for k, v in sd["meta"].items():
            setattr(self, k, v)
        self.vocab = sd["vocab"]
        self.t = {k: v.to(device) for k, v in sd["tensors"].items()}
        return self
