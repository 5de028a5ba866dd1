"""The CPU oracle at benchmark scale (1 M - 10 M chunks) -- TEST INFRASTRUCTURE ONLY.

Same arithmetic as ``oracle.dense`` / ``oracle.bm25`` (tier B), organised so that a 10 M-chunk corpus
can be checked for a SAMPLE of queries in a couple of minutes on the GPU box's host cores:

* ``BlockedDense`` -- ``oracle.dense.cosine_scores_batch`` + ``dense_search`` over row blocks: the corpus arrives
  block by block (never a second full copy), every block is upsert-normalised in fp32 exactly as ``normalize_rows``
  does, scored against the sample queries in fp64, and each query keeps its running top-k' under the oracle's tie
  rule (score descending, larger row first).  Restates qdrant-client local mode as called from
  ``src/indexing/vector_indexer.py:202-209``.
* ``SampledBm25`` -- ``oracle.bm25.Bm25Csr`` with postings kept only for the terms the sample queries use.  The
  statistics every score depends on (N, avgdl, df of EVERY term, the dict-insertion order the idf values are summed
  in, the epsilon floor) are computed over the whole corpus, block by block; ``get_scores`` is the same loop as
  ``Bm25Csr.get_scores``.  Restates ``rank_bm25.BM25Okapi`` as called from ``src/indexing/bm25_indexer.py:102-106,127``.
  ``tests/test_oracle_scale.py`` proves both bit-identical to the plain tier-B oracle on small corpora.
"""
from __future__ import annotations

import math
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import dense as _dense


def topk_desc_rows(scores: np.ndarray, k: int):
    """``np.argsort(scores, kind='stable')[::-1][:k]`` without sorting the whole array: (rows, scores)."""
    n = scores.shape[0]
    k = min(k, n)
    if k == 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.float64)
    thr = np.partition(scores, n - k)[n - k]                 # the k-th largest value
    cand = np.flatnonzero(scores >= thr)                     # everything that ties with it or beats it
    order = np.lexsort((cand, scores[cand]))[::-1][:k]       # score ascending then row ascending, reversed
    rows = cand[order]
    return rows.astype(np.int64), scores[rows].astype(np.float64)


class BlockedDense:
    def __init__(self, queries: np.ndarray, kprime: int):
        q = np.asarray(queries, dtype=np.float64)
        qn = np.linalg.norm(q, axis=1)
        self.q = q / np.where(qn != 0.0, qn, _dense.EPS)[:, None]          # as cosine_scores_batch
        self.k = int(kprime)
        self.rows = [np.zeros(0, dtype=np.int64) for _ in range(q.shape[0])]
        self.scores = [np.zeros(0, dtype=np.float64) for _ in range(q.shape[0])]

    def add_block(self, row0: int, normed_block: np.ndarray, sub: int = 1 << 15):
        """``normed_block``: fp32 rows already passed through ``oracle.dense.normalize_rows``."""
        n = normed_block.shape[0]
        sc = np.empty((self.q.shape[0], n), dtype=np.float64)
        for a in range(0, n, sub):
            sc[:, a:a + sub] = self.q @ normed_block[a:a + sub].astype(np.float64).T
        for i in range(self.q.shape[0]):
            r, s = topk_desc_rows(sc[i], self.k)
            rows = np.concatenate([self.rows[i], r + row0])
            scores = np.concatenate([self.scores[i], s])
            order = np.lexsort((rows, scores))[::-1][: self.k]
            self.rows[i], self.scores[i] = rows[order], scores[order]

    def result(self, i: int):
        return self.rows[i], self.scores[i]


class SampledBm25:
    """Tier-B BM25 restricted to the terms of a query sample; build with ``add_block`` in row order, then
    ``finalize``.  ``get_scores`` is valid only for queries made of the sampled terms (and OOV = -1)."""

    def __init__(self, vocab: int, terms, k1=1.5, b=0.75, epsilon=0.25):
        self.vocab, self.k1, self.b, self.epsilon = int(vocab), k1, b, epsilon
        self.keep = np.zeros(self.vocab, dtype=bool)
        t = np.asarray(sorted({int(x) for x in terms if int(x) >= 0}), dtype=np.int64)
        self.keep[t] = True
        self.df = np.zeros(self.vocab, dtype=np.int64)
        self.first_pos = np.full(self.vocab, np.iinfo(np.int64).max, dtype=np.int64)
        self.doc_len_parts, self.post_parts = [], []
        self.n_docs, self.n_tokens = 0, 0

    @staticmethod
    def _block_postings(args):
        doc_offsets, tokens, vocab, keep = args
        n = len(doc_offsets) - 1
        doc_len = np.diff(doc_offsets).astype(np.int64)
        tok64 = tokens.astype(np.int64)
        doc_of_tok = np.repeat(np.arange(n, dtype=np.int64), doc_len)
        ukey, tf = np.unique(tok64 * n + doc_of_tok, return_counts=True)     # (term, doc) pairs with tf
        term = ukey // n
        df = np.bincount(term, minlength=vocab)
        # first occurrence of every term in this block's token stream: reversed assignment, the last write (the
        # smallest position) wins
        first = np.full(vocab, np.iinfo(np.int64).max, dtype=np.int64)
        pos = np.arange(len(tokens), dtype=np.int64)
        first[tok64[::-1]] = pos[::-1]
        sel = keep[term]
        return doc_len, df, first, term[sel].astype(np.int32), (ukey[sel] % n).astype(np.int64), tf[sel].astype(np.int32)

    def add_blocks(self, blocks, threads: int = 8):
        """``blocks``: iterable of (doc_offsets int64 [n+1], tokens int32) in corpus order."""
        blocks = list(blocks)
        with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
            outs = list(ex.map(self._block_postings, [(o, t, self.vocab, self.keep) for o, t in blocks]))
        for (o, t), (doc_len, df, first, term, doc, tf) in zip(blocks, outs):
            self.df += df
            seen = first != np.iinfo(np.int64).max
            self.first_pos[seen] = np.minimum(self.first_pos[seen], first[seen] + self.n_tokens)
            self.post_parts.append((term, doc + self.n_docs, tf))
            self.doc_len_parts.append(doc_len)
            self.n_docs += len(doc_len)
            self.n_tokens += int(o[-1])

    def finalize(self):
        n = self.corpus_size = self.n_docs
        self.doc_len = np.concatenate(self.doc_len_parts)
        self.avgdl = self.n_tokens / n
        term = np.concatenate([p[0] for p in self.post_parts]).astype(np.int64)
        doc = np.concatenate([p[1] for p in self.post_parts])
        tf = np.concatenate([p[2] for p in self.post_parts]).astype(np.int64)
        order = np.lexsort((doc, term))
        self.post_doc, self.post_tf = doc[order], tf[order]
        self.term_offsets = np.zeros(self.vocab + 1, dtype=np.int64)
        np.add.at(self.term_offsets, term + 1, 1)
        np.cumsum(self.term_offsets, out=self.term_offsets)
        self.post_parts = self.doc_len_parts = None
        # idf exactly as Bm25Csr: dict-insertion order = order of first appearance in the token stream
        present = np.flatnonzero(self.df > 0)
        ins = present[np.argsort(self.first_pos[present], kind="stable")]
        self.idf = np.zeros(self.vocab, dtype=np.float64)
        idf_sum = 0
        neg = []
        for t in ins:
            v = math.log(n - int(self.df[t]) + 0.5) - math.log(int(self.df[t]) + 0.5)
            self.idf[t] = v
            idf_sum += v
            if v < 0:
                neg.append(t)
        self.average_idf = idf_sum / len(ins)
        self.idf[np.array(neg, dtype=np.int64)] = self.epsilon * self.average_idf
        return self

    def get_scores(self, query_terms):
        score = np.zeros(self.corpus_size)
        for t in query_terms:
            t = int(t)
            if t < 0:
                continue
            assert self.keep[t], "term outside the sampled set"
            a, e = self.term_offsets[t], self.term_offsets[t + 1]
            docs = self.post_doc[a:e]
            q_freq = self.post_tf[a:e]
            dl = self.doc_len[docs]
            score[docs] += self.idf[t] * (q_freq * (self.k1 + 1) /
                                          (q_freq + self.k1 * (1 - self.b + self.b * dl / self.avgdl)))
        return score


def bm25_topk(scores: np.ndarray, top_k: int):
    """``oracle.bm25.bm25_search`` without filters (argsort descending, larger row first among ties; the wrapper's
    ``top_k*2`` over-fetch only matters with a post-filter), via a partial selection."""
    return topk_desc_rows(scores, top_k)
