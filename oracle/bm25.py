"""BM25 restatement (oracle; test infrastructure only).

Two tiers:

* ``BM25OkapiRestated`` -- tier A, the published ``rank_bm25.BM25Okapi``
  algorithm (third-party, un-vendored; reference pin ``rank_bm25>=0.2.1``,
  ``requirements.txt:7``; call sites ``src/indexing/bm25_indexer.py:102-106,127``):
  one ``{term: tf}`` dict per document, Python loops, NumPy fp64.
* ``Bm25Csr`` -- tier B, the same arithmetic in the same operation order over a
  CSR term->postings layout so it runs at 1M-10M documents.  ``tests/`` prove
  it bit-identical to tier A on randomised corpora.

``bm25_search`` restates the reference's own wrapper
(``src/indexing/bm25_indexer.py:113-163``): argsort, ``top_k*2`` over-fetch,
equality post-filter, stop at ``top_k``.
"""
from __future__ import annotations

import math

import numpy as np


class BM25OkapiRestated:
    """rank_bm25.BM25Okapi, restated.  Attributes mirror the library's."""

    def __init__(self, corpus, tokenizer=None, k1=1.5, b=0.75, epsilon=0.25):
        self.k1, self.b, self.epsilon = k1, b, epsilon
        self.corpus_size = 0
        self.avgdl = 0
        self.doc_freqs = []
        self.idf = {}
        self.doc_len = []
        self.tokenizer = tokenizer
        if tokenizer:
            corpus = [tokenizer(doc) for doc in corpus]
        nd = {}
        total = 0
        for document in corpus:
            self.doc_len.append(len(document))
            total += len(document)
            freqs = {}
            for word in document:
                freqs[word] = freqs.get(word, 0) + 1
            self.doc_freqs.append(freqs)
            for word in freqs:
                nd[word] = nd.get(word, 0) + 1
            self.corpus_size += 1
        self.avgdl = total / self.corpus_size
        # idf with the epsilon floor; the running sum follows dict insertion order
        idf_sum = 0
        negative = []
        for word, df in nd.items():
            v = math.log(self.corpus_size - df + 0.5) - math.log(df + 0.5)
            self.idf[word] = v
            idf_sum += v
            if v < 0:
                negative.append(word)
        self.average_idf = idf_sum / len(self.idf)
        eps = self.epsilon * self.average_idf
        for word in negative:
            self.idf[word] = eps

    def get_scores(self, query):
        score = np.zeros(self.corpus_size)
        doc_len = np.array(self.doc_len)
        for q in query:  # every OCCURRENCE counts; OOV -> idf 0
            q_freq = np.array([(doc.get(q) or 0) for doc in self.doc_freqs])
            score += (self.idf.get(q) or 0) * (q_freq * (self.k1 + 1) /
                                               (q_freq + self.k1 * (1 - self.b + self.b * doc_len / self.avgdl)))
        return score


class Bm25Csr:
    """Tier B: identical arithmetic over term-major CSR postings.

    Build from token-id documents (``doc_offsets`` int64 [N+1], ``tokens`` int32).
    Term ids stand for strings; ``-1`` in a query is an OOV token.
    """

    def __init__(self, doc_offsets, tokens, vocab, k1=1.5, b=0.75, epsilon=0.25):
        self.k1, self.b, self.epsilon = k1, b, epsilon
        n = len(doc_offsets) - 1
        self.corpus_size = n
        self.doc_len = np.diff(doc_offsets).astype(np.int64)
        self.avgdl = int(doc_offsets[-1]) / n
        doc_of_tok = np.repeat(np.arange(n, dtype=np.int64), self.doc_len)
        key = tokens.astype(np.int64) * n + doc_of_tok            # (term, doc) pairs
        ukey, first_pos, tf = np.unique(key, return_index=True, return_counts=True)
        self.post_term = (ukey // n).astype(np.int64)
        self.post_doc = (ukey % n).astype(np.int64)
        self.post_tf = tf.astype(np.int64)
        self.term_offsets = np.zeros(vocab + 1, dtype=np.int64)
        np.add.at(self.term_offsets, self.post_term + 1, 1)
        np.cumsum(self.term_offsets, out=self.term_offsets)
        df = np.diff(self.term_offsets)
        # dict-insertion order of rank_bm25's ``nd``: first document containing the
        # term, then position of the term's first occurrence inside that document.
        first_tok_pos = np.full(vocab, np.iinfo(np.int64).max, dtype=np.int64)
        np.minimum.at(first_tok_pos, tokens.astype(np.int64), np.arange(len(tokens), dtype=np.int64))
        present = np.flatnonzero(df > 0)
        order = present[np.argsort(first_tok_pos[present], kind="stable")]
        self.idf = np.zeros(vocab, dtype=np.float64)
        idf_sum = 0
        neg = []
        for t in order:
            v = math.log(n - int(df[t]) + 0.5) - math.log(int(df[t]) + 0.5)
            self.idf[t] = v
            idf_sum += v
            if v < 0:
                neg.append(t)
        self.average_idf = idf_sum / len(order)
        self.idf[np.array(neg, dtype=np.int64)] = self.epsilon * self.average_idf
        self.vocab_present = len(order)

    def get_scores(self, query_terms):
        score = np.zeros(self.corpus_size)
        for t in query_terms:
            t = int(t)
            if t < 0:
                continue  # OOV: idf.get(q) or 0 -> adds +0.0 everywhere
            a, e = self.term_offsets[t], self.term_offsets[t + 1]
            docs = self.post_doc[a:e]
            q_freq = self.post_tf[a:e]
            dl = self.doc_len[docs]
            score[docs] += self.idf[t] * (q_freq * (self.k1 + 1) /
                                          (q_freq + self.k1 * (1 - self.b + self.b * dl / self.avgdl)))
        return score


def stable_desc_order(scores: np.ndarray) -> np.ndarray:
    """``np.argsort(scores)[::-1]`` with the tie order made deterministic:
    among equal scores the LARGER row index comes first."""
    return np.argsort(scores, kind="stable")[::-1]


def bm25_search(scores: np.ndarray, top_k: int, row_passes=None):
    """Reference wrapper ``BM25Indexer.search`` (bm25_indexer.py:127-163).

    ``top_indices = argsort(scores)[::-1][:top_k*2]``; walk them, skip rows failing
    the equality post-filter (``row_passes(idx) -> bool``), stop at ``top_k``.
    Zero-score rows ARE returned.  Returns (rows int64, scores float64).
    """
    top = stable_desc_order(scores)[: top_k * 2]
    rows, out = [], []
    for idx in top:
        if row_passes is not None and not row_passes(int(idx)):
            continue
        rows.append(int(idx))
        out.append(float(scores[idx]))
        if len(rows) >= top_k:
            break
    return np.array(rows, dtype=np.int64), np.array(out, dtype=np.float64)
