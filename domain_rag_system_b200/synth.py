"""Synthetic corpora for parity tests and benchmarks.

Nothing here is on the retrieval path; it only manufactures inputs of the
shape the reference's index holds:

* chunk embeddings: random unit vectors, 768-d, rounded to bf16 (the rounded
  values ARE the input; every consumer -- CUDA path and oracle -- sees
  ``bf16(x).float()``), as produced by ``MedicalEmbedder.encode``
  (reference ``src/indexing/embedder.py:50-73``, ``normalize_embeddings=True``).
* chunk token lists: Zipf(s=1.07) draws over a 50k vocabulary rendered
  ``w00000``..; these strings are fixed points of ``BM25Indexer.tokenize``
  (reference ``src/indexing/bm25_indexer.py:50-78``).
"""
from __future__ import annotations

import numpy as np
import torch

EMB_DIM = 768
VOCAB = 50_000
ZIPF_S = 1.07
DOC_LEN_MIN, DOC_LEN_MAX = 32, 96
QUERY_TOKENS = 8


# --------------------------------------------------------------------------
# embeddings
# --------------------------------------------------------------------------
def unit_embeddings(rows: int, dim: int = EMB_DIM, seed: int = 1234, device: str = "cpu",
                    dup_frac: float = 0.0) -> torch.Tensor:
    """bf16 [rows, dim] random unit vectors.

    ``dup_frac`` > 0 overwrites that fraction of rows with copies of other rows
    (seed 7) so exact score ties exist and the tie rule is exercised.
    """
    g = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty(rows, dim, dtype=torch.bfloat16, device=device)
    step = 1 << 18
    for s in range(0, rows, step):
        e = min(rows, s + step)
        x = torch.randn(e - s, dim, generator=g, device=device, dtype=torch.float32)
        x /= x.norm(dim=1, keepdim=True)
        out[s:e] = x.to(torch.bfloat16)
    if dup_frac > 0 and rows > 1:
        ndup = max(1, int(rows * dup_frac))
        g7 = torch.Generator(device="cpu").manual_seed(7)
        dst = torch.randint(0, rows, (ndup,), generator=g7)
        src = torch.randint(0, rows, (ndup,), generator=g7)
        out[dst.to(device)] = out[src.to(device)]
    return out


def query_embeddings(q: int, corpus: torch.Tensor | None = None, dim: int = EMB_DIM,
                     seed: int = 4321, near_frac: float = 0.1) -> torch.Tensor:
    """fp32 [q, dim] unit query vectors (the embedder hands out fp32).

    ``near_frac`` of them are noisy copies of corpus rows so that a few queries
    have a clear nearest neighbour.
    """
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(q, dim, generator=g, dtype=torch.float32)
    if corpus is not None and near_frac > 0 and corpus.shape[0] > 0:
        g99 = torch.Generator().manual_seed(99)
        nnear = int(q * near_frac)
        if nnear:
            pick = torch.randint(0, corpus.shape[0], (nnear,), generator=g99)
            base = corpus[pick.to(corpus.device)].float().cpu()
            x[:nnear] = base + 0.5 * x[:nnear] / x[:nnear].norm(dim=1, keepdim=True)
    x /= x.norm(dim=1, keepdim=True)
    return x


# --------------------------------------------------------------------------
# BM25 corpus
# --------------------------------------------------------------------------
def _zipf_cdf(vocab: int, s: float) -> np.ndarray:
    w = 1.0 / np.power(np.arange(1, vocab + 1, dtype=np.float64), s)
    c = np.cumsum(w)
    return c / c[-1]


def zipf_docs(rows: int, vocab: int = VOCAB, seed: int = 2025, s: float = ZIPF_S,
              len_min: int = DOC_LEN_MIN, len_max: int = DOC_LEN_MAX):
    """Token-id documents in CSR form: (doc_offsets int64 [rows+1], tokens int32 [total])."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(len_min, len_max + 1, size=rows, dtype=np.int64)
    offs = np.zeros(rows + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    cdf = _zipf_cdf(vocab, s)
    total = int(offs[-1])
    toks = np.empty(total, dtype=np.int32)
    step = 1 << 24
    for a in range(0, total, step):
        b = min(total, a + step)
        toks[a:b] = np.minimum(np.searchsorted(cdf, rng.random(b - a), side="left"), vocab - 1)
    return offs, toks


def zipf_queries(q: int, doc_offsets: np.ndarray, tokens: np.ndarray, vocab: int = VOCAB,
                 seed: int = 777, s: float = ZIPF_S, ntok: int = QUERY_TOKENS):
    """List of int32 arrays of term ids; -1 marks an out-of-vocabulary token.

    Half of the tokens are Zipf draws, half are copied from a random document.
    One query in 20 repeats a token and carries one OOV token.
    """
    rng = np.random.default_rng(seed)
    cdf = _zipf_cdf(vocab, s)
    rows = len(doc_offsets) - 1
    out = []
    for i in range(q):
        t = np.minimum(np.searchsorted(cdf, rng.random(ntok)), vocab - 1).astype(np.int32)
        if rows > 0:
            d = int(rng.integers(0, rows))
            lo, hi = int(doc_offsets[d]), int(doc_offsets[d + 1])
            copy = rng.random(ntok) < 0.5
            if hi > lo:
                t[copy] = tokens[rng.integers(lo, hi, size=int(copy.sum()))]
        if i % 20 == 19:
            t[1] = t[0]
            t[-1] = -1
        out.append(t)
    return out


def term_string(t: int) -> str:
    return "zzzz" if t < 0 else f"w{int(t):05d}"


def docs_as_text(doc_offsets: np.ndarray, tokens: np.ndarray, lo: int = 0, hi: int | None = None):
    """Render token-id docs as whitespace-joined strings (small corpora only)."""
    hi = len(doc_offsets) - 1 if hi is None else hi
    return [" ".join(term_string(t) for t in tokens[doc_offsets[d]:doc_offsets[d + 1]])
            for d in range(lo, hi)]


def make_chunks(texts, prefix: str = "chunk"):
    """Chunk dicts with the reference schema (``src/preprocessing/chunker.py:9-26``)."""
    return [{"chunk_id": f"{prefix}_{i}", "doc_id": f"doc_{i // 4}", "source": "synthetic",
             "chunk_type": "body", "text": t, "metadata": {"row": i}}
            for i, t in enumerate(texts)]


# --------------------------------------------------------------------------
# block-seeded corpora: the SAME corpus whatever the number of shards
# --------------------------------------------------------------------------
N_BLOCKS = 64


def block_bounds(n_total: int, block: int, n_blocks: int = N_BLOCKS):
    """Rows [lo, hi) of block `block` (contiguous, sizes differ by at most one)."""
    base, rem = divmod(n_total, n_blocks)
    lo = block * base + min(block, rem)
    return lo, lo + base + (1 if block < rem else 0)


def shard_blocks(world: int, rank: int, n_blocks: int = N_BLOCKS):
    """Blocks of rank `rank` when the corpus is row-sharded over `world` ranks (world must divide n_blocks, so that
    shard boundaries are block boundaries and every rank count sees the identical corpus)."""
    if n_blocks % world != 0:
        raise ValueError(f"world size {world} must divide {n_blocks}")
    per = n_blocks // world
    return range(rank * per, (rank + 1) * per)


def block_embeddings(n_total: int, block: int, dim: int = EMB_DIM, device: str = "cpu") -> torch.Tensor:
    lo, hi = block_bounds(n_total, block)
    return unit_embeddings(hi - lo, dim=dim, seed=1234 + block, device=device)


def zipf_docs_torch(rows: int, vocab: int = VOCAB, seed: int = 2025, s: float = ZIPF_S, len_min: int = DOC_LEN_MIN,
                    len_max: int = DOC_LEN_MAX, device: str = "cpu"):
    """Same distribution as ``zipf_docs`` drawn with torch's generator on `device` (the numpy version takes a minute
    per 10 M documents): (doc_offsets int64 [rows+1], tokens int32 [total]) as tensors on `device`."""
    g = torch.Generator(device=device).manual_seed(seed)
    lens = torch.randint(len_min, len_max + 1, (rows,), generator=g, device=device, dtype=torch.int64)
    offs = torch.zeros(rows + 1, dtype=torch.int64, device=device)
    torch.cumsum(lens, 0, out=offs[1:])
    cdf = torch.from_numpy(_zipf_cdf(vocab, s)).to(device)
    total = int(offs[-1])
    toks = torch.empty(total, dtype=torch.int32, device=device)
    step = 1 << 24
    for a in range(0, total, step):
        b = min(total, a + step)
        u = torch.rand(b - a, generator=g, device=device, dtype=torch.float64)
        toks[a:b] = torch.searchsorted(cdf, u).clamp_(max=vocab - 1).to(torch.int32)
    return offs, toks


def block_docs(n_total: int, block: int, vocab: int = VOCAB, device: str = "cpu"):
    lo, hi = block_bounds(n_total, block)
    return zipf_docs_torch(hi - lo, vocab=vocab, seed=2025 + block, device=device)


def shard_corpus(n_total: int, world: int, rank: int, vocab: int = VOCAB, dim: int = EMB_DIM, device: str = "cpu"):
    """This rank's shard of the block-seeded corpus: (embeddings bf16 [n_local, dim], doc_offsets int64 [n_local+1],
    tokens int32) on `device`, plus the global row of its first row."""
    blocks = shard_blocks(world, rank)
    lo = block_bounds(n_total, blocks[0])[0]
    hi = block_bounds(n_total, blocks[-1])[1]
    emb = torch.empty(hi - lo, dim, dtype=torch.bfloat16, device=device)
    offs_parts, tok_parts, base = [torch.zeros(1, dtype=torch.int64, device=device)], [], 0
    for b in blocks:
        a, e = block_bounds(n_total, b)
        emb[a - lo:e - lo] = block_embeddings(n_total, b, dim=dim, device=device)
        o, t = block_docs(n_total, b, vocab=vocab, device=device)
        offs_parts.append(o[1:] + base)
        tok_parts.append(t)
        base += int(o[-1])
    return emb, torch.cat(offs_parts), torch.cat(tok_parts), lo


def block_queries(q: int, n_total: int, vocab: int = VOCAB, dim: int = EMB_DIM, device: str = "cpu"):
    """The query batch of a block-seeded corpus: fp32 unit vectors (seed 4321) and token-id queries whose copied
    tokens come from documents of block 0 (so every rank can build them without seeing the other shards)."""
    qv = query_embeddings(q, dim=dim, seed=4321, near_frac=0.0)
    o, t = block_docs(n_total, 0, vocab=vocab, device=device)
    qt = zipf_queries(q, o.cpu().numpy(), t.cpu().numpy(), vocab=vocab, seed=777)
    return qv, qt
