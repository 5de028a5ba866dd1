"""ctypes binding of ``include/drs_b200.h`` (the C-ABI CUDA library).

PyTorch owns every buffer; this module only passes raw pointers, sizes and the current
CUDA stream.  There is NO fallback: if the library is missing the import of the product
fails loudly (mirrors the reference's ``ImportError`` when qdrant-client / rank_bm25 are
absent, ``src/indexing/vector_indexer.py:44-45``, ``bm25_indexer.py:36-37``).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libdrs_b200.so")

DENSE_AUTO, DENSE_SIMT, DENSE_UMMA, DENSE_UMMA_TS, DENSE_UMMA_PAIR = 0, 1, 2, 3, 4
MAX_KPRIME, MAX_DIM, MAX_POOL = 128, 1024, 1024

_vp, _i32, _i64, _f64, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_size_t


class Bm25IndexStruct(C.Structure):
    """``drs_bm25_index`` (include/drs_b200.h)."""
    _fields_ = [("n_docs", _i64), ("n_terms", C.c_int32), ("tile_docs", C.c_int32), ("n_tiles", C.c_int32),
                ("impact_scale_log2", C.c_int32), ("d_term_offsets", _vp), ("d_post_doc", _vp), ("d_post_impact", _vp),
                ("d_post_tf", _vp), ("d_tile_ptr", _vp), ("d_idf", _vp), ("d_doc_kd", _vp), ("k1_plus_1", _f64),
                ("d_dense_slot", _vp), ("d_dense_impact", _vp), ("dense_stride", _i64), ("dense_shift", C.c_int32),
                ("impacts_nonneg", C.c_int32)]


class SingleChars(C.Structure):
    """``drs_single_chars``: bit c set = the one-character token chr(c) is kept."""
    _fields_ = [("bits", C.c_uint32 * 8)]


class VocabTableStruct(C.Structure):
    """``drs_vocab_table`` (include/drs_b200.h)."""
    _fields_ = [("d_slot_hash", _vp), ("d_slot_term", _vp), ("d_str_offsets", _vp), ("d_pool", _vp),
                ("capacity", _i64), ("n_slots", _i64)]


EXPORTS = {
    "drs_version": (C.c_int, []),
    "drs_last_error": (C.c_char_p, []),
    "drs_device_sm_count": (C.c_int, []),
    "drs_dense_last_kernel": (C.c_int, []),
    "drs_kernel_launches": (C.c_ulonglong, []),
    "drs_profile_enable": (C.c_int, [_i32]),
    "drs_profile_read": (C.c_int, [_i32, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "drs_row_inv_norms": (C.c_int, [_vp, _i64, _i32, _vp, _vp]),
    "drs_dense_workspace_bytes": (_sz, [_i32, _i64, _i32, _i32, _i32]),
    "drs_dense_topk": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "drs_bm25_workspace_bytes": (_sz, [_i32, C.POINTER(Bm25IndexStruct), _i32]),
    "drs_bm25_topk": (C.c_int, [C.POINTER(Bm25IndexStruct), _vp, _vp, _i32, _i32, _vp, _vp, _vp, _sz, _vp]),
    "drs_bm25_topk_ex": (C.c_int, [C.POINTER(Bm25IndexStruct), _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _sz,
                                   _vp]),
    "drs_fuse_topk": (C.c_int, [_i32, _i32, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _f64, _f64, _f64, _i32, _vp, _vp, _vp, _vp]),
    "drs_rerank_cosine": (C.c_int, [_vp, _i32, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "drs_merge_candidates_f32": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "drs_merge_candidates_f64": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "drs_candidate_cosines": (C.c_int, [_vp, _i32, _vp, _vp, _i32, _i32, _vp, _vp]),
    "drs_merge_gathered": (C.c_int, [_vp, _i32, _sz, _sz, _i32, _sz, _sz, _i32, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "drs_rerank_from_cosines": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "drs_tokenize_count": (C.c_int, [_vp, _vp, _i32, C.POINTER(SingleChars), _vp, _vp]),
    "drs_tokenize_lookup": (C.c_int, [_vp, _vp, _i32, C.POINTER(SingleChars), C.POINTER(VocabTableStruct), _vp, _vp, _vp]),
}

_lib = None


class DrsError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises ImportError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: build it with `python -m domain_rag_system_b200.build` "
                              "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def _check(rc: int):
    if rc != 0:
        msg = load().drs_last_error().decode("utf-8", "replace")
        if rc == -2:
            raise NotImplementedError(f"drs: {msg}")
        if rc == -1:
            raise ValueError(f"drs: {msg}")
        raise DrsError(f"drs error {rc}: {msg}")


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise ValueError("drs: device tensors required (no CPU path exists)")


PROF_KINDS = {"dense_select": 0, "dense_tail": 1, "bm25_select": 2, "bm25_tail": 3, "fuse": 4, "rerank": 5}


def profile_enable(on: bool):
    _check(load().drs_profile_enable(1 if on else 0))


def profile_read():
    """{kind: (total_ms, launches)} since the last profile_enable(True); synchronises."""
    out = {}
    for name, k in PROF_KINDS.items():
        ms, n = C.c_double(0.0), C.c_int(0)
        _check(load().drs_profile_read(k, C.byref(ms), C.byref(n)))
        out[name] = (ms.value, n.value)
    return out


class Workspace:
    """Grow-only device scratch buffer handed to the library on every call."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes: int, device):
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        return self.buf


# ------------------------------------------------------------------------------------
# thin wrappers: tensors in, tensors out
# ------------------------------------------------------------------------------------
def row_inv_norms(corpus_bf16: torch.Tensor) -> torch.Tensor:
    _need_cuda(corpus_bf16)
    n, d = corpus_bf16.shape
    out = torch.empty(n, dtype=torch.float32, device=corpus_bf16.device)
    with torch.cuda.device(corpus_bf16.device):
        _check(load().drs_row_inv_norms(_ptr(corpus_bf16), n, d, _ptr(out), _stream()))
    return out


def dense_topk(queries_f32, corpus_bf16, inv_norm, kprime, row_mask=None, kernel=DENSE_AUTO, escalate=True,
               ws: Workspace | None = None, out=None):
    """Returns (scores fp32 [Q,k'], ids int64 [Q,k'], flags int32 [Q])."""
    _need_cuda(queries_f32, corpus_bf16, inv_norm, row_mask)
    lib = load()
    q, d = queries_f32.shape
    n = corpus_bf16.shape[0]
    assert queries_f32.dtype == torch.float32 and queries_f32.is_contiguous()
    assert corpus_bf16.dtype == torch.bfloat16 and corpus_bf16.is_contiguous() and corpus_bf16.shape[1] == d
    dev = corpus_bf16.device
    if out is None:
        scores = torch.empty(q, kprime, dtype=torch.float32, device=dev)
        ids = torch.empty(q, kprime, dtype=torch.int64, device=dev)
        flags = torch.empty(q, dtype=torch.int32, device=dev)
    else:
        scores, ids, flags = out
    ws = ws or Workspace()
    nbytes = lib.drs_dense_workspace_bytes(q, n, d, kprime, kernel)
    buf = ws.get(nbytes, dev)
    with torch.cuda.device(dev):
        _check(lib.drs_dense_topk(_ptr(queries_f32), q, _ptr(corpus_bf16), _ptr(inv_norm), _ptr(row_mask), n, d,
                                  kprime, kernel, 1 if escalate else 0, _ptr(scores), _ptr(ids), _ptr(flags),
                                  _ptr(buf), buf.numel(), _stream()))
    return scores, ids, flags


BM25_MODES = {"auto": 0, "no_escalation": 1, "escalate_all": 2}


def bm25_topk(index_struct: Bm25IndexStruct, q_terms, q_offsets, n_queries, kprime, ws: Workspace | None = None,
              out=None, mode: str = "auto", want_flags: bool = False):
    """Returns (scores fp64 [Q,k'], ids int64 [Q,k']) -- plus flags int32 [Q] with ``want_flags``:
    0 = the integer pre-selection was proven exact, 2 = resolved by the exact escalation pass,
    1 = unproven (only with mode "no_escalation")."""
    _need_cuda(q_terms, q_offsets)
    lib = load()
    dev = q_offsets.device
    if out is None:
        scores = torch.empty(n_queries, kprime, dtype=torch.float64, device=dev)
        ids = torch.empty(n_queries, kprime, dtype=torch.int64, device=dev)
    else:
        scores, ids = out
    ws = ws or Workspace()
    nbytes = lib.drs_bm25_workspace_bytes(n_queries, C.byref(index_struct), kprime)
    buf = ws.get(nbytes, dev)
    flags = torch.empty(n_queries, dtype=torch.int32, device=dev) if want_flags else None
    with torch.cuda.device(dev):
        _check(lib.drs_bm25_topk_ex(C.byref(index_struct), _ptr(q_terms), _ptr(q_offsets), n_queries, kprime,
                                    BM25_MODES[mode], _ptr(scores), _ptr(ids), _ptr(flags) if want_flags else None,
                                    _ptr(buf), buf.numel(), _stream()))
    return (scores, ids, flags) if want_flags else (scores, ids)


def tokenize_lookup(text: torch.Tensor, text_offsets: torch.Tensor, keep1: SingleChars, table: VocabTableStruct):
    """ASCII query strings (uint8 bytes back to back + int64 offsets, device tensors) -> (q_terms int32, q_offsets
    int32 [Q+1]) on the device, the pair ``bm25_topk`` reads: the reference tokeniser + vocabulary lookup."""
    _need_cuda(text_offsets)
    lib = load()
    dev = text_offsets.device
    n = text_offsets.numel() - 1
    counts = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _check(lib.drs_tokenize_count(_ptr(text), _ptr(text_offsets), n, C.byref(keep1), _ptr(counts), _stream()))
        offsets = torch.zeros(n + 1, dtype=torch.int32, device=dev)
        if n:
            offsets[1:] = torch.cumsum(counts[:n], 0)
        total = int(offsets[-1].item()) if n else 0
        terms = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
        _check(lib.drs_tokenize_lookup(_ptr(text), _ptr(text_offsets), n, C.byref(keep1), C.byref(table), _ptr(offsets),
                                       _ptr(terms), _stream()))
    return terms[:total], offsets


def fuse_topk(mode, v_scores, v_ids, b_scores, b_ids, top_k, fusion_key=None, rrf_k=60.0, w_vector=0.5, w_bm25=0.5,
              want_norm=False):
    """mode 'rrf' | 'weighted'.  Either side may be None (that retriever is absent).
    Returns (ids int64 [Q,K], scores fp64 [Q,K], norm fp64 [Q,K] or None)."""
    lib = load()
    ref = v_ids if v_ids is not None else b_ids
    _need_cuda(ref, fusion_key)
    q = ref.shape[0]
    dev = ref.device
    kv = 0 if v_ids is None else v_ids.shape[1]
    kb = 0 if b_ids is None else b_ids.shape[1]
    out_ids = torch.empty(q, top_k, dtype=torch.int64, device=dev)
    out_scores = torch.empty(q, top_k, dtype=torch.float64, device=dev)
    out_norm = torch.empty(q, top_k, dtype=torch.float64, device=dev) if want_norm else None
    with torch.cuda.device(dev):
        v_f64 = 1 if (v_scores is not None and v_scores.dtype == torch.float64) else 0
        _check(lib.drs_fuse_topk(0 if mode == "rrf" else 1, q, kv, kb, _ptr(v_scores), v_f64, _ptr(v_ids), _ptr(b_scores),
                                 _ptr(b_ids), _ptr(fusion_key), float(rrf_k), float(w_vector), float(w_bm25), top_k,
                                 _ptr(out_ids), _ptr(out_scores), _ptr(out_norm), _stream()))
    return out_ids, out_scores, out_norm


def rerank_cosine(corpus_bf16, rq_f32, pool_ids, top_k):
    """Returns (scores fp32 [Q,top_k], ids int64 [Q,top_k], pos int32 [Q,top_k])."""
    _need_cuda(corpus_bf16, rq_f32, pool_ids)
    lib = load()
    q, top_n = pool_ids.shape
    d = corpus_bf16.shape[1]
    dev = corpus_bf16.device
    scores = torch.empty(q, top_k, dtype=torch.float32, device=dev)
    ids = torch.empty(q, top_k, dtype=torch.int64, device=dev)
    pos = torch.empty(q, top_k, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _check(lib.drs_rerank_cosine(_ptr(corpus_bf16), d, _ptr(rq_f32), _ptr(pool_ids), q, top_n, top_k,
                                     _ptr(scores), _ptr(ids), _ptr(pos), _stream()))
    return scores, ids, pos


def merge_candidates(scores, ids, k_out):
    """scores/ids [P, Q, k] (fp32 or fp64 / int64, global ids) -> ([Q,k_out], [Q,k_out])."""
    _need_cuda(scores, ids)
    lib = load()
    p, q, k = ids.shape
    dev = ids.device
    out_s = torch.empty(q, k_out, dtype=scores.dtype, device=dev)
    out_i = torch.empty(q, k_out, dtype=torch.int64, device=dev)
    fn = lib.drs_merge_candidates_f32 if scores.dtype == torch.float32 else lib.drs_merge_candidates_f64
    with torch.cuda.device(dev):
        _check(fn(_ptr(scores.contiguous()), _ptr(ids.contiguous()), p, q, k, k_out, _ptr(out_s), _ptr(out_i),
                  _stream()))
    return out_s, out_i


def candidate_cosines(corpus_bf16, rq_f32, ids, out=None):
    """cos(rq[q], corpus[ids[q, j]]) in the simple reranker's exact fp32 arithmetic; ids [Q, n] LOCAL rows, -1 -> 0."""
    _need_cuda(corpus_bf16, rq_f32, ids)
    q, n = ids.shape
    out = torch.empty(q, n, dtype=torch.float32, device=ids.device) if out is None else out
    with torch.cuda.device(ids.device):
        _check(load().drs_candidate_cosines(_ptr(corpus_bf16), corpus_bf16.shape[1], _ptr(rq_f32), _ptr(ids), q, n,
                                            _ptr(out), _stream()))
    return out


def merge_gathered(gathered, n_parts, part_stride, off_scores, scores_f64, off_ids, off_cos, id_offset, q, k, k_out,
                   out=None):
    """Merge the parts' [Q, k] lists out of a gathered byte buffer -> (scores, ids int64 GLOBAL, cos fp32 or None)."""
    _need_cuda(gathered, id_offset)
    dev = gathered.device
    has_cos = off_cos is not None
    if out is None:
        out = (torch.empty(q, k_out, dtype=torch.float64 if scores_f64 else torch.float32, device=dev),
               torch.empty(q, k_out, dtype=torch.int64, device=dev),
               torch.empty(q, k_out, dtype=torch.float32, device=dev) if has_cos else None)
    with torch.cuda.device(dev):
        _check(load().drs_merge_gathered(_ptr(gathered), n_parts, part_stride, off_scores, 1 if scores_f64 else 0, off_ids,
                                         off_cos if has_cos else 0, 1 if has_cos else 0, _ptr(id_offset), q, k, k_out,
                                         _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _stream()))
    return out


def rerank_from_cosines(pool_ids, top_k, v_ids, v_cos, b_ids, b_cos, out=None):
    """Stable sort of the fused pool by cosines carried in the merged lists -> (scores fp32, ids int64, pos int32)."""
    ref = pool_ids
    _need_cuda(ref, v_ids, v_cos, b_ids, b_cos)
    q, top_n = pool_ids.shape
    dev = ref.device
    if out is None:
        out = (torch.empty(q, top_k, dtype=torch.float32, device=dev), torch.empty(q, top_k, dtype=torch.int64, device=dev),
               torch.empty(q, top_k, dtype=torch.int32, device=dev))
    kv = 0 if v_ids is None else v_ids.shape[1]
    kb = 0 if b_ids is None else b_ids.shape[1]
    with torch.cuda.device(dev):
        _check(load().drs_rerank_from_cosines(_ptr(pool_ids), q, top_n, top_k, _ptr(v_ids), _ptr(v_cos), kv, _ptr(b_ids),
                                              _ptr(b_cos), kb, _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _stream()))
    return out
