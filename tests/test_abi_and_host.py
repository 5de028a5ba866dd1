"""CPU-only checks: the C-ABI library loads and exports every symbol the header declares,
the host-side index builder matches the oracle, the tokenizer matches the reference."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from domain_rag_system_b200 import _lib, synth
from domain_rag_system_b200.indexing.bm25_index import Bm25DeviceIndex, choose_tile_docs
from helpers import load_golden
from oracle import bm25 as obm25

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "drs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(drs_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _header_functions()
    assert len(names) >= 12
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/drs_b200.h but not exported"
    assert set(_lib.EXPORTS) == set(names)
    assert _lib.load().drs_version() >= 100


def test_product_has_no_cpu_path():
    """Host tensors are refused; nothing in the package imports the oracle."""
    with pytest.raises(ValueError):
        _lib.row_inv_norms(torch.zeros(4, 8, dtype=torch.bfloat16))
    pkg = os.path.join(ROOT, "domain_rag_system_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"


def test_tokenizer_matches_reference_golden():
    from domain_rag_system_b200.indexing.bm25_indexer import BM25Indexer
    bi = BM25Indexer.__new__(BM25Indexer)          # tokenizer only; no device needed
    from domain_rag_system_b200.indexing.bm25_indexer import _NoAbbreviations
    bi.normalizer = _NoAbbreviations()
    for text, toks in load_golden("tokenizer_golden.json"):
        assert bi.tokenize(text) == toks

    class Abbr:   # a 1-character abbreviation key keeps that token alive (bm25_indexer.py:64-76)
        abbreviations = {"K": "potassium"}
    bi.normalizer = Abbr()
    assert bi.tokenize("serum K level") == ["serum", "k", "level"]


@pytest.mark.parametrize("n,vocab", [(1, 4), (37, 11), (600, 400), (5000, 900)])
def test_host_index_build_matches_oracle(n, vocab):
    offs, toks = synth.zipf_docs(n, vocab=vocab, seed=5 + n, len_min=1, len_max=40)
    ix = Bm25DeviceIndex.from_token_ids(offs, toks, vocab, k1=1.5, b=0.75, device="cpu", tile_docs=256)
    orc = obm25.Bm25Csr(offs, toks, vocab)
    assert ix.avgdl == orc.avgdl and ix.average_idf == orc.average_idf
    assert np.array_equal(ix.t["idf"].numpy(), orc.idf)
    assert np.array_equal(ix.t["term_offsets"].numpy(), orc.term_offsets)
    nnz = ix.nnz
    assert nnz == orc.post_doc.size and ix.t["post_doc"].numel() == nnz + 4   # 4 zeroed pad entries
    assert not ix.t["post_doc"][nnz:].any() and not ix.t["post_impact"][nnz:].any()
    assert np.array_equal(ix.t["post_doc"].numpy()[:nnz].astype(np.int64), orc.post_doc)
    assert np.array_equal(ix.t["post_tf"].numpy().astype(np.int64), orc.post_tf)
    kd = 1.5 * (1 - 0.75 + 0.75 * orc.doc_len / orc.avgdl)
    assert np.array_equal(ix.t["doc_kd"].numpy(), kd)
    # tile offsets: first posting of term t with doc >= j*tile
    tp = ix.t["tile_ptr"].numpy()
    for t in np.random.default_rng(0).integers(0, vocab, size=min(vocab, 25)):
        docs = orc.post_doc[orc.term_offsets[t]:orc.term_offsets[t + 1]]
        want = np.searchsorted(docs, np.arange(ix.n_tiles + 1) * ix.tile_docs, side="left")
        assert np.array_equal(tp[t], want)
    # stored impacts are the fp64 contributions in fixed point: round(impact * 2^s), |.| <= 2^26
    t_of = np.repeat(np.arange(vocab), np.diff(orc.term_offsets))
    imp = orc.idf[t_of] * (orc.post_tf * 2.5 / (orc.post_tf + kd[orc.post_doc]))
    s = ix.impact_scale_log2
    fx = ix.t["post_impact"].numpy()[:nnz].astype(np.int64)
    assert np.abs(fx).max() <= 2 ** 26 and np.abs(orc.idf).max() * 2.5 * 2.0 ** (s + 1) > 2 ** 26
    assert np.abs(fx - imp * 2.0 ** s).max() <= 0.5
    st = ix.struct()   # the C struct carries the scale (the kernel's exactness bound is computed from it)
    assert st.impact_scale_log2 == s and st.n_docs == n and st.n_tiles == ix.n_tiles


def test_string_corpus_build_assigns_ids_in_first_appearance_order():
    docs = [["bb", "aa", "bb"], ["cc", "aa"], []]
    ix = Bm25DeviceIndex.from_token_lists(docs, device="cpu", tile_docs=256)
    assert ix.vocab == {"bb": 0, "aa": 1, "cc": 2}
    ref = obm25.BM25OkapiRestated(docs)
    assert ix.avgdl == ref.avgdl and ix.average_idf == ref.average_idf
    for w, i in ix.vocab.items():
        assert ix.t["idf"][i].item() == ref.idf[w]
    qt, qo = ix.encode_queries([["aa", "zz", "aa"], []])
    assert qt.tolist() == [1, -1, 1] and qo.tolist() == [0, 3, 3]


def test_tile_size_choice():
    assert choose_tile_docs(600) == 256
    assert choose_tile_docs(20_000) == 2048
    assert choose_tile_docs(100_000) == 8192
    assert choose_tile_docs(1_000_000) == 8192
    assert choose_tile_docs(10_000_000) == 8192


def test_dense_head_term_rows_match_postings():
    offs, toks = synth.zipf_docs(3000, vocab=60, seed=77, len_min=5, len_max=40)
    ix = Bm25DeviceIndex.from_token_ids(offs, toks, 60, device="cpu", tile_docs=256, dense_frac=0.25)
    assert ix.n_dense > 0
    slot = ix.t["dense_slot"].numpy()
    df = np.diff(ix.t["term_offsets"].numpy())
    assert ((slot >= 0) == (df >= 0.25 * 3000)).all()
    dense = ix.t["dense_impact"]
    assert dense.dtype == torch.int16 and dense.shape[1] == ix.n_tiles * ix.tile_docs
    # undo the per-tile interleave: 16-byte group i = docs 4i..4i+3 of the first half + 4i..4i+3 of the second
    t8 = ix.tile_docs // 8
    plain = dense.view(-1, ix.n_tiles, t8, 2, 4).permute(0, 1, 3, 2, 4).reshape(dense.shape[0], -1).numpy()
    to, pd, pi = ix.t["term_offsets"].numpy(), ix.t["post_doc"].numpy(), ix.t["post_impact"].numpy().astype(np.int64)
    sh = ix.dense_shift
    assert 0 < plain.max() <= 16383 and (sh == 0 or plain.max() > 8191)     # 14-bit values, range used
    for t in np.flatnonzero(slot >= 0):
        row = np.zeros(plain.shape[1], dtype=np.int64)
        row[pd[to[t]:to[t + 1]]] = (pi[to[t]:to[t + 1]] + ((1 << sh) >> 1)) >> sh
        assert np.array_equal(plain[slot[t]].astype(np.int64), row)
    assert ix.struct().dense_shift == sh
    off = Bm25DeviceIndex.from_token_ids(offs, toks, 60, device="cpu", tile_docs=256, dense_frac=0.0)
    assert off.n_dense == 0


def test_load_reference_style_bm25_pickle_without_rank_bm25(tmp_path):
    """A bm25_index.pkl as the reference writes it (bm25_indexer.py:165-181) holds a rank_bm25.BM25Okapi
    object; rank_bm25 is not installed here.  load() must still work and rebuild the device index from the
    tokenised corpus in the pickle (statistics identical to the pickled object's)."""
    import pickle
    import sys
    import types
    from domain_rag_system_b200.indexing.bm25_indexer import BM25Indexer
    docs = [["aspirin", "pain", "relief"], ["aspirin", "side", "effects", "bleeding"], ["ibuprofen", "pain"]]
    chunks = [{"chunk_id": f"c{i}", "text": " ".join(d)} for i, d in enumerate(docs)]
    mod = types.ModuleType("rank_bm25")               # a throw-away rank_bm25 with the oracle's restatement

    class BM25Okapi(obm25.BM25OkapiRestated):
        pass
    BM25Okapi.__module__ = "rank_bm25"
    BM25Okapi.__qualname__ = "BM25Okapi"
    mod.BM25Okapi = BM25Okapi
    sys.modules["rank_bm25"] = mod
    try:
        ref_obj = BM25Okapi(docs, k1=1.2, b=0.6)
        with open(tmp_path / "bm25_index.pkl", "wb") as f:
            pickle.dump({"bm25": ref_obj, "chunks": chunks, "tokenized_corpus": docs, "k1": 1.2, "b": 0.6}, f)
    finally:
        del sys.modules["rank_bm25"]
    bi = BM25Indexer(k1=1.5, b=0.75, device="cpu")
    bi.load(str(tmp_path / "bm25_index.pkl"))
    assert (bi.k1, bi.b) == (1.2, 0.6) and bi.chunks == chunks and bi.tokenized_corpus == docs
    assert bi.bm25.avgdl == ref_obj.avgdl and bi.bm25.average_idf == ref_obj.average_idf
    for w, i in bi.bm25.vocab.items():
        assert bi.bm25.t["idf"][i].item() == ref_obj.idf[w]
    assert bi.get_stats() == {"num_documents": 3, "avg_doc_length": 3.0, "vocab_size": 7, "k1": 1.2, "b": 0.6}
    # round trip through our own format
    bi.save(str(tmp_path / "ours.pkl"))
    bj = BM25Indexer(device="cpu")
    bj.load(str(tmp_path / "ours.pkl"))
    assert bj.bm25.avgdl == bi.bm25.avgdl and torch.equal(bj.bm25.t["post_doc"], bi.bm25.t["post_doc"])


def test_bm25_struct_layout_matches_the_header(tmp_path):
    """ctypes mirrors `drs_bm25_index` by hand; a field added on one side only (or assigned under a stale name:
    ctypes silently accepts unknown attributes) shifts every pointer after it.  gcc is the referee."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    fields = [f[0] for f in _lib.Bm25IndexStruct._fields_]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "layout.c"
    body = "\n".join(f'  printf("{f} %zu\\n", offsetof(drs_bm25_index, {f}));' for f in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "drs_b200.h"\nint main(void) {\n' + body +
                   '\n  printf("sizeof %zu\\n", sizeof(drs_bm25_index));\n  return 0;\n}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)])
    out = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for f in fields:
        assert int(out[f]) == getattr(_lib.Bm25IndexStruct, f).offset, f
    assert int(out["sizeof"]) == ctypes.sizeof(_lib.Bm25IndexStruct)
    # every header field is mirrored (the C program above would not compile for a field unknown to the header)
    hdr = open(os.path.join(root, "include", "drs_b200.h")).read()
    struct_body = hdr[hdr.index("typedef struct drs_bm25_index {"):hdr.index("} drs_bm25_index;")]
    struct_body = re.sub(r"/\*.*?\*/", "", struct_body, flags=re.S)
    names = re.findall(r"(\w+)\s*;", struct_body)
    assert names == fields
    # and struct() fills every one of them from the index (no stale attribute names)
    offs, toks = synth.zipf_docs(300, vocab=40, seed=1, len_min=2, len_max=20)
    st = Bm25DeviceIndex.from_token_ids(offs, toks, 40, device="cpu", tile_docs=256).struct()
    assert not (set(vars(st)) - set(fields)), "attribute assigned that is not a struct field"


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver times next to ours) must print ONE JSON line with the
    contract's keys, on this CPU-only container too (tiny workload so the test stays fast)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--chunks", "20000",
                          "--queries", "16", "--steps", "1", "--warmup", "1", "--ref-queries-per-step", "2"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "hybrid_topk_queries_per_sec" and d["unit"] == "queries/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"]


def _write_qdrant_like_dir(root, name, vectors, payloads, ids, pydantic_v1_state=False):
    """A local-mode Qdrant directory as qdrant-client >= 1.7 lays it out (restated, see indexing/qdrant_import.py):
    the points are pickled instances of a class living in `qdrant_client.http.models.models`."""
    import base64
    import json
    import pickle
    import sqlite3
    import sys
    import types
    mods = ["qdrant_client", "qdrant_client.http", "qdrant_client.http.models", "qdrant_client.http.models.models"]
    created = [m for m in mods if m not in sys.modules]
    for m in created:
        sys.modules[m] = types.ModuleType(m)
    mod = sys.modules[mods[-1]]
    try:
        if pydantic_v1_state:
            class PointStruct:                                   # pickles with pydantic v1's state layout
                def __init__(self, **kw):
                    self.__dict__.update(kw)

                def __getstate__(self):
                    return {"__dict__": dict(self.__dict__), "__fields_set__": set(self.__dict__),
                            "__private_attribute_values__": {}}
        else:
            import pydantic

            class PointStruct(pydantic.BaseModel):               # a real pydantic (v2) model
                id: int
                vector: list
                payload: dict
        PointStruct.__module__ = mods[-1]
        PointStruct.__qualname__ = "PointStruct"
        mod.PointStruct = PointStruct
        d = root / "collection" / name
        d.mkdir(parents=True)
        con = sqlite3.connect(str(d / "storage.sqlite"))
        con.execute("CREATE TABLE IF NOT EXISTS points (id TEXT PRIMARY KEY, point BLOB)")
        for i, v, p in zip(ids, vectors, payloads):
            key = base64.b64encode(pickle.dumps(i)).decode()
            con.execute("INSERT INTO points VALUES (?, ?)",
                        (key, pickle.dumps(PointStruct(id=i, vector=[float(x) for x in v], payload=p))))
        con.commit()
        con.close()
        (root / "meta.json").write_text(json.dumps(
            {"collections": {name: {"vectors": {"size": len(vectors[0]), "distance": "Cosine"}}}, "aliases": {}}))
    finally:
        for m in created:
            sys.modules.pop(m, None)


@pytest.mark.parametrize("v1_state", [False, True])
def test_qdrant_local_directory_reader(tmp_path, v1_state):
    """The importer must work WITHOUT qdrant-client (not installed here): stand-in classes take the pickled points'
    state, rows come back in point-id order whatever the insertion order was."""
    import sys
    from domain_rag_system_b200.indexing.qdrant_import import read_qdrant_local
    rng = np.random.default_rng(3)
    n, d = 23, 16
    vecs = rng.standard_normal((n, d)).astype(np.float32)
    payloads = [{"chunk_id": f"c{i}", "doc_id": f"d{i // 3}", "source": "s", "chunk_type": "t", "text": f"text {i}",
                 "metadata": {"k": i}} for i in range(n)]
    order = rng.permutation(n)                                   # upserts in arbitrary order
    _write_qdrant_like_dir(tmp_path, "medical_docs", vecs[order], [payloads[i] for i in order],
                           [int(i) for i in order], pydantic_v1_state=v1_state)
    assert "qdrant_client" not in sys.modules
    got, pl, cfg = read_qdrant_local(tmp_path, "medical_docs")
    assert got.dtype == np.float32 and np.array_equal(got, vecs)
    assert pl == payloads
    assert cfg["vectors"]["size"] == d
    with pytest.raises(FileNotFoundError):
        read_qdrant_local(tmp_path, "no_such_collection")


def test_qdrant_reader_refuses_foreign_classes(tmp_path):
    import pickle
    import sqlite3
    from domain_rag_system_b200.indexing.qdrant_import import read_qdrant_local
    d = tmp_path / "collection" / "c"
    d.mkdir(parents=True)
    con = sqlite3.connect(str(d / "storage.sqlite"))
    con.execute("CREATE TABLE points (id TEXT PRIMARY KEY, point BLOB)")
    con.execute("INSERT INTO points VALUES (?, ?)", ("k", pickle.dumps(os.system)))   # a callable from `os`/`posix`
    con.commit()
    con.close()
    with pytest.raises(pickle.UnpicklingError):
        read_qdrant_local(tmp_path, "c")


def test_built_objects_contain_blackwell_native_instructions():
    """The Blackwell-native claim, regression-proof: the built tensor kernels must contain tcgen05 MMAs
    (UTC*MMA in SASS, the pair kernel in its .2CTA form), TMA tensor loads (UTMALDG) and tensor-memory loads
    (LDTM) -- and no legacy HMMA tensor path.  Runs where nvcc's cuobjdump is (no GPU needed)."""
    import shutil
    import subprocess
    from domain_rag_system_b200 import build as _build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    _build.build()
    want = {"dense_pair.o": ("UTCHMMA.2CTA", "UTMALDG.2D.2CTA", "LDTM", "UTCBAR.2CTA.MULTICAST"),
            "dense_umma.o": ("UTCHMMA", "UTMALDG.2D", "LDTM", "STTM"),
            "dense_simt.o": ("UBLKCP.S.G", "SYNCS.ARRIVE.TRANS64", "SYNCS.PHASECHK.TRANS64.TRYWAIT")}   # bulk-copy ring
    for obj, needles in want.items():
        sass = subprocess.run([cuobjdump, "-sass", os.path.join(_build.CSRC, obj)], capture_output=True, text=True).stdout
        for n in needles:
            assert n in sass, f"{obj}: no {n} in SASS"
        assert " HMMA" not in sass and "HGMMA" not in sass, f"{obj}: legacy tensor-core path present"


# ---- device tokeniser (csrc/tokenize.cu): CPU models of what the kernel does ---------------------------------------
def _ascii_tokens_model(text: str, keep1=()):
    """Byte-level restatement of tokenize_kernel: maximal runs of [A-Za-z0-9_], lower-cased, one-character runs kept
    only when listed."""
    out, cur = [], []
    for ch in text + " ":
        if ch.isascii() and (ch.isalnum() or ch == "_"):
            cur.append(ch.lower())
        elif cur:
            tok = "".join(cur)
            if len(tok) >= 2 or tok in keep1:
                out.append(tok)
            cur = []
    return out


def test_ascii_tokeniser_model_equals_the_reference_style_tokeniser():
    """For ASCII text the reference tokeniser (lower(), \\b\\w+\\b, one-character tokens only when they are abbreviation
    keys; reproduced by BM25Indexer.tokenize, which the golden test pins) is byte-level: the model the CUDA kernel
    implements yields the same tokens."""
    from domain_rag_system_b200.indexing.bm25_indexer import BM25Indexer

    class Norm:
        abbreviations = {"K": "potassium", "MI": "myocardial infarction", "x": "times", "IV": "intravenous"}

    bi = BM25Indexer.__new__(BM25Indexer)         # tokenize() only reads .normalizer
    bi.normalizer = Norm()
    keep1 = {a.lower() for a in Norm.abbreviations if len(a.lower()) == 1}
    rng = np.random.default_rng(11)
    alphabet = list("abcXYZ019_ -.,;:/()[]'\"\t\n!?%+k")
    texts = ["", " ", "a", "K", "k+ 5 mmol/L", "Patient_1 has MI; IV fluids x 2 (BP 120/80).", "__init__ __", "a_b-c_d",
             "3.5mg/kg  q8h", "x" * 300, "END."]
    texts += ["".join(rng.choice(alphabet, size=int(rng.integers(0, 60)))) for _ in range(300)]
    for t in texts:
        assert _ascii_tokens_model(t, keep1) == bi.tokenize(t), repr(t)


def test_vocab_table_layout_finds_every_word_and_only_those():
    """Bm25DeviceIndex.vocab_table_arrays: replay the kernel's probe (home slot forward to the first empty slot, hash
    then bytes) on the host arrays."""
    rng = np.random.default_rng(5)
    words = [None if i % 17 == 3 else "w%05d" % i for i in range(3000)]
    words += ["".join(rng.choice(list("abcdefgh"), size=int(rng.integers(1, 9)))) + "_%d" % i for i in range(500)]
    words += ["naïve", "µg"]                                    # non-ASCII vocabulary entries: present, never matched by ASCII
    sh, st, offs, pool, cap = Bm25DeviceIndex.vocab_table_arrays(words)
    assert cap & (cap - 1) == 0 and sh.size == st.size and sh.size >= cap + sum(w is not None for w in words)
    mask64 = (1 << 64) - 1

    def fnv(b):
        h = 0xCBF29CE484222325
        for c in b:
            h = ((h ^ c) * 0x100000001B3) & mask64
        return h

    def lookup(tok: bytes):
        h = fnv(tok)
        s = h & (cap - 1)
        while s < sh.size:
            t = int(st[s])
            if t < 0:
                return -1
            if int(sh[s]) == h and bytes(pool[offs[t]:offs[t + 1]]) == tok:
                return t
            s += 1
        return -1

    for i, w in enumerate(words):
        if w is not None:
            assert lookup(w.encode("utf-8")) == i
    for miss in (b"", b"w99999", b"zzz", b"W00001", b"w0000"):
        assert lookup(miss) == -1
    # an empty vocabulary is a table that answers -1
    sh0, st0, _, _, cap0 = Bm25DeviceIndex.vocab_table_arrays([])
    assert (st0 == -1).all() and cap0 >= 16
