"""Times the selection kernel of drs_dense_topk alone for both tensor-core variants and checks that they return
the same ids:   [N=rows] [Q=queries] [KP=k'] [VARIANTS=ts,pair] [DRS_PAIR_DEBUG=bits] [DRS_PAIR_STAGES=n]
[DRS_PAIR_GROUPS=1] python profiles/pair_timing.py
DRS_PAIR_DEBUG bits (timing experiments; results are garbage): 1 no heap inserts, 2 no epilogue work."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from domain_rag_system_b200 import _lib, synth
dev = torch.device("cuda:0")
n = int(os.environ.get("N", "1000000")); q = int(os.environ.get("Q", "1000")); kp = int(os.environ.get("KP", "10"))
variants = os.environ.get("VARIANTS", "ts,pair").split(",")
emb = synth.unit_embeddings(n, 768, seed=1234, device=dev)
inv = _lib.row_inv_norms(emb)
qv = synth.query_embeddings(q, None, 768, seed=4321).to(dev)
code = {"ts": _lib.DENSE_UMMA_TS, "pair": _lib.DENSE_UMMA_PAIR, "simt": _lib.DENSE_SIMT}
res = {}
for v in variants:
    for _ in range(3): out = _lib.dense_topk(qv, emb, inv, kp, kernel=code[v], escalate=False)
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    for _ in range(10): out = _lib.dense_topk(qv, emb, inv, kp, kernel=code[v], escalate=False)
    torch.cuda.synchronize()
    pr = _lib.profile_read(); _lib.profile_enable(False)
    ms = pr["dense_select"][0] / pr["dense_select"][1]
    tail = pr["dense_tail"][0] / max(pr["dense_tail"][1], 1)
    res[v] = out
    print(f"variant {v} N {n} Q {q} KP {kp} debug {os.environ.get('DRS_PAIR_DEBUG')} stages {os.environ.get('DRS_PAIR_STAGES')} "
          f"groups {os.environ.get('DRS_PAIR_GROUPS')}: dense_select {ms:.4f} ms/launch = {2.0*q*n*768/ms/1e9:.1f} TFLOP/s, "
          f"tail {tail:.4f} ms, unproven {int((out[2] == 1).sum())}", flush=True)
if len(res) == 2 and not os.environ.get("DRS_PAIR_DEBUG"):
    a, b = list(res.values())
    print("ids equal:", bool(torch.equal(a[1], b[1])), "scores equal:", bool(torch.equal(a[0], b[0])))
