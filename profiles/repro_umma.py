"""Small reproducer: one query, k' = 100 (single epilogue group, 16 KB ring stages) through the tensor kernel."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from domain_rag_system_b200 import _lib, synth
dev = torch.device("cuda:0")
n = int(os.environ.get("N", "300000")); kp = int(os.environ.get("KP", "100")); q = int(os.environ.get("Q", "1"))
rows = synth.unit_embeddings(n, 768, seed=1234, device=dev)
inv = _lib.row_inv_norms(rows)
qv = synth.query_embeddings(q, None, 768, seed=4321).to(dev)
s1, i1, _ = _lib.dense_topk(qv, rows, inv, kp, kernel=_lib.DENSE_SIMT)
s2, i2, f2 = _lib.dense_topk(qv, rows, inv, kp, kernel=_lib.DENSE_UMMA)
torch.cuda.synchronize()
print("ISSUERS", os.environ.get("DRS_UMMA_ISSUERS"), "N", n, "KP", kp, "Q", q, "equal", bool(torch.equal(i1, i2) and torch.equal(s1, s2)), "flags", f2.tolist()[:4])
