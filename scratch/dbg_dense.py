import os, sys, torch, time
sys.path.insert(0, os.getcwd())
from domain_rag_system_b200 import _lib, synth
dev = torch.device("cuda:0")
emb = synth.unit_embeddings(1_000_000, 768, seed=1234, device=dev)
inv = _lib.row_inv_norms(emb)
qv = synth.query_embeddings(1000, None, 768, seed=4321).to(dev)
for _ in range(3): _lib.dense_topk(qv, emb, inv, 10, kernel=_lib.DENSE_UMMA, escalate=False)
torch.cuda.synchronize()
_lib.profile_enable(True)
for _ in range(10): _lib.dense_topk(qv, emb, inv, 10, kernel=_lib.DENSE_UMMA, escalate=False)
torch.cuda.synchronize()
pr = _lib.profile_read(); _lib.profile_enable(False)
print("DEBUG", os.environ.get("DRS_UMMA_DEBUG"), "dense_select ms/launch", pr["dense_select"][0] / pr["dense_select"][1])
