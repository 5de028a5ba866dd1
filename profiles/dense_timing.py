"""Times drs_dense_topk's selection kernel alone (1 M x 768, 1000 queries) -- the driver of the timing experiments
in DESIGN.md 4.1:   DRS_UMMA_DEBUG=<bits> [DRS_UMMA_STAGES=n] [DRS_UMMA_PAIR=1] [KP=k'] python profiles/dense_timing.py
bits: 2 no heap inserts, 4 half of every TMA stage, 8 every other MMA, 16 no pass 1, 32 no accumulator read,
64 / 128 MMA N = 32 / 128, 256 alternate accumulators.  Results are garbage with any bit set; only the clock matters."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from domain_rag_system_b200 import _lib, synth
dev = torch.device("cuda:0")
kp = int(os.environ.get("KP", "10"))
emb = synth.unit_embeddings(1_000_000, 768, seed=1234, device=dev)
inv = _lib.row_inv_norms(emb)
qv = synth.query_embeddings(1000, None, 768, seed=4321).to(dev)
for _ in range(3): _lib.dense_topk(qv, emb, inv, kp, kernel=_lib.DENSE_UMMA, escalate=False)
torch.cuda.synchronize()
_lib.profile_enable(True)
for _ in range(10): _lib.dense_topk(qv, emb, inv, kp, kernel=_lib.DENSE_UMMA, escalate=False)
torch.cuda.synchronize()
pr = _lib.profile_read(); _lib.profile_enable(False)
print("DEBUG", os.environ.get("DRS_UMMA_DEBUG"), "STAGES", os.environ.get("DRS_UMMA_STAGES"), "KP", kp, "dense_select ms/launch", pr["dense_select"][0] / pr["dense_select"][1])
