"""CPU model of the tensor path's exactness proof (csrc/select.cu: dense_rescore_kernel, DESIGN.md "Exactness of
the tensor-core path").  The GEMM ranks rows by  a(row) = bf16(q) . row  (fp32 accumulate) * inv_norm[row];  the
proof assumes  |a(row) / ||q|| - cos(q, row)| <= ||q - bf16(q)|| / ||q|| + 2e-5  (Cauchy-Schwarz: the row scaled by
its inverse norm is a unit vector).  Here the approximate scores are recomputed with torch on the CPU and the
inequality -- and what the proof concludes from it -- is checked against the fp64 oracle."""
import numpy as np
import pytest
import torch

from oracle import dense as odense

APPROX_ERR = 2e-5      # api.cu: launch_dense_rescore(..., 2e-5f, ...) for the tensor kernel


@pytest.mark.parametrize("kind", ["unit", "wide_dynamic_range", "tiny_and_huge_rows", "near_duplicates"])
def test_bf16_query_error_bound_and_proof(kind):
    rng = np.random.default_rng(7)
    n, d, kprime = 4000, 768, 10
    kk = kprime + 24 + kprime // 4
    rows = rng.standard_normal((n, d)).astype(np.float32)
    if kind == "tiny_and_huge_rows":
        rows *= (10.0 ** rng.uniform(-3, 3, size=(n, 1))).astype(np.float32)
    if kind == "near_duplicates":
        rows[1:400] = rows[0] + 1e-3 * rng.standard_normal((399, d)).astype(np.float32)
    rows_bf = torch.from_numpy(rows).to(torch.bfloat16)
    stored = rows_bf.float().numpy()                                   # what the index holds
    inv_norm = (1.0 / np.maximum(np.linalg.norm(stored.astype(np.float64), axis=1), 1e-30)).astype(np.float32)
    normed = odense.normalize_rows(stored)
    for qi in range(6):
        q = rng.standard_normal(d).astype(np.float32)
        if kind == "wide_dynamic_range":
            q *= (10.0 ** rng.uniform(-4, 2, size=d)).astype(np.float32)
        if kind == "near_duplicates":
            q = stored[0] + 0.05 * rng.standard_normal(d).astype(np.float32)
        q_bf = torch.from_numpy(q).to(torch.bfloat16).float().numpy()
        approx = (stored @ q_bf) * inv_norm                            # fp32, like the accumulators
        exact = odense.cosine_scores(normed, q)                        # fp64 oracle (normalised query)
        qn = np.linalg.norm(q.astype(np.float64))
        bound = np.linalg.norm(q.astype(np.float64) - q_bf.astype(np.float64)) / qn * 1.0001 + APPROX_ERR
        assert np.abs(approx.astype(np.float64) / qn - exact).max() <= bound
        # the proof: rows outside the top-kk by approximate score cannot reach the exact k'-th candidate score
        order = np.lexsort((np.arange(n), approx))[::-1][:kk]
        cand = np.sort(exact[order])[::-1]
        tau = cand[kprime - 1]
        outside_best = approx[order[-1]] / qn + bound
        want = np.argsort(-exact, kind="stable")[:kprime]
        if outside_best < tau:                                          # flag 0
            assert set(want.tolist()) <= set(order.tolist())
