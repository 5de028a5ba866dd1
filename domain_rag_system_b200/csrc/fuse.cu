// Score fusion + final top-K: one CTA per query, everything in shared memory, fp64 in the
// reference's operation order.  Restates HybridIndexer._reciprocal_rank_fusion and
// _weighted_fusion (reference src/indexing/hybrid_indexer.py:112-156, :158-218):
//   * entries are visited vector list first, then the BM25 list ("insertion order");
//   * entries sharing a fusion key accumulate, in that order, starting from the first one;
//   * the sort is stable: ties keep insertion order of the key's first occurrence.
#include "common.cuh"

namespace drs {

constexpr int FUSE_MAX = 2 * (DRS_MAX_KPRIME + 8 + DRS_MAX_KPRIME / 4) ;  // both lists at full length
constexpr int FUSE_THREADS = 128;

__global__ void __launch_bounds__(FUSE_THREADS) fuse_kernel(
    int mode, int kv, int kb, const void* __restrict__ v_scores_raw, int v_f64, const int64_t* __restrict__ v_ids,
    const double* __restrict__ b_scores, const int64_t* __restrict__ b_ids,
    const int64_t* __restrict__ fusion_key, double rrf_k, double w_v, double w_b, int top_k,
    int64_t* __restrict__ out_ids, double* __restrict__ out_scores, double* __restrict__ out_norm) {
  __shared__ int64_t s_key[FUSE_MAX];
  __shared__ int64_t s_row[FUSE_MAX];
  __shared__ double s_contrib[FUSE_MAX];
  __shared__ double s_normed[FUSE_MAX];   // normalised score of the entry (weighted), NaN if unset
  __shared__ double s_total[FUSE_MAX];
  __shared__ int s_first[FUSE_MAX];       // 1 if this entry is the first occurrence of its key
  __shared__ double s_mm[4];              // vmin, vmax, bmin, bmax
  __shared__ int s_cnt[2];

  const int q = blockIdx.x, tid = threadIdx.x;
  const int n = kv + kb;
  auto vscore = [&](size_t at) -> double {
    return v_f64 ? reinterpret_cast<const double*>(v_scores_raw)[at]
                 : double(reinterpret_cast<const float*>(v_scores_raw)[at]);
  };

  // valid prefix lengths (lists are -1 padded at the tail)
  if (tid == 0) {
    int nv = 0, nb = 0;
    while (nv < kv && v_ids[size_t(q) * kv + nv] >= 0) ++nv;
    while (nb < kb && b_ids[size_t(q) * kb + nb] >= 0) ++nb;
    s_cnt[0] = nv;
    s_cnt[1] = nb;
    if (mode == 1) {
      // min / max exactly as Python's min()/max() over the list (values, so order is irrelevant)
      double lo = 0, hi = 0;
      for (int i = 0; i < nv; ++i) {
        const double s = vscore(size_t(q) * kv + i);
        if (i == 0 || s < lo) lo = s;
        if (i == 0 || s > hi) hi = s;
      }
      s_mm[0] = lo;
      s_mm[1] = hi;
      lo = hi = 0;
      for (int i = 0; i < nb; ++i) {
        const double s = b_scores[size_t(q) * kb + i];
        if (i == 0 || s < lo) lo = s;
        if (i == 0 || s > hi) hi = s;
      }
      s_mm[2] = lo;
      s_mm[3] = hi;
    }
  }
  __syncthreads();
  const int nv = s_cnt[0], nb = s_cnt[1];

  for (int i = tid; i < n; i += FUSE_THREADS) {
    const bool is_v = i < kv;
    const int r = is_v ? i : i - kv;
    const bool valid = is_v ? (r < nv) : (r < nb);
    int64_t row = -1, key = -1;
    double contrib = 0.0, normed = __longlong_as_double(0x7ff8000000000000ll);
    if (valid) {
      row = is_v ? v_ids[size_t(q) * kv + r] : b_ids[size_t(q) * kb + r];
      key = fusion_key ? fusion_key[row] : row;
      if (mode == 0) {
        contrib = __ddiv_rn(1.0, __dadd_rn(rrf_k, double(r + 1)));   // 1 / (k + rank), rank 1-based
      } else {
        const double s = is_v ? vscore(size_t(q) * kv + r) : b_scores[size_t(q) * kb + r];
        const double lo = is_v ? s_mm[0] : s_mm[2], hi = is_v ? s_mm[1] : s_mm[3];
        double used = s;                                              // raw score when max == min
        if (hi != lo) {
          used = __ddiv_rn(__dsub_rn(s, lo), __dsub_rn(hi, lo));
          normed = used;
        }
        contrib = __dmul_rn(is_v ? w_v : w_b, used);
      }
    }
    s_row[i] = row;
    s_key[i] = key;
    s_contrib[i] = contrib;
    s_normed[i] = normed;
  }
  __syncthreads();

  // first occurrence + sequential accumulation in insertion order
  for (int i = tid; i < n; i += FUSE_THREADS) {
    int first = 0;
    double total = 0.0;
    if (s_row[i] >= 0) {
      first = 1;
      const int64_t key = s_key[i];
      for (int j = 0; j < i; ++j)
        if (s_row[j] >= 0 && s_key[j] == key) {
          first = 0;
          break;
        }
      if (first) {
        total = s_contrib[i];  // Python: get(key, 0) + c  ==  c
        for (int j = i + 1; j < n; ++j)
          if (s_row[j] >= 0 && s_key[j] == key) total = __dadd_rn(total, s_contrib[j]);
      }
    }
    s_first[i] = first;
    s_total[i] = total;
  }
  __syncthreads();

  // stable descending rank among first occurrences
  for (int i = tid; i < n; i += FUSE_THREADS) {
    if (!s_first[i]) continue;
    const double t = s_total[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      if (!s_first[j] || j == i) continue;
      const double u = s_total[j];
      rank += (u > t) || (u == t && j < i);
    }
    if (rank < top_k) {
      // representative row: RRF keeps the first dict seen; weighted keeps the LAST vector
      // occurrence if the key appears in the vector list, else the first BM25 occurrence.
      int rep = i;
      if (mode == 1) {
        const int64_t key = s_key[i];
        for (int j = 0; j < kv; ++j)
          if (s_row[j] >= 0 && s_key[j] == key) rep = j;
      }
      out_ids[size_t(q) * top_k + rank] = s_row[rep];
      out_scores[size_t(q) * top_k + rank] = t;
      if (out_norm) out_norm[size_t(q) * top_k + rank] = s_normed[rep];
    }
  }
  // pad
  __shared__ int s_unique;
  if (tid == 0) s_unique = 0;
  __syncthreads();
  int mine = 0;
  for (int i = tid; i < n; i += FUSE_THREADS) mine += s_first[i];
  if (mine) atomicAdd(&s_unique, mine);
  __syncthreads();
  for (int i = s_unique + tid; i < top_k; i += FUSE_THREADS) {
    out_ids[size_t(q) * top_k + i] = -1;
    out_scores[size_t(q) * top_k + i] = -INFINITY;
    if (out_norm) out_norm[size_t(q) * top_k + i] = __longlong_as_double(0x7ff8000000000000ll);
  }
}

}  // namespace drs

extern "C" int drs_fuse_topk(int mode, int n_queries, int kv, int kb, const void* d_v_scores, int v_scores_f64,
                             const int64_t* d_v_ids, const double* d_b_scores, const int64_t* d_b_ids,
                             const int64_t* d_fusion_key, double rrf_k, double w_vector, double w_bm25,
                             int top_k, int64_t* d_out_ids, double* d_out_scores, double* d_out_norm,
                             void* stream) {
  using namespace drs;
  DRS_REQUIRE(mode == 0 || mode == 1, DRS_ERR_ARG, "fuse: mode must be 0 (rrf) or 1 (weighted)");
  DRS_REQUIRE(n_queries >= 0 && kv >= 0 && kb >= 0 && top_k > 0, DRS_ERR_ARG, "fuse: bad sizes");
  DRS_REQUIRE(kv + kb <= FUSE_MAX, DRS_ERR_UNSUPPORTED, "fuse: lists too long");
  DRS_REQUIRE((kv == 0 || (d_v_scores && d_v_ids)) && (kb == 0 || (d_b_scores && d_b_ids)) && d_out_ids &&
                  d_out_scores,
              DRS_ERR_ARG, "fuse: null pointer");
  if (n_queries == 0) return DRS_OK;
  ProfScope ps(DRS_PROF_FUSE, (cudaStream_t)stream);
  fuse_kernel<<<n_queries, FUSE_THREADS, 0, (cudaStream_t)stream>>>(
      mode, kv, kb, d_v_scores, v_scores_f64, d_v_ids, d_b_scores, d_b_ids, d_fusion_key, rrf_k, w_vector, w_bm25, top_k,
      d_out_ids, d_out_scores, d_out_norm);
  count_launches(1);
  DRS_CUDA_CHECK(cudaGetLastError());
  return DRS_OK;
}
