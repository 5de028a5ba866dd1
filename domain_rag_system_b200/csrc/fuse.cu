// Score fusion + final top-K: one CTA per query, everything in shared memory, fp64 in the
// reference's operation order.  Restates HybridIndexer._reciprocal_rank_fusion and
// _weighted_fusion (reference src/indexing/hybrid_indexer.py:112-156, :158-218):
//   * entries are visited vector list first, then the BM25 list ("insertion order");
//   * entries sharing a fusion key accumulate, in that order, starting from the first one;
//   * the sort is stable: ties keep insertion order of the key's first occurrence.
// The O(n^2) pair work (duplicate detection, ranking) is laid out so that a WARP owns an entry and
// its LANES stride over the other entries: every shared-memory load is independent of the previous
// one (a per-thread serial scan of the list is a ~250-cycle dependent chain per step).
#include "common.cuh"

namespace drs {

constexpr int FUSE_MAX = 2 * (DRS_MAX_KPRIME + 8 + DRS_MAX_KPRIME / 4);  // both lists at full length
constexpr int FUSE_THREADS = 256;
constexpr int FUSE_WARPS = FUSE_THREADS / 32;

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
  __shared__ int s_dup[FUSE_MAX];         // 1 if a LATER entry shares this entry's key
  __shared__ int s_vrep[FUSE_MAX];        // weighted: index of the last vector entry with this key, else -1
  __shared__ double s_red[4][FUSE_WARPS]; // per-warp min / max partials
  __shared__ int s_cnt[2];
  __shared__ int s_unique;

  const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = kv + kb;
  const double qnan = __longlong_as_double(0x7ff8000000000000ll);
  auto vscore = [&](size_t at) -> double {
    return v_f64 ? reinterpret_cast<const double*>(v_scores_raw)[at]
                 : double(reinterpret_cast<const float*>(v_scores_raw)[at]);
  };

  // ---- stage both lists (coalesced); valid prefix lengths (lists are -1 padded at the tail)
  if (tid < 2) s_cnt[tid] = tid == 0 ? kv : kb;
  if (tid == 0) s_unique = 0;
  __syncthreads();
  for (int i = tid; i < n; i += FUSE_THREADS) {
    const bool is_v = i < kv;
    const int r = is_v ? i : i - kv;
    const int64_t row = is_v ? v_ids[size_t(q) * kv + r] : b_ids[size_t(q) * kb + r];
    s_row[i] = row;
    s_contrib[i] = is_v ? vscore(size_t(q) * kv + r) : b_scores[size_t(q) * kb + r];   // raw score for now
    if (row < 0) atomicMin(&s_cnt[is_v ? 0 : 1], r);
  }
  __syncthreads();
  const int nv = s_cnt[0], nb = s_cnt[1];

  // ---- weighted: min / max of each list (values only, so any reduction order gives Python's result)
  double vlo = 0, vhi = 0, blo = 0, bhi = 0;
  if (mode == 1) {
    double lo_v = INFINITY, hi_v = -INFINITY, lo_b = INFINITY, hi_b = -INFINITY;
    for (int i = tid; i < n; i += FUSE_THREADS) {
      const bool is_v = i < kv;
      const int r = is_v ? i : i - kv;
      if (is_v ? (r < nv) : (r < nb)) {
        const double s = s_contrib[i];
        if (is_v) { lo_v = fmin(lo_v, s); hi_v = fmax(hi_v, s); }
        else { lo_b = fmin(lo_b, s); hi_b = fmax(hi_b, s); }
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      lo_v = fmin(lo_v, __shfl_xor_sync(0xffffffffu, lo_v, o));
      hi_v = fmax(hi_v, __shfl_xor_sync(0xffffffffu, hi_v, o));
      lo_b = fmin(lo_b, __shfl_xor_sync(0xffffffffu, lo_b, o));
      hi_b = fmax(hi_b, __shfl_xor_sync(0xffffffffu, hi_b, o));
    }
    if (lane == 0) {
      s_red[0][warp] = lo_v;
      s_red[1][warp] = hi_v;
      s_red[2][warp] = lo_b;
      s_red[3][warp] = hi_b;
    }
    __syncthreads();
    vlo = s_red[0][0]; vhi = s_red[1][0]; blo = s_red[2][0]; bhi = s_red[3][0];
    for (int w = 1; w < FUSE_WARPS; ++w) {
      vlo = fmin(vlo, s_red[0][w]); vhi = fmax(vhi, s_red[1][w]);
      blo = fmin(blo, s_red[2][w]); bhi = fmax(bhi, s_red[3][w]);
    }
  }

  // ---- per-entry key and contribution
  for (int i = tid; i < n; i += FUSE_THREADS) {
    const bool is_v = i < kv;
    const int r = is_v ? i : i - kv;
    const bool valid = is_v ? (r < nv) : (r < nb);
    int64_t row = -1, key = -1;
    double contrib = 0.0, normed = qnan;
    if (valid) {
      row = s_row[i];
      key = fusion_key ? fusion_key[row] : row;
      if (mode == 0) {
        contrib = __ddiv_rn(1.0, __dadd_rn(rrf_k, double(r + 1)));   // 1 / (k + rank), rank 1-based
      } else {
        const double s = s_contrib[i];
        const double lo = is_v ? vlo : blo, hi = is_v ? vhi : bhi;
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
    s_first[i] = valid ? 1 : 0;
    s_dup[i] = 0;
    s_vrep[i] = -1;
  }
  __syncthreads();

  // ---- duplicate detection: warp owns entry i, lanes stride over the other entries
  for (int i = warp; i < n; i += FUSE_WARPS) {
    if (s_row[i] < 0) continue;
    const int64_t key = s_key[i];
    bool seen = false;
    int last_v = -1;
    for (int j = lane; j < n; j += 32) {
      if (j != i && s_row[j] >= 0 && s_key[j] == key) {
        if (j < i) {
          seen = true;
          s_dup[j] = 1;              // benign race: everybody writes 1
        }
        if (j < kv) last_v = max(last_v, j);
      }
    }
    if (i < kv) last_v = max(last_v, i);
    seen = __any_sync(0xffffffffu, seen);
    for (int o = 16; o > 0; o >>= 1) last_v = max(last_v, __shfl_xor_sync(0xffffffffu, last_v, o));
    if (lane == 0) {
      if (seen) s_first[i] = 0;
      s_vrep[i] = last_v;
    }
  }
  __syncthreads();

  // ---- totals: Python's `get(key, 0) + c` chain, sequential in insertion order (only keys that repeat)
  for (int i = tid; i < n; i += FUSE_THREADS) {
    double total = s_contrib[i];
    if (s_first[i] && s_dup[i]) {
      const int64_t key = s_key[i];
      for (int j = i + 1; j < n; ++j)
        if (s_row[j] >= 0 && s_key[j] == key) total = __dadd_rn(total, s_contrib[j]);
    }
    s_total[i] = total;
  }
  __syncthreads();

  // ---- stable descending rank among first occurrences: warp owns entry i, lanes count over j
  for (int i = warp; i < n; i += FUSE_WARPS) {
    if (!s_first[i]) continue;
    const double t = s_total[i];
    int cnt = 0;
    for (int j = lane; j < n; j += 32) {
      if (j != i && s_first[j]) {
        const double u = s_total[j];
        cnt += (u > t) || (u == t && j < i);
      }
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) {
      atomicAdd(&s_unique, 1);
      if (cnt < top_k) {
        // representative row: RRF keeps the first dict seen; weighted keeps the LAST vector occurrence
        // of the key if it appears in the vector list, else the first BM25 occurrence
        const int rep = (mode == 1 && s_vrep[i] >= 0) ? s_vrep[i] : i;
        out_ids[size_t(q) * top_k + cnt] = s_row[rep];
        out_scores[size_t(q) * top_k + cnt] = t;
        if (out_norm) out_norm[size_t(q) * top_k + cnt] = s_normed[rep];
      }
    }
  }
  __syncthreads();
  for (int i = s_unique + tid; i < top_k; i += FUSE_THREADS) {
    out_ids[size_t(q) * top_k + i] = -1;
    out_scores[size_t(q) * top_k + i] = -INFINITY;
    if (out_norm) out_norm[size_t(q) * top_k + i] = qnan;
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
