// Candidate-list kernels: k-way merge of partial top-k lists, exact re-scoring of dense
// candidates, and the C-ABI merge used after the multi-GPU all-gather.
#include <cstring>

#include "common.cuh"

namespace drs {

// ------------------------------------------------------------------------------------------
// merge_keys: CTA (q, g) merges parts [g*ppc, (g+1)*ppc) of query q.  Input: lists of kk u64 keys
// (unsorted, 0 = empty) laid out parts[(p * q_stride + q) * kk + c].  Output: the kk largest keys,
// descending, at out[(g * out_qstride + q) * kk ..].  Chunked block bitonic sort.  With many parts
// (single-query latency mode: hundreds of CTAs each hand in a list) the merge runs in two levels
// so that it is spread over many CTAs instead of one.
// ------------------------------------------------------------------------------------------
constexpr int MERGE_BUF = 2048;
constexpr int MERGE_FANIN = 16;       // parts per CTA in the first level of a two-level merge
constexpr int MERGE_TWO_LEVEL = 48;   // more parts than this => two levels

// Optional pre-filter (the tensor-core kernel's thresholds): tau32[q] and min over pool[s * pool_stride + q]
// (if every entry is set) are ordered-u32 LOWER bounds of query q's kk-th best score, so keys scoring below them
// cannot be among the kk best.  Usually a few dozen of the ~1300 keys survive, and a rank-by-counting pass
// (keys are distinct: rank = number of larger keys) replaces the 2048-key sorting network.
__global__ void __launch_bounds__(256) merge_keys_kernel(const uint64_t* __restrict__ parts, int n_parts,
                                                         int q_stride, int kk, uint64_t* __restrict__ out,
                                                         int out_qstride, const int32_t* __restrict__ active,
                                                         int ppc, const uint32_t* __restrict__ tau32,
                                                         const uint32_t* __restrict__ pool, int n_pool,
                                                         int pool_stride) {
  __shared__ uint64_t buf[MERGE_BUF];
  __shared__ unsigned s_bound;
  __shared__ int s_n;
  const int q = blockIdx.x;
  if (active != nullptr && active[q] == 0) return;
  const int p0 = blockIdx.y * ppc;
  const int np = min(ppc, n_parts - p0);
  const int n = np * kk;
  // Many lists and no published threshold (single-query latency mode: every CTA of the selection kernel hands in
  // a list): the kk-th largest of the lists' MAXIMA is a lower bound of the kk-th best key (kk different lists
  // each hold a key at least that large), so everything below it is dropped before any sorting.
  if (tau32 == nullptr && np >= kk && np > MERGE_TWO_LEVEL && np <= MERGE_BUF && kk <= 256) {
    __shared__ uint64_t s_kbound;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int p = warp; p < np; p += nw) {
      uint64_t mx = 0;
      for (int c = lane; c < kk; c += 32) mx = max(mx, parts[(size_t(p0 + p) * q_stride + q) * kk + c]);
      for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if (lane == 0) buf[p] = mx;
    }
    const int mp = next_pow2(np);
    for (int i = np + threadIdx.x; i < mp; i += blockDim.x) buf[i] = 0;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    block_bitonic_desc(buf, mp, [](uint64_t a, uint64_t b) { return a < b; });
    if (threadIdx.x == 0) s_kbound = buf[kk - 1];
    __syncthreads();
    const uint64_t kb = s_kbound;
    if (kb != 0) {
      for (int f = threadIdx.x; f < n; f += blockDim.x) {
        const int p = f / kk, c = f - p * kk;
        const uint64_t v = parts[(size_t(p0 + p) * q_stride + q) * kk + c];
        if (v >= kb) {
          const int pos = atomicAdd(&s_n, 1);
          if (pos < MERGE_BUF) buf[pos] = v;
        }
      }
      __syncthreads();
      const int ns = s_n;
      if (ns <= MERGE_BUF) {
        const int m = next_pow2(max(ns, kk));
        for (int i = ns + threadIdx.x; i < m; i += blockDim.x) buf[i] = 0;
        __syncthreads();
        block_bitonic_desc(buf, m, [](uint64_t a, uint64_t b) { return a < b; });
        for (int i = threadIdx.x; i < kk; i += blockDim.x) out[(size_t(blockIdx.y) * out_qstride + q) * kk + i] = buf[i];
        return;
      }
    }
    __syncthreads();   // fewer than kk non-empty lists, or too many survivors: the general path below
  }
  if (tau32 != nullptr && kk <= 256) {
    if (threadIdx.x == 0) {
      s_bound = 0xffffffffu;
      s_n = 0;
    }
    __syncthreads();
    if (pool != nullptr && n_pool > 0) {
      unsigned mn = 0xffffffffu;
      for (int i = threadIdx.x; i < n_pool; i += blockDim.x) mn = min(mn, pool[size_t(i) * pool_stride + q]);
      for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      if ((threadIdx.x & 31) == 0) atomicMin(&s_bound, mn);
    }
    __syncthreads();
    unsigned bound = tau32[q];
    if (pool != nullptr && n_pool > 0 && s_bound != 0u) bound = max(bound, s_bound);   // 0: a slab never reported
    for (int f = threadIdx.x; f < n; f += blockDim.x) {
      const int p = f / kk, c = f - p * kk;
      const uint64_t v = parts[(size_t(p0 + p) * q_stride + q) * kk + c];
      if (v != 0 && unsigned(v >> 32) >= bound) {
        const int pos = atomicAdd(&s_n, 1);
        if (pos < MERGE_BUF) buf[pos] = v;
      }
    }
    __syncthreads();
    const int ns = s_n;
    if (ns <= MERGE_BUF) {
      if (ns <= 256) {
        const uint64_t key = int(threadIdx.x) < ns ? buf[threadIdx.x] : 0;
        int rank = 0;
        if (key != 0)
          for (int j = 0; j < ns; ++j) rank += buf[j] > key;
        for (int i = ns + threadIdx.x; i < kk; i += blockDim.x) out[(size_t(blockIdx.y) * out_qstride + q) * kk + i] = 0;
        if (key != 0 && rank < kk) out[(size_t(blockIdx.y) * out_qstride + q) * kk + rank] = key;
      } else {
        const int m = next_pow2(ns);
        for (int i = ns + threadIdx.x; i < m; i += blockDim.x) buf[i] = 0;
        __syncthreads();
        block_bitonic_desc(buf, m, [](uint64_t a, uint64_t b) { return a < b; });
        for (int i = threadIdx.x; i < kk; i += blockDim.x) out[(size_t(blockIdx.y) * out_qstride + q) * kk + i] = buf[i];
      }
      return;
    }
    __syncthreads();   // more survivors than the buffer holds (thresholds unset): the general path below
  }
  for (int i = threadIdx.x; i < kk; i += blockDim.x) buf[i] = 0;
  int pos = 0;
  while (pos < n) {
    const int take = min(MERGE_BUF - kk, n - pos);
    const int m = next_pow2(kk + take);
    for (int i = threadIdx.x; i < m - kk; i += blockDim.x) {
      uint64_t v = 0;
      if (i < take) {
        const int f = pos + i;
        const int p = f / kk, c = f - p * kk;
        v = parts[(size_t(p0 + p) * q_stride + q) * kk + c];
      }
      buf[kk + i] = v;
    }
    __syncthreads();
    block_bitonic_desc(buf, m, [](uint64_t a, uint64_t b) { return a < b; });
    pos += take;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kk; i += blockDim.x)
    out[(size_t(blockIdx.y) * out_qstride + q) * kk + i] = buf[i];
}

// ------------------------------------------------------------------------------------------
// merge_appended: the CTA-pair tensor kernel APPENDS every row whose approximate score reached the query's shared
// threshold at the time to one buffer per (slab, query): parts[(slab * q_stride + q) * cap + i], i < counts[..].
// The final thresholds (min over the pool parts, pool[part * q_stride + q]) are the tightest of all: keys below
// them are dropped, the kk largest survivors come out descending.  A buffer that overflowed (count > cap: keys
// were lost) or more survivors than the merge buffer holds set overflow[q]: the query is then treated as
// unproven and re-run through the exact kernel.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) merge_appended_kernel(const uint64_t* __restrict__ parts,
                                                             const uint32_t* __restrict__ counts, int n_slabs,
                                                             int q_stride, int cap, int kk, uint64_t* __restrict__ out,
                                                             const uint32_t* __restrict__ pool, int n_pool,
                                                             int32_t* __restrict__ overflow) {
  __shared__ uint64_t buf[MERGE_BUF];
  __shared__ unsigned s_off[160];
  __shared__ unsigned s_bound;
  __shared__ int s_n, s_ovf, s_dead;
  const int q = blockIdx.x;
  if (threadIdx.x == 0) {
    s_bound = 0xffffffffu;
    s_n = 0;
    s_dead = 0;
    unsigned acc = 0;
    int ovf = 0;
    for (int s = 0; s < n_slabs; ++s) {
      unsigned c = counts[size_t(s) * q_stride + q];
      if (c > unsigned(cap)) {
        ovf = 1;
        c = unsigned(cap);
      }
      s_off[s] = acc;
      acc += c;
    }
    s_off[n_slabs] = acc;
    s_ovf = ovf;
  }
  __syncthreads();
  {   // the kernel's rule: the kk-th largest maximum among the parts that hold a valid row (-infinity = none, 0 = never
      // reported) bounds the kk-th best score from below; 0 if fewer than kk such parts exist
    unsigned cur = 0;
    for (int bit = 31; bit >= 0; --bit) {
      const unsigned x = cur | (1u << bit);
      int c = 0;
      for (int i = threadIdx.x; i < n_pool; i += blockDim.x) {
        const unsigned v = pool[size_t(i) * q_stride + q];
        c += (v != 0x007fffffu) && (v >= x);
      }
      // block-wide sum of c
      c = __reduce_add_sync(0xffffffffu, c);
      if (threadIdx.x == 0) s_dead = 0;
      __syncthreads();
      if ((threadIdx.x & 31) == 0) atomicAdd(&s_dead, c);
      __syncthreads();
      if (s_dead >= kk) cur = x;
      __syncthreads();
    }
    if (threadIdx.x == 0) s_bound = cur > 0x007fffffu ? cur : 0u;
  }
  __syncthreads();
  const unsigned bound = s_bound;   // 0 keeps everything
  const unsigned total = s_off[n_slabs];
  uint64_t* dst = out + size_t(q) * kk;
  // Rounds: the kk best so far sit in buf[0, carry); a round reads at most MERGE_BUF - kk appended keys, keeps the
  // ones the bound lets through behind them, and sorts.  One round unless the thresholds never tightened.
  const unsigned batch = unsigned(MERGE_BUF - kk);
  int carry = 0;
  for (unsigned base = 0; base == 0 || base < total; base += batch) {
    const unsigned end = min(total, base + batch);
    for (unsigned f = base + threadIdx.x; f < end; f += blockDim.x) {
      int lo = 0, hi = n_slabs;   // slab of flat index f: the last s with s_off[s] <= f
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_off[mid] <= f) lo = mid;
        else hi = mid;
      }
      const uint64_t v = parts[(size_t(lo) * q_stride + q) * cap + (f - s_off[lo])];
      if (unsigned(v >> 32) >= bound) buf[atomicAdd(&s_n, 1)] = v;
    }
    __syncthreads();
    const int ns = s_n;
    const bool last = end >= total;
    if (last && carry == 0 && ns <= 256) {   // the common case: rank by counting (keys are distinct)
      const uint64_t key = int(threadIdx.x) < ns ? buf[threadIdx.x] : 0;
      int rank = 0;
      if (key != 0)
        for (int j = 0; j < ns; ++j) rank += buf[j] > key;
      for (int i = ns + threadIdx.x; i < kk; i += blockDim.x) dst[i] = 0;
      if (key != 0 && rank < kk) dst[rank] = key;
    } else {
      const int m = next_pow2(max(ns, 1));
      for (int i = ns + threadIdx.x; i < m; i += blockDim.x) buf[i] = 0;
      __syncthreads();
      block_bitonic_desc(buf, m, [](uint64_t a, uint64_t b) { return a < b; });
      if (last) {
        for (int i = threadIdx.x; i < kk; i += blockDim.x) dst[i] = i < ns ? buf[i] : 0;
      } else {
        carry = min(ns, kk);
        __syncthreads();
        if (threadIdx.x == 0) s_n = carry;
        __syncthreads();
      }
    }
  }
  if (threadIdx.x == 0) overflow[q] = s_ovf;
}

int launch_merge_appended(const uint64_t* d_parts, const uint32_t* d_counts, int n_slabs, int q_stride, int cap,
                          int n_queries, int kk, uint64_t* d_out, const uint32_t* d_pool, int n_pool,
                          int32_t* d_overflow, cudaStream_t st) {
  DRS_REQUIRE(kk > 0 && kk <= MERGE_BUF / 2 && n_slabs > 0 && n_slabs < 160, DRS_ERR_UNSUPPORTED, "merge: sizes out of range");
  merge_appended_kernel<<<n_queries, 256, 0, st>>>(d_parts, d_counts, n_slabs, q_stride, cap, kk, d_out, d_pool, n_pool,
                                                   d_overflow);
  count_launches(1);
  DRS_CUDA_CHECK(cudaGetLastError());
  return DRS_OK;
}

size_t merge_scratch_keys(int n_parts, int n_queries, int kk) {
  if (n_parts <= MERGE_TWO_LEVEL) return 0;
  return size_t((n_parts + MERGE_FANIN - 1) / MERGE_FANIN) * n_queries * kk;
}

int launch_merge_keys_filtered(const uint64_t* d_parts, int n_parts, int q_stride, int n_queries, int kk,
                               uint64_t* d_out, const int32_t* d_active, uint64_t* d_scratch, const uint32_t* d_tau32,
                               const uint32_t* d_pool, int n_pool, int pool_stride, cudaStream_t st) {
  DRS_REQUIRE(kk > 0 && kk <= MERGE_BUF / 2, DRS_ERR_UNSUPPORTED, "merge: list length out of range");
  const bool bounded_single = d_tau32 == nullptr && n_parts >= kk && n_parts <= MERGE_BUF && kk <= 256;
  if (n_parts > MERGE_TWO_LEVEL && d_scratch != nullptr && !bounded_single) {
    const int g1 = (n_parts + MERGE_FANIN - 1) / MERGE_FANIN;
    merge_keys_kernel<<<dim3(n_queries, g1), 256, 0, st>>>(d_parts, n_parts, q_stride, kk, d_scratch, n_queries,
                                                            d_active, MERGE_FANIN, d_tau32, d_pool, n_pool, pool_stride);
    count_launches(1);
    merge_keys_kernel<<<dim3(n_queries, 1), 256, 0, st>>>(d_scratch, g1, n_queries, kk, d_out, n_queries, d_active,
                                                           g1, nullptr, nullptr, 0, 0);
    count_launches(1);
  } else {
    merge_keys_kernel<<<dim3(n_queries, 1), 256, 0, st>>>(d_parts, n_parts, q_stride, kk, d_out, n_queries, d_active,
                                                           n_parts, d_tau32, d_pool, n_pool, pool_stride);
    count_launches(1);
  }
  DRS_CUDA_CHECK(cudaGetLastError());
  return DRS_OK;
}

int launch_merge_keys(const uint64_t* d_parts, int n_parts, int q_stride, int n_queries, int kk,
                      uint64_t* d_out, const int32_t* d_active, uint64_t* d_scratch, cudaStream_t st) {
  return launch_merge_keys_filtered(d_parts, n_parts, q_stride, n_queries, kk, d_out, d_active, d_scratch, nullptr,
                                    nullptr, 0, 0, st);
}

// ------------------------------------------------------------------------------------------
// dense_rescore: exact cosine of the kk surviving candidates of each query.
// score = fp32( dot_fp64(q_fp32, row_bf16) * (1/||q||) * inv_norm[row] ), i.e. the reference's
// renormalise-then-dot (qdrant local mode) with the dot product carried in fp64.
// Then sort by (score desc, row desc) and emit the first k'.
// Flag: the selection stages ranked rows by an APPROXIMATE score whose error is bounded by
// approx_err (in normalised-score units).  Every row outside the candidate list has
// approximate score <= the list's smallest; if that plus the bound is still below the exact
// k'-th score, no outside row can belong to the top-k' => proven exact (flag 0), else flag 1.
// ------------------------------------------------------------------------------------------
constexpr int RESCORE_MAXC = 256;

__global__ void __launch_bounds__(256) dense_rescore_kernel(
    const float* __restrict__ queries, const __nv_bfloat16* __restrict__ corpus,
    const float* __restrict__ inv_norm, int dim, const uint64_t* __restrict__ cand, int kk, int kprime,
    float approx_err, float* __restrict__ out_scores, int64_t* __restrict__ out_ids,
    int32_t* __restrict__ flags, const int32_t* __restrict__ active, int query_mode,
    int flag_value_unproven, int flag_value_proven, const int32_t* __restrict__ force_unproven) {
  __shared__ float sq[DRS_MAX_DIM];
  __shared__ uint64_t ekey[RESCORE_MAXC];
  __shared__ double red[8];
  __shared__ double red2[8];
  __shared__ float s_amin;
  __shared__ int s_nvalid;
  const int q = blockIdx.x;
  if (active != nullptr && active[q] == 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // query -> smem; ||q||^2 in fp64; and the L2 norm of the error the selection stage made by
  // rounding the query (bf16 of the raw query)
  double part = 0.0;
  for (int i = threadIdx.x; i < dim; i += blockDim.x) {
    float v = queries[size_t(q) * dim + i];
    sq[i] = v;
    part += double(v) * double(v);
  }
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) red[warp] = part;
  if (threadIdx.x == 0) s_nvalid = 0;
  __syncthreads();
  double qn2 = 0.0;
  for (int w = 0; w < 8; ++w) qn2 += red[w];
  const double inv_q = qn2 > 0.0 ? 1.0 / sqrt(qn2) : 0.0;
  double perr = 0.0;
  if (query_mode != DRS_QMODE_EXACT) {
    for (int i = threadIdx.x; i < dim; i += blockDim.x) {
      const float v = sq[i];
      const double d = (double(v) - double(__bfloat162float(__float2bfloat16_rn(v)))) * inv_q;
      perr += d * d;
    }
    for (int o = 16; o > 0; o >>= 1) perr += __shfl_xor_sync(0xffffffffu, perr, o);
  }
  __syncthreads();
  if (lane == 0) red2[warp] = perr;
  __syncthreads();
  double qe2 = 0.0;
  for (int w = 0; w < 8; ++w) qe2 += red2[w];
  // Cauchy-Schwarz: |(q_used - q/||q||) . c| * inv_norm[c] <= ||q_used - q/||q|| ||   (||c||*inv_norm = 1)
  const float err_bound = approx_err + float(sqrt(qe2) * 1.0001);
  const double amin_scale = inv_q;  // approximate scores are not scaled by 1/||q||

  const int m = next_pow2(kk);
  for (int i = threadIdx.x; i < m; i += blockDim.x) ekey[i] = 0;
  __syncthreads();

  float amin = INFINITY;
  int nvalid = 0;
  for (int c = warp; c < kk; c += 8) {
    const uint64_t key = cand[size_t(q) * kk + c];
    if (key == 0) continue;
    const uint32_t row = key_row(key);
    const uint4* rp = reinterpret_cast<const uint4*>(corpus + size_t(row) * dim);
    double acc = 0.0;
    for (int e = lane * 8; e < dim; e += 256) {
      const uint4 w = __ldg(rp + (e >> 3));
      const float* qq = sq + e;
      acc += double(bf16_lo(w.x)) * double(qq[0]);
      acc += double(bf16_hi(w.x)) * double(qq[1]);
      acc += double(bf16_lo(w.y)) * double(qq[2]);
      acc += double(bf16_hi(w.y)) * double(qq[3]);
      acc += double(bf16_lo(w.z)) * double(qq[4]);
      acc += double(bf16_hi(w.z)) * double(qq[5]);
      acc += double(bf16_lo(w.w)) * double(qq[6]);
      acc += double(bf16_hi(w.w)) * double(qq[7]);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      const float s = float(acc * inv_q * double(inv_norm[row]));
      ekey[c] = make_key(s, row);
      amin = fminf(amin, key_score(key));
      ++nvalid;
    }
  }
  if (lane == 0) atomicAdd(&s_nvalid, nvalid);
  __shared__ float wmin[8];
  if (lane == 0) wmin[warp] = amin;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = INFINITY;
    for (int w = 0; w < 8; ++w) a = fminf(a, wmin[w]);
    s_amin = a;
  }
  __syncthreads();
  block_bitonic_desc(ekey, m, [](uint64_t a, uint64_t b) { return a < b; });

  for (int i = threadIdx.x; i < kprime; i += blockDim.x) {
    const uint64_t k = (i < m) ? ekey[i] : 0;
    out_scores[size_t(q) * kprime + i] = k ? key_score(k) : -INFINITY;
    out_ids[size_t(q) * kprime + i] = k ? int64_t(key_row(k)) : int64_t(-1);
  }
  if (threadIdx.x == 0 && flags != nullptr) {
    int f = flag_value_proven;
    if (s_nvalid >= kk && kprime <= kk) {  // list was full: rows outside it exist
      const float tau = key_score(ekey[kprime - 1]);
      const float outside_best = float(double(s_amin) * amin_scale) + err_bound;
      f = (outside_best < tau) ? 0 : flag_value_unproven;
    }
    // the selection stage lost candidates for this query (a candidate buffer overflowed): nothing is proven
    if (force_unproven != nullptr && force_unproven[q] != 0) f = flag_value_unproven;
    flags[q] = f;
  }
}

int launch_dense_rescore(const float* d_q, int n_queries, const void* d_corpus, const float* d_inv_norm,
                         int dim, const uint64_t* d_cand, int kk, int kprime, float approx_err,
                         float* d_out_scores, int64_t* d_out_ids, int32_t* d_flags,
                         const int32_t* d_active, int query_mode, int flag_unproven, int flag_proven,
                         cudaStream_t st, const int32_t* d_force_unproven) {
  DRS_REQUIRE(kk <= RESCORE_MAXC, DRS_ERR_UNSUPPORTED, "rescore: candidate list too long");
  dense_rescore_kernel<<<n_queries, 256, 0, st>>>(d_q, reinterpret_cast<const __nv_bfloat16*>(d_corpus),
                                                   d_inv_norm, dim, d_cand, kk, kprime, approx_err,
                                                   d_out_scores, d_out_ids, d_flags, d_active,
                                                   query_mode, flag_unproven, flag_proven, d_force_unproven);
  count_launches(1);
  DRS_CUDA_CHECK(cudaGetLastError());
  return DRS_OK;
}

// ------------------------------------------------------------------------------------------
// C-ABI merges (post all-gather).  (score, id) pairs, order (score desc, id desc).
// ------------------------------------------------------------------------------------------
// flags (optional): only queries with flags[q] == 1 are merged; their flag becomes 2 afterwards (the BM25
// escalation pass: "resolved exactly").
template <typename ScoreT>
__global__ void __launch_bounds__(256) merge_pairs_kernel(const ScoreT* __restrict__ scores,
                                                          const int64_t* __restrict__ ids, int n_parts,
                                                          int n_queries, int k, int k_out,
                                                          ScoreT* __restrict__ out_scores,
                                                          int64_t* __restrict__ out_ids, int32_t* __restrict__ flags) {
  constexpr int BUF = 1024;
  __shared__ Pair64 buf[BUF];
  const int q = blockIdx.x;
  if (flags != nullptr && flags[q] != 1) return;
  const int n = n_parts * k;
  for (int i = threadIdx.x; i < k_out; i += blockDim.x) buf[i] = Pair64{0, -1};
  int pos = 0;
  while (pos < n) {
    const int take = min(BUF - k_out, n - pos);
    const int m = next_pow2(k_out + take);
    for (int i = threadIdx.x; i < m - k_out; i += blockDim.x) {
      Pair64 v{0, -1};
      if (i < take) {
        const int f = pos + i;
        const int p = f / k, c = f - p * k;
        const size_t at = (size_t(p) * n_queries + q) * k + c;
        const int64_t id = ids[at];
        if (id >= 0) {
          if constexpr (sizeof(ScoreT) == 4) v.s = uint64_t(ord32(float(scores[at]))) << 32;
          else v.s = ord64(double(scores[at]));
          v.id = id;
        }
      }
      buf[k_out + i] = v;
    }
    __syncthreads();
    block_bitonic_desc(buf, m, pair_less);
    pos += take;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < k_out; i += blockDim.x) {
    const Pair64 v = buf[i];
    ScoreT s;
    if (v.id < 0) s = ScoreT(-INFINITY);
    else if constexpr (sizeof(ScoreT) == 4) s = ScoreT(unord32(uint32_t(v.s >> 32)));
    else s = ScoreT(unord64(v.s));
    out_scores[size_t(q) * k_out + i] = s;
    out_ids[size_t(q) * k_out + i] = v.id;
  }
  __syncthreads();
  if (flags != nullptr && threadIdx.x == 0) flags[q] = 2;
}

template <typename ScoreT>
static int merge_pairs(const ScoreT* d_scores, const int64_t* d_ids, int n_parts, int n_queries, int k,
                       int k_out, ScoreT* d_out_scores, int64_t* d_out_ids, void* stream, int32_t* d_flags = nullptr) {
  DRS_REQUIRE(d_scores && d_ids && d_out_scores && d_out_ids, DRS_ERR_ARG, "merge: null pointer");
  DRS_REQUIRE(n_parts > 0 && n_queries >= 0 && k > 0 && k_out > 0, DRS_ERR_ARG, "merge: bad sizes");
  DRS_REQUIRE(k_out <= 512, DRS_ERR_UNSUPPORTED, "merge: k_out > 512");
  if (n_queries == 0) return DRS_OK;
  merge_pairs_kernel<ScoreT><<<n_queries, 256, 0, (cudaStream_t)stream>>>(
      d_scores, d_ids, n_parts, n_queries, k, k_out, d_out_scores, d_out_ids, d_flags);
  count_launches(1);
  DRS_CUDA_CHECK(cudaGetLastError());
  return DRS_OK;
}

int launch_merge_pairs_f64(const double* d_scores, const int64_t* d_ids, int n_parts, int n_queries, int k, int k_out,
                           double* d_out_scores, int64_t* d_out_ids, int32_t* d_flags, cudaStream_t st) {
  return merge_pairs<double>(d_scores, d_ids, n_parts, n_queries, k, k_out, d_out_scores, d_out_ids, (void*)st, d_flags);
}

}  // namespace drs

extern "C" int drs_merge_candidates_f32(const float* d_scores, const int64_t* d_ids, int n_parts,
                                        int n_queries, int k, int k_out, float* d_out_scores,
                                        int64_t* d_out_ids, void* stream) {
  return drs::merge_pairs<float>(d_scores, d_ids, n_parts, n_queries, k, k_out, d_out_scores, d_out_ids,
                                 stream);
}
extern "C" int drs_merge_candidates_f64(const double* d_scores, const int64_t* d_ids, int n_parts,
                                        int n_queries, int k, int k_out, double* d_out_scores,
                                        int64_t* d_out_ids, void* stream) {
  return drs::merge_pairs<double>(d_scores, d_ids, n_parts, n_queries, k, k_out, d_out_scores, d_out_ids,
                                  stream);
}
