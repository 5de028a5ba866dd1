// Row-sharded retrieval: the kernels around the one exchange step.
//
// Every rank computes its local top-k' lists; ONE all-gather moves them (no row data crosses NVLink); then every rank
// merges the gathered lists IN PLACE (no unpacking copies), fuses, and -- for the "simple" reranker -- orders the
// fused pool by cosines that were computed where the rows live and travelled with the candidates:
//   drs_candidate_cosines    cos(rerank query, row) of every local candidate, the reranker's exact fp32 arithmetic
//   drs_merge_gathered       k-way merge straight out of the gathered buffer: strided parts, local -> global row ids,
//                            the cosine carried as payload
//   drs_rerank_from_cosines  BaseReranker.rerank's stable sort of the fused pool, cosines looked up in the merged lists
// The reference is single-process; the results equal drs_rerank_cosine / drs_merge_candidates_* on the unsharded
// corpus bit for bit (tests/dist_gpu_check.py, tests/test_gpu_parity.py).
#include "common.cuh"

namespace drs {

// the reranker's cosine, bit for bit (see rerank.cu): sequential fp32 sums, fp64 sqrt, fp32 divide
__device__ __forceinline__ float rerank_cos_seq(const __nv_bfloat16* __restrict__ row, const float* __restrict__ x, int dim) {
  const uint4* yp = reinterpret_cast<const uint4*>(row);
  float num = 0.0f, sxx = 0.0f, syy = 0.0f;
  for (int e = 0; e < dim; e += 8) {
    const uint4 w = __ldg(yp + (e >> 3));
    const float y[8] = {bf16_lo(w.x), bf16_hi(w.x), bf16_lo(w.y), bf16_hi(w.y),
                        bf16_lo(w.z), bf16_hi(w.z), bf16_lo(w.w), bf16_hi(w.w)};
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float xv = __ldg(x + e + t);
      num = __fadd_rn(num, __fmul_rn(xv, y[t]));
      sxx = __fadd_rn(sxx, __fmul_rn(xv, xv));
      syy = __fadd_rn(syy, __fmul_rn(y[t], y[t]));
    }
  }
  const double da = sqrt(double(sxx)), db = sqrt(double(syy));
  if (da == 0.0 || db == 0.0) return 0.0f;
  return __fdiv_rn(num, __double2float_rn(__dmul_rn(da, db)));
}

__global__ void __launch_bounds__(128) candidate_cos_kernel(const __nv_bfloat16* __restrict__ corpus, int dim,
                                                            const float* __restrict__ rq,
                                                            const int64_t* __restrict__ ids, int n, int64_t total,
                                                            float* __restrict__ out) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t row = ids[i];
  out[i] = row < 0 ? 0.0f : rerank_cos_seq(corpus + size_t(row) * dim, rq + size_t(i / n) * dim, dim);
}

// (score, global id) ordered by (score desc, id desc); src = where the candidate sits in the gathered buffer
struct Cand {
  uint64_t s;
  int64_t id;
  int32_t src;
  int32_t pad;
};
__device__ __forceinline__ bool cand_less(const Cand& a, const Cand& b) {
  return (a.s < b.s) || (a.s == b.s && a.id < b.id);
}

template <typename ScoreT>
__global__ void __launch_bounds__(256) merge_gathered_kernel(const uint8_t* __restrict__ base, int n_parts,
                                                             size_t part_stride, size_t off_scores, size_t off_ids,
                                                             size_t off_cos, int has_cos,
                                                             const int64_t* __restrict__ id_offset, int n_queries, int k,
                                                             int k_out, ScoreT* __restrict__ out_scores,
                                                             int64_t* __restrict__ out_ids, float* __restrict__ out_cos) {
  constexpr int BUF = 1024;
  __shared__ Cand buf[BUF];
  const int q = blockIdx.x;
  const int n = n_parts * k;
  for (int i = threadIdx.x; i < k_out; i += blockDim.x) buf[i] = Cand{0, -1, -1, 0};
  int pos = 0;
  while (pos < n) {
    const int take = min(BUF - k_out, n - pos);
    const int m = next_pow2(k_out + take);
    for (int i = threadIdx.x; i < m - k_out; i += blockDim.x) {
      Cand v{0, -1, -1, 0};
      if (i < take) {
        const int f = pos + i;
        const int p = f / k, c = f - p * k;
        const uint8_t* part = base + size_t(p) * part_stride;
        const size_t at = size_t(q) * k + c;
        const int64_t id = reinterpret_cast<const int64_t*>(part + off_ids)[at];
        if (id >= 0) {
          const ScoreT sc = reinterpret_cast<const ScoreT*>(part + off_scores)[at];
          if constexpr (sizeof(ScoreT) == 4) v.s = uint64_t(ord32(float(sc))) << 32;
          else v.s = ord64(double(sc));
          v.id = id + (id_offset ? id_offset[p] : 0);
          v.src = f;
        }
      }
      buf[k_out + i] = v;
    }
    __syncthreads();
    block_bitonic_desc(buf, m, cand_less);
    pos += take;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < k_out; i += blockDim.x) {
    const Cand v = buf[i];
    ScoreT s;
    float cs = 0.0f;
    if (v.id < 0) {
      s = ScoreT(-INFINITY);
    } else {
      if constexpr (sizeof(ScoreT) == 4) s = ScoreT(unord32(uint32_t(v.s >> 32)));
      else s = ScoreT(unord64(v.s));
      if (has_cos) {
        const int p = v.src / k, c = v.src - p * k;
        cs = reinterpret_cast<const float*>(base + size_t(p) * part_stride + off_cos)[size_t(q) * k + c];
      }
    }
    out_scores[size_t(q) * k_out + i] = s;
    out_ids[size_t(q) * k_out + i] = v.id;
    if (out_cos) out_cos[size_t(q) * k_out + i] = cs;
  }
}

__global__ void __launch_bounds__(1024) rerank_from_cos_kernel(const int64_t* __restrict__ pool, int top_n, int top_k,
                                                               const int64_t* __restrict__ v_ids,
                                                               const float* __restrict__ v_cos, int kv,
                                                               const int64_t* __restrict__ b_ids,
                                                               const float* __restrict__ b_cos, int kb,
                                                               float* __restrict__ out_scores,
                                                               int64_t* __restrict__ out_ids,
                                                               int32_t* __restrict__ out_pos) {
  __shared__ float s[DRS_MAX_POOL];
  __shared__ int s_valid;
  const int q = blockIdx.x;
  const int i = threadIdx.x;
  if (i == 0) {
    int nv = 0;
    while (nv < top_n && pool[size_t(q) * top_n + nv] >= 0) ++nv;
    s_valid = nv;
  }
  __syncthreads();
  const int nv = s_valid;
  int64_t row = -1;
  float c = -INFINITY;
  if (i < nv) {
    row = pool[size_t(q) * top_n + i];
    // a fused entry comes from the vector list, the BM25 list or both; the cosine is the same wherever it is found
    bool found = false;
    for (int j = 0; j < kv && !found; ++j)
      if (v_ids[size_t(q) * kv + j] == row) {
        c = v_cos[size_t(q) * kv + j];
        found = true;
      }
    for (int j = 0; j < kb && !found; ++j)
      if (b_ids[size_t(q) * kb + j] == row) {
        c = b_cos[size_t(q) * kb + j];
        found = true;
      }
    s[i] = c;
  }
  __syncthreads();
  if (i < nv) {
    int rank = 0;
    for (int j = 0; j < nv; ++j) rank += (s[j] > c) || (s[j] == c && j < i);   // stable, descending
    if (rank < top_k) {
      out_scores[size_t(q) * top_k + rank] = c;
      out_ids[size_t(q) * top_k + rank] = row;
      if (out_pos) out_pos[size_t(q) * top_k + rank] = i;
    }
  }
  for (int k = nv + i; k < top_k; k += blockDim.x) {
    out_scores[size_t(q) * top_k + k] = -INFINITY;
    out_ids[size_t(q) * top_k + k] = -1;
    if (out_pos) out_pos[size_t(q) * top_k + k] = -1;
  }
}

}  // namespace drs

using namespace drs;

extern "C" int drs_candidate_cosines(const void* d_corpus_bf16, int dim, const float* d_rq_f32, const int64_t* d_ids,
                                     int n_queries, int n, float* d_out_cos, void* stream) {
  DRS_REQUIRE(d_corpus_bf16 && d_rq_f32 && d_ids && d_out_cos, DRS_ERR_ARG, "candidate cosines: null pointer");
  DRS_REQUIRE(dim > 0 && dim % 8 == 0 && n > 0 && n_queries >= 0, DRS_ERR_UNSUPPORTED, "candidate cosines: bad sizes");
  if (n_queries == 0) return DRS_OK;
  const int64_t total = int64_t(n_queries) * n;
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope ps(DRS_PROF_RERANK, st);
  candidate_cos_kernel<<<unsigned((total + 127) / 128), 128, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(d_corpus_bf16),
                                                                      dim, d_rq_f32, d_ids, n, total, d_out_cos);
  count_launches(1);
  DRS_CUDA_CHECK(cudaGetLastError());
  return DRS_OK;
}

extern "C" int drs_merge_gathered(const void* d_gathered, int n_parts, size_t part_stride_bytes, size_t off_scores,
                                  int scores_f64, size_t off_ids, size_t off_cos, int has_cos,
                                  const int64_t* d_id_offset, int n_queries, int k, int k_out, void* d_out_scores,
                                  int64_t* d_out_ids, float* d_out_cos, void* stream) {
  DRS_REQUIRE(d_gathered && d_out_scores && d_out_ids, DRS_ERR_ARG, "merge: null pointer");
  DRS_REQUIRE(n_parts > 0 && n_queries >= 0 && k > 0 && k_out > 0, DRS_ERR_ARG, "merge: bad sizes");
  DRS_REQUIRE(k_out <= 512, DRS_ERR_UNSUPPORTED, "merge: k_out > 512");
  DRS_REQUIRE((off_scores % 8) == 0 && (off_ids % 8) == 0 && (part_stride_bytes % 8) == 0 && (!has_cos || off_cos % 4 == 0),
              DRS_ERR_ARG, "merge: misaligned offsets");
  if (n_queries == 0) return DRS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const uint8_t* base = reinterpret_cast<const uint8_t*>(d_gathered);
  if (scores_f64)
    merge_gathered_kernel<double><<<n_queries, 256, 0, st>>>(base, n_parts, part_stride_bytes, off_scores, off_ids, off_cos,
                                                             has_cos, d_id_offset, n_queries, k, k_out,
                                                             reinterpret_cast<double*>(d_out_scores), d_out_ids, d_out_cos);
  else
    merge_gathered_kernel<float><<<n_queries, 256, 0, st>>>(base, n_parts, part_stride_bytes, off_scores, off_ids, off_cos,
                                                            has_cos, d_id_offset, n_queries, k, k_out,
                                                            reinterpret_cast<float*>(d_out_scores), d_out_ids, d_out_cos);
  count_launches(1);
  DRS_CUDA_CHECK(cudaGetLastError());
  return DRS_OK;
}

extern "C" int drs_rerank_from_cosines(const int64_t* d_pool_ids, int n_queries, int top_n, int top_k,
                                       const int64_t* d_v_ids, const float* d_v_cos, int kv, const int64_t* d_b_ids,
                                       const float* d_b_cos, int kb, float* d_out_scores, int64_t* d_out_ids,
                                       int32_t* d_out_pos, void* stream) {
  DRS_REQUIRE(d_pool_ids && d_out_scores && d_out_ids, DRS_ERR_ARG, "rerank: null pointer");
  DRS_REQUIRE((kv == 0 || (d_v_ids && d_v_cos)) && (kb == 0 || (d_b_ids && d_b_cos)), DRS_ERR_ARG, "rerank: null list");
  DRS_REQUIRE(top_n > 0 && top_n <= DRS_MAX_POOL && top_k > 0 && n_queries >= 0, DRS_ERR_UNSUPPORTED,
              "rerank: pool size out of range");
  if (n_queries == 0) return DRS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int threads = ((top_n + 31) / 32) * 32;
  ProfScope ps(DRS_PROF_RERANK, st);
  rerank_from_cos_kernel<<<n_queries, threads, 0, st>>>(d_pool_ids, top_n, top_k, d_v_ids, d_v_cos, kv, d_b_ids, d_b_cos,
                                                        kb, d_out_scores, d_out_ids, d_out_pos);
  count_launches(1);
  DRS_CUDA_CHECK(cudaGetLastError());
  return DRS_OK;
}
