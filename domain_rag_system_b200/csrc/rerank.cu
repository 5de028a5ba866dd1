// "simple" cosine rerank: gather the pool rows, cosine against the raw-query embedding, stable
// sort.  Restates SimpleSimilarityReranker._cosine (reference src/reranker/rerank.py:53-58) as it
// evaluates under NumPy >= 2 on fp32 vectors, bit for bit:
//   num = sequential fp32 sum of fp32 products x*y      (Python sum() over np.float32 scalars)
//   da  = sqrt(double(sequential fp32 sum of x*x)),  db likewise          (math.sqrt -> fp64)
//   cos = fp32(num) / fp32(da*db)   (np.float32 / Python float -> fp32 divide);  0 if da or db == 0
// and BaseReranker.rerank's stable sort by rerank_score descending (:42), ties keep pool order.
// One thread per (query, pool entry) walks the row sequentially -- the sums are order-dependent,
// so they are not parallelised; the pool is at most DRS_MAX_POOL rows.
#include "common.cuh"

namespace drs {

__global__ void __launch_bounds__(1024) rerank_kernel(const __nv_bfloat16* __restrict__ corpus, int dim,
                                                      const float* __restrict__ rq,
                                                      const int64_t* __restrict__ pool, int top_n, int top_k,
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
    const float* x = rq + size_t(q) * dim;
    const uint4* yp = reinterpret_cast<const uint4*>(corpus + size_t(row) * dim);
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
    c = 0.0f;
    if (da != 0.0 && db != 0.0) c = __fdiv_rn(num, __double2float_rn(__dmul_rn(da, db)));
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

__global__ void row_inv_norm_kernel(const __nv_bfloat16* __restrict__ corpus, int64_t n_rows, int dim,
                                    float* __restrict__ inv_norm) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < n_rows; r += warps) {
    const uint4* rp = reinterpret_cast<const uint4*>(corpus + size_t(r) * dim);
    double acc = 0.0;
    for (int e = lane * 8; e < dim; e += 256) {
      const uint4 w = __ldg(rp + (e >> 3));
      const float y[8] = {bf16_lo(w.x), bf16_hi(w.x), bf16_lo(w.y), bf16_hi(w.y),
                          bf16_lo(w.z), bf16_hi(w.z), bf16_lo(w.w), bf16_hi(w.w)};
#pragma unroll
      for (int t = 0; t < 8; ++t) acc += double(y[t]) * double(y[t]);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) inv_norm[r] = acc > 0.0 ? float(1.0 / sqrt(acc)) : 0.0f;
  }
}

}  // namespace drs

extern "C" int drs_rerank_cosine(const void* d_corpus_bf16, int dim, const float* d_rq_f32,
                                 const int64_t* d_pool_ids, int n_queries, int top_n, int top_k,
                                 float* d_out_scores, int64_t* d_out_ids, int32_t* d_out_pos, void* stream) {
  using namespace drs;
  DRS_REQUIRE(d_corpus_bf16 && d_rq_f32 && d_pool_ids && d_out_scores && d_out_ids, DRS_ERR_ARG,
              "rerank: null pointer");
  DRS_REQUIRE(dim > 0 && dim % 8 == 0, DRS_ERR_UNSUPPORTED, "rerank: dim must be a multiple of 8");
  DRS_REQUIRE(top_n > 0 && top_n <= DRS_MAX_POOL && top_k > 0 && n_queries >= 0, DRS_ERR_UNSUPPORTED,
              "rerank: pool size out of range");
  if (n_queries == 0) return DRS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int threads = ((top_n + 31) / 32) * 32;
  ProfScope ps(DRS_PROF_RERANK, st);
  rerank_kernel<<<n_queries, threads, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(d_corpus_bf16), dim, d_rq_f32,
                                               d_pool_ids, top_n, top_k, d_out_scores, d_out_ids, d_out_pos);
  count_launches(1);
  cudaError_t e = cudaGetLastError();
  DRS_CUDA_CHECK(e);
  return DRS_OK;
}

extern "C" int drs_row_inv_norms(const void* d_corpus_bf16, int64_t n_rows, int dim, float* d_inv_norm,
                                 void* stream) {
  using namespace drs;
  DRS_REQUIRE(d_corpus_bf16 && d_inv_norm, DRS_ERR_ARG, "inv_norms: null pointer");
  DRS_REQUIRE(dim > 0 && dim % 8 == 0 && n_rows >= 0, DRS_ERR_UNSUPPORTED, "inv_norms: dim must be a multiple of 8");
  if (n_rows == 0) return DRS_OK;
  const int blocks = int(std::min<int64_t>((n_rows + 7) / 8, int64_t(sm_count()) * 16));
  row_inv_norm_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(d_corpus_bf16),
                                                               n_rows, dim, d_inv_norm);
  count_launches(1);
  DRS_CUDA_CHECK(cudaGetLastError());
  return DRS_OK;
}
