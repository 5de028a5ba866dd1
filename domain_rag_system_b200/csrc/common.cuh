// Shared device/host helpers for the hybrid-retrieval kernels (sm_100a).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstring>
#include <string>

#include "../../include/drs_b200.h"

namespace drs {

// ---- error plumbing (thread-local message, C ABI returns codes) --------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define DRS_CUDA_CHECK(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess)                                                                \
      return ::drs::fail(DRS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

// Raise a kernel's dynamic shared-memory limit only when it grows: the attribute is per device and sticky, so the
// per-call cost (a driver call per launch in latency mode) is paid once per (call site, device, size).
#define DRS_ENSURE_SMEM(kern, bytes)                                                                  \
  do {                                                                                                \
    static std::atomic<int> drs_hw_[64];                                                              \
    int drs_dev_ = 0;                                                                                 \
    DRS_CUDA_CHECK(cudaGetDevice(&drs_dev_));                                                         \
    const int drs_b_ = int(bytes) + 1;                                                                \
    if (drs_b_ > drs_hw_[drs_dev_ & 63].load(std::memory_order_relaxed)) {                            \
      DRS_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, drs_b_ - 1)); \
      drs_hw_[drs_dev_ & 63].store(drs_b_, std::memory_order_relaxed);                                \
    }                                                                                                 \
  } while (0)

#define DRS_REQUIRE(cond, code, msg) \
  do {                               \
    if (!(cond)) return ::drs::fail((code), (msg)); \
  } while (0)

int sm_count();
void count_launches(int n);   // bumps the process-wide kernel-launch counter (drs_kernel_launches)

// measurement hooks (api.cu): no-ops unless drs_profile_enable(1) was called
void prof_begin(int kind, cudaStream_t st);
void prof_end(int kind, cudaStream_t st);
struct ProfScope {
  int kind;
  cudaStream_t st;
  ProfScope(int k, cudaStream_t s) : kind(k), st(s) { prof_begin(k, s); }
  ~ProfScope() { prof_end(kind, st); }
};

// ---- sortable keys -------------------------------------------------------------------
// A candidate is (score, row).  Order everywhere: score descending, then row descending
// (the determinised `np.argsort(scores)[::-1]`).  Packing the order-preserving image of
// the fp32 score above the row gives a u64 whose plain integer order is that order.
// Key 0 is "empty slot" (no finite or infinite score maps to 0).
__host__ __device__ __forceinline__ uint32_t ord32(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f + 0.0f);  // -0.0 -> +0.0 so the two zeros tie
#else
  float g = f + 0.0f;
  uint32_t u;
  memcpy(&u, &g, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float unord32(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float score, uint32_t row) {
  return (uint64_t(ord32(score)) << 32) | row;
}
__host__ __device__ __forceinline__ float key_score(uint64_t k) { return unord32(uint32_t(k >> 32)); }
__host__ __device__ __forceinline__ uint32_t key_row(uint64_t k) { return uint32_t(k); }
// fixed-point (signed integer) scores: same (score desc, row desc) order, 0 stays "empty" because the
// BM25 scale keeps |score| < 2^31
__host__ __device__ __forceinline__ uint64_t make_key_fx(int32_t score, uint32_t row) {
  return (uint64_t(uint32_t(score) ^ 0x80000000u) << 32) | row;
}
__host__ __device__ __forceinline__ int32_t key_fx(uint64_t k) { return int32_t(uint32_t(k >> 32) ^ 0x80000000u); }

__device__ __forceinline__ uint64_t ord64(double d) {
  uint64_t u = (uint64_t)__double_as_longlong(d + 0.0);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

__device__ __forceinline__ double unord64(uint64_t k) {
  const uint64_t u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}

// (score, id) pair ordered by (score, id); s is the ordered image of the score, id -1 = absent
struct Pair64 {
  uint64_t s;
  int64_t id;
};
__device__ __forceinline__ bool pair_less(const Pair64& a, const Pair64& b) {
  return (a.s < b.s) || (a.s == b.s && a.id < b.id);
}

// ---- block-wide bitonic sort, descending, n a power of two, data in shared memory -------
template <typename T, typename Less>
__device__ __forceinline__ void block_bitonic_desc(T* s, int n, Less less) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          T a = s[i], b = s[ixj];
          bool desc = ((i & k) == 0);
          // descending block: larger first
          if (desc ? less(a, b) : less(b, a)) {
            s[i] = b;
            s[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

__host__ __device__ __forceinline__ int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// how the selection stage represented the query (drives the exactness bound in dense_rescore)
#define DRS_QMODE_EXACT 0      /* fp32 query, fp32 FMA accumulation (streaming kernel)            */
#define DRS_QMODE_BF16_RAW 1   /* query rounded to bf16, approximate scores NOT scaled by 1/||q|| */

// candidate-list length carried through the approximate stages: k' plus a margin that the
// exact re-scoring stage consumes (see DESIGN.md "exactness").
__host__ __device__ __forceinline__ int cand_len(int kprime) { return kprime + 8 + (kprime >> 2); }
// The tensor-core kernel ranks rows with a bf16-rounded query (score error up to ~2e-3), so it
// carries a wider margin: the exactness proof needs the gap between the k'-th exact and the
// kk-th approximate score to exceed that error, which k'+24 makes near-certain.
__host__ __device__ __forceinline__ int cand_len_umma(int kprime) { return kprime + 24 + (kprime >> 2); }
// The CTA-pair kernel keeps no list on chip, so its candidate list can afford a margin that makes the proof fail
// only on genuine near-ties: the gap between the k'-th and the (2k'+32)-th best score of a corpus is several times
// the bf16 rounding bound, where k'+24 left ~3 % of the queries of a 10 M-chunk corpus at k' = 60 unproven (each
// costing a pass of the exact kernel).  Bounded by the re-scoring kernel's list (256).
__host__ __device__ __forceinline__ int cand_len_pair(int kprime) {
  const int c = 2 * kprime + 32;
  return c < 64 ? 64 : (c > 256 ? 256 : c);
}

// ---- launchers implemented in the .cu files ---------------------------------------------
// select.cu
size_t merge_scratch_keys(int n_parts, int n_queries, int kk);   // scratch a two-level merge needs (keys)
int launch_merge_keys(const uint64_t* d_parts, int n_parts, int q_stride, int n_queries, int kk,
                      uint64_t* d_out, const int32_t* d_active, uint64_t* d_scratch, cudaStream_t st);
int launch_merge_keys_filtered(const uint64_t* d_parts, int n_parts, int q_stride, int n_queries, int kk,
                               uint64_t* d_out, const int32_t* d_active, uint64_t* d_scratch, const uint32_t* d_tau32,
                               const uint32_t* d_pool, int n_pool, int pool_stride, cudaStream_t st);
int launch_dense_rescore(const float* d_q, int n_queries, const void* d_corpus, const float* d_inv_norm,
                         int dim, const uint64_t* d_cand, int kk, int kprime, float approx_err,
                         float* d_out_scores, int64_t* d_out_ids, int32_t* d_flags,
                         const int32_t* d_active, int query_mode, int flag_unproven, int flag_proven,
                         cudaStream_t st, const int32_t* d_force_unproven = nullptr);
int launch_merge_pairs_f64(const double* d_scores, const int64_t* d_ids, int n_parts, int n_queries, int k, int k_out,
                           double* d_out_scores, int64_t* d_out_ids, int32_t* d_flags, cudaStream_t st);
// dense_simt.cu
int dense_simt_grid();
int launch_dense_simt(const float* d_q, int n_queries, const void* d_corpus, const float* d_inv_norm,
                      const uint8_t* d_mask, int64_t n_rows, int dim, int kk, uint64_t* d_parts,
                      int q_stride, const int32_t* d_active, cudaStream_t st);
// dense_umma.cu
bool dense_umma_supported(int n_queries, int64_t n_rows, int dim, int kk);
int dense_umma_parts(int n_queries, int64_t n_rows, int dim, int kk);
int dense_umma_query_mode();  // DRS_QMODE_* the tensor-core kernel will use
int launch_dense_umma(const float* d_q, int n_queries, const void* d_corpus, const float* d_inv_norm,
                      const uint8_t* d_mask, int64_t n_rows, int dim, int kk, uint64_t* d_parts,
                      int q_stride, uint32_t* d_gtau, cudaStream_t st);

// dense_pair.cu
struct DensePairGeometry {
  int n_slabs, cap, pool_parts, query_rows;
  size_t zero_words;
};
bool dense_pair_supported(int n_queries, int64_t n_rows, int dim, int kk);
bool dense_pair_geometry(int n_queries, int64_t n_rows, int dim, int kk, int q_stride, DensePairGeometry& geo);
int launch_dense_pair(const float* d_q, int n_queries, const void* d_corpus, const float* d_inv_norm,
                      const uint8_t* d_mask, int64_t n_rows, int dim, int kk, uint64_t* d_parts, uint32_t* d_counts,
                      int q_stride, uint32_t* d_zero_area, void* d_qbf16, cudaStream_t st);
// select.cu: top-kk of the keys the pair kernel appended (filtered by its final thresholds); overflow marks
int launch_merge_appended(const uint64_t* d_parts, const uint32_t* d_counts, int n_slabs, int q_stride, int cap,
                          int n_queries, int kk, uint64_t* d_out, const uint32_t* d_pool, int n_pool,
                          int32_t* d_overflow, cudaStream_t st);

}  // namespace drs
