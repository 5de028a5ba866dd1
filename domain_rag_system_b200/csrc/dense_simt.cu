// Dense cosine candidate selection on CUDA cores: an HBM-bound streaming kernel.
//
// Used when the query batch is small (latency mode, Q <= 8: the corpus is read once and the
// bound is HBM bandwidth, 2*N*D bytes), for dimensions the tensor-core kernel does not
// cover, and to re-run the few queries whose tensor-core result could not be proven exact.
//
// Layout: corpus bf16 row-major [N, D].  One warp owns one row at a time: lane l holds the 8
// consecutive elements [ch*256 + 8l, +8) of every 256-element chunk (one 16-byte load per chunk,
// 512 contiguous bytes per warp per chunk).  The QT queries of the pass live in registers in
// the same element order, so a row costs NCH 16-byte loads and 8*NCH*QT FFMAs per lane plus
// one 5-step shuffle reduction per query.  Each warp keeps a private top-kk list per query in
// shared memory (replace-min with a threshold reject); lists are merged per CTA at the end and
// written as one partial list per (CTA, query).
#include "common.cuh"

namespace drs {

constexpr int SIMT_WARPS = 8;
constexpr int SIMT_THREADS = SIMT_WARPS * 32;
constexpr int SIMT_SORT_BUF = 2048;  // >= next_pow2(SIMT_WARPS * cand_len(DRS_MAX_KPRIME))

__device__ __forceinline__ void warp_argmin(const uint64_t* lst, int kk, int lane, uint64_t& mk, int& mp) {
  uint64_t bk = ~0ull;
  int bp = 0;
  for (int i = lane; i < kk; i += 32) {
    const uint64_t v = lst[i];
    if (v < bk) {
      bk = v;
      bp = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint64_t ok = __shfl_xor_sync(0xffffffffu, bk, o);
    const int op = __shfl_xor_sync(0xffffffffu, bp, o);
    if (ok < bk || (ok == bk && op < bp)) {
      bk = ok;
      bp = op;
    }
  }
  mk = bk;
  mp = bp;
}

template <int QT, int NCH>
__global__ void __launch_bounds__(SIMT_THREADS, 2) dense_simt_kernel(
    const float* __restrict__ queries, int n_queries, int q0, const __nv_bfloat16* __restrict__ corpus,
    const float* __restrict__ inv_norm, const uint8_t* __restrict__ mask, int64_t n_rows, int dim,
    int64_t rows_per_cta, int kk, uint64_t* __restrict__ parts, int q_stride,
    const int32_t* __restrict__ active, int n_passes) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint64_t* lists = reinterpret_cast<uint64_t*>(smem_raw);              // [WARPS][QT][kk]
  uint64_t* sortbuf = lists + SIMT_WARPS * QT * kk;                     // [SIMT_SORT_BUF]
  __shared__ int qidx[QT];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // active != nullptr: up to n_passes passes over the corpus, QT flagged queries each (the escalation of
  // queries whose tensor-core result is unproven); the loop ends at the first pass with nothing to do.
  for (int pass = 0; pass < n_passes; ++pass) {
    // which queries does this pass serve?
    if (threadIdx.x < 32) {
      if (active == nullptr) {
        if (lane < QT) qidx[lane] = (q0 + lane < n_queries) ? q0 + lane : -1;
      } else {
        // the (pass*QT + j)-th active query, j < QT
        if (lane < QT) qidx[lane] = -1;
        __syncwarp();
        int seen = 0;
        const int want0 = pass * QT;
        for (int base = 0; base < n_queries && seen < want0 + QT; base += 32) {
          const int q = base + lane;
          const bool a = (q < n_queries) && (active[q] != 0);
          const unsigned b = __ballot_sync(0xffffffffu, a);
          if (a) {
            const int ord = seen + __popc(b & ((1u << lane) - 1u));
            if (ord >= want0 && ord < want0 + QT) qidx[ord - want0] = q;
          }
          seen += __popc(b);
        }
      }
    }
    __syncthreads();
    bool any = false;
  #pragma unroll
    for (int j = 0; j < QT; ++j) any |= (qidx[j] >= 0);
    if (!any) break;   // no (more) flagged queries

    // queries -> registers (element order of the lane)
    float qr[QT][NCH * 8];
  #pragma unroll
    for (int j = 0; j < QT; ++j) {
      const int q = qidx[j];
  #pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const int e = ch * 256 + lane * 8;
  #pragma unroll
        for (int t = 0; t < 8; ++t)
          qr[j][ch * 8 + t] = (q >= 0 && e + t < dim) ? queries[size_t(q) * dim + e + t] : 0.0f;
      }
    }
    uint64_t* mylist = lists + size_t(warp) * QT * kk;
    for (int i = lane; i < QT * kk; i += 32) mylist[i] = 0;
    __syncwarp();
    float tau[QT];
    uint64_t tau_key[QT];
  #pragma unroll
    for (int j = 0; j < QT; ++j) {
      tau[j] = -INFINITY;
      tau_key[j] = 0;
    }

    const int64_t r_begin = int64_t(blockIdx.x) * rows_per_cta;
    const int64_t r_end = min(n_rows, r_begin + rows_per_cta);

    auto load_row = [&](int64_t r, uint4 (&w)[NCH]) {
      const uint4* rp = reinterpret_cast<const uint4*>(corpus + size_t(r) * dim);
  #pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const int e = ch * 256 + lane * 8;
        w[ch] = (e < dim) ? __ldg(rp + (e >> 3)) : make_uint4(0, 0, 0, 0);
      }
    };
    auto score_row = [&](int64_t r, const uint4 (&w)[NCH]) {
      if (mask != nullptr && mask[r] == 0) return;
      float acc[QT];
  #pragma unroll
      for (int j = 0; j < QT; ++j) acc[j] = 0.0f;
  #pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const float c0 = bf16_lo(w[ch].x), c1 = bf16_hi(w[ch].x), c2 = bf16_lo(w[ch].y), c3 = bf16_hi(w[ch].y);
        const float c4 = bf16_lo(w[ch].z), c5 = bf16_hi(w[ch].z), c6 = bf16_lo(w[ch].w), c7 = bf16_hi(w[ch].w);
  #pragma unroll
        for (int j = 0; j < QT; ++j) {
          const float* qq = &qr[j][ch * 8];
          acc[j] = fmaf(c0, qq[0], acc[j]);
          acc[j] = fmaf(c1, qq[1], acc[j]);
          acc[j] = fmaf(c2, qq[2], acc[j]);
          acc[j] = fmaf(c3, qq[3], acc[j]);
          acc[j] = fmaf(c4, qq[4], acc[j]);
          acc[j] = fmaf(c5, qq[5], acc[j]);
          acc[j] = fmaf(c6, qq[6], acc[j]);
          acc[j] = fmaf(c7, qq[7], acc[j]);
        }
      }
      const float nrm = __ldg(inv_norm + r);
  #pragma unroll
      for (int j = 0; j < QT; ++j) {
        float a = acc[j];
  #pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        const float s = a * nrm;
        if (s >= tau[j] && qidx[j] >= 0) {  // warp-uniform
          const uint64_t key = make_key(s, uint32_t(r));
          if (key > tau_key[j]) {
            uint64_t* lst = mylist + j * kk;
            uint64_t mk;
            int mp;
            warp_argmin(lst, kk, lane, mk, mp);
            if (lane == 0) lst[mp] = key;
            __syncwarp();
            warp_argmin(lst, kk, lane, mk, mp);
            tau_key[j] = mk;
            tau[j] = mk ? key_score(mk) : -INFINITY;
          }
        }
      }
    };

    // rows in flight per warp: four when the queries leave registers for it (latency mode: the kernel is
    // purely HBM-bound and needs ~64 KB of loads outstanding per SM), else two.  (A rolling window that issues a
    // row's loads three rows ahead of its scoring, with L1::no_allocate loads, measured SLOWER: 4.06 vs 2.79 ms
    // at 10 M x 768 -- the batch of twelve 16-byte loads per warp is what keeps HBM busy.  L1::no_allocate /
    // L2::evict_first loads in THIS loop: 2.786 vs 2.786 ms, no difference.)
    int64_t r = r_begin + warp;
    if (QT <= 2) {
      for (; r + 3 * SIMT_WARPS < r_end; r += 4 * SIMT_WARPS) {
        uint4 w0[NCH], w1[NCH], w2[NCH], w3[NCH];
        load_row(r, w0);
        load_row(r + SIMT_WARPS, w1);
        load_row(r + 2 * SIMT_WARPS, w2);
        load_row(r + 3 * SIMT_WARPS, w3);
        score_row(r, w0);
        score_row(r + SIMT_WARPS, w1);
        score_row(r + 2 * SIMT_WARPS, w2);
        score_row(r + 3 * SIMT_WARPS, w3);
      }
    }
    for (; r + SIMT_WARPS < r_end; r += 2 * SIMT_WARPS) {
      uint4 w0[NCH], w1[NCH];
      load_row(r, w0);
      load_row(r + SIMT_WARPS, w1);
      score_row(r, w0);
      score_row(r + SIMT_WARPS, w1);
    }
    if (r < r_end) {
      uint4 w0[NCH];
      load_row(r, w0);
      score_row(r, w0);
    }
    __syncthreads();

    // per-CTA merge of the warps' lists, one query at a time
    const int n = SIMT_WARPS * kk;
    const int m = next_pow2(n);
  #pragma unroll 1
    for (int j = 0; j < QT; ++j) {
      const int q = qidx[j];
      if (q < 0) continue;
      for (int i = threadIdx.x; i < m; i += blockDim.x) {
        uint64_t v = 0;
        if (i < n) {
          const int w = i / kk, c = i - w * kk;
          v = lists[(size_t(w) * QT + j) * kk + c];
        }
        sortbuf[i] = v;
      }
      __syncthreads();
      block_bitonic_desc(sortbuf, m, [](uint64_t a, uint64_t b) { return a < b; });
      for (int i = threadIdx.x; i < kk; i += blockDim.x)
        parts[(size_t(blockIdx.x) * q_stride + q) * kk + i] = sortbuf[i];
      __syncthreads();
    }
    __syncthreads();
  }
}

int dense_simt_grid() { return 2 * sm_count(); }

template <int QT, int NCH>
static int launch_one(const float* d_q, int n_queries, int q0, const void* d_corpus, const float* d_inv_norm,
                      const uint8_t* d_mask, int64_t n_rows, int dim, int kk, uint64_t* d_parts, int q_stride,
                      const int32_t* d_active, int pass, int grid, int64_t rows_per_cta, cudaStream_t st) {
  const size_t smem = (size_t(SIMT_WARPS) * QT * kk + SIMT_SORT_BUF) * sizeof(uint64_t);
  auto kern = dense_simt_kernel<QT, NCH>;
  DRS_ENSURE_SMEM((kern), smem);
  kern<<<grid, SIMT_THREADS, smem, st>>>(d_q, n_queries, q0, reinterpret_cast<const __nv_bfloat16*>(d_corpus),
                                          d_inv_norm, d_mask, n_rows, dim, rows_per_cta, kk, d_parts, q_stride,
                                          d_active, pass);
  count_launches(1);
  DRS_CUDA_CHECK(cudaGetLastError());
  return DRS_OK;
}

// parts must hold dense_simt_grid() * q_stride * kk keys.  With d_active == nullptr every query is
// served (ceil(Q/QT) passes over the corpus); with d_active set, one pass per 4 flagged queries.
int launch_dense_simt(const float* d_q, int n_queries, const void* d_corpus, const float* d_inv_norm,
                      const uint8_t* d_mask, int64_t n_rows, int dim, int kk, uint64_t* d_parts, int q_stride,
                      const int32_t* d_active, cudaStream_t st) {
  DRS_REQUIRE(dim % 8 == 0 && dim > 0 && dim <= DRS_MAX_DIM, DRS_ERR_UNSUPPORTED,
              "dense: dim must be a multiple of 8 in (0, 1024]");
  DRS_REQUIRE(SIMT_WARPS * kk <= SIMT_SORT_BUF, DRS_ERR_UNSUPPORTED, "dense: k' too large");
  const int grid = dense_simt_grid();
  int64_t rows_per_cta = (n_rows + grid - 1) / grid;
  rows_per_cta = ((rows_per_cta + 15) / 16) * 16;
  if (rows_per_cta < 16) rows_per_cta = 16;
  // every CTA writes its partial list (possibly empty), so the merge can read all `grid` parts
  const int nch = (dim + 255) / 256;
#define DRS_SIMT_CASE(QT_, NCH_)                                                                         \
  if (qt == QT_ && nch == NCH_)                                                                          \
    rc = launch_one<QT_, NCH_>(d_q, n_queries, q0, d_corpus, d_inv_norm, d_mask, n_rows, dim, kk, d_parts, \
                               q_stride, d_active, pass, grid, rows_per_cta, st);
  if (d_active == nullptr) {
    for (int q0 = 0; q0 < n_queries;) {
      const int left = n_queries - q0;
      const int qt = left >= 4 ? 4 : (left >= 2 ? 2 : 1);
      const int pass = 1;
      int rc = DRS_ERR_UNSUPPORTED;
      DRS_SIMT_CASE(1, 1) DRS_SIMT_CASE(1, 2) DRS_SIMT_CASE(1, 3) DRS_SIMT_CASE(1, 4)
      DRS_SIMT_CASE(2, 1) DRS_SIMT_CASE(2, 2) DRS_SIMT_CASE(2, 3) DRS_SIMT_CASE(2, 4)
      DRS_SIMT_CASE(4, 1) DRS_SIMT_CASE(4, 2) DRS_SIMT_CASE(4, 3) DRS_SIMT_CASE(4, 4)
      if (rc != DRS_OK) return rc;
      q0 += qt;
    }
  } else {
    // one launch; inside the kernel one pass over the corpus per four flagged queries, ending at the first pass
    // with nothing left to do (normally the first or second)
    const int qt = 4, q0 = 0, pass = (n_queries + 3) / 4;
    int rc = DRS_ERR_UNSUPPORTED;
    DRS_SIMT_CASE(4, 1) DRS_SIMT_CASE(4, 2) DRS_SIMT_CASE(4, 3) DRS_SIMT_CASE(4, 4)
    if (rc != DRS_OK) return rc;
  }
#undef DRS_SIMT_CASE
  return DRS_OK;
}

}  // namespace drs
