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
#include <cstdlib>

#include "common.cuh"
#include "umma_ptx.cuh"

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

// ------------------------------------------------------------------------------------------------------------------
// Latency mode (one or two queries): the same selection, fed by bulk asynchronous copies.  One producer lane streams
// the CTA's contiguous row range through a ring of shared-memory stages (16 rows = one contiguous block per
// cp.async.bulk, completion on an mbarrier), eight consumer warps score the rows out of shared memory: the copy
// engine keeps a full ring per CTA in flight whatever the consumers are doing, which a load-then-score loop cannot.
// Measured at 10 M x 768, one query (ms per launch, in-process A/B against dense_simt_kernel, three boxes):
// 2.68-2.71 vs 2.785 | 2.80 vs 2.787 | 2.815 vs 2.801 -- 5.5-5.7 TB/s either way: both kernels sit at what a
// read-only stream gets out of HBM on these boxes (0.84-0.88 of the measured copy bandwidth).  Ring depth: 2 stages
// 3.10 ms, 3 stages (two CTAs per SM) 2.80, >= 4 stages (one CTA per SM: eight consumer warps) 4.2.
// ------------------------------------------------------------------------------------------------------------------
constexpr int STREAM_ROWS = 16;
constexpr int STREAM_STAGES = 3;       // default ring depth: two CTAs per SM, 2 x 72 KB in flight at dim 768
constexpr int STREAM_STAGES_MAX = 8;   // DRS_STREAM_STAGES (experiments): > 4 stages of 24 KB leave one CTA per SM
constexpr int STREAM_THREADS = SIMT_THREADS + 32;   // + the producer warp

__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   umma::smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(umma::smem_u32(bar))
               : "memory");
}

template <int QT, int NCH>
__global__ void __launch_bounds__(STREAM_THREADS, 2) dense_stream_kernel(
    const float* __restrict__ queries, int n_queries, int q0, const __nv_bfloat16* __restrict__ corpus,
    const float* __restrict__ inv_norm, const uint8_t* __restrict__ mask, int64_t n_rows, int dim,
    int64_t rows_per_cta, int kk, uint64_t* __restrict__ parts, int q_stride, const int n_stages) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int stage_bytes = STREAM_ROWS * dim * 2;
  uint8_t* ring = smem_raw;                                                       // [n_stages][16 rows][dim] bf16
  uint64_t* lists = reinterpret_cast<uint64_t*>(ring + n_stages * stage_bytes);   // [WARPS][QT][kk]
  uint64_t* sortbuf = lists + SIMT_WARPS * QT * kk;                               // [SIMT_SORT_BUF]
  uint64_t* full = sortbuf + SIMT_SORT_BUF;                                       // [STREAM_STAGES_MAX]
  uint64_t* empty = full + STREAM_STAGES_MAX;                                     // [STREAM_STAGES_MAX]
  __shared__ int qidx[QT];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < n_stages; ++s) {
      umma::mbar_init(&full[s], 1);
      umma::mbar_init(&empty[s], SIMT_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < QT) qidx[threadIdx.x] = (q0 + int(threadIdx.x) < n_queries) ? q0 + int(threadIdx.x) : -1;
  __syncthreads();

  const int64_t r_begin = int64_t(blockIdx.x) * rows_per_cta;
  const int64_t r_end = min(n_rows, r_begin + rows_per_cta);
  const int n_chunks = r_begin < r_end ? int((r_end - r_begin + STREAM_ROWS - 1) / STREAM_ROWS) : 0;

  if (warp == SIMT_WARPS) {
    // ---- producer ----
    if (lane == 0) {
      for (int c = 0; c < n_chunks; ++c) {
        const int s = c % n_stages;
        if (c >= n_stages) umma::mbar_wait(&empty[s], ((c / n_stages) - 1) & 1, 100 + s);
        const int64_t r0 = r_begin + int64_t(c) * STREAM_ROWS;
        const int rows = (r_end - r0) < int64_t(STREAM_ROWS) ? int(r_end - r0) : STREAM_ROWS;
        const uint32_t bytes = uint32_t(rows) * uint32_t(dim) * 2u;
        umma::mbar_arrive_expect_tx(&full[s], bytes);
        bulk_copy_g2s(ring + s * stage_bytes, corpus + size_t(r0) * dim, bytes, &full[s]);
      }
    }
    __syncwarp();
  } else {
    // ---- consumers: queries in registers (element order of the lane), per-warp top-kk lists ----
    float qr[QT][NCH * 8];
#pragma unroll
    for (int j = 0; j < QT; ++j) {
      const int q = qidx[j];
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const int e = ch * 256 + lane * 8;
#pragma unroll
        for (int t = 0; t < 8; ++t) qr[j][ch * 8 + t] = (q >= 0 && e + t < dim) ? queries[size_t(q) * dim + e + t] : 0.0f;
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
    // Two rows at a time, two partial sums per (row, query): four independent FMA chains of NCH*4 per query and the
    // two rows' shuffle reductions interleaved -- the consumers must not become the bottleneck of the ring.
    auto offer = [&](int j, float sc, int64_t r) {
      if (sc >= tau[j] && qidx[j] >= 0) {  // warp-uniform
        const uint64_t key = make_key(sc, uint32_t(r));
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
    };
    auto score_pair = [&](int64_t ra_, const uint8_t* a_smem, float na_, bool ha_, int64_t rb_, const uint8_t* b_smem,
                          float nb_, bool hb_) {
      float a0[QT], a1[QT], b0[QT], b1[QT];
#pragma unroll
      for (int j = 0; j < QT; ++j) a0[j] = a1[j] = b0[j] = b1[j] = 0.0f;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const int e = ch * 256 + lane * 8;
        const uint4 z = make_uint4(0, 0, 0, 0);
        const uint4 wa = (e < dim && ha_) ? *reinterpret_cast<const uint4*>(a_smem + size_t(e) * 2) : z;
        const uint4 wb = (e < dim && hb_) ? *reinterpret_cast<const uint4*>(b_smem + size_t(e) * 2) : z;
#pragma unroll
        for (int j = 0; j < QT; ++j) {
          const float* qq = &qr[j][ch * 8];
          a0[j] = fmaf(bf16_lo(wa.x), qq[0], a0[j]);
          a1[j] = fmaf(bf16_hi(wa.x), qq[1], a1[j]);
          b0[j] = fmaf(bf16_lo(wb.x), qq[0], b0[j]);
          b1[j] = fmaf(bf16_hi(wb.x), qq[1], b1[j]);
          a0[j] = fmaf(bf16_lo(wa.y), qq[2], a0[j]);
          a1[j] = fmaf(bf16_hi(wa.y), qq[3], a1[j]);
          b0[j] = fmaf(bf16_lo(wb.y), qq[2], b0[j]);
          b1[j] = fmaf(bf16_hi(wb.y), qq[3], b1[j]);
          a0[j] = fmaf(bf16_lo(wa.z), qq[4], a0[j]);
          a1[j] = fmaf(bf16_hi(wa.z), qq[5], a1[j]);
          b0[j] = fmaf(bf16_lo(wb.z), qq[4], b0[j]);
          b1[j] = fmaf(bf16_hi(wb.z), qq[5], b1[j]);
          a0[j] = fmaf(bf16_lo(wa.w), qq[6], a0[j]);
          a1[j] = fmaf(bf16_hi(wa.w), qq[7], a1[j]);
          b0[j] = fmaf(bf16_lo(wb.w), qq[6], b0[j]);
          b1[j] = fmaf(bf16_hi(wb.w), qq[7], b1[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < QT; ++j) {
        float sa = a0[j] + a1[j], sb = b0[j] + b1[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          sa += __shfl_xor_sync(0xffffffffu, sa, o);
          sb += __shfl_xor_sync(0xffffffffu, sb, o);
        }
        if (ha_) offer(j, sa * na_, ra_);
        if (hb_) offer(j, sb * nb_, rb_);
      }
    };
    for (int c = 0; c < n_chunks; ++c) {
      const int s = c % n_stages;
      const int64_t r0 = r_begin + int64_t(c) * STREAM_ROWS;
      const int rows = (r_end - r0) < int64_t(STREAM_ROWS) ? int(r_end - r0) : STREAM_ROWS;
      // my two rows of the chunk; their norms (and mask bytes) are fetched before the wait for the stage
      const int ra = warp, rb = warp + SIMT_WARPS;
      const bool ha = ra < rows && (mask == nullptr || mask[r0 + ra] != 0);
      const bool hb = rb < rows && (mask == nullptr || mask[r0 + rb] != 0);
      const float na = ha ? __ldg(inv_norm + r0 + ra) : 0.0f, nb = hb ? __ldg(inv_norm + r0 + rb) : 0.0f;
      umma::mbar_wait(&full[s], (c / n_stages) & 1, 200 + s);
      const uint8_t* st = ring + s * stage_bytes;
      if (ha || hb)
        score_pair(r0 + ra, st + size_t(ra) * dim * 2, na, ha, r0 + rb, st + size_t(rb) * dim * 2, nb, hb);
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&empty[s]);
    }
  }
  __syncthreads();

  // per-CTA merge of the warps' lists, one query at a time (as dense_simt_kernel)
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
    for (int i = threadIdx.x; i < kk; i += blockDim.x) parts[(size_t(blockIdx.x) * q_stride + q) * kk + i] = sortbuf[i];
    __syncthreads();
  }
}

// DRS_DENSE_STREAM=0 falls back to the load-then-score kernel for one- and two-query calls (A/B); read per call.
static bool dense_stream_enabled() {
  const char* e = getenv("DRS_DENSE_STREAM");
  return e == nullptr || atoi(e) != 0;
}
static size_t dense_stream_smem(int n_stages, int dim, int qt, int kk) {
  return size_t(n_stages) * STREAM_ROWS * dim * 2 +
         (size_t(SIMT_WARPS) * qt * kk + SIMT_SORT_BUF + 2 * STREAM_STAGES_MAX) * sizeof(uint64_t);
}
// ring depth: the default, or DRS_STREAM_STAGES, trimmed to what fits one CTA
static int dense_stream_stages(int dim, int qt, int kk) {
  const char* e = getenv("DRS_STREAM_STAGES");
  int n = e ? atoi(e) : STREAM_STAGES;
  n = n < 2 ? 2 : (n > STREAM_STAGES_MAX ? STREAM_STAGES_MAX : n);
  while (n > 2 && dense_stream_smem(n, dim, qt, kk) > 220 * 1024) --n;
  return n;
}

template <int QT, int NCH>
static int launch_stream(const float* d_q, int n_queries, int q0, const void* d_corpus, const float* d_inv_norm,
                         const uint8_t* d_mask, int64_t n_rows, int dim, int kk, uint64_t* d_parts, int q_stride,
                         int grid, int64_t rows_per_cta, cudaStream_t st) {
  const int n_stages = dense_stream_stages(dim, QT, kk);
  const size_t smem = dense_stream_smem(n_stages, dim, QT, kk);
  auto kern = dense_stream_kernel<QT, NCH>;
  DRS_ENSURE_SMEM((kern), smem);
  kern<<<grid, STREAM_THREADS, smem, st>>>(d_q, n_queries, q0, reinterpret_cast<const __nv_bfloat16*>(d_corpus),
                                           d_inv_norm, d_mask, n_rows, dim, rows_per_cta, kk, d_parts, q_stride,
                                           n_stages);
  count_launches(1);
  DRS_CUDA_CHECK(cudaGetLastError());
  return DRS_OK;
}

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
      // one or two queries: the bulk-copy streaming kernel (rows are contiguous 16-byte multiples: dim % 8 == 0)
      const bool stream = qt <= 2 && dense_stream_enabled() && (reinterpret_cast<uintptr_t>(d_corpus) & 15) == 0 &&
                          dense_stream_smem(2, dim, qt, kk) <= 220 * 1024;
#define DRS_STREAM_CASE(QT_, NCH_)                                                                          \
  if (stream && qt == QT_ && nch == NCH_)                                                                   \
    rc = launch_stream<QT_, NCH_>(d_q, n_queries, q0, d_corpus, d_inv_norm, d_mask, n_rows, dim, kk, d_parts, \
                                  q_stride, grid, rows_per_cta, st);
      DRS_STREAM_CASE(1, 1) DRS_STREAM_CASE(1, 2) DRS_STREAM_CASE(1, 3) DRS_STREAM_CASE(1, 4)
      DRS_STREAM_CASE(2, 1) DRS_STREAM_CASE(2, 2) DRS_STREAM_CASE(2, 3) DRS_STREAM_CASE(2, 4)
#undef DRS_STREAM_CASE
      if (rc == DRS_OK) {
        q0 += qt;
        continue;
      }
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
