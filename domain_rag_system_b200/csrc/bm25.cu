// BM25 over a term-major CSR inverted index, document-partitioned into tiles of 8192 documents.
//
// Stage 1 (bm25_select_kernel<false>): INTEGER pre-selection.  Impacts are stored in fixed point
// (round(impact * 2^s), |.| <= 2^26), so a document's pre-selection score is an integer sum: the order of the
// additions is free (shared-memory atomics, no barrier between terms) and documents with the same
// (term, impact) multiset tie exactly, as they do in the reference's fp64 arithmetic.
//   * work item = (slice of <= 32 consecutive tiles, query); persistent CTAs pull items from one atomic counter
//     in SLICE-MAJOR order, so the whole grid works on one slice at a time and that slice's rows and postings
//     stay in L2 while every query passes over them;
//   * per (tile, query), on an int32 score tile in shared memory:
//       pass A -- the query's head terms (df >= N/4: ~94 % of a Zipfian query's postings) are added from dense
//                 rows of packed 14-bit values (16-byte loads = 8 documents, four rows summed as packed 16-bit
//                 lanes), by the thread that owns the documents; the first sweep also initialises the tile
//                 (bm25_dense_sweep, specialised on the number of rows);
//       pass B -- the other terms' postings inside the tile are cut into chunks of 128 (aligned 16-byte loads of
//                 doc ids and impacts), the warps take chunks round-robin over two register sets with the next
//                 chunk's loads in flight, and add with shared-memory integer atomics;
//       selection -- with a warm positive threshold and no negative impacts, sums only grow: the thread that
//                 takes a document across the threshold (old < tau <= new) notes it, and the tile is never
//                 scanned.  Otherwise (cold list, negative-idf corpora): a vectorised scan, highest document
//                 first, with geometric steps while the list is young and a segment-wise overflow path;
//       merge  -- survivors + list <= 256 keys: rank by counting (keys are distinct); else a block bitonic sort.
//                 A global per-query threshold (atomicMax) lets the CTAs prune each other.
// Stage 2 (bm25_rescore_kernel): the kk candidates per query are re-scored in fp64 with rank_bm25's exact
// operation order (one thread per (candidate, token) looks the posting up by binary search), ranked by
// (score desc, doc desc), and the selection is PROVEN with an error bound or flagged.
// Stage 3 (bm25_select_kernel<true>, predicated on the flags): every document that can still belong to the
// top-k' of a flagged query is re-scored exactly; exact (score, doc) top-k' per slice, merged by merge_pairs.
//
// The kernel's algorithmic traffic is 8 bytes per posting visited; real DRAM traffic is ~100x lower
// (2-byte head-term rows shared by all queries through L2).  What bounds it (profiles/r2t_summary.txt): the
// dependent chain of a (tile, query) pass -- rows, barrier, postings, barrier, selection; ~11.4 k cycles per tile,
// four passes in flight per SM -- with the SM saturated by the mix (issue slots 56 %, LSU 56 %), no unit at its peak.
// DRS_BM25_TRACE prints the phase timeline of one CTA, DRS_BM25_DEBUG switches phases off for timing experiments,
// DRS_BM25_CTAS=5 selects the 48-register build (five CTAs per SM).
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace drs {

constexpr int BM25_THREADS = 256;
// merge buffer (u64 keys): list + candidate append area.  CPS = resident CTAs per SM the kernel is compiled for: 4 (64
// registers, 16 KB buffer; the default) or 5 (48 registers, 8 KB buffer: 25 % more documents in flight per SM).
// Measured at 1 M x 1000, k'=10, in one process: 3 CTAs (79 registers) 1.494 ms, 4 CTAs 1.287 ms, 5 CTAs 1.337 ms --
// the occupancy curve is flat at the top: the SM is saturated by the mix (issue slots 56 %, LSU 56 %, latency).
__host__ __device__ constexpr int bm25_mbuf(int cps) { return cps >= 5 ? 1024 : 2048; }
constexpr int BM25_DEFAULT_CPS = 4;
constexpr int BM25_TCHUNK = 64;  // query term occurrences whose posting ranges are staged per tile at once
constexpr int BM25_SEG = 1024;   // documents scanned between overflow checks
constexpr int BM25_MERGE_AT = 192;  // survivors that justify a re-sort (warm list): list + survivors fit a 256-sort
constexpr int BM25_COLD_STEP = 128; // documents scanned before the first merge of a cold list

// exact fp64 score of one document; rank_bm25 operation order:
//   score += idf[t] * ( tf*(k1+1) / ( tf + k1*(1-b+b*dl/avgdl) ) )     for every query token occurrence
__device__ __forceinline__ double bm25_exact_score(const drs_bm25_index& ix, const int32_t* __restrict__ q_terms,
                                                   int t0, int t1, int64_t doc) {
  const int j = int(doc / ix.tile_docs);
  const double kd = ix.d_doc_kd[doc];
  double score = 0.0;
  for (int ti = t0; ti < t1; ++ti) {
    const int t = q_terms[ti];
    if (t < 0 || t >= ix.n_terms) continue;
    const int64_t base = ix.d_term_offsets[t];
    const uint32_t* tp = ix.d_tile_ptr + size_t(t) * (ix.n_tiles + 1) + j;
    const int64_t end = base + tp[1];
    int64_t lo = base + tp[0], hi = end;  // [lo, hi)
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (ix.d_post_doc[mid] < uint32_t(doc)) lo = mid + 1;
      else hi = mid;
    }
    if (lo < end && ix.d_post_doc[lo] == uint32_t(doc)) {
      const double tf = double(ix.d_post_tf[lo]);
      const double num = __dmul_rn(tf, ix.k1_plus_1);
      const double den = __dadd_rn(tf, kd);
      score = __dadd_rn(score, __dmul_rn(ix.d_idf[t], __ddiv_rn(num, den)));
    }
  }
  return score;
}

// shift applied to the stored impacts of a query with nt token occurrences (see bm25_select_kernel)
__host__ __device__ __forceinline__ int bm25_query_shift(int nt) {
  int sh = 0;
  while ((16 << sh) < nt) ++sh;   // nt * 2^(26 - sh) <= 2^30
  return sh;
}

// shared-memory integer atomics by 32-bit shared address (one IMAD for the address, no generic-pointer arithmetic)
__device__ __forceinline__ int smem_atom_add_s32(uint32_t addr, int v) {
  int r;
  asm volatile("atom.shared.add.s32 %0, [%1], %2;" : "=r"(r) : "r"(addr), "r"(v) : "memory");
  return r;
}
__device__ __forceinline__ void smem_red_add_s32(uint32_t addr, int v) {
  asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// Pass A of bm25_select_kernel on its fast path (tile_docs % 4096 == 0, rows not finer than the score unit):
// NG head-term rows are summed as packed 16-bit lanes and added to the score tile, two 16-byte groups (16
// documents) per thread and iteration, all 2*NG loads issued before the first use.  ACC = the tile already holds
// this query's sums (second and later sweeps); CROSS = note the documents a sweep takes across the threshold.
// Specialised on NG so that the loop carries no per-row predicates: ~40 instructions per group of 8 documents
// (98 in the general loop it replaces).
template <int NG, bool CROSS, class NoteFn>
__device__ __forceinline__ void bm25_dense_sweep(int4* __restrict__ dst, const int T8, const uint16_t* const* rowp,
                                                 const int64_t doc0, const int mul, const int tau_s0, const int tid,
                                                 const bool ACC, NoteFn&& note) {
  const uint4* r[NG];
#pragma unroll
  for (int k = 0; k < NG; ++k) r[k] = reinterpret_cast<const uint4*>(rowp[k] + doc0);
  auto finish = [&](const uint4 w, const int i) {
    int4 lo4 = make_int4(0, 0, 0, 0), hi4 = lo4;
    if (ACC) {
      lo4 = dst[i];
      hi4 = dst[T8 + i];
    }
    lo4.x += int(w.x & 0xffffu) * mul; lo4.y += int(w.x >> 16) * mul;
    lo4.z += int(w.y & 0xffffu) * mul; lo4.w += int(w.y >> 16) * mul;
    hi4.x += int(w.z & 0xffffu) * mul; hi4.y += int(w.z >> 16) * mul;
    hi4.z += int(w.w & 0xffffu) * mul; hi4.w += int(w.w >> 16) * mul;
    dst[i] = lo4;
    dst[T8 + i] = hi4;
    if (CROSS) {
      if (max(max(max(lo4.x, lo4.y), max(lo4.z, lo4.w)), max(max(hi4.x, hi4.y), max(hi4.z, hi4.w))) >= tau_s0) {
        // rare: which of my eight documents did THIS sweep take across the threshold?  (one note() site: the
        // kernel is instruction-cache sensitive)
        const int nv[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
        const unsigned ws[4] = {w.x, w.y, w.z, w.w};
        unsigned hit = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int cv = int((u & 1) ? (ws[u >> 1] >> 16) : (ws[u >> 1] & 0xffffu)) * mul;   // this sweep's share
          hit |= (nv[u] >= tau_s0 && nv[u] - cv < tau_s0) ? (1u << u) : 0u;
        }
#pragma unroll 1
        while (hit) {
          const int u = __ffs(hit) - 1;
          hit &= hit - 1;
          note((u < 4 ? 0 : 4 * T8) + 4 * i + (u & 3));
        }
      }
    }
  };
  // Software pipeline: the next iteration's 2*NG loads are issued as soon as this iteration's rows are folded into
  // two packed words, so the unpack / accumulate / store work of one iteration hides the L2 latency of the next.
  // (Three and four rows in flight twice over do not fit the 64-register budget: those sweeps load per iteration.)
  constexpr bool PIPE = NG <= 2;
  uint4 a[NG], b[NG];
  if (PIPE) {
#pragma unroll
    for (int k = 0; k < NG; ++k) {
      a[k] = __ldg(r[k] + tid);
      b[k] = __ldg(r[k] + tid + BM25_THREADS);
    }
  }
  for (int i = tid; i < T8; i += 2 * BM25_THREADS) {
    if (!PIPE) {
#pragma unroll
      for (int k = 0; k < NG; ++k) {
        a[k] = __ldg(r[k] + i);
        b[k] = __ldg(r[k] + i + BM25_THREADS);
      }
    }
    uint4 wa = a[0], wb = b[0];
#pragma unroll
    for (int k = 1; k < NG; ++k) {
      wa.x += a[k].x; wa.y += a[k].y; wa.z += a[k].z; wa.w += a[k].w;
      wb.x += b[k].x; wb.y += b[k].y; wb.z += b[k].z; wb.w += b[k].w;
    }
    const int in = i + 2 * BM25_THREADS;
    if (PIPE && in < T8) {
#pragma unroll
      for (int k = 0; k < NG; ++k) {
        a[k] = __ldg(r[k] + in);
        b[k] = __ldg(r[k] + in + BM25_THREADS);
      }
    }
    finish(wa, i);
    finish(wb, i + BM25_THREADS);
  }
}

// EXACT = false: the pre-selection pass (integer scores, top-kk keys per (slice, query)).
// EXACT = true : the escalation pass for the queries whose pre-selection could not be proven (flags[q] == 1):
//   same accumulation, but EVERY document whose integer score reaches fx_low[q] -- the smallest integer score
//   a member of the true top-k' can have, see bm25_rescore_kernel -- is re-scored exactly on the spot and the
//   exact (score, doc) top-k' of the slice is kept.
template <bool EXACT, int CPS>
__global__ void __launch_bounds__(BM25_THREADS, CPS) bm25_select_kernel(
    drs_bm25_index ix, const int32_t* __restrict__ q_terms, const int32_t* __restrict__ q_offsets,
    int n_queries, int kk, int tiles_per_cta, uint64_t* __restrict__ parts, int q_stride,
    unsigned long long* __restrict__ gtau, const int32_t* __restrict__ flags, const int32_t* __restrict__ fx_low,
    int kprime, double* __restrict__ e_scores, int64_t* __restrict__ e_ids, int* __restrict__ next_query,
    unsigned long long* __restrict__ trace, int dbg) {
  constexpr int BM25_MBUF = bm25_mbuf(CPS);
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint64_t* mbuf = reinterpret_cast<uint64_t*>(smem_raw);                // [BM25_MBUF]
  Pair64* pbuf = reinterpret_cast<Pair64*>(smem_raw);                    // EXACT: [BM25_MBUF / 2] (score, doc) pairs
  __shared__ Pair64 s_kth;                                               // EXACT: current k'-th best pair
  int32_t* sc = reinterpret_cast<int32_t*>(mbuf + BM25_MBUF);            // [tile_docs] fixed-point scores
  __shared__ int s_cnt;
  __shared__ int s_note[2];   // crossing mode: documents noted in the current / previous tile (alternating)
  __shared__ int s_q;         // the work item (slice, query) this CTA works on (handed out by *next_item)
  __shared__ unsigned long long s_tau;
  __shared__ int s_slot[BM25_TCHUNK];                        // dense row of the staged term, or -1
  __shared__ int s_drow[BM25_TCHUNK];                        // compacted dense rows, query order
  __shared__ int s_nd;
  __shared__ const uint16_t* s_rowp[BM25_TCHUNK];   // start of the compacted dense rows
  __shared__ int s_term[BM25_TCHUNK];
  __shared__ long long s_base[BM25_TCHUNK];

  const int T = ix.tile_docs;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_slices = (ix.n_tiles + tiles_per_cta - 1) / tiles_per_cta;
  const int cap = BM25_MBUF - kk;

  auto merge = [&](int q, int cnt_in) {
    // callers guarantee all threads arrive here with the same cnt_in; cnt_in candidates sit at mbuf[kk ..)
    const int cnt = min(cnt_in, cap);
    const int n = kk + cnt;
    if (n <= BM25_THREADS) {
      // Small merges (the common case: a few dozen survivors) rank by counting -- keys are distinct, so an
      // element's rank is the number of larger keys: three barriers instead of a 28..36-stage sorting network.
      // (Sharing a key's comparisons among 256 / n adjacent threads and dropping one barrier was measured: no
      // gain, 1.309 vs 1.296 ms in an in-process A/B -- the merge is not what the CTA waits for.)
      const uint64_t key = tid < n ? mbuf[tid] : 0;
      int rank = 0;
      if (key != 0)
        for (int j = 0; j < n; ++j) rank += mbuf[j] > key;
      __syncthreads();
      if (tid < kk) mbuf[tid] = 0;
      __syncthreads();
      if (key != 0 && rank < kk) mbuf[rank] = key;
      __syncthreads();
    } else {
      const int m = next_pow2(n);
      for (int i = n + tid; i < m; i += BM25_THREADS) mbuf[i] = 0;
      __syncthreads();
      block_bitonic_desc(mbuf, m, [](uint64_t a, uint64_t b) { return a < b; });
    }
    if (tid == 0) {
      s_cnt = 0;
      unsigned long long t = mbuf[kk - 1];  // 0 while the list is not full
      // one global round trip: the atomic's return value is the other CTAs' best so far
      const unsigned long long g = t ? atomicMax(gtau + q, t) : *reinterpret_cast<volatile unsigned long long*>(gtau + q);
      s_tau = t > g ? t : g;
    }
    __syncthreads();
  };

  // Term metadata of up to BM25_TCHUNK query-term occurrences: term id, dense row (or -1), start of the
  // postings list (or -1 when the term is dense, out of vocabulary, or absent).  Tile independent, so a
  // query that fits in one chunk stages it ONCE per (CTA, query) instead of once per tile.
  auto stage_terms = [&](int c0, int nstaged) {
    __syncthreads();
    if (tid < nstaged) {
      const int t = q_terms[c0 + tid];
      int slot = -1;
      long long base = -1;
      if (t >= 0 && t < ix.n_terms) {  // OOV: contributes nothing
        slot = ix.d_dense_slot ? ix.d_dense_slot[t] : -1;
        if (slot < 0) base = ix.d_term_offsets[t];
      }
      s_term[tid] = t;
      s_slot[tid] = slot;
      s_base[tid] = base;
    }
    __syncthreads();
    if (tid < 32) {   // warp 0 compacts the dense rows, keeping query order
      int nd = 0;
      for (int b0 = 0; b0 < nstaged; b0 += 32) {
        const int i = b0 + tid;
        const bool f = i < nstaged && s_slot[i] >= 0;
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (f) {
          const int at = nd + __popc(m & ((1u << tid) - 1u));
          s_drow[at] = s_slot[i];
          s_rowp[at] = ix.d_dense_impact + size_t(s_slot[i]) * ix.dense_stride;
        }
        nd += __popc(m);
      }
      if (tid == 0) s_nd = nd;
    }
    __syncthreads();
  };

  if (tid < 2) s_note[tid] = 0;
  int par = 0;
  // Work items = (slice of consecutive tiles, query), handed out dynamically (an atomic counter) in SLICE-MAJOR
  // order: all CTAs of the grid work on the same slice (or two adjacent ones) at any time, so the head-term rows
  // and postings of that slice -- a few MB -- stay in L2 while every query passes over them, instead of every
  // query streaming its rows from HBM (at 12.5 M documents the rows alone are 0.5 GB, far beyond L2).  Slices are
  // taken from the top of the corpus downwards; after a query's first slice its global threshold is warm.
  // Dynamic hand-out also evens out the cost differences between queries.
  for (;;) {
    __syncthreads();
    if (tid == 0) s_q = atomicAdd(next_query, 1);
    __syncthreads();
    const int item = s_q;
    if (item >= n_slices * n_queries) break;
    const int slice = n_slices - 1 - item / n_queries, q = item - (item / n_queries) * n_queries;
    const int tile_lo = slice * tiles_per_cta, tile_hi = min(ix.n_tiles, tile_lo + tiles_per_cta) - 1;
    if (EXACT && flags[q] != 1) continue;
    const int t0 = q_offsets[q], t1 = q_offsets[q + 1];
    const bool one_chunk = (t1 - t0) <= BM25_TCHUNK;
    // Stored impacts are at most 2^26 in magnitude; with more than 16 token occurrences they are shifted
    // down so that no document's sum can leave the signed 31-bit range.  The shift depends on the query
    // alone, so scores stay comparable across tiles, CTAs and shards of the launch.
    const int sh = bm25_query_shift(t1 - t0);
    long long seen = 0;   // documents scanned so far for this (slice, query)
    int cnt_list = 0;     // survivors waiting in the append area (the same value in every thread)
    if (one_chunk) stage_terms(t0, t1 - t0);
    if constexpr (EXACT) {
      for (int i = tid; i < kprime; i += BM25_THREADS) pbuf[i] = Pair64{0, -1};
      if (tid == 0) {
        s_cnt = 0;
        s_kth = Pair64{0, -1};
      }
    } else {
      for (int i = tid; i < kk; i += BM25_THREADS) mbuf[i] = 0;
      if (tid == 0) {
        s_cnt = 0;
        s_tau = *reinterpret_cast<volatile unsigned long long*>(gtau + q);
      }
    }
    __syncthreads();

    for (int j = tile_hi; j >= tile_lo; --j) {
      const int64_t doc0 = int64_t(j) * T;
      const int64_t rem = ix.n_docs - doc0;
      const int tile_n = rem < int64_t(T) ? int(rem) : T;
      const uint32_t d0u = uint32_t(doc0);  // tile-local index = document id - d0u
      // Crossing mode (warm positive threshold, no negative impacts in the index): sums only grow, so a document
      // reaches the threshold at exactly one addition -- the thread that performs it (old < tau <= new) notes the
      // document, and the 8192-document scan after the accumulation is replaced by a look at the noted few.
      const unsigned long long tau0 = s_tau;
      const int tau_s0 = tau0 ? key_fx(tau0) : INT_MIN;
      const bool cross = !EXACT && ix.impacts_nonneg != 0 && tau_s0 > 0;
      // The noted documents are counted in s_note[par], par alternating from tile to tile: the count of tile j
      // is read by every thread after tile j's last barrier while the early threads may already be noting for
      // tile j+1 -- in the other counter.  Thread 0 clears a counter one tile after its use.
      par ^= 1;
      // DRS_BM25_TRACE (debugging aid): clock of CTA 0 / thread 0 at the phase boundaries of its first 400 tiles
      auto mark = [&](int k) {
        if (trace != nullptr && blockIdx.x == 0 && tid == 0) {
          const unsigned long long n = trace[0];
          if (n < 400) trace[1 + n * 7 + k] = clock64();
          if (k == 6) trace[0] = n + 1;
        }
      };
      mark(0);
      auto note = [&](int dl) {
        const int pos = cnt_list + atomicAdd(&s_note[par], 1);
        if (pos < cap) mbuf[kk + pos] = uint64_t(dl);
      };
      // ---- accumulate ---------------------------------------------------------------------
      // Scores are integer sums of fixed-point impacts, so the order of the additions is free: the head
      // terms' dense rows are added by the thread that owns the documents (plain read-modify-write, the
      // first sweep also initialises the tile), then the other terms' postings are added with shared-
      // memory integer atomics by whichever warp picks the postings up -- no barrier between terms.
      bool tile_init = false;   // false: the tile still holds the previous query's scores
      for (int c0 = t0; c0 < t1; c0 += BM25_TCHUNK) {
        const int nstaged = min(BM25_TCHUNK, t1 - c0);
        if (!one_chunk) stage_terms(c0, nstaged);   // long queries: term metadata is restaged per tile
        // Every warp fetches the posting ranges of the first 32 staged terms inside this tile for itself
        // (term id -> tile pointer is a dependent load; issued here, consumed after the dense sweeps).
        uint32_t tp0 = 0, tp1 = 0;
        if (lane < nstaged && s_base[lane] >= 0) {
          const uint32_t* tp = ix.d_tile_ptr + size_t(s_term[lane]) * (ix.n_tiles + 1) + j;
          tp0 = __ldg(tp);
          tp1 = __ldg(tp + 1);
        }
        // dbg (DRS_BM25_DEBUG, timing experiments only -- results are garbage): 1 no head-term rows, 2 no postings
        // pass, 4 postings fetched but not added, 8 no selection
        const int nd = (dbg & 1) ? 0 : s_nd;
        // -- pass A: dense rows (packed 14-bit values, 8 documents per 16-byte load), up to four rows summed as
        // packed 16-bit lanes per sweep; the first sweep also initialises the tile.  16-byte group i of a row
        // holds documents 4i..4i+3 of the tile's first half and 4i..4i+3 of its second half, so a thread's two
        // shared-memory accesses (int4 i and int4 T/8 + i) are conflict free.
        int4* dst = reinterpret_cast<int4*>(sc);
        const int T8 = T >> 3;
        const int dk = ix.dense_shift - sh;   // row value -> score units: << dk (>> -dk when negative)
        auto unpack_add = [&](int4& lo4, int4& hi4, const uint4 w, int i) {
          int4 cl, ch;   // this sweep's contribution to my eight documents
          if (dk >= 0) {   // value * 2^dk: the multiply folds into the add below (one IMAD per document)
            const int sc2 = 1 << dk;
            cl.x = int(w.x & 0xffffu) * sc2; cl.y = int(w.x >> 16) * sc2;
            cl.z = int(w.y & 0xffffu) * sc2; cl.w = int(w.y >> 16) * sc2;
            ch.x = int(w.z & 0xffffu) * sc2; ch.y = int(w.z >> 16) * sc2;
            ch.z = int(w.w & 0xffffu) * sc2; ch.w = int(w.w >> 16) * sc2;
          } else {
            cl.x = int((w.x & 0xffffu) >> -dk); cl.y = int((w.x >> 16) >> -dk);
            cl.z = int((w.y & 0xffffu) >> -dk); cl.w = int((w.y >> 16) >> -dk);
            ch.x = int((w.z & 0xffffu) >> -dk); ch.y = int((w.z >> 16) >> -dk);
            ch.z = int((w.w & 0xffffu) >> -dk); ch.w = int((w.w >> 16) >> -dk);
          }
          lo4.x += cl.x; lo4.y += cl.y; lo4.z += cl.z; lo4.w += cl.w;
          hi4.x += ch.x; hi4.y += ch.y; hi4.z += ch.z; hi4.w += ch.w;
          if (cross && max(max(max(lo4.x, lo4.y), max(lo4.z, lo4.w)), max(max(hi4.x, hi4.y), max(hi4.z, hi4.w))) >= tau_s0) {
            const int nv[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
            const int cv[8] = {cl.x, cl.y, cl.z, cl.w, ch.x, ch.y, ch.z, ch.w};
            unsigned hit = 0;
#pragma unroll
            for (int u = 0; u < 8; ++u) hit |= (nv[u] >= tau_s0 && nv[u] - cv[u] < tau_s0) ? (1u << u) : 0u;
#pragma unroll 1
            while (hit) {
              const int u = __ffs(hit) - 1;
              hit &= hit - 1;
              note((u < 4 ? 0 : T / 2) + 4 * i + (u & 3));
            }
          }
        };
        // Fast path: full-size tiles and rows at least as coarse as the score unit (every benchmark shape).  The
        // multiplier is hidden from the compiler so that scale + accumulate stays one IMAD per document.
        const bool fast_a = (T8 % (2 * BM25_THREADS)) == 0 && dk >= 0;
        int mul = 1 << (dk >= 0 ? dk : 0);
        asm volatile("" : "+r"(mul));
        for (int g0 = 0; fast_a && g0 < nd; g0 += 4) {
          const int ng = min(4, nd - g0);
          const uint16_t* const* rp = s_rowp + g0;
#define DRS_SWEEP(NG_)                                                                                 \
  do {                                                                                                 \
    if (cross) bm25_dense_sweep<NG_, true>(dst, T8, rp, doc0, mul, tau_s0, tid, tile_init, note);      \
    else bm25_dense_sweep<NG_, false>(dst, T8, rp, doc0, mul, tau_s0, tid, tile_init, note);           \
  } while (0)
          if (ng == 4) DRS_SWEEP(4);
          else if (ng == 3) DRS_SWEEP(3);
          else if (ng == 2) DRS_SWEEP(2);
          else DRS_SWEEP(1);
#undef DRS_SWEEP
          tile_init = true;   // the same thread owns the same documents in every sweep: no barrier between sweeps
        }
        for (int g0 = 0; !fast_a && g0 < nd; g0 += 4) {   // general path: small tiles, rows finer than the unit
          const int ng = min(4, nd - g0);
          const uint4* r0 = reinterpret_cast<const uint4*>(s_rowp[g0] + doc0);
          const uint4* r1 = ng > 1 ? reinterpret_cast<const uint4*>(s_rowp[g0 + 1] + doc0) : r0;
          const uint4* r2 = ng > 2 ? reinterpret_cast<const uint4*>(s_rowp[g0 + 2] + doc0) : r0;
          const uint4* r3 = ng > 3 ? reinterpret_cast<const uint4*>(s_rowp[g0 + 3] + doc0) : r0;
          const uint4 zu = make_uint4(0, 0, 0, 0);
          const int4 z = make_int4(0, 0, 0, 0);
          for (int i = tid; i < T8; i += 2 * BM25_THREADS) {
            const int i2 = i + BM25_THREADS;
            const bool two = i2 < T8;
            const uint4 a0 = __ldg(r0 + i), a1 = ng > 1 ? __ldg(r1 + i) : zu, a2 = ng > 2 ? __ldg(r2 + i) : zu,
                        a3 = ng > 3 ? __ldg(r3 + i) : zu;
            const uint4 b0 = two ? __ldg(r0 + i2) : zu, b1 = (two && ng > 1) ? __ldg(r1 + i2) : zu,
                        b2 = (two && ng > 2) ? __ldg(r2 + i2) : zu, b3 = (two && ng > 3) ? __ldg(r3 + i2) : zu;
            {
              const uint4 w = make_uint4(a0.x + a1.x + a2.x + a3.x, a0.y + a1.y + a2.y + a3.y,
                                         a0.z + a1.z + a2.z + a3.z, a0.w + a1.w + a2.w + a3.w);
              int4 lo4 = tile_init ? dst[i] : z, hi4 = tile_init ? dst[T8 + i] : z;
              unpack_add(lo4, hi4, w, i);
              dst[i] = lo4;
              dst[T8 + i] = hi4;
            }
            if (two) {
              const uint4 w = make_uint4(b0.x + b1.x + b2.x + b3.x, b0.y + b1.y + b2.y + b3.y,
                                         b0.z + b1.z + b2.z + b3.z, b0.w + b1.w + b2.w + b3.w);
              int4 lo4 = tile_init ? dst[i2] : z, hi4 = tile_init ? dst[T8 + i2] : z;
              unpack_add(lo4, hi4, w, i2);
              dst[i2] = lo4;
              dst[T8 + i2] = hi4;
            }
          }
          tile_init = true;   // the same thread owns the same documents in every sweep: no barrier between sweeps
        }
        if (!tile_init) {
          for (int i = tid; i < T / 4; i += BM25_THREADS) dst[i] = make_int4(0, 0, 0, 0);
          tile_init = true;
        }
        // The barrier that separates the plain stores of pass A from the atomics of pass B sits AFTER the first
        // chunk's loads are issued (they read global memory only), so the wait for the slowest warp and the L2
        // latency of the postings overlap.
        if (nstaged <= 0) __syncthreads();
        mark(1);
        // -- pass B: postings-list terms.  Rounds of 32 staged terms; inside a round the terms' postings are
        // cut into chunks of 128 (aligned groups of four per lane) and the warps take chunks round-robin.
        if (dbg & 2) {
          __syncthreads();
          if (tid == 0) s_note[par ^ 1] = 0;
        }
        for (int r0 = 0; r0 < ((dbg & 2) ? 0 : nstaged); r0 += 32) {
          if (r0 > 0) {   // rounds after the first fetch their ranges here
            tp0 = tp1 = 0;
            if (r0 + lane < nstaged && s_base[r0 + lane] >= 0) {
              const uint32_t* tp = ix.d_tile_ptr + size_t(s_term[r0 + lane]) * (ix.n_tiles + 1) + j;
              tp0 = __ldg(tp);
              tp1 = __ldg(tp + 1);
            }
          }
          long long a4 = 0;      // first posting of my term in this tile, rounded down to a group of four
          int lo = 0, len4 = 0;  // postings [a4+lo, a4+len4) are the term's
          if (tp1 > tp0) {
            const long long a = s_base[r0 + lane] + tp0;
            a4 = a & ~3ll;
            lo = int(a - a4);
            len4 = lo + int(tp1 - tp0);
          }
          const int nch = (len4 + 127) >> 7;
          int incl = nch;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
          }
          const int total = __shfl_sync(0xffffffffu, incl, 31);
          const int cstart = incl - nch;
          // Software pipeline over two register sets (A, B): the loads of chunk c + 8 are in flight while chunk c
          // is added, and the sets swap roles by unrolling instead of by moving registers.
          // lo (2 bits), len4 (<= tile + 3: 14 bits) and the owner's first chunk (<= 32 * 65: 12 bits) travel in one shuffle
          const unsigned pk_mine = unsigned(lo) | (unsigned(len4) << 2) | (unsigned(cstart) << 16);
          auto fetch = [&](int c, uint4& fd, int4& fi, int& foff, int& flo, int& flen) {
            const unsigned m = __ballot_sync(0xffffffffu, c >= cstart && c < incl);   // exactly one lane owns chunk c
            const int src = __ffs(m) - 1;
            const long long sa4 = __shfl_sync(0xffffffffu, a4, src);
            const unsigned pk = __shfl_sync(0xffffffffu, pk_mine, src);
            flo = int(pk & 3u);
            flen = int((pk >> 2) & 0x3fffu);
            foff = ((c - int(pk >> 16)) << 7) + (lane << 2);
            if (foff < flen) {
              fd = __ldg(reinterpret_cast<const uint4*>(ix.d_post_doc + sa4 + foff));
              fi = __ldg(reinterpret_cast<const int4*>(ix.d_post_impact + sa4 + foff));
            }
          };
          // shared address of score[doc - d0u] = sc_rel + 4 * doc (mod 2^32)
          // (kept opaque: recomputing it from the CTA's shared window and the tile number costs more than a register)
          uint32_t sc_rel = uint32_t(__cvta_generic_to_shared(sc)) - 4u * d0u;
          asm volatile("" : "+r"(sc_rel));
          const int taum1 = tau_s0 - 1;
          auto add_chunk = [&](const uint4 d, int4 im, const int off, const int slo, const int slen) {
            if (off >= slen) return;
            if (dbg & 4) {
              if ((d.x ^ d.y ^ d.z ^ d.w ^ unsigned(im.x ^ im.y ^ im.z ^ im.w)) == 0x9e3779b9u) s_cnt = 1;   // keeps the loads
              return;
            }
            im.x >>= sh; im.y >>= sh; im.z >>= sh; im.w >>= sh;
            if (off >= slo && off + 3 < slen) {   // all four postings are the term's: no per-posting predicates
              if (cross) {
                // before < tau <= before + v  <=>  0 <= tau - 1 - before < v   (v >= 0 in crossing mode)
                const int b0 = smem_atom_add_s32(sc_rel + 4u * d.x, im.x), b1 = smem_atom_add_s32(sc_rel + 4u * d.y, im.y);
                const int b2 = smem_atom_add_s32(sc_rel + 4u * d.z, im.z), b3 = smem_atom_add_s32(sc_rel + 4u * d.w, im.w);
                const bool c0 = unsigned(taum1 - b0) < unsigned(im.x), c1 = unsigned(taum1 - b1) < unsigned(im.y);
                const bool c2 = unsigned(taum1 - b2) < unsigned(im.z), c3 = unsigned(taum1 - b3) < unsigned(im.w);
                unsigned hit = (c0 ? 1u : 0u) | (c1 ? 2u : 0u) | (c2 ? 4u : 0u) | (c3 ? 8u : 0u);
                if (hit) {
                  const uint32_t ds[4] = {d.x, d.y, d.z, d.w};
#pragma unroll 1
                  while (hit) {
                    const int u = __ffs(hit) - 1;
                    hit &= hit - 1;
                    note(int((u == 0 ? ds[0] : u == 1 ? ds[1] : u == 2 ? ds[2] : ds[3]) - d0u));
                  }
                }
              } else {
                smem_red_add_s32(sc_rel + 4u * d.x, im.x);
                smem_red_add_s32(sc_rel + 4u * d.y, im.y);
                smem_red_add_s32(sc_rel + 4u * d.z, im.z);
                smem_red_add_s32(sc_rel + 4u * d.w, im.w);
              }
            } else {   // a term's first or last group
              const uint32_t ds[4] = {d.x, d.y, d.z, d.w};
              const int vs[4] = {im.x, im.y, im.z, im.w};
#pragma unroll 1
              for (int u = 0; u < 4; ++u) {
                if (off + u < slo || off + u >= slen) continue;
                const uint32_t doc = u == 0 ? ds[0] : u == 1 ? ds[1] : u == 2 ? ds[2] : ds[3];
                const int v = u == 0 ? vs[0] : u == 1 ? vs[1] : u == 2 ? vs[2] : vs[3];
                if (cross) {
                  const int before = smem_atom_add_s32(sc_rel + 4u * doc, v);
                  if (unsigned(taum1 - before) < unsigned(v)) note(int(doc - d0u));
                } else {
                  smem_red_add_s32(sc_rel + 4u * doc, v);
                }
              }
            }
          };
          constexpr int NW = BM25_THREADS / 32;
          uint4 dA = make_uint4(0, 0, 0, 0), dB = dA;
          int4 iA = make_int4(0, 0, 0, 0), iB = iA;
          int offA = 0, loA = 0, lenA = 0, offB = 0, loB = 0, lenB = 0;
          int c = warp;
          if (warp < total) fetch(c, dA, iA, offA, loA, lenA);
          if (r0 == 0) {
            __syncthreads();   // the tile is initialised: atomics from any warp may follow
            if (tid == 0) s_note[par ^ 1] = 0;
            mark(2);
          }
          if (warp < total) {
            for (;;) {
              const bool more_b = c + NW < total;
              if (more_b) fetch(c + NW, dB, iB, offB, loB, lenB);
              add_chunk(dA, iA, offA, loA, lenA);
              if (!more_b) break;
              c += NW;
              const bool more_a = c + NW < total;
              if (more_a) fetch(c + NW, dA, iA, offA, loA, lenA);
              add_chunk(dB, iB, offB, loB, lenB);
              if (!more_a) break;
              c += NW;
            }
          }
        }
        mark(3);
        __syncthreads();  // all postings of this chunk of terms are in; the staging arrays may be rewritten
        mark(4);
      }
      if (!tile_init) {   // query without terms: every score is 0
        for (int i = tid; i < T / 4; i += BM25_THREADS) reinterpret_cast<int4*>(sc)[i] = make_int4(0, 0, 0, 0);
        __syncthreads();
      }
      if constexpr (EXACT) {
        // ---- escalation: exact re-scoring of every document that can still belong to the top-k' ----
        const int cap2 = BM25_MBUF / 2 - kprime;
        const int fl = fx_low[q];
        auto merge2 = [&]() {   // all threads; sorts list + appended pairs, keeps the best k'
          const int n = kprime + min(s_cnt, cap2);
          const int m = next_pow2(n);
          for (int i = n + tid; i < m; i += BM25_THREADS) pbuf[i] = Pair64{0, -1};
          __syncthreads();
          block_bitonic_desc(pbuf, m, pair_less);
          if (tid == 0) {
            s_cnt = 0;
            s_kth = pbuf[kprime - 1];
          }
          __syncthreads();
        };
        for (int seg = 0; seg < T; seg += BM25_THREADS) {
          const int dl = T - 1 - seg - tid;   // highest document first
          if (dl < tile_n && sc[dl] >= fl) {
            const int64_t doc = doc0 + dl;
            const Pair64 p{ord64(bm25_exact_score(ix, q_terms, t0, t1, doc)), doc};
            if (pair_less(s_kth, p)) pbuf[kprime + atomicAdd(&s_cnt, 1)] = p;
          }
          __syncthreads();
          if (s_cnt > cap2 - BM25_THREADS) merge2();
        }
        if (j == tile_lo) {
          if (s_cnt > 0) merge2();
        }
      } else {
        // ---- scan the tile, highest document first -----------------------------------------
        // Four consecutive documents per thread: one 16-byte shared load and a max test reject almost every
        // group once the threshold is warm.  While this CTA has no threshold yet for the query (cold list,
        // nothing published by other CTAs) only BM25_COLD_STEP documents are scanned before the first
        // merge, so the first sort is small; a warm CTA sweeps the rest of the tile in one go.
        auto scan_range = [&](int from, int count) {   // documents T-1-from downwards
          const unsigned long long tau = s_tau;
          const int tau_s = tau == 0 ? INT_MIN : key_fx(tau);
          for (int g = 4 * tid; g < count && from + g < T; g += 4 * BM25_THREADS) {
            const int hi = T - 1 - from - g;   // my group is documents hi-3 .. hi (hi-3 is 4-aligned)
            const int4 v = *reinterpret_cast<const int4*>(sc + hi - 3);
            if (max(max(v.x, v.y), max(v.z, v.w)) >= tau_s) {
              const int vs[4] = {v.w, v.z, v.y, v.x};   // highest document first
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int dl = hi - u;
                if (vs[u] >= tau_s && dl < tile_n) {
                  const uint64_t key = make_key_fx(vs[u], uint32_t(doc0 + dl));
                  if (key > tau) {
                    const int pos = atomicAdd(&s_cnt, 1);
                    if (pos < cap) mbuf[kk + pos] = key;
                  }
                }
              }
            }
          }
        };
        // Merge (sort list + survivors, tighten the threshold) when enough survivors piled up that a tighter
        // threshold pays for the sort, as soon as a cold list can be filled, when the append area is half full,
        // and at the very end of this CTA's tiles.
        auto maybe_merge = [&](bool cold, bool last) {
          const int cnt = cnt_list;
          if (cnt > 0 && (last || cnt > cap / 2 || cnt >= BM25_MERGE_AT || ((cold || seen < 4ll * T) && cnt >= kk))) {
            if (trace != nullptr && blockIdx.x == 0 && tid == 0) {
              trace[2 + 400 * 7] += 1;
              trace[3 + 400 * 7] += unsigned(kk + cnt);
            }
            merge(q, cnt);
            cnt_list = 0;
          }
        };
        int done = (dbg & 8) ? T : 0;
        if (cross && !(dbg & 8)) {
          const int cnt_now = cnt_list + s_note[par];   // the accumulation ended with a barrier
          if (cnt_now <= cap) {        // (else: more crossings than the append area holds -- scan the ordinary way)
            if (cnt_now > cnt_list) {  // noted documents -> keys, with their FINAL scores
              // (Letting every thread convert the notes of the documents it owns in pass A -- which needs no barrier
              // before the next tile -- was measured SLOWER, 1.335 vs 1.303 ms: the barrier keeps the warps in phase.)
              for (int i = cnt_list + tid; i < cnt_now; i += BM25_THREADS) {
                const int dl = int(mbuf[kk + i]);
                const uint64_t key = make_key_fx(sc[dl], uint32_t(doc0 + dl));
                mbuf[kk + i] = key > tau0 ? key : 0;
              }
              __syncthreads();
            }
            if (trace != nullptr && blockIdx.x == 0 && tid == 0) trace[1 + 400 * 7] += unsigned(cnt_now - cnt_list);
            cnt_list = cnt_now;
            done = T;
            seen += T;
            mark(5);
            maybe_merge(false, j == tile_lo);
          }
        }
        while (done < T) {
          const bool cold = (s_tau == 0);
          // Steps grow geometrically with the number of documents this (slice, query) has seen: a step of 3x the
          // documents seen yields about 3*kk survivors, so every early merge is a small sort.
          const long long grow = 3 * seen;
          long long want = grow > BM25_COLD_STEP ? grow : BM25_COLD_STEP;
          if (want > T - done) want = T - done;
          const int step = cold ? BM25_COLD_STEP : int(want);
          if (tid == 0) s_cnt = cnt_list;
          __syncthreads();
          scan_range(done, step);
          __syncthreads();
          if (s_cnt > cap) {
            // more survivors than the append area holds (some were dropped): redo the range in segments
            // that cannot overflow, merging whenever the area fills up
            __syncthreads();
            if (tid == 0) s_cnt = cnt_list;
            __syncthreads();
            for (int d2 = done; d2 < done + step; d2 += BM25_SEG) {
              if (s_cnt > cap - BM25_SEG) merge(q, s_cnt);
              scan_range(d2, min(BM25_SEG, done + step - d2));
              __syncthreads();
            }
          }
          cnt_list = s_cnt;
          done += step;
          seen += step;
          maybe_merge(cold, (done >= T) && (j == tile_lo));
        }
      }
      mark(6);
    }
    if constexpr (EXACT) {
      for (int i = tid; i < kprime; i += BM25_THREADS) {
        const Pair64 v = pbuf[i];
        e_scores[(size_t(slice) * q_stride + q) * kprime + i] = v.id < 0 ? -INFINITY : unord64(v.s);
        e_ids[(size_t(slice) * q_stride + q) * kprime + i] = v.id;
      }
    } else {
      for (int i = tid; i < kk; i += BM25_THREADS) parts[(size_t(slice) * q_stride + q) * kk + i] = mbuf[i];
    }
    __syncthreads();
  }
}

// Exact fp64 re-scoring of the kk candidates (bm25_exact_score), ranking by (score desc, doc desc), and the
// PROOF that the pre-selection did not lose a member of the true top-k'.
//   Stored impacts are round(impact * 2^s); the kernel adds (stored >> sh) for postings terms and the coarser
//   dense-row value for head terms.  For a document d and a query of nt token occurrences:
//   exact(d) <= fx(d) * 2^(sh-s) + E,   E = sum over the tokens of their rounding + shift error (below).
//   Let L be the exact k'-th best score among the candidates (a lower bound of the true k'-th best, these are
//   real documents).  A document can belong to the true top-k' only if exact(d) >= L, i.e. only if
//   fx(d) >= fx_low := floor((L - E) * 2^(s-sh)) - 1.  Documents outside the candidate list have
//   (fx, doc) below the list's smallest key; if that key's fx is < fx_low none of them qualifies:
//   flag 0 = proven.  Otherwise flag 1: the escalation pass re-scores every document with fx >= fx_low.
//   A list that is not full holds every document: proven.
__global__ void __launch_bounds__(256) bm25_rescore_kernel(drs_bm25_index ix, const int32_t* __restrict__ q_terms,
                                                           const int32_t* __restrict__ q_offsets,
                                                           const uint64_t* __restrict__ cand, int kk, int kprime,
                                                           double* __restrict__ out_scores,
                                                           int64_t* __restrict__ out_ids, int32_t* __restrict__ flags,
                                                           int32_t* __restrict__ fx_low, int force_flag) {
  __shared__ uint64_t skey[256];
  __shared__ int64_t sdoc[256];
  __shared__ double s_kth_score;
  __shared__ int s_valid;
  const int q = blockIdx.x, c = threadIdx.x;
  const int t0 = q_offsets[q], t1 = q_offsets[q + 1];
  // One thread per (candidate, query token): the posting lookup is a ~13-step binary search of dependent loads,
  // so the candidates' terms are looked up side by side (8 tokens per round) and then added by the candidate's
  // own thread in query order -- the same fp64 operations in the same order as bm25_exact_score.
  constexpr int RT = 8;
  __shared__ double s_contrib[256 * RT];   // NaN = the document does not contain the term
  int64_t doc = -1;
  if (c < kk) {
    const uint64_t key = cand[size_t(q) * kk + c];
    if (key != 0) doc = key_row(key);
  }
  sdoc[c] = doc;
  if (c == 0) s_valid = 0;
  __syncthreads();
  double score = 0.0;
  for (int r0 = t0; r0 < t1; r0 += RT) {
    const int nr = min(RT, t1 - r0);
    for (int pair = c; pair < kk * nr; pair += blockDim.x) {
      const int ci = pair / nr, ti = pair - ci * nr;
      const int64_t d = sdoc[ci];
      double contrib = __longlong_as_double(0x7ff8000000000000ll);
      const int t = q_terms[r0 + ti];
      if (d >= 0 && t >= 0 && t < ix.n_terms) {
        const int j = int(d / ix.tile_docs);
        const int64_t base = ix.d_term_offsets[t];
        const uint32_t* tp = ix.d_tile_ptr + size_t(t) * (ix.n_tiles + 1) + j;
        const int64_t end = base + tp[1];
        int64_t lo = base + tp[0], hi = end;  // [lo, hi)
        while (lo < hi) {
          const int64_t mid = (lo + hi) >> 1;
          if (ix.d_post_doc[mid] < uint32_t(d)) lo = mid + 1;
          else hi = mid;
        }
        if (lo < end && ix.d_post_doc[lo] == uint32_t(d)) {
          const double tf = double(ix.d_post_tf[lo]);
          contrib = __dmul_rn(ix.d_idf[t], __ddiv_rn(__dmul_rn(tf, ix.k1_plus_1), __dadd_rn(tf, ix.d_doc_kd[d])));
        }
      }
      s_contrib[ci * RT + ti] = contrib;
    }
    __syncthreads();
    if (c < kk && doc >= 0) {
      for (int ti = 0; ti < nr; ++ti) {
        const double v = s_contrib[c * RT + ti];
        if (v == v) score = __dadd_rn(score, v);
      }
    }
    __syncthreads();
  }
  const uint64_t okey = doc >= 0 ? ord64(score) : 0;
  skey[c] = okey;
  __syncthreads();
  if (c < kk && doc >= 0) {
    int rank = 0;
    for (int o = 0; o < kk; ++o) {
      const int64_t od = sdoc[o];
      if (od < 0) continue;
      const uint64_t ok = skey[o];
      rank += (ok > okey) || (ok == okey && od > doc);
    }
    if (rank < kprime) {
      out_scores[size_t(q) * kprime + rank] = score;
      out_ids[size_t(q) * kprime + rank] = doc;
    }
    if (rank == kprime - 1) s_kth_score = score;
    atomicAdd(&s_valid, 1);
  }
  __syncthreads();
  // pad the tail when fewer than k' candidates exist
  for (int i = s_valid + c; i < kprime; i += blockDim.x) {
    out_scores[size_t(q) * kprime + i] = -INFINITY;
    out_ids[size_t(q) * kprime + i] = -1;
  }
  if (c == 0 && flags != nullptr) {
    int flag = 0, fl = INT_MIN;
    const uint64_t kmin = cand[size_t(q) * kk + kk - 1];   // the merged list is sorted: non-zero <=> full
    if (kmin != 0) {
      const int nt = t1 - t0, sh = bm25_query_shift(nt), s = ix.impact_scale_log2;
      // per token occurrence, in units of 2^-s: a postings term is off by < 2^sh - 0.5 (rounding when stored +
      // the floor of the shift); a head term by its row rounding 2^(dense_shift-1) plus a floor of < 2^sh
      double units = 0.0;
      for (int ti = t0; ti < t1; ++ti) {
        const int t = q_terms[ti];
        if (t < 0 || t >= ix.n_terms) continue;
        const bool dense = ix.d_dense_slot != nullptr && ix.d_dense_slot[t] >= 0;
        units += ldexp(1.0, sh) + (dense ? (ix.dense_shift > 0 ? ldexp(1.0, ix.dense_shift - 1) : 0.0) : -0.5) + 1e-6;
      }
      const double E = units * ldexp(1.0, -s);
      const double L = s_kth_score;
      const double thr = ldexp(L - E - fabs(L) * 1e-12, s - sh);
      fl = thr <= -2147483000.0 ? INT_MIN : (thr >= 2147483000.0 ? INT_MAX : int(floor(thr)) - 1);
      flag = key_fx(kmin) < fl ? 0 : 1;
    }
    flags[q] = force_flag ? 1 : flag;
    fx_low[q] = fl;
  }
}

// Work decomposition: gx slices of tiles_per_cta consecutive tiles; a work item is (slice, query).  Slices are
// short enough (<= 32 tiles) for a slice's rows and postings to stay in L2 while all queries pass over it, and
// numerous enough that a small batch still spreads over the GPU (a single query: one tile per item).
// n_ctas persistent CTAs pull items from a counter.
// Resident CTAs per SM the selection kernel runs with (see bm25_mbuf): DRS_BM25_CTAS=4|5 overrides, read on every
// call so that one process can compare (workspace sizing and launch see the same value).
static int bm25_cps() {
  const char* e = getenv("DRS_BM25_CTAS");
  const int v = e ? atoi(e) : BM25_DEFAULT_CPS;
  return v >= 5 ? 5 : 4;
}

static void bm25_grid(const drs_bm25_index* ix, int n_queries, int& gx, int& n_ctas, int& tiles_per_cta) {
  const int target = bm25_cps() * sm_count();
  const int q = n_queries > 0 ? n_queries : 1;
  const int gx_l2 = (ix->n_tiles + 31) / 32;        // L2 locality
  int gx_par = (8 * target + q - 1) / q;             // ~8 items per resident CTA
  if (gx_par > target) gx_par = target;              // never more slices than resident CTAs: each hands in a list
  gx = gx_l2 > gx_par ? gx_l2 : gx_par;
  if (gx > ix->n_tiles) gx = ix->n_tiles;
  if (gx < 1) gx = 1;
  tiles_per_cta = (ix->n_tiles + gx - 1) / gx;
  if (tiles_per_cta < 1) tiles_per_cta = 1;
  gx = (ix->n_tiles + tiles_per_cta - 1) / tiles_per_cta;
  if (gx < 1) gx = 1;
  const long long items = (long long)gx * q;
  n_ctas = items < target ? int(items) : target;
}

}  // namespace drs

using namespace drs;

// DRS_BM25_DEBUG: timing experiments (see the kernel); read on every call so that one process can compare variants
static int bm25_debug_flags() {
  const char* e = getenv("DRS_BM25_DEBUG");
  return e ? atoi(e) : 0;
}

// workspace layout (bytes): parts | cand | gtau | merge scratch | flags | fx_low | e_scores | e_ids
struct Bm25Ws {
  size_t parts, cand, gtau, mscratch, flags, fxlow, escores, eids, total;
};
static Bm25Ws bm25_ws(int n_queries, int gx, int kk, int kprime) {
  Bm25Ws w;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    const size_t at = o;
    o += (bytes + 255) & ~size_t(255);
    return at;
  };
  w.parts = take(size_t(gx) * n_queries * kk * 8);
  w.cand = take(size_t(n_queries) * kk * 8);
  w.gtau = take(size_t(n_queries) * 8 + 2 * 4);   // thresholds, then the two passes' work-item counters
  w.mscratch = take(merge_scratch_keys(gx, n_queries, kk) * 8);
  w.flags = take(size_t(n_queries) * 4);
  w.fxlow = take(size_t(n_queries) * 4);
  w.escores = take(size_t(gx) * n_queries * kprime * 8);
  w.eids = take(size_t(gx) * n_queries * kprime * 8);
  w.total = o + 256;
  return w;
}

extern "C" size_t drs_bm25_workspace_bytes(int n_queries, const drs_bm25_index* index, int kprime) {
  if (!index || n_queries < 0 || kprime <= 0) return 0;
  int gx, gy, tpc;
  bm25_grid(index, n_queries > 0 ? n_queries : 1, gx, gy, tpc);
  return bm25_ws(n_queries, gx, cand_len(kprime), kprime).total;
}

extern "C" int drs_bm25_topk_ex(const drs_bm25_index* index, const int32_t* d_q_terms, const int32_t* d_q_offsets,
                                int n_queries, int kprime, int mode, double* d_out_scores, int64_t* d_out_ids,
                                int32_t* d_out_flags, void* d_workspace, size_t workspace_bytes, void* stream) {
  DRS_REQUIRE(index && d_q_offsets && d_out_scores && d_out_ids && d_workspace, DRS_ERR_ARG, "bm25: null pointer");
  DRS_REQUIRE(n_queries >= 0 && kprime > 0, DRS_ERR_ARG, "bm25: bad sizes");
  DRS_REQUIRE(mode >= 0 && mode <= 2, DRS_ERR_ARG, "bm25: mode must be 0 (auto), 1 (no escalation) or 2 (escalate all)");
  DRS_REQUIRE(kprime <= DRS_MAX_KPRIME, DRS_ERR_UNSUPPORTED, "bm25: k' exceeds DRS_MAX_KPRIME");
  const drs_bm25_index& ix = *index;
  DRS_REQUIRE(ix.n_docs > 0 && ix.n_docs < (int64_t(1) << 32), DRS_ERR_UNSUPPORTED, "bm25: n_docs out of range");
  DRS_REQUIRE(ix.tile_docs >= 256 && ix.tile_docs <= 8192 && (ix.tile_docs & (ix.tile_docs - 1)) == 0,
              DRS_ERR_UNSUPPORTED, "bm25: tile_docs must be a power of two in [256, 8192]");
  DRS_REQUIRE(ix.n_tiles == int((ix.n_docs + ix.tile_docs - 1) / ix.tile_docs), DRS_ERR_ARG, "bm25: n_tiles mismatch");
  DRS_REQUIRE(ix.impact_scale_log2 >= -64 && ix.impact_scale_log2 <= 64, DRS_ERR_ARG, "bm25: impact_scale_log2 out of range");
  DRS_REQUIRE(workspace_bytes >= drs_bm25_workspace_bytes(n_queries, index, kprime), DRS_ERR_WORKSPACE,
              "bm25: workspace too small");
  if (n_queries == 0) return DRS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int kk = cand_len(kprime);
  int gx, gy, tpc;
  bm25_grid(index, n_queries, gx, gy, tpc);
  const Bm25Ws w = bm25_ws(n_queries, gx, kk, kprime);
  uint8_t* base = reinterpret_cast<uint8_t*>(d_workspace);
  uint64_t* parts = reinterpret_cast<uint64_t*>(base + w.parts);
  uint64_t* cand = reinterpret_cast<uint64_t*>(base + w.cand);
  unsigned long long* gtau = reinterpret_cast<unsigned long long*>(base + w.gtau);
  uint64_t* mscratch = merge_scratch_keys(gx, n_queries, kk) ? reinterpret_cast<uint64_t*>(base + w.mscratch) : nullptr;
  int32_t* flags = reinterpret_cast<int32_t*>(base + w.flags);
  int32_t* fxlow = reinterpret_cast<int32_t*>(base + w.fxlow);
  double* escores = reinterpret_cast<double*>(base + w.escores);
  int64_t* eids = reinterpret_cast<int64_t*>(base + w.eids);
  int* next_query = reinterpret_cast<int*>(gtau + n_queries);   // [2]: one counter per pass
  DRS_CUDA_CHECK(cudaMemsetAsync(gtau, 0, size_t(n_queries) * 8 + 2 * 4, st));
  const int cps = bm25_cps();
  const size_t smem = size_t(bm25_mbuf(cps)) * 8 + size_t(ix.tile_docs) * 4;
  if (cps == 5) {
    DRS_ENSURE_SMEM((bm25_select_kernel<false, 5>), smem);
    DRS_ENSURE_SMEM((bm25_select_kernel<true, 5>), smem);
  } else {
    DRS_ENSURE_SMEM((bm25_select_kernel<false, 4>), smem);
    DRS_ENSURE_SMEM((bm25_select_kernel<true, 4>), smem);
  }
  // DRS_BM25_TRACE=1 (debugging aid, synchronises): phase timeline of CTA 0 on stderr
  static const bool want_trace = getenv("DRS_BM25_TRACE") != nullptr;
  static unsigned long long* d_trace = nullptr;
  if (want_trace) {
    if (!d_trace) cudaMalloc(&d_trace, (1 + 400 * 7 + 4) * sizeof(unsigned long long));
    cudaMemsetAsync(d_trace, 0, (1 + 400 * 7 + 4) * sizeof(unsigned long long), st);
  }
  {
    ProfScope ps(DRS_PROF_BM25_SELECT, st);
    if (cps == 5)
      bm25_select_kernel<false, 5><<<gy, BM25_THREADS, smem, st>>>(
          ix, d_q_terms, d_q_offsets, n_queries, kk, tpc, parts, n_queries, gtau, nullptr, nullptr, kprime, nullptr,
          nullptr, next_query, want_trace ? d_trace : nullptr, bm25_debug_flags());
    else
      bm25_select_kernel<false, 4><<<gy, BM25_THREADS, smem, st>>>(
          ix, d_q_terms, d_q_offsets, n_queries, kk, tpc, parts, n_queries, gtau, nullptr, nullptr, kprime, nullptr,
          nullptr, next_query, want_trace ? d_trace : nullptr, bm25_debug_flags());
    count_launches(1);
  }
  if (want_trace) {
    static int dumps = 0;
    std::vector<unsigned long long> h(1 + 400 * 7 + 4);
    cudaStreamSynchronize(st);
    cudaMemcpy(h.data(), d_trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    if (dumps++ == 3) {   // a warm call
      const int n = int(h[0] < 400 ? h[0] : 400);
      double ph[6] = {0, 0, 0, 0, 0, 0}, gap = 0;
      int crossing = 0;
      for (int i = 0; i < n; ++i) {
        const unsigned long long* t = &h[1 + i * 7];
        const bool cr = t[5] != 0;                  // mark 5 is only reached in crossing mode
        crossing += cr;
        for (int k = 0; k < 4; ++k) ph[k] += double(t[k + 1] - t[k]);
        if (cr) {
          ph[4] += double(t[5] - t[4]);
          ph[5] += double(t[6] - t[5]);
        } else {
          ph[5] += double(t[6] - t[4]);
        }
        if (i) gap += double(t[0] - h[1 + (i - 1) * 7 + 6]);
      }
      fprintf(stderr, "[bm25 trace] noted documents %llu, merges %llu (mean size %.0f keys)\n", h[1 + 400 * 7], h[2 + 400 * 7],
              h[2 + 400 * 7] ? double(h[3 + 400 * 7]) / double(h[2 + 400 * 7]) : 0.0);
      fprintf(stderr, "[bm25 trace] %d tiles of CTA 0 (%d in crossing mode), cycles per tile: passA %.0f | barrier1(+fetch) "
                      "%.0f | passB %.0f | barrier2 %.0f | noted->keys %.0f | merge/scan %.0f | between tiles %.0f\n", n,
              crossing, ph[0] / n, ph[1] / n, ph[2] / n, ph[3] / n, ph[4] / n, ph[5] / n, gap / (n > 1 ? n - 1 : 1));
    }
  }
  DRS_CUDA_CHECK(cudaGetLastError());
  ProfScope pt(DRS_PROF_BM25_TAIL, st);
  int rc = launch_merge_keys(parts, gx, n_queries, n_queries, kk, cand, nullptr, mscratch, st);
  if (rc != DRS_OK) return rc;
  bm25_rescore_kernel<<<n_queries, 256, 0, st>>>(ix, d_q_terms, d_q_offsets, cand, kk, kprime, d_out_scores, d_out_ids,
                                                 flags, fxlow, mode == 2 ? 1 : 0);
  count_launches(1);
  DRS_CUDA_CHECK(cudaGetLastError());
  if (mode != 1) {
    // escalation, predicated on the device-side flags (no host round trip): CTAs of proven queries exit at once
    if (cps == 5)
      bm25_select_kernel<true, 5><<<gy, BM25_THREADS, smem, st>>>(
          ix, d_q_terms, d_q_offsets, n_queries, kk, tpc, nullptr, n_queries, nullptr, flags, fxlow, kprime, escores, eids,
          next_query + 1, nullptr, 0);
    else
      bm25_select_kernel<true, 4><<<gy, BM25_THREADS, smem, st>>>(
          ix, d_q_terms, d_q_offsets, n_queries, kk, tpc, nullptr, n_queries, nullptr, flags, fxlow, kprime, escores, eids,
          next_query + 1, nullptr, 0);
    count_launches(1);
    DRS_CUDA_CHECK(cudaGetLastError());
    rc = launch_merge_pairs_f64(escores, eids, gx, n_queries, kprime, kprime, d_out_scores, d_out_ids, flags, st);
    if (rc != DRS_OK) return rc;
  }
  if (d_out_flags) DRS_CUDA_CHECK(cudaMemcpyAsync(d_out_flags, flags, size_t(n_queries) * 4, cudaMemcpyDeviceToDevice, st));
  return DRS_OK;
}

extern "C" int drs_bm25_topk(const drs_bm25_index* index, const int32_t* d_q_terms, const int32_t* d_q_offsets,
                             int n_queries, int kprime, double* d_out_scores, int64_t* d_out_ids,
                             void* d_workspace, size_t workspace_bytes, void* stream) {
  return drs_bm25_topk_ex(index, d_q_terms, d_q_offsets, n_queries, kprime, 0, d_out_scores, d_out_ids, nullptr,
                          d_workspace, workspace_bytes, stream);
}
