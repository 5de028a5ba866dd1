// Dense cosine candidate selection on the 5th-gen tensor cores (sm_100a): a persistent,
// warp-specialised bf16 GEMM  scores[q, row] = queries[q, :] . corpus[row, :]  whose epilogue
// never writes the scores -- it keeps a per-query running top-kk instead.
//
//   A operand  = 128 queries x D, converted fp32 -> bf16 and parked in TENSOR MEMORY for the
//                whole work item (D/2 <= 384 of the 512 TMEM columns; tcgen05.st, once per item).
//                Shared memory is therefore free to be a deep ring for the streamed operand.
//   B operand  = corpus rows, K-major, TMA-loaded (cp.async.bulk.tensor, 128B swizzle) as
//                64-row x 128-element stages (16 KB) into an mbarrier ring.
//   D          = 128 x 64 fp32 accumulator tile in the remaining 128 TMEM columns, double
//                buffered: the epilogue of tile t overlaps the MMAs of tile t+1.
//   epilogue   = 4 warps; after tcgen05.ld thread L holds query L's 64 scores for the tile, scales
//                them by inv_norm[row] and compares against its private threshold (the kk-th best so
//                far); the rare survivors replace the minimum of a private list in shared memory.
//
// Roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane; optionally warps 3 and 2 as well: ring stages
// dealt round-robin, every MMA accumulates, the epilogue zeroes the accumulator it drained), warp 2 = TMEM allocator,
// warps 4..11 = A loader + two epilogue groups.  Work item = (block of 128 queries, slab of corpus tiles);
// items are dealt round-robin so CTAs that run concurrently share corpus tiles through L2.
// Output: one partial list of kk keys per (slab, query); merged and exactly re-scored downstream.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "umma_ptx.cuh"

namespace drs {
namespace umma {

constexpr int TILE_N = 64;                        // corpus rows per accumulator tile
constexpr int MIN_STAGE_K = 128;                  // K elements per smem stage: 128 (2 TMA boxes) or 256 (4 boxes)
constexpr int BOX_K = 64;                         // 64 bf16 = 128 bytes = one swizzle row
constexpr int BOX_BYTES = TILE_N * BOX_K * 2;     // 8192
constexpr int MIN_STAGE_BYTES = 2 * BOX_BYTES;    // 16384
constexpr int MAX_D = 768;
constexpr int TMEM_COLS = 512;
constexpr int ACC_COL0 = MAX_D / 2;               // accumulators live above the A operand
constexpr int THREADS = 384;                       // 4 service warps + 2 epilogue groups of 4 warps
constexpr int MIN_STAGES = 4;
constexpr size_t SMEM_LIMIT = 227 * 1024;
// instruction descriptor: D=f32, A=B=bf16, both operands K-major, N=64, M=128.  (A and B must share
// the format: the hardware rejects fp16 x bf16 inside one kind::f16 MMA as an illegal instruction,
// so the query cannot be carried at fp16 precision while the corpus is bf16.)
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(TILE_N >> 3) << 17) |
                           (uint32_t(BLOCK_M >> 4) << 24);

struct Params {
  const float* queries;
  const float* inv_norm;
  const uint8_t* mask;
  uint64_t* parts;
  uint32_t* gtau;  // [q_stride] zero-initialised shared thresholds
  uint32_t* gslab; // [n_slabs][q_stride] zero-initialised: each slab's m_rank-th best inserted score
  int64_t n_rows;
  int n_queries, nqb, dim, kk, q_stride, n_slabs, n_tiles, n_stages;
  uint32_t debug;
  int m_rank;      // ceil(kk / (n_groups*n_slabs)) if <= 4 and all slabs run concurrently, else 0 (pooling off)
  int n_groups;    // epilogue warp groups (candidate heaps per query per CTA): 2, or 1 for large kk
  int n_issuers;   // warps issuing MMAs (1..3): ring stages are dealt round-robin among them
};

// NBOX = 64-wide TMA boxes per ring stage.  A stage is the unit of the full/empty handshake: with 4 boxes
// (K = 256 per stage) the MMA warp pays one barrier round trip per 16 MMAs instead of per 8.
// PAIR = two CTAs of a cluster (one TPC) share every MMA (cta_group::2): each holds ITS OWN block of 128 queries in
// its tensor memory and HALF of every 64-row corpus tile in its shared memory; the leader CTA issues M = 256 MMAs
// whose accumulator rows 0..127 land in the leader's TMEM and 128..255 in the peer's.  The fixed per-MMA cost is
// paid once for two query blocks and every shared-memory byte feeds twice the math.
// MULTI = several MMA-issuing warps (p.n_issuers = 2 or 3); a template parameter so that the single-issuer loop
// keeps its handful of instructions per MMA.
template <int NBOX, bool PAIR, bool MULTI>
__global__ void __launch_bounds__(THREADS, 1) dense_umma_kernel(const __grid_constant__ CUtensorMap tmap, Params p) {
  constexpr int STAGE_K = NBOX * BOX_K;
  constexpr int BOXB = PAIR ? BOX_BYTES / 2 : BOX_BYTES;   // a box is 64 K-elements x 64 (32 in a pair) rows
  constexpr int STAGE_BYTES = NBOX * BOXB;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;     // 0 = leader
  const int cta = PAIR ? int(blockIdx.x >> 1) : int(blockIdx.x);       // work-item stream this CTA (pair) walks
  const int n_ctas = PAIR ? int(gridDim.x >> 1) : int(gridDim.x);
  const int nqbx = PAIR ? (p.nqb + 1) / 2 : p.nqb;         // query blocks (pairs of blocks) per slab
  extern __shared__ uint8_t smem_dyn[];
  // carve: [ring | lists | barriers | tmem slot]; the swizzled ring needs 1024-byte alignment
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint64_t* lists = reinterpret_cast<uint64_t*>(ring + size_t(p.n_stages) * STAGE_BYTES);
  uint64_t* bars = lists + size_t(p.n_groups) * p.kk * BLOCK_M;
  uint64_t* full = bars;
  uint64_t* empty = bars + p.n_stages;
  uint64_t* tfull = bars + 2 * p.n_stages;
  uint64_t* tempty = tfull + 2;
  uint64_t* a_ready = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 1);
  float* norm_strips = reinterpret_cast<float*>(                  // [4 warps][2][TILE_N], 16-byte aligned
      (reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = nqbx * p.n_slabs;
  const int kst = p.dim / STAGE_K;  // smem stages per accumulator tile

  if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
  if (warp == 1 && lane == 0) {
    // In a pair the leader's full / tempty / a_ready barriers collect the arrivals of BOTH CTAs; empty and tfull
    // are signalled in each CTA by the multicast commit.
    for (int s = 0; s < p.n_stages; ++s) {
      mbar_init(full + s, PAIR ? 2 : 1);
      mbar_init(empty + s, 1);
    }
    mbar_init(tfull + 0, MULTI ? p.n_issuers : 1);   // every issuing warp commits the accumulator once per tile
    mbar_init(tfull + 1, MULTI ? p.n_issuers : 1);
    mbar_init(tempty + 0, PAIR ? 8 : 4);
    mbar_init(tempty + 1, PAIR ? 8 : 4);
    mbar_init(a_ready, (PAIR ? 8 : 4) * p.n_groups);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    if constexpr (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(uint32_t(TMEM_COLS))
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(uint32_t(TMEM_COLS))
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();   // the peer's barriers are initialised before anyone arrives on them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  auto slab_tiles = [&](int slab, int& t0, int& t1) {
    t0 = int((int64_t(slab) * p.n_tiles) / p.n_slabs);
    t1 = int((int64_t(slab + 1) * p.n_tiles) / p.n_slabs);
  };

  if (warp == 0) {
    // ===================================================================== TMA producer
    uint32_t s = 0, ph = 0;
    for (int item = cta; item < n_items; item += n_ctas) {
      int t0, t1;
      slab_tiles(item / nqbx, t0, t1);
      for (int t = t0; t < t1; ++t) {
        for (int ks = 0; ks < kst; ++ks) {
          mbar_wait(empty + s, ph ^ 1, 1);
          if (elect_one()) {
            // debug bit 4 (timing experiment only, results are garbage): fetch half of every stage
            const int nb = (p.debug & 4) ? NBOX / 2 : NBOX;
            uint8_t* dst = ring + size_t(s) * STAGE_BYTES;
            if constexpr (PAIR) {
              // my half of the tile (rows 32*rank ..) goes to MY ring; the bytes are counted by the LEADER's barrier
              const uint32_t lbar = mapa_u32(smem_u32(full + s), 0);
              mbar_arrive_expect_tx_remote(lbar, uint32_t(nb) * BOXB);
#pragma unroll
              for (int b = 0; b < NBOX; ++b)
                if (b < nb)
                  tma_load_2d_pair(&tmap, lbar, dst + b * BOXB, ks * STAGE_K + b * BOX_K, t * TILE_N + int(rank) * (TILE_N / 2));
            } else {
              mbar_arrive_expect_tx(full + s, uint32_t(nb) * BOXB);
#pragma unroll
              for (int b = 0; b < NBOX; ++b)
                if (b < nb) tma_load_2d(&tmap, full + s, dst + b * BOXB, ks * STAGE_K + b * BOX_K, t * TILE_N);
            }
          }
          __syncwarp();
          if (++s == uint32_t(p.n_stages)) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (rank == 0 && (warp == 1 || (MULTI && (warp == 3 || (warp == 2 && p.n_issuers > 2))))) {
    // ===================================================================== MMA issuer (the leader CTA of a pair)
    // The WHOLE warp walks the loops (so every address below is warp-uniform and lives in uniform
    // registers); one elected lane issues the tcgen05 instructions.  A tile's 48 MMAs take 32 tensor
    // cycles each, so the issue path must stay within a handful of instructions per MMA.
    uint32_t st = 0, it = 0, item_idx = 0;
    uint32_t s = 0, ph = 0;   // ring slot / phase, stepped without divisions
    // Several issuing warps (1, 3, 2): a tcgen05.mma with A in tensor memory costs ~26 cycles on top of its 32
    // tensor cycles when one thread issues them back to back; issued from different warps the fixed parts overlap.
    // The ring stages are dealt round-robin.  Because the stages of one tile are then issued in no particular
    // order, no MMA may be "the first" (overwrite): every MMA accumulates, and the epilogue zeroes an accumulator
    // after reading it.
    constexpr bool dual = MULTI;
    const uint32_t me = warp == 1 ? 0u : (warp == 3 ? 1u : 2u);
    const uint32_t n_iss = MULTI ? uint32_t(p.n_issuers) : 1u;
    constexpr uint32_t IDESC_PAIR = (IDESC & ~(0x1fu << 24)) | (uint32_t((2 * BLOCK_M) >> 4) << 24);   // M = 256
    for (int item = cta; item < n_items; item += n_ctas, ++item_idx) {
      int t0, t1;
      slab_tiles(item / nqbx, t0, t1);
      if constexpr (PAIR) mbar_wait_cluster(a_ready, item_idx & 1, 2);
      else mbar_wait(a_ready, item_idx & 1, 2);
      tc_fence_after();
      for (int t = t0; t < t1; ++t, ++it) {
        const uint32_t as = it & 1, aph = (it >> 1) & 1;
        if constexpr (PAIR) mbar_wait_cluster(tempty + as, aph ^ 1, 3);
        else mbar_wait(tempty + as, aph ^ 1, 3);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ACC_COL0 + as * TILE_N;
        for (int ks = 0; ks < kst; ++ks, ++st) {
          const bool mine = !dual || (st % n_iss) == me;
          if (mine) {
            if constexpr (PAIR) mbar_wait_cluster(full + s, ph, 4);
            else mbar_wait(full + s, ph, 4);
            tc_fence_after();
          }
          if (mine && elect_one()) {
            const uint64_t desc0 = make_b_desc(smem_u32(ring + size_t(s) * STAGE_BYTES));
            const uint32_t a0 = tmem_base + uint32_t(ks) * (STAGE_K / 2);
#pragma unroll
            for (int k16 = 0; k16 < STAGE_K / 16; ++k16) {
              // +32 bytes per K=16 step inside a 128-byte swizzle row, +8 KB for the second box
              const uint64_t b_desc = desc0 + uint64_t((k16 >> 2) * (BOXB >> 4) + (k16 & 3) * 2);
              // debug bit 8 (timing experiment only): issue every other MMA
              // debug bits 64 / 128 (timing experiments): N = 32 / N = 128 instead of 64 per MMA
              const uint32_t idesc = (p.debug & 64) ? ((IDESC & ~(0x3fu << 17)) | (4u << 17))
                                                    : ((p.debug & 128) ? ((IDESC & ~(0x3fu << 17)) | (16u << 17)) : IDESC);
              // debug bit 256 (timing experiment): consecutive MMAs alternate between the two accumulator stages
              const uint32_t dd = (p.debug & 128) ? tmem_base + ACC_COL0
                                                  : ((p.debug & 256) ? tmem_base + ACC_COL0 + (k16 & 1) * TILE_N : d_tmem);
              if constexpr (PAIR) {
                umma_ts_pair(d_tmem, a0 + k16 * 8, b_desc, IDESC_PAIR, (ks | k16) != 0 ? 1u : 0u);
              } else {
                if (!(p.debug & 8) || (k16 & 1) == 0)
                  umma_ts(dd, a0 + k16 * 8, b_desc, idesc, (dual || (ks | k16) != 0) ? 1u : 0u);
              }
            }
            if constexpr (PAIR) {
              tc_commit_pair(empty + s);                      // frees the stage in both CTAs
              if (ks == kst - 1) tc_commit_pair(tfull + as);  // both CTAs' halves of the accumulator are complete
            } else {
              tc_commit(empty + s);                      // the stage is free once these MMAs have read it
              if (ks == kst - 1 && !dual) tc_commit(tfull + as);  // accumulator tile complete
            }
          }
          __syncwarp();
          if (++s == uint32_t(p.n_stages)) {
            s = 0;
            ph ^= 1;
          }
        }
        if (dual) {
          if (elect_one()) tc_commit(tfull + as);
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4 && warp < 4 + 4 * p.n_groups) {
    // ===================================================================== A loader + epilogue
    // Two epilogue groups of four warps.  Group g owns accumulator stage g and therefore every other
    // tile, so each group has TWO tile-times of tensor work to drain one tile; each keeps its own
    // candidate heap per query (a "virtual slab" of the corpus: the tiles of its parity).
    const int ew = warp & 3;                 // the TMEM sub-partition (lanes 32*ew ..) this warp may touch
    const int grp = (warp - 4) >> 2;         // 0 or 1
    const int L = ew * 32 + lane;            // TMEM lane == query within the block
    const uint32_t lane_addr = uint32_t(ew * 32) << 16;
    const int ng = p.n_groups;               // 2 normally; 1 when two heaps per query would not fit in smem
    const uint32_t mylist = smem_u32(lists) + uint32_t(grp * p.kk * BLOCK_M + L) * 8;  // my heap's slot 0
    float* norm_strip = norm_strips + (warp - 4) * 2 * TILE_N;
    uint32_t it = 0;                         // tiles issued so far (both groups), mirrors the MMA warp
    const uint32_t a_ready_leader = PAIR ? mapa_u32(smem_u32(a_ready), 0) : 0u;
    const uint32_t tempty_leader = PAIR ? mapa_u32(smem_u32(tempty), 0) : 0u;
    // several issuing warps => every MMA accumulates: the accumulators start at zero and are zeroed again by the
    // epilogue that drained them (group g owns stage g; a single group owns both)
    constexpr bool zacc = MULTI;
    if (zacc) {
      uint32_t z[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) z[j] = 0u;
      for (int stg = (ng == 2 ? grp : 0); stg < (ng == 2 ? grp + 1 : 2); ++stg) {
        tmem_st32(tmem_base + lane_addr + ACC_COL0 + stg * TILE_N, z);
        tmem_st32(tmem_base + lane_addr + ACC_COL0 + stg * TILE_N + 32, z);
      }
      // (the wait::st + fence before the first a_ready arrival below covers these stores)
    }
    for (int item = cta; item < n_items; item += n_ctas) {
      const int qb = PAIR ? 2 * (item % nqbx) + int(rank) : item % nqbx, slab = item / nqbx;
      int t0, t1;
      slab_tiles(slab, t0, t1);
      // ---- queries -> bf16 -> TMEM (A operand), 64 elements (32 columns) at a time
      const int q = qb * BLOCK_M + L;
      const bool live = q < p.n_queries;
      const float* qrow = p.queries + size_t(live ? q : 0) * p.dim;
      for (int c0 = grp * 32; c0 < p.dim / 2; c0 += 32 * ng) {   // the groups interleave 32-column chunks
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
          if (live) f = __ldg(reinterpret_cast<const float4*>(qrow + (c0 + j) * 2));
          const __nv_bfloat162 lo = __floats2bfloat162_rn(f.x, f.y);
          const __nv_bfloat162 hi = __floats2bfloat162_rn(f.z, f.w);
          v[j] = *reinterpret_cast<const uint32_t*>(&lo);
          v[j + 1] = *reinterpret_cast<const uint32_t*>(&hi);
        }
        tmem_st32(tmem_base + lane_addr + c0, v);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_remote(a_ready_leader);
        else mbar_arrive(a_ready);
      }
      // ---- private top-kk min-heap; its root is the kk-th best candidate once the list is full
      for (int c = 0; c < p.kk; ++c) sts_u64(mylist + c * LSTRIDE, 0);
      uint64_t root = 0;
      float tau_loc = -INFINITY;
      // gtau[q]: a lower bound on query q's kk-th best score over the WHOLE corpus, published by
      // every slab working on q (ordered-u32 image, atomicMax).  A row scoring below it cannot be
      // in the global top-kk, so slabs warm each other's thresholds.  Padding lanes never insert.
      uint32_t* gt = p.gtau + (live ? q : 0);
      uint32_t g_seen = live ? *reinterpret_cast<volatile uint32_t*>(gt) : 0u;
      // Pooled threshold.  All n_slabs slabs of a query block run CONCURRENTLY (one item per CTA), each
      // publishes the m-th best score it has inserted so far (m = ceil(kk / n_slabs) <= 4) in
      // gslab[slab][q].  If every slab has m rows scoring >= x, the corpus has >= kk such rows, so
      // x = min over slabs is a valid lower bound on the global kk-th best -- and a far tighter one than
      // any single slab's own kk-th best while all the lists are still warming up.
      const int vslab = slab * ng + grp;
      volatile uint32_t* gs = p.gslab + size_t(vslab) * p.q_stride + (live ? q : 0);
      float tm0 = -INFINITY, tm1 = -INFINITY, tm2 = -INFINITY, tm3 = -INFINITY;
      float tau_pool = -INFINITY;
      const int n_vslabs = ng * p.n_slabs;
      const int pool_every = n_vslabs > 32 ? n_vslabs / 4 : 8;
      const bool no_insert = (p.debug & 2) != 0;
      // inverse norms of the tile's 64 rows: lane l fetches rows 2l, 2l+1 (prefetched one tile ahead),
      // parks them in this warp's shared-memory strip, and every lane reads all 64 by broadcast
      auto load_norms = [&](int t, float& n0, float& n1) {
        const int64_t r = int64_t(t) * TILE_N + 2 * lane;
        n0 = (t < t1 && r < p.n_rows) ? __ldg(p.inv_norm + r) : 0.f;
        n1 = (t < t1 && r + 1 < p.n_rows) ? __ldg(p.inv_norm + r + 1) : 0.f;
      };
      float nn0, nn1;
      const int t_first = (ng == 1) ? t0 : t0 + (((it & 1u) == uint32_t(grp)) ? 0 : 1);   // my first tile of this item
      load_norms(t_first, nn0, nn1);
      int nt = 0;
      for (int t = t_first; t < t1; t += ng, ++nt) {
        const uint32_t itt = it + uint32_t(t - t0);      // this tile's global sequence number (parity == grp)
        float* nbuf = norm_strip + (nt & 1) * TILE_N;
        reinterpret_cast<float2*>(nbuf)[lane] = make_float2(nn0, nn1);
        __syncwarp();
        load_norms(t + ng, nn0, nn1);  // prefetch for my next tile
        // refreshed before each of the first tiles (thresholds move fastest while the heaps fill), then periodically
        if (live && p.m_rank > 0 && (nt < 12 || (nt % pool_every) == pool_every - 1)) {
          uint32_t mn = 0xffffffffu;
          const volatile uint32_t* col = p.gslab + q;
          for (int sl = 0; sl < n_vslabs; ++sl) mn = min(mn, col[size_t(sl) * p.q_stride]);
          if (mn != 0u) tau_pool = fmaxf(tau_pool, unord32(mn));   // 0 = some slab has not reported yet
        }
        float tau = live ? fmaxf(fmaxf(tau_loc, tau_pool), g_seen ? unord32(g_seen) : -INFINITY) : INFINITY;
        if (live) g_seen = *reinterpret_cast<volatile uint32_t*>(gt);  // lands while this tile is processed
        const uint32_t as = itt & 1u, aph = (itt >> 1) & 1;   // with two groups as == grp
        mbar_wait(tfull + as, aph, 5);
        tc_fence_after();
        uint32_t acc0[32], acc1[32];
        const uint32_t taddr = tmem_base + lane_addr + ACC_COL0 + as * TILE_N;
        if (!(p.debug & 32)) {   // debug bit 32 (timing experiment): skip the accumulator read
          tmem_ld32(taddr, acc0);
          tmem_ld32(taddr + 32, acc1);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        // ---- pass 1, branch-free: which of my 64 scores reach the threshold?
        uint32_t m0 = 0, m1 = 0;
        const float4* nb4 = reinterpret_cast<const float4*>(nbuf);
        if (!(p.debug & 16)) {   // debug bit 16 (timing experiment): skip pass 1
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 n = nb4[g];
          m0 |= (__uint_as_float(acc0[4 * g + 0]) * n.x >= tau) ? (1u << (4 * g + 0)) : 0u;
          m0 |= (__uint_as_float(acc0[4 * g + 1]) * n.y >= tau) ? (1u << (4 * g + 1)) : 0u;
          m0 |= (__uint_as_float(acc0[4 * g + 2]) * n.z >= tau) ? (1u << (4 * g + 2)) : 0u;
          m0 |= (__uint_as_float(acc0[4 * g + 3]) * n.w >= tau) ? (1u << (4 * g + 3)) : 0u;
        }
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 n = nb4[8 + g];
          m1 |= (__uint_as_float(acc1[4 * g + 0]) * n.x >= tau) ? (1u << (4 * g + 0)) : 0u;
          m1 |= (__uint_as_float(acc1[4 * g + 1]) * n.y >= tau) ? (1u << (4 * g + 1)) : 0u;
          m1 |= (__uint_as_float(acc1[4 * g + 2]) * n.z >= tau) ? (1u << (4 * g + 2)) : 0u;
          m1 |= (__uint_as_float(acc1[4 * g + 3]) * n.w >= tau) ? (1u << (4 * g + 3)) : 0u;
        }
        }
        if (no_insert) m0 = m1 = 0;
        // ---- pass 2, rare and rolled.  2a: for every column some lane flagged, all lanes re-read that one
        // accumulator column from TMEM (the load is warp-collective); a flagged lane only STASHES its score
        // (four register slots).  The accumulator stage is then handed back to the MMA warp, and the heap
        // work (2b) -- the expensive part -- runs off the tensor pipe's critical path.
        auto insert = [&](float s, int j) {
          const int64_t row = int64_t(t) * TILE_N + j;
          if (s >= tau && row < p.n_rows && (p.mask == nullptr || p.mask[row] != 0)) {
            const uint64_t key = make_key(s, uint32_t(row));
            if (key > root) {
              root = heap_replace_root(mylist, p.kk, key);
              if (p.m_rank > 0 && s > tm3) {   // keep this slab's four best scores, publish the m-th
                if (s > tm0) { tm3 = tm2; tm2 = tm1; tm1 = tm0; tm0 = s; }
                else if (s > tm1) { tm3 = tm2; tm2 = tm1; tm1 = s; }
                else if (s > tm2) { tm3 = tm2; tm2 = s; }
                else tm3 = s;
                const float mth = p.m_rank == 1 ? tm0 : (p.m_rank == 2 ? tm1 : (p.m_rank == 3 ? tm2 : tm3));
                if (mth > -INFINITY) *gs = ord32(mth);
              }
              if (root) {
                tau_loc = key_score(root);
                tau = fmaxf(tau, tau_loc);
                atomicMax(gt, ord32(tau_loc));
              }
            }
          }
        };
        uint32_t u0 = __reduce_or_sync(0xffffffffu, m0), u1 = __reduce_or_sync(0xffffffffu, m1);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int j0 = 0, j1 = 0, j2 = 0, j3 = 0, nst = 0;
        while (u0 | u1) {
          int j;
          if (u0) {
            j = __ffs(u0) - 1;
            u0 &= u0 - 1;
          } else {
            j = 32 + __ffs(u1) - 1;
            u1 &= u1 - 1;
          }
          const uint32_t v = tmem_ld1(taddr + j);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const bool mine = (((j < 32) ? (m0 >> j) : (m1 >> (j - 32))) & 1u) != 0;
          if (mine) {
            const float s = __uint_as_float(v) * nbuf[j];
            if (nst == 0) { s0 = s; j0 = j; }
            else if (nst == 1) { s1 = s; j1 = j; }
            else if (nst == 2) { s2 = s; j2 = j; }
            else if (nst == 3) { s3 = s; j3 = j; }
            else insert(s, j);                 // stash full (warm-up tiles only): insert in place
            ++nst;
          }
        }
        if (zacc) {   // hand the accumulator back zeroed
          uint32_t z[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) z[j] = 0u;
          tmem_st32(taddr, z);
          tmem_st32(taddr + 32, z);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {                          // accumulator stage may be overwritten
          if constexpr (PAIR) mbar_arrive_remote(tempty_leader + as * 8);
          else mbar_arrive(tempty + as);
        }
        // ---- 2b
        if (nst > 0) insert(s0, j0);
        if (nst > 1) insert(s1, j1);
        if (nst > 2) insert(s2, j2);
        if (nst > 3) insert(s3, j3);
      }
      it += uint32_t(t1 - t0);
      // ---- hand the list over
      if (qb < p.nqb) {   // (the odd block out of a pair carries no queries)
        uint64_t* dst = p.parts + (size_t(vslab) * p.q_stride + size_t(qb) * BLOCK_M + L) * p.kk;
        for (int c = 0; c < p.kk; ++c) dst[c] = lds_u64(mylist + c * LSTRIDE);
      }
    }
  }

  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();   // nobody leaves while the peer may still touch its barriers / TMEM
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(TMEM_COLS))
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(TMEM_COLS))
                   : "memory");
  }
}

// ---------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

struct Plan {
  int nqb, n_tiles, n_slabs, n_stages, grid, m_rank, n_groups, nbox, n_issuers;
  bool pair;   // CTA pairs (cta_group::2): two query blocks share every MMA
  size_t smem;
};

// CTA pairs are opt-in for now (DRS_UMMA_PAIR=1)
static bool pair_enabled() {
  static const int on = [] { const char* e = getenv("DRS_UMMA_PAIR"); return e ? atoi(e) : 0; }();
  return on != 0;
}

static bool make_plan(int n_queries, int64_t n_rows, int dim, int kk, Plan& pl) {
  if (dim % MIN_STAGE_K != 0 || dim <= 0 || dim > MAX_D) return false;
  if (n_queries <= 0 || n_rows <= 0 || n_rows >= (int64_t(1) << 31)) return false;
  // two candidate heaps per query (one per epilogue group) when they leave room for a >= 6-stage ring
  size_t fixed = 1024 /*align*/ + 2 * size_t(kk) * BLOCK_M * 8 + 6144 /*barriers, tmem slot, norm strips*/;
  pl.n_groups = 2;
  if (fixed + 6 * size_t(MIN_STAGE_BYTES) > SMEM_LIMIT) {
    fixed -= size_t(kk) * BLOCK_M * 8;
    pl.n_groups = 1;
  }
  const int g = sm_count();
  pl.nqb = (n_queries + BLOCK_M - 1) / BLOCK_M;
  pl.pair = pair_enabled() && pl.nqb >= 2 && (g % 2) == 0 && dim % 256 == 0 && (pl.nqb + 1) / 2 <= g / 2;
  // 32 KB stages (K = 256) when the dimension allows and at least three of them fit; else 16 KB stages.  In a pair
  // every CTA holds half of each tile: K = 256 stages of 16 KB.
  pl.nbox = (dim % 256 == 0 && (pl.pair || fixed + 3 * size_t(4 * BOX_BYTES) <= SMEM_LIMIT)) ? 4 : 2;
  const size_t STAGE_BYTES = size_t(pl.nbox) * (pl.pair ? BOX_BYTES / 2 : BOX_BYTES);
  if (fixed + size_t(pl.nbox == 4 && !pl.pair ? 3 : MIN_STAGES) * STAGE_BYTES > SMEM_LIMIT) return false;
  int ns = int((SMEM_LIMIT - fixed) / STAGE_BYTES);
  if (ns > (pl.pair ? 48 : 24) / pl.nbox) ns = (pl.pair ? 48 : 24) / pl.nbox;
  static const int cap_stages = [] { const char* e = getenv("DRS_UMMA_STAGES"); return e ? atoi(e) : 0; }();   // tuning hook
  if (cap_stages > 0 && ns > cap_stages) ns = cap_stages;
  // MMA-issuing warps.  The ring stages are dealt round-robin, and a warp waits only for ITS stages' `full`
  // barriers -- by phase parity.  That is sound only if a ring slot is always consumed by the same warp
  // (n_stages % n_issuers == 0): otherwise a warp can reach a slot whose PREVIOUS lap (another warp's stage) has
  // not even landed yet, the parity test passes on the stale phase, and it issues MMAs on a buffer still being
  // filled.  (This is what faulted with 3 issuers on a 4-stage ring; it needs TMA completions one stage out of
  // order, so it only showed on long runs.)  The ring is therefore trimmed to a multiple of the issuer count.
  // Several issuers are used in the class they were validated in: 32 KB stages, two epilogue groups.
  static const int want_iss = [] { const char* e = getenv("DRS_UMMA_ISSUERS"); return e ? atoi(e) : 3; }();
  pl.n_issuers = 1;
  if (!pl.pair && pl.nbox == 4 && pl.n_groups == 2 && want_iss > 1) {
    const int ni = want_iss > 3 ? 3 : want_iss;
    if (ns >= ni) {
      pl.n_issuers = ni;
      ns = (ns / ni) * ni;
    }
  }
  pl.n_stages = ns;
  pl.smem = fixed + size_t(ns) * STAGE_BYTES;
  pl.n_tiles = int((n_rows + TILE_N - 1) / TILE_N);
  if (pl.pair) {
    // one item per CTA pair: (pair of query blocks, slab)
    const int nqb2 = (pl.nqb + 1) / 2, gp = g / 2;
    int s = gp / nqb2;
    if (s > pl.n_tiles) s = pl.n_tiles;
    pl.n_slabs = s < 1 ? 1 : s;
    const int items = nqb2 * pl.n_slabs;
    pl.grid = 2 * (items < gp ? items : gp);
    const int m = (kk + pl.n_groups * pl.n_slabs - 1) / (pl.n_groups * pl.n_slabs);
    pl.m_rank = (items <= gp && m <= 4) ? m : 0;
    return true;
  }
  // One item per CTA whenever the query blocks fit: every slab of a query block is then resident at
  // the same time, which the pooled threshold relies on (and no CTA reloads its A operand).
  int s = pl.nqb <= g ? g / pl.nqb : 1;
  if (s > pl.n_tiles) s = pl.n_tiles;
  pl.n_slabs = s < 1 ? 1 : s;
  const int items = pl.nqb * pl.n_slabs;
  pl.grid = items < g ? items : g;
  const int m = (kk + pl.n_groups * pl.n_slabs - 1) / (pl.n_groups * pl.n_slabs);   // heaps == virtual slabs
  pl.m_rank = (items <= g && m <= 4) ? m : 0;
  return true;
}

}  // namespace umma

bool dense_umma_supported(int n_queries, int64_t n_rows, int dim, int kk) {
  umma::Plan pl;
  return umma::make_plan(n_queries, n_rows, dim, kk, pl) && umma::encode_tiled() != nullptr;
}

int dense_umma_query_mode() { return DRS_QMODE_BF16_RAW; }

int dense_umma_parts(int n_queries, int64_t n_rows, int dim, int kk) {
  umma::Plan pl;
  return umma::make_plan(n_queries, n_rows, dim, kk, pl) ? pl.n_groups * pl.n_slabs : 0;
}

int launch_dense_umma(const float* d_q, int n_queries, const void* d_corpus, const float* d_inv_norm,
                      const uint8_t* d_mask, int64_t n_rows, int dim, int kk, uint64_t* d_parts, int q_stride,
                      uint32_t* d_gtau, cudaStream_t st) {
  using namespace umma;
  Plan pl;
  DRS_REQUIRE(make_plan(n_queries, n_rows, dim, kk, pl), DRS_ERR_UNSUPPORTED,
              "dense(tcgen05): needs dim % 128 == 0, dim <= 768 and k' small enough for shared memory");
  EncodeTiledFn enc = encode_tiled();
  DRS_REQUIRE(enc != nullptr, DRS_ERR_UNSUPPORTED, "dense(tcgen05): cuTensorMapEncodeTiled unavailable");
  DRS_REQUIRE((reinterpret_cast<uintptr_t>(d_q) & 15) == 0, DRS_ERR_ARG, "dense: queries not 16-byte aligned");
  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {cuuint64_t(dim), cuuint64_t(n_rows)};
  const cuuint64_t gstride[1] = {cuuint64_t(dim) * 2};
  const cuuint32_t box[2] = {BOX_K, cuuint32_t(pl.pair ? TILE_N / 2 : TILE_N)};   // a CTA of a pair loads half a tile
  const cuuint32_t estride[2] = {1, 1};
  CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(d_corpus), gdim, gstride, box, estride,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DRS_REQUIRE(r == CUDA_SUCCESS, DRS_ERR_CUDA, "dense(tcgen05): cuTensorMapEncodeTiled failed (" + std::to_string(int(r)) + ")");
  Params p;
  p.queries = d_q;
  p.inv_norm = d_inv_norm;
  p.mask = d_mask;
  p.parts = d_parts;
  p.gtau = d_gtau;
  p.gslab = d_gtau + q_stride;
  p.m_rank = pl.m_rank;
  p.n_groups = pl.n_groups;
  p.n_issuers = pl.n_issuers;
  p.n_rows = n_rows;
  p.n_queries = n_queries;
  p.nqb = pl.nqb;
  p.dim = dim;
  p.kk = kk;
  p.q_stride = q_stride;
  p.n_slabs = pl.n_slabs;
  p.n_tiles = pl.n_tiles;
  p.n_stages = pl.n_stages;
  static const uint32_t dbg = [] { const char* e = getenv("DRS_UMMA_DEBUG"); return e ? uint32_t(atoi(e)) : 0u; }();
  p.debug = dbg;
  DRS_CUDA_CHECK(cudaMemsetAsync(d_gtau, 0, size_t(q_stride) * (1 + pl.n_groups * pl.n_slabs) * sizeof(uint32_t), st));
  if (pl.pair) {
    auto kern = dense_umma_kernel<4, true, false>;
    DRS_ENSURE_SMEM((kern), pl.smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(pl.grid);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DRS_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tmap, p));
  } else if (pl.nbox == 4 && p.n_issuers > 1) {
    DRS_ENSURE_SMEM((dense_umma_kernel<4, false, true>), pl.smem);
    dense_umma_kernel<4, false, true><<<pl.grid, THREADS, pl.smem, st>>>(tmap, p);
  } else if (pl.nbox == 4) {
    DRS_ENSURE_SMEM((dense_umma_kernel<4, false, false>), pl.smem);
    dense_umma_kernel<4, false, false><<<pl.grid, THREADS, pl.smem, st>>>(tmap, p);
  } else {
    DRS_ENSURE_SMEM((dense_umma_kernel<2, false, false>), pl.smem);
    dense_umma_kernel<2, false, false><<<pl.grid, THREADS, pl.smem, st>>>(tmap, p);
  }
  count_launches(1);
  DRS_CUDA_CHECK(cudaGetLastError());
  return DRS_OK;
}

}  // namespace drs
