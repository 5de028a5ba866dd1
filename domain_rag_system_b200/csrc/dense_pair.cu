// Dense cosine candidate selection, CTA-pair form (sm_100a): the main tensor-core kernel.
//
//   scores[q, row] = queries[q, :] . corpus[row, :]   as a persistent, warp-specialised bf16 GEMM issued as
//   tcgen05.mma.cta_group::2 with M = 256 (two blocks of 128 queries, one per CTA of the pair), N = 256 (corpus rows,
//   128 per CTA) and BOTH operands read from shared memory (SS form).  Every MMA therefore keeps the tensor pipes of
//   both SMs busy for 128 cycles, which amortises the fixed issue cost that capped the A-in-tensor-memory kernel
//   (dense_umma.cu, N = 64) at ~55 % pipe occupancy, and each shared-memory byte feeds 256 flop.
//
//   operands   = K-chunks of 64 elements of BOTH matrices stream through one TMA ring: per stage and CTA 16 KB of
//                queries (128 x 64, bf16, pre-converted once per call, L2-resident) + 16 KB of corpus rows
//                (128 x 64), 128-byte swizzle, `cp.async.bulk.tensor...cta_group::2` signalling the leader's barrier.
//   D          = 2 x (128 lanes x 256 columns) fp32 in tensor memory: ALL 512 columns, double buffered, so the
//                epilogue of tile t overlaps the MMAs of tile t+1.
//   epilogue   = lane L of a CTA is query L of its block; after tcgen05.ld a thread holds its query's 256 scores of the
//                tile, scales them by inv_norm[row] and compares against its threshold -- a proven lower bound of
//                the query's kk-th best score over the whole corpus, shared by all CTAs (see "pool parts").  The rare
//                survivors are APPENDED to the query's candidate buffer in global memory: no list, no heap and no
//                score is ever kept on chip, so shared memory is all ring, and k' is not limited by it.
//
// Roles per CTA (384 threads): warp 0 TMA producer (each CTA loads ITS halves), warp 1 MMA issuer (leader CTA only,
// one elected lane), warp 2 tensor-memory allocator, warps 4..11 two epilogue groups that alternate tiles.
// Work item = (pair of query blocks, slab of corpus tiles), ONE per CTA pair (the threshold scheme needs every slab
// of a query block resident at once); the pairs that run concurrently cover the same slab for different query
// blocks, so the corpus leaves HBM about once.
// Output: per (slab, query) the appended keys and their count; select.cu drops what the final thresholds rule out,
// keeps the kk best and re-scores them exactly.  A buffer that overflows (thresholds that cannot tighten: e.g. an
// all-zero query, every score ties) marks its query "unproven", which sends it through the exact fp32 kernel.
#include <cuda.h>

#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "umma_ptx.cuh"

namespace drs {
namespace umma {
namespace ss2 {

constexpr int TILE_N = 256;                          // corpus rows per accumulator tile (128 per CTA)
constexpr int HALF_N = 128;
constexpr int STAGE_K = 64;                          // K elements per ring stage: 64 bf16 = one 128-byte swizzle row
constexpr int A_BYTES = BLOCK_M * STAGE_K * 2;       // 16384: my 128 queries x 64
constexpr int B_BYTES = HALF_N * STAGE_K * 2;        // 16384: my 128 corpus rows x 64
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;       // 32768 per CTA
constexpr int MAX_STAGES = 6;
constexpr size_t SMEM_LIMIT = 227 * 1024;
// D = f32, A = B = bf16, both K-major, N = 256, M = 256 (over the pair)
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(TILE_N >> 3) << 17) |
                           (uint32_t((2 * BLOCK_M) >> 4) << 24);

struct Params {
  const float* inv_norm;
  const uint8_t* mask;
  uint64_t* parts;     // [n_slabs][q_stride][cap] appended keys
  uint32_t* counts;    // [n_slabs][q_stride] keys appended (may exceed cap: the excess was dropped = overflow)
  uint32_t* gtau;      // [q_stride] zero-initialised: the queries' shared thresholds (ordered-u32, atomicMax)
  uint32_t* gslab;     // [n_slabs * pool_g][q_stride] zero-initialised: running maximum of every pool part
  uint32_t* seed_cnt;  // [2][2 * nqb2] zero-initialised: epilogue warps of a query block past the two seed steps
  int64_t n_rows;
  int n_queries, nqb, nqb2, dim, kk, cap, q_stride, n_slabs, n_tiles, n_stages;
  int pool_g;          // pool parts per slab (1, 2, .. 32: groups of 256 / pool_g columns of a tile)
  int mask_vec;        // the row mask may be read 8 bytes at a time (8-byte aligned)
  uint32_t debug;      // timing experiments only (results are garbage): 1 no survivors, 2 no epilogue work,
                       // 4 no threshold publishing, 16 no seed pass
  unsigned long long* trace;   // DRS_PAIR_TRACE: per tile of block 0 / warp 4: {clock at accumulator ready, clock at
                               // tile done, drains}; NULL otherwise
};

// D[tmem, 256 x 256 over the pair] (+)= A[smem, 128 rows per CTA] * B[smem, 128 rows per CTA]^T
__device__ __forceinline__ void umma_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// fp32 queries -> bf16 [nqb2 * 256, dim], rows beyond n_queries zero (the TMA source of the A operand)
__global__ void __launch_bounds__(256) queries_to_bf16_kernel(const float* __restrict__ q, int n_queries, int n_pad,
                                                              int dim, __nv_bfloat16* __restrict__ out) {
  const int64_t n8 = int64_t(n_pad) * dim / 8;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n8; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t e = i * 8;
    const int row = int(e / dim);
    uint4 w = make_uint4(0u, 0u, 0u, 0u);
    if (row < n_queries) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(q + e));
      const float4 b = __ldg(reinterpret_cast<const float4*>(q + e + 4));
      const __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
      const __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
      w.x = *reinterpret_cast<const uint32_t*>(&p0);
      w.y = *reinterpret_cast<const uint32_t*>(&p1);
      w.z = *reinterpret_cast<const uint32_t*>(&p2);
      w.w = *reinterpret_cast<const uint32_t*>(&p3);
    }
    *reinterpret_cast<uint4*>(out + e) = w;
  }
}

constexpr int THREADS = 384;                         // 4 service warps + 2 epilogue groups of 4 warps
constexpr int POOL_GMAX = 32;                        // pool parts per slab at most (groups of 8 columns of a tile)
constexpr int EPI_BAR = 1;                           // named barrier of the 8 epilogue warps
constexpr uint32_t ORD_NEG_INF = 0x007fffffu;        // ord32(-infinity)
constexpr int POOL_PER_LANE = 10;                    // pool parts a lane holds in the warp-wide selection
constexpr int POOL_MAX = 32 * POOL_PER_LANE;         // pool parts per query at most
constexpr int STRIP_WORDS = TILE_N + 8;              // per warp: the tile's 256 inverse norms + 256 validity bits

__global__ void __launch_bounds__(THREADS, 1)
dense_pair_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c, Params p) {
  const uint32_t rank = cluster_ctarank();            // 0 = leader
  const int cta = int(blockIdx.x >> 1);               // work-item stream this pair walks
  const int n_ctas = int(gridDim.x >> 1);
  extern __shared__ uint8_t smem_dyn[];
  // carve: [ring | append counters | part maxima | norm strips | barriers | tmem slot]; the swizzled ring needs
  // 1024-byte alignment
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  uint32_t* acnt = reinterpret_cast<uint32_t*>(ring + size_t(p.n_stages) * STAGE_BYTES);   // [128] keys appended
  uint32_t* pmem = acnt + BLOCK_M;                                                           // [pool_g][128] ord32
  float* norm_strips = reinterpret_cast<float*>(pmem + size_t(p.pool_g) * BLOCK_M);         // [8 warps][STRIP_WORDS]
  uint64_t* bars = reinterpret_cast<uint64_t*>(norm_strips + 8 * STRIP_WORDS);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.n_stages;
  uint64_t* tfull = bars + 2 * p.n_stages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = p.nqb2 * p.n_slabs;
  const int kst = p.dim / STAGE_K;   // ring stages per accumulator tile

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_q)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_c)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    // The leader's full barriers count the bytes of BOTH CTAs' TMA loads (one arrival: the leader's expect_tx), its
    // tempty barriers the epilogue warps of both CTAs; empty and tfull are signalled in each CTA by the leader's
    // multicast commit.
    for (int s = 0; s < p.n_stages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    mbar_init(tfull + 0, 1);
    mbar_init(tfull + 1, 1);
    mbar_init(tempty + 0, 8);   // 4 epilogue warps x 2 CTAs
    mbar_init(tempty + 1, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();   // the peer's barriers are initialised before anyone arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  auto slab_tiles = [&](int slab, int& t0, int& t1) {
    t0 = int((int64_t(slab) * p.n_tiles) / p.n_slabs);
    t1 = int((int64_t(slab + 1) * p.n_tiles) / p.n_slabs);
  };

  if (warp == 0) {
    // ===================================================================== TMA producer (both CTAs)
    uint32_t s = 0, ph = 0;
    for (int item = cta; item < n_items; item += n_ctas) {
      int t0, t1;
      slab_tiles(item / p.nqb2, t0, t1);
      const int qrow = (2 * (item % p.nqb2) + int(rank)) * BLOCK_M;   // my block of 128 queries
      for (int t = t0; t < t1; ++t) {
        const int crow = t * TILE_N + int(rank) * HALF_N;             // my half of the corpus tile
        for (int ks = 0; ks < kst; ++ks) {
          mbar_wait(empty + s, ph ^ 1, 1);
          if (elect_one()) {
            uint8_t* dst = ring + size_t(s) * STAGE_BYTES;
            // My bytes land in MY ring; they are counted by the LEADER's barrier, which expects both CTAs' bytes
            // with ONE plain local arrival.  (No cluster-scope release here: that costs a MEMBAR.GPU per stage,
            // which waits for the TMA loads in flight and serialises the ring.)  The peer's bytes may land before
            // the leader's expect_tx: the transaction count goes negative for a moment, the phase cannot complete
            // before the leader's arrival.  The peer cannot run a lap ahead: its `empty` barrier is signalled by
            // MMAs the leader issued after that lap's `full` phase completed.
            const uint32_t lbar = mapa_u32(smem_u32(full + s), 0);
            if (rank == 0) mbar_arrive_expect_tx(full + s, 2u * uint32_t(STAGE_BYTES));
            tma_load_2d_pair(&tmap_q, lbar, dst, ks * STAGE_K, qrow);
            tma_load_2d_pair(&tmap_c, lbar, dst + A_BYTES, ks * STAGE_K, crow);
          }
          __syncwarp();
          if (++s == uint32_t(p.n_stages)) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===================================================================== MMA issuer (leader CTA)
    // The whole warp walks the loops (addresses stay warp-uniform); one elected lane issues.  A stage is four
    // M256 x N256 x K16 MMAs of 128 tensor cycles each.
    uint32_t s = 0, ph = 0, it = 0;
    for (int item = cta; item < n_items; item += n_ctas) {
      int t0, t1;
      slab_tiles(item / p.nqb2, t0, t1);
      for (int t = t0; t < t1; ++t, ++it) {
        const uint32_t as = it & 1, aph = (it >> 1) & 1;
        mbar_wait(tempty + as, aph ^ 1, 3);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * TILE_N;
        for (int ks = 0; ks < kst; ++ks) {
          mbar_wait(full + s, ph, 4);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t base = smem_u32(ring + size_t(s) * STAGE_BYTES);
            const uint64_t a0 = make_b_desc(base), b0 = make_b_desc(base + A_BYTES);
#pragma unroll
            for (int k16 = 0; k16 < STAGE_K / 16; ++k16)   // +32 bytes per K = 16 step inside the 128-byte swizzle row
              umma_ss_pair(d_tmem, a0 + uint64_t(k16 * 2), b0 + uint64_t(k16 * 2), IDESC, (ks | k16) != 0 ? 1u : 0u);
            tc_commit_pair(empty + s);                       // frees the stage in both CTAs
            if (ks == kst - 1) tc_commit_pair(tfull + as);   // both CTAs' halves of the accumulator are complete
          }
          __syncwarp();
          if (++s == uint32_t(p.n_stages)) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================================================================== epilogue
    // Two groups of four warps; group g owns accumulator stage g and every other tile, i.e. has two tile-times to
    // drain one tile.  Lane L of BOTH groups is query L of the block.
    const int ew = warp & 3;                 // the TMEM sub-partition (lanes 32*ew ..) this warp may touch
    const int grp = (warp - 4) >> 2;
    const int L = ew * 32 + lane;            // TMEM lane == query within my block
    const uint32_t lane_addr = uint32_t(ew * 32) << 16;
    const uint32_t cnt_addr = smem_u32(acnt) + uint32_t(L) * 4;         // keys appended for query L (both groups)
    const uint32_t pm_addr = smem_u32(pmem) + uint32_t(L) * 4;          // part g at + g * 512
    const uint32_t nstrip = smem_u32(norm_strips + (warp - 4) * STRIP_WORDS);   // this warp's norm strip ...
    const uint32_t vstrip = nstrip + TILE_N * 4;                                // ... and row validity bits
    const uint32_t tempty_leader = mapa_u32(smem_u32(tempty), 0);
    const bool no_insert = (p.debug & 1) != 0;
    const int G = p.pool_g;
    const int n_pool = p.n_slabs * G;
    uint32_t it = 0;                         // tiles issued so far, mirrors the MMA warp
    for (int item = cta; item < n_items; item += n_ctas) {
      const int qb = 2 * (item % p.nqb2) + int(rank), slab = item / p.nqb2;
      int t0, t1;
      slab_tiles(slab, t0, t1);
      const int q = qb * BLOCK_M + L;
      const bool live = q < p.n_queries;
      if (grp == 0) {
        sts_u32(cnt_addr, 0u);
        for (int g = 0; g < G; ++g) sts_u32(pm_addr + g * 512, 0u);
      }
      named_bar_sync(EPI_BAR, 256);
      uint64_t* mybuf = p.parts + (size_t(slab) * p.q_stride + (live ? q : 0)) * p.cap;
      // Pool parts and the shared threshold.  The corpus is cut into n_pool = (#slabs x G) disjoint PARTS: part (v, g)
      // = the rows of slab v that sit in column group g (256 / G columns) of their tile.  All slabs of a query block
      // run concurrently and publish the running MAXIMUM of each of their parts in gslab[part][q].  If kk different
      // parts each hold a row scoring >= x, the corpus has >= kk such rows: the kk-th LARGEST part maximum is a valid
      // lower bound of the query's kk-th best score, and with about twice as many parts as kk it is close to it (a
      // row below it can be dropped for good).  Selecting it costs a few hundred instructions, so it is computed
      // for every query by ONE warp (the query's owner among the slabs' epilogue warps, all 32 lanes cooperating)
      // and shared through gtau[q] (atomicMax); everybody else reads gtau[q] once per tile.
      uint32_t* gs = p.gslab + size_t(slab) * G * p.q_stride + (live ? q : 0);
      const volatile uint32_t* gt = p.gtau + (live ? q : 0);
      float tau = live ? -INFINITY : INFINITY;   // padding lanes never survive
      const bool owner = live && (L % (2 * p.n_slabs)) == 2 * slab + grp;
      auto publish_bounds = [&]() {
        uint32_t todo = __ballot_sync(0xffffffffu, owner);
        while (todo) {
          const int src = __ffs(todo) - 1;
          todo &= todo - 1;
          const int qq = __shfl_sync(0xffffffffu, q, src);
          // my share of query qq's part maxima; parts without a valid row so far (-infinity) or that have not
          // reported (0) count for nothing
          uint32_t v[POOL_PER_LANE];
#pragma unroll
          for (int j = 0; j < POOL_PER_LANE; ++j) {
            const int part = lane + 32 * j;
            uint32_t x = part < n_pool ? __ldcg(p.gslab + size_t(part) * p.q_stride + qq) : 0u;
            v[j] = x == ORD_NEG_INF ? 0u : x;
          }
          uint32_t cur = 0;   // the largest x with at least kk parts >= x, bit by bit
#pragma unroll 1
          for (int bit = 31; bit >= 0; --bit) {
            const uint32_t x = cur | (1u << bit);
            int c = 0;
#pragma unroll
            for (int j = 0; j < POOL_PER_LANE; ++j) c += v[j] >= x;
            if (__reduce_add_sync(0xffffffffu, c) >= p.kk) cur = x;
          }
          if (lane == src && cur > ORD_NEG_INF) atomicMax(p.gtau + qq, cur);
        }
      };
      // inverse norms and validity (inside the corpus, passes the row mask) of the tile's 256 rows: lane l fetches
      // rows 8l..8l+7 one tile ahead, parks them in this warp's strip, and every lane reads all of them by broadcast
      float4 nA, nB;
      uint32_t vbyte;
      auto load_norms = [&](int t) {
        const int64_t r = int64_t(t) * TILE_N + 8 * lane;
        if (t < t1 && r + 7 < p.n_rows) {
          nA = __ldg(reinterpret_cast<const float4*>(p.inv_norm + r));
          nB = __ldg(reinterpret_cast<const float4*>(p.inv_norm + r + 4));
          vbyte = 0xffu;
          if (p.mask != nullptr) {
            uint2 mk;
            if (p.mask_vec) {
              mk = __ldg(reinterpret_cast<const uint2*>(p.mask + r));
            } else {
              mk.x = uint32_t(p.mask[r]) | (uint32_t(p.mask[r + 1]) << 8) | (uint32_t(p.mask[r + 2]) << 16) | (uint32_t(p.mask[r + 3]) << 24);
              mk.y = uint32_t(p.mask[r + 4]) | (uint32_t(p.mask[r + 5]) << 8) | (uint32_t(p.mask[r + 6]) << 16) | (uint32_t(p.mask[r + 7]) << 24);
            }
            vbyte = 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              vbyte |= ((mk.x >> (8 * j)) & 0xffu) ? (1u << j) : 0u;
              vbyte |= ((mk.y >> (8 * j)) & 0xffu) ? (16u << j) : 0u;
            }
          }
        } else {
          float v[8];
          vbyte = 0u;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const bool in = t < t1 && r + j < p.n_rows;
            v[j] = in ? __ldg(p.inv_norm + r + j) : 0.f;
            if (in && (p.mask == nullptr || p.mask[r + j] != 0)) vbyte |= 1u << j;
          }
          nA = make_float4(v[0], v[1], v[2], v[3]);
          nB = make_float4(v[4], v[5], v[6], v[7]);
        }
      };
      const int t_first = t0 + (((it & 1u) == uint32_t(grp)) ? 0 : 1);   // my first tile of the item
      // Seed rendezvous: every epilogue warp working on this block of 128 queries counts itself in once its first
      // tile's part maxima are published (or at once if it has no tile / cannot seed).
      const uint32_t seed_expect = uint32_t(p.n_slabs * 8);
      uint32_t* seed_ctr = p.seed_cnt + qb;                    // step 1: part maxima of the first tiles published
      uint32_t* seed_ctr2 = p.seed_cnt + 2 * p.nqb2 + qb;      // step 2: the owners' bounds published
      const bool seeding = !(p.debug & 16);
      auto rendezvous = [&](uint32_t* ctr) {   // count this warp in, wait (bounded) for all the others
        __threadfence();
        __syncwarp();
        int all_in = 0;
        if (lane == 0) {
          atomicAdd(ctr, 1u);
          const long long w0 = clock64();
          do {
            all_in = *reinterpret_cast<volatile uint32_t*>(ctr) >= seed_expect ? 1 : 0;
          } while (!all_in && clock64() - w0 < 60000);
        }
        all_in = __shfl_sync(0xffffffffu, all_in, 0);
        __threadfence();
        return all_in != 0;
      };
      if (seeding && t_first >= t1) {   // a warp without a tile in this item still takes part in both steps
        if (rendezvous(seed_ctr)) publish_bounds();
        rendezvous(seed_ctr2);
      }
      load_norms(t_first);
      int nt = 0, next_refresh = 1;
      for (int t = t_first; t < t1; t += 2, ++nt) {
        const uint32_t itt = it + uint32_t(t - t0);      // this tile's global sequence number
        __syncwarp();                                    // everyone is done reading the previous tile's strip
        sts_f32x4(nstrip + 32 * lane, nA);
        sts_f32x4(nstrip + 32 * lane + 16, nB);
        sts_u8(vstrip + lane, vbyte);
        __syncwarp();
        load_norms(t + 2);                               // prefetch for my next tile
        // the bound moves fastest early on: refreshed before tiles 1, 2, 3, 4, 6, 9, 13, ... then every 32nd
        if (nt == next_refresh) {
          publish_bounds();
          next_refresh += min(32, max(1, nt >> 1));
        }
        const uint32_t as = itt & 1u, aph = (itt >> 1) & 1;
        mbar_wait(tfull + as, aph, 5);
        tc_fence_after();
        const uint32_t taddr0 = tmem_base + lane_addr + as * TILE_N;
        if (seeding && nt == 0) {
          // ---- seed pass over my FIRST tile: nothing is appended; the maximum of each group of 8 columns is
          // folded into the part maxima and published, then the warp waits (bounded) until every slab of this
          // query block has done the same.  The first real pass then starts from a threshold that only
          // ~kk / (rows seen by all slabs) of the scores reach, instead of from -infinity.
          {
#pragma unroll 1
            for (int ch = 0; ch < TILE_N / 64; ++ch) {
              uint32_t acc0[32], acc1[32];
              tmem_ld32(taddr0 + ch * 64, acc0);
              tmem_ld32(taddr0 + ch * 64 + 32, acc1);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
              const uint32_t vb0 = lds_u32_volatile(vstrip + 8 * ch), vb1 = lds_u32_volatile(vstrip + 8 * ch + 4);
#pragma unroll
              for (int g = 0; g < 4; ++g) {   // column group of columns c..c+7: ((c >> 3) * G) >> 5
                const float4 n0 = lds_f32x4(nstrip + (ch * 64 + 8 * g) * 4), n1 = lds_f32x4(nstrip + (ch * 64 + 8 * g + 4) * 4);
                const float nn[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
                float a = -INFINITY;
#pragma unroll
                for (int e = 0; e < 8; ++e)
                  a = fmaxf(a, ((vb0 >> (8 * g + e)) & 1u) ? __uint_as_float(acc0[8 * g + e]) * nn[e] : -INFINITY);
                smem_atomic_max_u32(pm_addr + (((ch * 8 + g) * G) >> 5) * 512, ord32(a));
              }
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const float4 n0 = lds_f32x4(nstrip + (ch * 64 + 32 + 8 * g) * 4), n1 = lds_f32x4(nstrip + (ch * 64 + 32 + 8 * g + 4) * 4);
                const float nn[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
                float a = -INFINITY;
#pragma unroll
                for (int e = 0; e < 8; ++e)
                  a = fmaxf(a, ((vb1 >> (8 * g + e)) & 1u) ? __uint_as_float(acc1[8 * g + e]) * nn[e] : -INFINITY);
                smem_atomic_max_u32(pm_addr + (((ch * 8 + 4 + g) * G) >> 5) * 512, ord32(a));
              }
            }
            if (live)
              for (int g = 0; g < G; ++g) atomicMax(gs + size_t(g) * p.q_stride, lds_u32_volatile(pm_addr + g * 512));
          }
          if (rendezvous(seed_ctr)) publish_bounds();
          rendezvous(seed_ctr2);
        }
        {   // the query's shared threshold (whoever published it last)
          const uint32_t g = *gt;
          if (g != 0u) tau = fmaxf(tau, unord32(g));
        }
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int j0 = 0, j1 = 0, j2 = 0, j3 = 0, nst = 0;
        const bool tracing = p.trace != nullptr && blockIdx.x == 0 && warp == 4 && nt < 1000;
        long long tr0 = 0;
        int trc = 0;
        if (tracing) tr0 = clock64();
        // Survivors of pass 1 are taken out of the REGISTERS the accumulator chunk was loaded into: every lane
        // that has flagged columns picks ITS OWN lowest one through a select tree, so a round serves up to 32
        // different columns at once and a chunk costs max-over-lanes(#flagged) rounds, not #distinct columns.
        // A survivor is first parked in a four-entry register stash and appended in ONE place (the drain at the
        // bottom of the loop) when some lane's stash is full or the tile is done; iteration TILE_N/64 of the chunk
        // loop only drains.
#pragma unroll 1
        for (int ch = 0; ch <= TILE_N / 64; ++ch) {   // 64 accumulator columns at a time
          uint32_t acc0[32], acc1[32];
          uint32_t m0 = 0, m1 = 0;
          const bool last = ch == TILE_N / 64;
          if (!last) {
            const uint32_t taddr = taddr0 + ch * 64;
            if (!(p.debug & 2)) {
              tmem_ld32(taddr, acc0);
              tmem_ld32(taddr + 32, acc1);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            }
            if (ch == TILE_N / 64 - 1) {   // the whole tile has been read: hand the accumulator back
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_cluster_relaxed(tempty_leader + as * 8);
            }
            if (!(p.debug & 2)) {
              // ---- pass 1, branch-free: which of these 64 scores reach the threshold?
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float4 n = lds_f32x4(nstrip + (ch * 64 + 4 * g) * 4);
                m0 |= (__uint_as_float(acc0[4 * g + 0]) * n.x >= tau) ? (1u << (4 * g + 0)) : 0u;
                m0 |= (__uint_as_float(acc0[4 * g + 1]) * n.y >= tau) ? (1u << (4 * g + 1)) : 0u;
                m0 |= (__uint_as_float(acc0[4 * g + 2]) * n.z >= tau) ? (1u << (4 * g + 2)) : 0u;
                m0 |= (__uint_as_float(acc0[4 * g + 3]) * n.w >= tau) ? (1u << (4 * g + 3)) : 0u;
              }
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float4 n = lds_f32x4(nstrip + (ch * 64 + 32 + 4 * g) * 4);
                m1 |= (__uint_as_float(acc1[4 * g + 0]) * n.x >= tau) ? (1u << (4 * g + 0)) : 0u;
                m1 |= (__uint_as_float(acc1[4 * g + 1]) * n.y >= tau) ? (1u << (4 * g + 1)) : 0u;
                m1 |= (__uint_as_float(acc1[4 * g + 2]) * n.z >= tau) ? (1u << (4 * g + 2)) : 0u;
                m1 |= (__uint_as_float(acc1[4 * g + 3]) * n.w >= tau) ? (1u << (4 * g + 3)) : 0u;
              }
              m0 &= lds_u32_volatile(vstrip + 8 * ch);       // rows outside the corpus / excluded by the row mask
              m1 &= lds_u32_volatile(vstrip + 8 * ch + 4);
              if (no_insert) m0 = m1 = 0;
            }
          }
          // ---- pass 2, rare and rolled
          bool more;
          do {
            if ((m0 | m1) != 0 && nst < 4) {
              const bool hi = m0 == 0;
              const uint32_t mm = hi ? m1 : m0;
              const int j = __ffs(mm) - 1;
              if (hi) m1 = mm & (mm - 1);
              else m0 = mm & (mm - 1);
              const uint32_t v0 = pick32(acc0, j), v1 = pick32(acc1, j);
              const int jj = ch * 64 + (hi ? 32 : 0) + j;
              const float s = __uint_as_float(hi ? v1 : v0) * lds_f32(nstrip + jj * 4);
              if (nst == 0) { s0 = s; j0 = jj; }
              else if (nst == 1) { s1 = s; j1 = jj; }
              else if (nst == 2) { s2 = s; j2 = jj; }
              else { s3 = s; j3 = jj; }
              ++nst;
            }
            more = __any_sync(0xffffffffu, (m0 | m1) != 0);
            const bool drain = more ? __any_sync(0xffffffffu, nst == 4) : (last && __any_sync(0xffffffffu, nst > 0));
            if (drain) {
              ++trc;
              for (int i = 0; i < nst; ++i) {
                const float s = i == 0 ? s0 : (i == 1 ? s1 : (i == 2 ? s2 : s3));
                const int j = i == 0 ? j0 : (i == 1 ? j1 : (i == 2 ? j2 : j3));
                {
                  const uint32_t idx = smem_atomic_add_u32(cnt_addr, 1u);
                  if (idx < uint32_t(p.cap)) mybuf[idx] = make_key(s, uint32_t(t) * TILE_N + uint32_t(j));
                  // running maximum of the row's pool part
                  const int g = ((j >> 3) * G) >> 5;
                  const uint32_t o = ord32(s);
                  if (o > smem_atomic_max_u32(pm_addr + g * 512, o) && !(p.debug & 4)) atomicMax(gs + size_t(g) * p.q_stride, o);
                }
              }
              nst = 0;
            }
          } while (more);
        }
        if (tracing && lane == 0) {
          p.trace[3 * nt + 0] = (unsigned long long)tr0;
          p.trace[3 * nt + 1] = (unsigned long long)clock64();
          p.trace[3 * nt + 2] = (unsigned long long)trc;
        }
      }
      it += uint32_t(t1 - t0);
      // ---- both groups are done with the item: publish how many keys query L's buffer holds
      named_bar_sync(EPI_BAR, 256);
      if (grp == 0 && live) p.counts[size_t(slab) * p.q_stride + q] = lds_u32_volatile(cnt_addr);
      named_bar_sync(EPI_BAR, 256);   // ... before the next item's counters are zeroed
    }
  }

  tc_fence_before();
  cluster_sync_all();   // nobody leaves while the peer may still touch its barriers / tensor memory
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      return reinterpret_cast<EncodeTiledFn>(sym);
    return EncodeTiledFn(nullptr);
  }();
  return fn;
}

// Tensor maps depend only on (pointer, rows, dim): the last few are kept so that a steady stream of calls against
// the same index (latency mode, every benchmark step) does not re-encode them.
struct MapKey {
  const void* ptr;
  int64_t rows;
  int dim;
};
static bool cached_map(const void* ptr, int64_t rows, int dim, CUtensorMap& out) {
  constexpr int SLOTS = 8;
  static std::mutex mu;
  static MapKey keys[SLOTS];
  static CUtensorMap maps[SLOTS];
  static int next = 0, used = 0;
  std::lock_guard<std::mutex> lk(mu);
  for (int i = 0; i < used; ++i)
    if (keys[i].ptr == ptr && keys[i].rows == rows && keys[i].dim == dim) {
      out = maps[i];
      return true;
    }
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return false;
  const cuuint64_t gdim[2] = {cuuint64_t(dim), cuuint64_t(rows)};
  const cuuint64_t gstride[1] = {cuuint64_t(dim) * 2};
  const cuuint32_t box[2] = {STAGE_K, HALF_N};
  const cuuint32_t estride[2] = {1, 1};
  CUtensorMap m;
  if (enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estride,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  const int slot = used < SLOTS ? used++ : (next++ % SLOTS);
  keys[slot] = MapKey{ptr, rows, dim};
  maps[slot] = m;
  out = m;
  return true;
}

struct Plan {
  int nqb, nqb2, n_tiles, n_slabs, n_stages, grid, pool_g, cap;
  size_t smem;
};

static size_t smem_bytes(int n_stages, int pool_g) {
  return 1024 /*align*/ + size_t(n_stages) * STAGE_BYTES + BLOCK_M * 4 /*append counters*/ +
         size_t(pool_g) * BLOCK_M * 4 /*part maxima*/ + 8 * STRIP_WORDS * 4 /*norm + validity strips*/ +
         (2 * size_t(MAX_STAGES) + 4) * 8 + 64;
}

// Keys a (slab, query) buffer holds.  A query typically appends a few dozen keys per slab (more while the
// thresholds warm up, and more for a large k'); a buffer that overflows only costs that query the fast path.
static int buffer_cap(int kk) {
  int c = 256;
  while (c < 4 * kk) c <<= 1;
  return c;
}

static bool make_plan(int n_queries, int64_t n_rows, int dim, int kk, Plan& pl) {
  if (dim <= 0 || dim % STAGE_K != 0 || dim > DRS_MAX_DIM) return false;
  if (n_queries <= 0 || n_rows <= 0 || n_rows >= (int64_t(1) << 31)) return false;
  const int g = sm_count();
  if (g < 2) return false;
  pl.nqb = (n_queries + BLOCK_M - 1) / BLOCK_M;
  pl.nqb2 = (pl.nqb + 1) / 2;
  pl.n_tiles = int((n_rows + TILE_N - 1) / TILE_N);
  // One item per CTA pair: every slab of a query block is resident at the same time, which the shared thresholds
  // rely on.  More query blocks than pairs: not this kernel's shape (the caller splits the batch or falls back).
  const int gp = g / 2;
  if (pl.nqb2 > gp) return false;
  // (at least two tiles per slab: a ragged last tile then shares its pool parts with a full one, and both
  // epilogue groups of every CTA have work)
  int s = gp / pl.nqb2;
  if (s > pl.n_tiles / 2) s = pl.n_tiles / 2;
  pl.n_slabs = s < 1 ? 1 : s;
  pl.grid = 2 * pl.nqb2 * pl.n_slabs;
  // pool parts per slab: a power of two G <= 32 with kk <= (#slabs x G) <= POOL_MAX, about 2 kk parts if possible (the
  // kk-th largest of ~2 kk part maxima tracks the kk-th best score closely; of exactly kk it is their minimum)
  pl.pool_g = 0;
  for (int pg = 1; pg <= POOL_GMAX; pg <<= 1) {
    const int parts = pl.n_slabs * pg;
    if (parts > POOL_MAX) break;
    if (parts >= kk) pl.pool_g = pg;
    if (parts >= 2 * kk) break;
  }
  if (pl.pool_g == 0) return false;
  pl.cap = buffer_cap(kk);
  int ns = 2;
  while (ns < MAX_STAGES && smem_bytes(ns + 1, pl.pool_g) <= SMEM_LIMIT) ++ns;
  static const int cap_stages = [] { const char* e = getenv("DRS_PAIR_STAGES"); return e ? atoi(e) : 0; }();   // tuning hook
  if (cap_stages >= 2 && ns > cap_stages) ns = cap_stages;
  pl.n_stages = ns;
  pl.smem = smem_bytes(ns, pl.pool_g);
  return pl.smem <= SMEM_LIMIT;
}

}  // namespace ss2
}  // namespace umma

bool dense_pair_supported(int n_queries, int64_t n_rows, int dim, int kk) {
  umma::ss2::Plan pl;
  return umma::ss2::make_plan(n_queries, n_rows, dim, kk, pl) && umma::ss2::encode_tiled() != nullptr;
}

// Workspace geometry of a call: slabs (= candidate buffers per query), keys per buffer, pool parts (rows of the
// threshold table) and the u32 words of the zero-initialised area [gtau[q_stride] | gslab[pool parts][q_stride] |
// seed counters].
bool dense_pair_geometry(int n_queries, int64_t n_rows, int dim, int kk, int q_stride, DensePairGeometry& geo) {
  umma::ss2::Plan pl;
  if (!umma::ss2::make_plan(n_queries, n_rows, dim, kk, pl)) return false;
  geo.n_slabs = pl.n_slabs;
  geo.cap = pl.cap;
  geo.pool_parts = pl.n_slabs * pl.pool_g;
  geo.zero_words = size_t(q_stride) * (1 + geo.pool_parts) + 4 * size_t(pl.nqb2) + 8;
  geo.query_rows = ((n_queries + 255) / 256) * 256;
  return true;
}

int launch_dense_pair(const float* d_q, int n_queries, const void* d_corpus, const float* d_inv_norm,
                      const uint8_t* d_mask, int64_t n_rows, int dim, int kk, uint64_t* d_parts, uint32_t* d_counts,
                      int q_stride, uint32_t* d_zero_area, void* d_qbf16, cudaStream_t st) {
  using namespace umma;
  using namespace umma::ss2;
  Plan pl;
  DRS_REQUIRE(make_plan(n_queries, n_rows, dim, kk, pl), DRS_ERR_UNSUPPORTED,
              "dense(tcgen05 pair): needs dim % 64 == 0 and at most 256 queries per SM pair");
  DRS_REQUIRE((reinterpret_cast<uintptr_t>(d_q) & 15) == 0, DRS_ERR_ARG, "dense: queries not 16-byte aligned");
  DRS_REQUIRE((reinterpret_cast<uintptr_t>(d_qbf16) & 127) == 0, DRS_ERR_ARG, "dense: query image not 128-byte aligned");
  const int q_rows = ((n_queries + 255) / 256) * 256;
  CUtensorMap tmap_q, tmap_c;
  DRS_REQUIRE(cached_map(d_qbf16, q_rows, dim, tmap_q) && cached_map(d_corpus, n_rows, dim, tmap_c), DRS_ERR_CUDA,
              "dense(tcgen05 pair): cuTensorMapEncodeTiled failed or is unavailable");
  const size_t pool_parts = size_t(pl.n_slabs) * pl.pool_g;
  Params p;
  p.inv_norm = d_inv_norm;
  p.mask = d_mask;
  p.mask_vec = (reinterpret_cast<uintptr_t>(d_mask) & 7) == 0 ? 1 : 0;
  p.parts = d_parts;
  p.counts = d_counts;
  p.gtau = d_zero_area;
  p.gslab = d_zero_area + q_stride;
  p.seed_cnt = d_zero_area + size_t(q_stride) * (1 + pool_parts);
  p.n_rows = n_rows;
  p.n_queries = n_queries;
  p.nqb = pl.nqb;
  p.nqb2 = pl.nqb2;
  p.dim = dim;
  p.kk = kk;
  p.cap = pl.cap;
  p.q_stride = q_stride;
  p.n_slabs = pl.n_slabs;
  p.n_tiles = pl.n_tiles;
  p.n_stages = pl.n_stages;
  p.pool_g = pl.pool_g;
  static const uint32_t dbg = [] { const char* e = getenv("DRS_PAIR_DEBUG"); return e ? uint32_t(atoi(e)) : 0u; }();
  p.debug = dbg;
  // DRS_PAIR_TRACE=1 (debugging aid, synchronises): per-tile timeline of one epilogue warp on stderr
  static const bool want_trace = getenv("DRS_PAIR_TRACE") != nullptr;
  static unsigned long long* d_trace = nullptr;
  if (want_trace && !d_trace) cudaMalloc(&d_trace, 3000 * sizeof(unsigned long long));
  p.trace = want_trace ? d_trace : nullptr;
  if (want_trace) cudaMemsetAsync(d_trace, 0, 3000 * sizeof(unsigned long long), st);
  DRS_ENSURE_SMEM(dense_pair_kernel, SMEM_LIMIT);
  DRS_CUDA_CHECK(cudaMemsetAsync(d_zero_area, 0, (size_t(q_stride) * (1 + pool_parts) + 4 * size_t(pl.nqb2) + 8) * sizeof(uint32_t), st));
  {
    const int64_t n8 = int64_t(q_rows) * dim / 8;
    const int blocks = int(std::min<int64_t>((n8 + 255) / 256, 4 * sm_count()));
    queries_to_bf16_kernel<<<blocks, 256, 0, st>>>(d_q, n_queries, q_rows, dim, reinterpret_cast<__nv_bfloat16*>(d_qbf16));
    count_launches(1);
  }
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
  DRS_CUDA_CHECK(cudaLaunchKernelEx(&cfg, dense_pair_kernel, tmap_q, tmap_c, p));
  count_launches(1);
  DRS_CUDA_CHECK(cudaGetLastError());
  if (want_trace) {
    static unsigned long long h[3000];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[pair trace] tiles/slab %d stages %d pool_g %d cap %d: tile: wait_done(+cycles since first) busy drains\n",
            pl.n_tiles / pl.n_slabs, pl.n_stages, pl.pool_g, pl.cap);
    for (int i = 0; i < 1000 && h[3 * i + 1]; ++i)
      fprintf(stderr, "[pair trace] %d %llu %llu %llu\n", i, h[3 * i] - h[0], h[3 * i + 1] - h[3 * i], h[3 * i + 2]);
  }
  return DRS_OK;
}

}  // namespace drs
