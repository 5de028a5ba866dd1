// C-ABI plumbing: error string, device query, and the dense top-k orchestration
// (candidate selection -> k-way merge -> exact re-score -> optional fp32 re-run of unproven queries).
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace drs {

static thread_local std::string g_last_error;
static thread_local int g_last_dense_kernel = 0;
static std::atomic<unsigned long long> g_launches{0};
void count_launches(int n) { g_launches += (unsigned long long)n; }

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ---- measurement hooks -------------------------------------------------------------------
struct ProfKind {
  std::vector<cudaEvent_t> start, stop;
  size_t used = 0;
};
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static ProfKind g_prof[DRS_PROF_KINDS];
constexpr size_t PROF_MAX_LAUNCHES = 1 << 14;

void prof_begin(int kind, cudaStream_t st) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfKind& k = g_prof[kind];
  if (k.used >= PROF_MAX_LAUNCHES) return;
  if (k.used == k.start.size()) {
    cudaEvent_t a, b;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
    k.start.push_back(a);
    k.stop.push_back(b);
  }
  cudaEventRecord(k.start[k.used], st);
}
void prof_end(int kind, cudaStream_t st) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfKind& k = g_prof[kind];
  if (k.used >= k.start.size()) return;
  cudaEventRecord(k.stop[k.used], st);
  ++k.used;
}

// Which tensor-core kernel serves a shape: the CTA-pair SS-form kernel (dense_pair.cu, M = 256 x N = 256 MMAs) for
// two or more blocks of 128 queries, the single-CTA kernel with the queries parked in tensor memory (dense_umma.cu)
// for a single block, 0 = neither.  DRS_TENSOR_KERNEL=ts|pair forces one (A/B measurements).
enum { TENSOR_NONE = 0, TENSOR_TS = 1, TENSOR_PAIR = 2 };
static int tensor_variant(int kernel, int n_queries, int64_t n_rows, int dim, int kprime) {
  static const int env_forced = [] {
    const char* e = getenv("DRS_TENSOR_KERNEL");
    if (!e) return 0;
    return !strcmp(e, "ts") ? TENSOR_TS : (!strcmp(e, "pair") ? TENSOR_PAIR : 0);
  }();
  const bool ts_ok = dense_umma_supported(n_queries, n_rows, dim, cand_len_umma(kprime));
  const bool pair_ok = dense_pair_supported(n_queries, n_rows, dim, cand_len_pair(kprime));
  if (kernel == DRS_DENSE_UMMA_TS) return ts_ok ? TENSOR_TS : TENSOR_NONE;
  if (kernel == DRS_DENSE_UMMA_PAIR) return pair_ok ? TENSOR_PAIR : TENSOR_NONE;
  const int forced = env_forced;
  if (forced == TENSOR_TS && ts_ok) return TENSOR_TS;
  if (forced == TENSOR_PAIR && pair_ok) return TENSOR_PAIR;
  if (pair_ok && (n_queries > 128 || !ts_ok)) return TENSOR_PAIR;
  return ts_ok ? TENSOR_TS : (pair_ok ? TENSOR_PAIR : TENSOR_NONE);
}

static int pick_dense_kernel(int kernel, int n_queries, int64_t n_rows, int dim, int kprime) {
  const bool umma_ok = tensor_variant(kernel, n_queries, n_rows, dim, kprime) != TENSOR_NONE;
  if (kernel == DRS_DENSE_UMMA || kernel == DRS_DENSE_UMMA_TS || kernel == DRS_DENSE_UMMA_PAIR)
    return umma_ok ? DRS_DENSE_UMMA : -1;
  if (kernel == DRS_DENSE_SIMT) return DRS_DENSE_SIMT;
  // AUTO: the streaming fp32 kernel reads the corpus once per 4 queries; beyond 8 queries the
  // tensor-core kernels (one pass per 128 / 256 queries) win.
  return (umma_ok && n_queries > 8) ? DRS_DENSE_UMMA : DRS_DENSE_SIMT;
}

struct DenseWs {
  size_t parts_keys, cand_keys, scratch_keys, flags_off, gtau_off, qimg_off, counts_off, ovf_off, total;
  int simt_parts, umma_parts, umma_qstride, variant, pool_parts;
  DensePairGeometry geo;
};

static DenseWs dense_ws(int n_queries, int64_t n_rows, int dim, int kprime, int kernel) {
  DenseWs w{};
  const int which = pick_dense_kernel(kernel, n_queries, n_rows, dim, kprime);
  w.simt_parts = dense_simt_grid();
  w.variant = (which == DRS_DENSE_UMMA) ? tensor_variant(kernel, n_queries, n_rows, dim, kprime) : TENSOR_NONE;
  // the longest candidate list any stage of this call carries sizes every buffer
  const size_t kk = w.variant == TENSOR_PAIR ? size_t(cand_len_pair(kprime)) : size_t(cand_len_umma(kprime));
  w.umma_qstride = ((n_queries + 127) / 128) * 128;
  w.umma_parts = w.variant == TENSOR_TS ? dense_umma_parts(n_queries, n_rows, dim, int(kk)) : 0;
  size_t umma_keys = size_t(w.umma_parts) * w.umma_qstride * kk, thr_words = size_t(w.umma_qstride) * (1 + size_t(w.umma_parts));
  w.pool_parts = w.umma_parts;
  if (w.variant == TENSOR_PAIR && dense_pair_geometry(n_queries, n_rows, dim, int(kk), w.umma_qstride, w.geo)) {
    umma_keys = size_t(w.geo.n_slabs) * w.umma_qstride * w.geo.cap;   // one append buffer per (slab, query)
    thr_words = w.geo.zero_words;
    w.pool_parts = w.geo.pool_parts;
  }
  const size_t simt_keys = size_t(w.simt_parts) * n_queries * kk;
  w.parts_keys = std::max(simt_keys, umma_keys);
  w.cand_keys = size_t(n_queries) * kk;
  w.scratch_keys = std::max(merge_scratch_keys(w.simt_parts, n_queries, int(kk)),
                            merge_scratch_keys(w.umma_parts, n_queries, int(kk)));
  w.flags_off = (w.parts_keys + w.cand_keys + w.scratch_keys) * 8;
  w.gtau_off = w.flags_off + ((size_t(n_queries) * 4 + 255) / 256) * 256;
  // pair kernel: bf16 image of the queries (its TMA source, 128-byte aligned), append counts, overflow marks
  w.qimg_off = ((w.gtau_off + thr_words * 4 + 255) / 256) * 256;
  w.counts_off = w.qimg_off + (w.variant == TENSOR_PAIR ? size_t(w.geo.query_rows) * dim * 2 : 0);
  w.ovf_off = w.counts_off + (w.variant == TENSOR_PAIR ? size_t(w.geo.n_slabs) * w.umma_qstride * 4 : 0);
  w.total = w.ovf_off + (w.variant == TENSOR_PAIR ? size_t(w.umma_qstride) * 4 : 0) + 256;
  return w;
}

}  // namespace drs

using namespace drs;

extern "C" int drs_version(void) { return 100; }
extern "C" const char* drs_last_error(void) { return g_last_error.c_str(); }
extern "C" int drs_device_sm_count(void) { return sm_count(); }
extern "C" int drs_dense_last_kernel(void) { return g_last_dense_kernel; }

extern "C" unsigned long long drs_kernel_launches(void) { return g_launches.load(); }

extern "C" int drs_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = on != 0;
  for (auto& k : g_prof) k.used = 0;
  return DRS_OK;
}
extern "C" int drs_profile_read(int kind, double* total_ms, int* launches) {
  DRS_REQUIRE(kind >= 0 && kind < DRS_PROF_KINDS && total_ms && launches, DRS_ERR_ARG, "profile: bad arguments");
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfKind& k = g_prof[kind];
  double ms = 0.0;
  for (size_t i = 0; i < k.used; ++i) {
    DRS_CUDA_CHECK(cudaEventSynchronize(k.stop[i]));
    float t = 0.f;
    DRS_CUDA_CHECK(cudaEventElapsedTime(&t, k.start[i], k.stop[i]));
    ms += t;
  }
  *total_ms = ms;
  *launches = int(k.used);
  return DRS_OK;
}

extern "C" size_t drs_dense_workspace_bytes(int n_queries, int64_t n_rows, int dim, int kprime, int kernel) {
  if (n_queries < 0 || n_rows < 0 || kprime <= 0 || kprime > DRS_MAX_KPRIME) return 0;
  return dense_ws(n_queries, n_rows, dim, kprime, kernel).total;
}

extern "C" int drs_dense_topk(const float* d_queries_f32, int n_queries, const void* d_corpus_bf16,
                              const float* d_inv_norm, const uint8_t* d_row_mask, int64_t n_rows, int dim,
                              int kprime, int kernel, int escalate, float* d_out_scores, int64_t* d_out_ids,
                              int32_t* d_out_flags, void* d_workspace, size_t workspace_bytes, void* stream) {
  DRS_REQUIRE(d_queries_f32 && d_corpus_bf16 && d_inv_norm && d_out_scores && d_out_ids && d_workspace,
              DRS_ERR_ARG, "dense: null pointer");
  DRS_REQUIRE(n_queries >= 0 && n_rows > 0 && kprime > 0, DRS_ERR_ARG, "dense: bad sizes");
  DRS_REQUIRE(kprime <= DRS_MAX_KPRIME, DRS_ERR_UNSUPPORTED, "dense: k' exceeds DRS_MAX_KPRIME");
  DRS_REQUIRE(dim > 0 && dim % 8 == 0 && dim <= DRS_MAX_DIM, DRS_ERR_UNSUPPORTED,
              "dense: dim must be a multiple of 8 in (0, 1024]");
  DRS_REQUIRE(n_rows < (int64_t(1) << 32), DRS_ERR_UNSUPPORTED, "dense: more than 2^32-1 rows in one shard");
  DRS_REQUIRE((reinterpret_cast<uintptr_t>(d_corpus_bf16) & 15) == 0, DRS_ERR_ARG, "dense: corpus not 16-byte aligned");
  if (n_queries == 0) return DRS_OK;
  const int kk_s = cand_len(kprime);
  const int which = pick_dense_kernel(kernel, n_queries, n_rows, dim, kprime);
  DRS_REQUIRE(which > 0, DRS_ERR_UNSUPPORTED, "dense: tensor-core kernel does not support this shape");
  const DenseWs w = dense_ws(n_queries, n_rows, dim, kprime, kernel);
  DRS_REQUIRE(workspace_bytes >= w.total, DRS_ERR_WORKSPACE, "dense: workspace too small");
  const int kk = w.variant == TENSOR_PAIR ? cand_len_pair(kprime) : cand_len_umma(kprime);
  cudaStream_t st = (cudaStream_t)stream;
  uint64_t* parts = reinterpret_cast<uint64_t*>(d_workspace);
  uint64_t* cand = parts + w.parts_keys;
  uint64_t* scratch = w.scratch_keys ? cand + w.cand_keys : nullptr;
  int32_t* flags = d_out_flags ? d_out_flags
                               : reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(d_workspace) + w.flags_off);
  g_last_dense_kernel = (which == DRS_DENSE_UMMA && w.variant == TENSOR_PAIR) ? 3 : which;
  int rc;
  if (which == DRS_DENSE_SIMT) {
    {
      ProfScope ps(DRS_PROF_DENSE_SELECT, st);
      rc = launch_dense_simt(d_queries_f32, n_queries, d_corpus_bf16, d_inv_norm, d_row_mask, n_rows, dim, kk_s, parts,
                             n_queries, nullptr, st);
    }
    if (rc != DRS_OK) return rc;
    ProfScope pt(DRS_PROF_DENSE_TAIL, st);
    rc = launch_merge_keys(parts, w.simt_parts, n_queries, n_queries, kk_s, cand, nullptr, scratch, st);
    if (rc != DRS_OK) return rc;
    return launch_dense_rescore(d_queries_f32, n_queries, d_corpus_bf16, d_inv_norm, dim, cand, kk_s, kprime, 4e-6f,
                                d_out_scores, d_out_ids, flags, nullptr, DRS_QMODE_EXACT, 1, 0, st);
  }
  uint8_t* ws8 = reinterpret_cast<uint8_t*>(d_workspace);
  uint32_t* gtau = reinterpret_cast<uint32_t*>(ws8 + w.gtau_off);
  if (w.variant == TENSOR_PAIR) {
    uint32_t* counts = reinterpret_cast<uint32_t*>(ws8 + w.counts_off);
    int32_t* ovf = reinterpret_cast<int32_t*>(ws8 + w.ovf_off);
    {
      ProfScope ps(DRS_PROF_DENSE_SELECT, st);
      rc = launch_dense_pair(d_queries_f32, n_queries, d_corpus_bf16, d_inv_norm, d_row_mask, n_rows, dim, kk, parts,
                             counts, w.umma_qstride, gtau, ws8 + w.qimg_off, st);
    }
    if (rc != DRS_OK) return rc;
    ProfScope pt(DRS_PROF_DENSE_TAIL, st);
    rc = launch_merge_appended(parts, counts, w.geo.n_slabs, w.umma_qstride, w.geo.cap, n_queries, kk, cand,
                               gtau + w.umma_qstride, w.geo.pool_parts, ovf, st);
    if (rc != DRS_OK) return rc;
    rc = launch_dense_rescore(d_queries_f32, n_queries, d_corpus_bf16, d_inv_norm, dim, cand, kk, kprime, 2e-5f,
                              d_out_scores, d_out_ids, flags, nullptr, dense_umma_query_mode(), 1, 0, st, ovf);
  } else {
    {
      ProfScope ps(DRS_PROF_DENSE_SELECT, st);
      rc = launch_dense_umma(d_queries_f32, n_queries, d_corpus_bf16, d_inv_norm, d_row_mask, n_rows, dim, kk, parts,
                             w.umma_qstride, gtau, st);
    }
    if (rc != DRS_OK) return rc;
    ProfScope pt(DRS_PROF_DENSE_TAIL, st);
    // the kernel's own thresholds (gtau, and the per-slab pooled bounds when pooling was on) pre-filter the merge
    rc = launch_merge_keys_filtered(parts, w.umma_parts, w.umma_qstride, n_queries, kk, cand, nullptr, scratch, gtau,
                                    gtau + w.umma_qstride, w.pool_parts, w.umma_qstride, st);
    if (rc != DRS_OK) return rc;
    rc = launch_dense_rescore(d_queries_f32, n_queries, d_corpus_bf16, d_inv_norm, dim, cand, kk, kprime, 2e-5f,
                              d_out_scores, d_out_ids, flags, nullptr, dense_umma_query_mode(), 1, 0, st);
  }
  if (rc != DRS_OK || !escalate) return rc;
  // Re-run the (few) queries whose tensor-core result is not proven exact through the fp32 kernel.
  // Everything is predicated on flags[] on the device; with no flagged query the kernels exit at once.
  ProfScope pe(DRS_PROF_DENSE_TAIL, st);
  rc = launch_dense_simt(d_queries_f32, n_queries, d_corpus_bf16, d_inv_norm, d_row_mask, n_rows, dim, kk_s, parts,
                         n_queries, flags, st);
  if (rc != DRS_OK) return rc;
  rc = launch_merge_keys(parts, w.simt_parts, n_queries, n_queries, kk_s, cand, flags, scratch, st);
  if (rc != DRS_OK) return rc;
  return launch_dense_rescore(d_queries_f32, n_queries, d_corpus_bf16, d_inv_norm, dim, cand, kk_s, kprime, 4e-6f,
                              d_out_scores, d_out_ids, flags, flags, DRS_QMODE_EXACT, 3, 2, st);
}
