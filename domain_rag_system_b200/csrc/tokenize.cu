// Query tokenisation and vocabulary lookup on the device.
//
// Replaces, for ASCII text, BM25Indexer.tokenize (src/indexing/bm25_indexer.py:50-78: text.lower(), re.findall(r'\b\w+\b'),
// tokens of one character kept only when they are abbreviation keys) followed by the vocabulary lookups
// rank_bm25.BM25Okapi.get_scores performs per query token (idf.get(token), doc_freqs[...].get(token)): a batch of query
// strings goes in as one byte buffer, term ids come out in the layout drs_bm25_topk reads (q_terms, q_offsets).
//
// For ASCII input the reference's semantics are byte-level: lower() maps A-Z to a-z, \w is [A-Za-z0-9_], and
// \b\w+\b yields the maximal runs of word characters.  Text with any byte >= 0x80 is Unicode-defined (str.lower()
// and \w know thousands of code points) and is NOT handled here: the host caller routes such strings through the
// Python tokeniser (indexing/bm25_index.py: encode_texts).
//
// Lookup: open-addressing table keyed by 64-bit FNV-1a of the token, entries placed in order of their home slot with
// no wrap-around (the table has a tail region), so a probe walks forward from the home slot until it meets the token
// or an empty slot.  A hash match is confirmed by comparing the token's bytes with the vocabulary string -- the
// result is exact, not probabilistic.  One thread per query string: queries are short and the whole batch is a few
// hundred KB; the kernel exists to take the per-token Python work off the host path, not to saturate anything.
#include "common.cuh"

namespace drs {

__device__ __forceinline__ bool is_word(uint8_t c) {
  return (c >= 'a' && c <= 'z') || (c >= '0' && c <= '9') || c == '_' || (c >= 'A' && c <= 'Z');
}
__device__ __forceinline__ uint8_t lower(uint8_t c) { return (c >= 'A' && c <= 'Z') ? uint8_t(c + 32) : c; }

// FILL = false: counts[q] = number of tokens the reference's tokeniser yields for text q.
// FILL = true : the tokens' term ids (or -1: not in the vocabulary) are written at out_terms[out_offsets[q] ..).
template <bool FILL>
__global__ void __launch_bounds__(128) tokenize_kernel(const uint8_t* __restrict__ text,
                                                       const int64_t* __restrict__ text_offsets, int n_texts,
                                                       drs_single_chars keep1, drs_vocab_table tab,
                                                       int32_t* __restrict__ counts,
                                                       const int32_t* __restrict__ out_offsets,
                                                       int32_t* __restrict__ out_terms) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_texts) return;
  const int64_t a = text_offsets[q], e = text_offsets[q + 1];
  int n = 0;
  int32_t* out = FILL ? out_terms + out_offsets[q] : nullptr;
  int64_t i = a;
  while (i < e) {
    while (i < e && !is_word(text[i])) ++i;
    if (i >= e) break;
    const int64_t t0 = i;
    uint64_t h = 0xcbf29ce484222325ull;                  // FNV-1a 64
    while (i < e && is_word(text[i])) {
      h = (h ^ uint64_t(lower(text[i]))) * 0x100000001b3ull;
      ++i;
    }
    const int64_t len = i - t0;
    if (len == 1) {   // kept only when it is an abbreviation key (bm25_indexer.py:72-76)
      const uint8_t c = lower(text[t0]);
      if (!((keep1.bits[c >> 5] >> (c & 31)) & 1u)) continue;
    }
    if (FILL) {
      int32_t term = -1;
      if (tab.capacity > 0) {
        for (int64_t s = int64_t(h & uint64_t(tab.capacity - 1)); s < tab.n_slots; ++s) {
          const int32_t t = tab.d_slot_term[s];
          if (t < 0) break;                              // empty slot: the token is not in the vocabulary
          if (tab.d_slot_hash[s] != h) continue;
          const int64_t so = tab.d_str_offsets[t];
          if (tab.d_str_offsets[t + 1] - so != len) continue;
          bool same = true;
          for (int64_t k = 0; k < len && same; ++k) same = tab.d_pool[so + k] == lower(text[t0 + k]);
          if (same) {
            term = t;
            break;
          }
        }
      }
      out[n] = term;
    }
    ++n;
  }
  if (!FILL) counts[q] = n;
}

}  // namespace drs

using namespace drs;

extern "C" int drs_tokenize_count(const uint8_t* d_text, const int64_t* d_text_offsets, int n_texts,
                                  const drs_single_chars* keep1, int32_t* d_counts, void* stream) {
  DRS_REQUIRE(n_texts >= 0 && keep1 != nullptr, DRS_ERR_ARG, "tokenize: bad arguments");
  if (n_texts == 0) return DRS_OK;
  drs_vocab_table none = {};
  tokenize_kernel<false><<<(n_texts + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_text, d_text_offsets, n_texts, *keep1,
                                                                                   none, d_counts, nullptr, nullptr);
  count_launches(1);
  DRS_CUDA_CHECK(cudaGetLastError());
  return DRS_OK;
}

extern "C" int drs_tokenize_lookup(const uint8_t* d_text, const int64_t* d_text_offsets, int n_texts,
                                   const drs_single_chars* keep1, const drs_vocab_table* table,
                                   const int32_t* d_out_offsets, int32_t* d_out_terms, void* stream) {
  DRS_REQUIRE(n_texts >= 0 && keep1 != nullptr && table != nullptr, DRS_ERR_ARG, "tokenize: bad arguments");
  DRS_REQUIRE(table->capacity == 0 || (table->capacity & (table->capacity - 1)) == 0, DRS_ERR_ARG,
              "tokenize: table capacity must be a power of two");
  if (n_texts == 0) return DRS_OK;
  tokenize_kernel<true><<<(n_texts + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_text, d_text_offsets, n_texts, *keep1,
                                                                                  *table, nullptr, d_out_offsets, d_out_terms);
  count_launches(1);
  DRS_CUDA_CHECK(cudaGetLastError());
  return DRS_OK;
}
