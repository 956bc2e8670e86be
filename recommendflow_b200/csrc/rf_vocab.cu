// rf_vocab.cu -- vocabulary lookup and bucketisation feeding the bag kernel with ids
// (SURVEY.md §8f rank 4).
//
// Replaces, for LookupEmbedding / DiscreteEmbedding (backend/layers/preprocess_layers.py:134-200):
//   * Keras StringLookup / IntegerLookup(vocabulary=vocabs, output_mode="int") as built at
//     preprocess_layers.py:148-150: term i of the vocabulary -> index i + 1, anything else -> 0
//     (one OOV bucket at index 0, no mask token);
//   * Keras Discretization(bin_boundaries=vocabs) (preprocess_layers.py:187): index = number of
//     boundaries <= x, i.e. upper_bound with the predicate `x < boundary` (NaN -> n_boundaries).
//
// The vocabulary is an open-addressing table in HBM built on the device: slot = (hash >> 32) << 32
// | term index, position = low hash bits, linear probing, empty = all ones.  A lookup hashes the
// key (FarmHash Fingerprint64 for strings -- the routine the bag kernel uses -- or a 64-bit mixer
// for ints), probes, and on a tag match compares the key with the term itself, so the result is
// exact whatever the hash does.  One thread per key; both kernels are latency/HBM-bound integer
// work, no shared memory needed (the vocabularies of the reference's configs have <= 1e5 terms and
// live in L2).
#include <atomic>

#include "rf_common.h"
#include "rf_hash.cuh"

namespace rf {
extern std::atomic<int64_t> g_launches;

namespace {

constexpr unsigned long long kEmptySlot = ~0ull;
constexpr int kVocabThreads = 256;

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 30;
    x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27;
    x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}

struct DevVocab {
    const uint8_t *term_bytes;
    const int32_t *term_offsets;
    const int64_t *term_ints;
    unsigned long long *slots;
    uint32_t mask;
    int32_t n_terms;
};

__device__ __forceinline__ uint64_t hash_bytes(const uint8_t *arena, uint32_t begin, uint32_t len) {
    const uintptr_t addr = reinterpret_cast<uintptr_t>(arena) + begin;
    WordSrcGlobal src{reinterpret_cast<const uint32_t *>(addr & ~(uintptr_t)3), (uint32_t)(addr & 3)};
    return fingerprint64(src, len);
}

__global__ void __launch_bounds__(kVocabThreads) vocab_insert_kernel(DevVocab v) {
    const int i = blockIdx.x * kVocabThreads + threadIdx.x;
    if (i >= v.n_terms) return;
    uint64_t h;
    if (v.term_ints != nullptr) {
        h = mix64((uint64_t)v.term_ints[i]);
    } else {
        const uint32_t b = (uint32_t)v.term_offsets[i];
        h = hash_bytes(v.term_bytes, b, (uint32_t)v.term_offsets[i + 1] - b);
    }
    const unsigned long long slot = (h & 0xffffffff00000000ull) | (uint32_t)i;
    uint32_t pos = (uint32_t)h & v.mask;
    while (atomicCAS(v.slots + pos, kEmptySlot, slot) != kEmptySlot) pos = (pos + 1) & v.mask;
}

__global__ void __launch_bounds__(kVocabThreads)
vocab_lookup_strings_kernel(DevVocab v, const uint8_t *__restrict__ bytes, const int32_t *__restrict__ offs, int64_t n,
                            int64_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * kVocabThreads + threadIdx.x;
    if (i >= n) return;
    const uint32_t b = (uint32_t)offs[i], len = (uint32_t)offs[i + 1] - b;
    const uint64_t h = hash_bytes(bytes, b, len);
    const uint32_t tag = (uint32_t)(h >> 32);
    uint32_t pos = (uint32_t)h & v.mask;
    int64_t id = 0;
    for (;;) {
        const unsigned long long s = v.slots[pos];
        if (s == kEmptySlot) break;
        if ((uint32_t)(s >> 32) == tag) {
            const uint32_t j = (uint32_t)s;
            const uint32_t tb = (uint32_t)v.term_offsets[j];
            if ((uint32_t)v.term_offsets[j + 1] - tb == len) {
                uint32_t k = 0;
                while (k < len && __ldg(bytes + b + k) == __ldg(v.term_bytes + tb + k)) ++k;
                if (k == len) {
                    id = (int64_t)j + 1;
                    break;
                }
            }
        }
        pos = (pos + 1) & v.mask;
    }
    out[i] = id;
}

__global__ void __launch_bounds__(kVocabThreads)
vocab_lookup_int64_kernel(DevVocab v, const int64_t *__restrict__ values, int64_t n, int64_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * kVocabThreads + threadIdx.x;
    if (i >= n) return;
    const int64_t x = values[i];
    const uint64_t h = mix64((uint64_t)x);
    const uint32_t tag = (uint32_t)(h >> 32);
    uint32_t pos = (uint32_t)h & v.mask;
    int64_t id = 0;
    for (;;) {
        const unsigned long long s = v.slots[pos];
        if (s == kEmptySlot) break;
        if ((uint32_t)(s >> 32) == tag && v.term_ints[(uint32_t)s] == x) {
            id = (int64_t)(uint32_t)s + 1;
            break;
        }
        pos = (pos + 1) & v.mask;
    }
    out[i] = id;
}

__global__ void __launch_bounds__(kVocabThreads)
bucketize_kernel(const float *__restrict__ x, int64_t n, const float *__restrict__ edges, int32_t n_edges,
                 int64_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * kVocabThreads + threadIdx.x;
    if (i >= n) return;
    const float v = x[i];
    int lo = 0, hi = n_edges;                 // first edge with v < edge
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (v < __ldg(edges + mid)) hi = mid; else lo = mid + 1;
    }
    out[i] = lo;
}

int check_vocab(const rf_vocab_desc *d, DevVocab &v) {
    if (!d) return set_error(RF_ERR_INVALID, "vocabulary descriptor is NULL");
    if (d->n_terms < 0 || d->n_terms > (int64_t)1 << 30) return set_error(RF_ERR_INVALID, "n_terms %lld out of range", (long long)d->n_terms);
    if (d->capacity < 2 || d->capacity > (int64_t)1 << 31 || (d->capacity & (d->capacity - 1)) != 0 || d->capacity < 2 * d->n_terms)
        return set_error(RF_ERR_INVALID, "capacity %lld must be a power of two >= max(2, 2 * n_terms)", (long long)d->capacity);
    if (!d->slots) return set_error(RF_ERR_INVALID, "slots is NULL");
    const bool strings = d->term_bytes != nullptr || d->term_offsets != nullptr;
    if (d->n_terms > 0) {
        if (strings == (d->term_ints != nullptr))
            return set_error(RF_ERR_INVALID, "give either term_bytes + term_offsets or term_ints");
        if (strings && (!d->term_bytes || !d->term_offsets)) return set_error(RF_ERR_INVALID, "term_bytes / term_offsets is NULL");
    }
    v.term_bytes = d->term_bytes;
    v.term_offsets = d->term_offsets;
    v.term_ints = d->term_ints;
    v.slots = reinterpret_cast<unsigned long long *>(d->slots);
    v.mask = (uint32_t)(d->capacity - 1);
    v.n_terms = (int32_t)d->n_terms;
    return RF_OK;
}

inline unsigned blocks_for(int64_t n) { return (unsigned)((n + kVocabThreads - 1) / kVocabThreads); }

}  // namespace
}  // namespace rf

using namespace rf;

extern "C" {

int rf_vocab_build(const rf_vocab_desc *vocab, void *stream) {
    DevVocab v;
    int rc = check_vocab(vocab, v);
    if (rc != RF_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    RF_CUDA(cudaMemsetAsync(v.slots, 0xff, sizeof(unsigned long long) * (size_t)vocab->capacity, st));
    if (v.n_terms == 0) return RF_OK;
    vocab_insert_kernel<<<blocks_for(v.n_terms), kVocabThreads, 0, st>>>(v);
    RF_CUDA(cudaGetLastError());
    g_launches.fetch_add(1);
    return RF_OK;
}

int rf_vocab_lookup_strings(const rf_vocab_desc *vocab, const uint8_t *d_bytes, const int32_t *d_str_offsets, int64_t n_items,
                            int64_t *d_ids_out, void *stream) {
    DevVocab v;
    int rc = check_vocab(vocab, v);
    if (rc != RF_OK) return rc;
    if (n_items < 0 || n_items > INT32_MAX) return set_error(RF_ERR_INVALID, "n_items out of range");
    if (n_items == 0) return RF_OK;
    if (v.n_terms > 0 && v.term_ints) return set_error(RF_ERR_INVALID, "string keys looked up in an integer vocabulary");
    if (!d_bytes || !d_str_offsets || !d_ids_out) return set_error(RF_ERR_INVALID, "bytes / str_offsets / ids_out is NULL");
    vocab_lookup_strings_kernel<<<blocks_for(n_items), kVocabThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        v, d_bytes, d_str_offsets, n_items, d_ids_out);
    RF_CUDA(cudaGetLastError());
    g_launches.fetch_add(1);
    return RF_OK;
}

int rf_vocab_lookup_int64(const rf_vocab_desc *vocab, const int64_t *d_values, int64_t n_items, int64_t *d_ids_out,
                          void *stream) {
    DevVocab v;
    int rc = check_vocab(vocab, v);
    if (rc != RF_OK) return rc;
    if (n_items < 0 || n_items > INT32_MAX) return set_error(RF_ERR_INVALID, "n_items out of range");
    if (n_items == 0) return RF_OK;
    if (v.n_terms > 0 && !v.term_ints) return set_error(RF_ERR_INVALID, "integer keys looked up in a string vocabulary");
    if (!d_values || !d_ids_out) return set_error(RF_ERR_INVALID, "values / ids_out is NULL");
    vocab_lookup_int64_kernel<<<blocks_for(n_items), kVocabThreads, 0, static_cast<cudaStream_t>(stream)>>>(v, d_values, n_items,
                                                                                                           d_ids_out);
    RF_CUDA(cudaGetLastError());
    g_launches.fetch_add(1);
    return RF_OK;
}

int rf_bucketize_f32(const float *d_values, int64_t n_items, const float *d_boundaries, int32_t n_boundaries,
                     int64_t *d_ids_out, void *stream) {
    if (n_items < 0 || n_items > INT32_MAX) return set_error(RF_ERR_INVALID, "n_items out of range");
    if (n_boundaries < 0) return set_error(RF_ERR_INVALID, "n_boundaries is negative");
    if (n_items == 0) return RF_OK;
    if (!d_values || !d_ids_out || (n_boundaries > 0 && !d_boundaries))
        return set_error(RF_ERR_INVALID, "values / boundaries / ids_out is NULL");
    bucketize_kernel<<<blocks_for(n_items), kVocabThreads, 0, static_cast<cudaStream_t>(stream)>>>(d_values, n_items, d_boundaries,
                                                                                                   n_boundaries, d_ids_out);
    RF_CUDA(cudaGetLastError());
    g_launches.fetch_add(1);
    return RF_OK;
}

}  // extern "C"
