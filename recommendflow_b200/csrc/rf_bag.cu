// rf_bag.cu -- fused hash + gather + pool for sm_100a (B200), and its C-ABI.
//
// One launch covers every feature field of a batch.  It replaces, per field, the op chain
//   Hashing x T -> Embedding gather x T -> reduce(axis=1) x T -> concat
// of /root/reference/backend/layers/preprocess_layers.py:94-97 (DoubleHashingEmbedding.call,
// T = 2) and :66-68 (EmbeddingBag.call), and, across fields, the per-feature Python loop of
// models/matching/que2search.py:68,76-79.  The [B, L, D] gather result is never materialised.
//
// Work decomposition (HBM-bound byte/integer work -- no tensor cores on this path):
//   tile  = (field, range of bags), ~kChunk keys; one CTA per tile, all fields in one grid.
//   round = <= kChunk keys whose bags fit the round (a single bag longer than kChunk is
//           walked in sub-rounds with the accumulators kept in registers).
//   phase A: key offsets -> smem (coalesced), key bytes -> smem (16-byte vector copies of
//            the tile's contiguous slice of the string arena).
//   phase B: one key per thread: FarmHash64 / SipHash-2-4 out of shared memory, bucket via a
//            multiply-high reciprocal, ids -> smem (and optionally to global as int64).
//   phase C: one lane group per (bag, table): ids broadcast from smem, rows fetched with
//            128-bit loads (4-8 rows in flight per lane), pooled in bag order in registers,
//            one coalesced 128-bit store per lane.  Pooling is sequential in index order so
//            fp32 results are bit-identical to the oracle's.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rf_b200.h"
#include "rf_common.h"
#include "rf_hash.cuh"

namespace rf {

constexpr int kThreads = 256;
constexpr int kChunk = 2048;              // keys (bucket ids) held per round
constexpr int kSub = 1024;                // keys hashed per sub-chunk of a round (offsets + staged bytes)
constexpr int kStageBytes = 16 * 1024;    // key bytes staged per sub-chunk
constexpr int kMaxTileBags = 512;         // bags per tile (jagged: begin / end of every bag kept in smem)
constexpr int kMaxFieldsSmem = 512;       // tile-prefix table kept in smem up to this many fields
constexpr int kMaxDim = 512;

struct DevTable {
    const float *w;
    HashSpec h;
};

struct DevField {
    const uint8_t *bytes;
    const int32_t *soffs;
    const int64_t *ints;
    const int64_t *ids_in;
    const int32_t *boffs;
    const int32_t *bends;   // optional: bag b = [boffs[b], bends[b]) instead of [boffs[b], boffs[b+1])
    float *out;
    int64_t *ids_out;
    int64_t out_stride;
    int64_t n_items;
    int64_t int_mask;
    int32_t batch;
    int32_t bag_len;
    int32_t dim;
    int32_t n_tables;
    int32_t combiner;
    int32_t mask_mode;
    int32_t bags_per_tile;
    int32_t tile_begin;
    int32_t vec_ok;      // 128-bit path usable (dim % 4 == 0, aligned pointers/strides)
    int32_t flags;
    uint32_t mask_words[RF_MAX_MASK_BYTES / 4];   // RF_MASK_STRING_VALUE: the mask string, zero padded
    int32_t mask_len;
    int32_t pad_;
    DevTable t[RF_MAX_TABLES_PER_FIELD];
};

// key == mask string?  (RF_MASK_STRING_VALUE; lengths already equal)
template <class Src>
__device__ __forceinline__ bool equals_mask(const Src &src, uint32_t len, const uint32_t *mask_words) {
    bool same = true;
    for (uint32_t p = 0; p < len; p += 4) {
        const uint32_t rem = len - p;
        const uint32_t keep = rem >= 4 ? 0xffffffffu : ((1u << (rem * 8)) - 1u);
        same = same && ((fetch32(src, p) & keep) == (mask_words[p >> 2] & keep));
    }
    return same;
}

// ------------------------------------------------------------------------------------------
// pooling primitives.  Two code shapes only: ADD (sum / avg) and SELECT (min / max, picked by a
// launch-uniform flag) -- fewer template instances keeps the fused kernel small.
// ------------------------------------------------------------------------------------------
constexpr int kAdd = 0, kSelect = 1;

struct PoolOp {
    int combiner;
    int partial;   // RF_FIELD_PARTIAL: an empty bag yields the combiner's identity, not 0
    __device__ __forceinline__ bool is_max() const { return combiner == RF_COMBINER_MAX; }
    __device__ __forceinline__ bool is_avg() const { return combiner == RF_COMBINER_AVG; }
    template <int K>
    __device__ __forceinline__ float init() const {
        if (K == kAdd) return 0.0f;
        return is_max() ? __int_as_float(0xff800000) : __int_as_float(0x7f800000);
    }
    template <int K>
    __device__ __forceinline__ float apply(float acc, float x) const {
        if (K == kAdd) return acc + x;
        const bool take = is_max() ? (x > acc) : (x < acc);     // same select as the oracle
        return take ? x : acc;
    }
    template <int K>
    __device__ __forceinline__ void apply4(float4 &a, const float4 &x) const {
        a.x = apply<K>(a.x, x.x);
        a.y = apply<K>(a.y, x.y);
        a.z = apply<K>(a.z, x.z);
        a.w = apply<K>(a.w, x.w);
    }
    template <int K>
    __device__ __forceinline__ float finish(float acc, int64_t count) const {
        if (count == 0) return (K == kSelect && partial) ? acc : 0.0f;
        if (K == kAdd && is_avg()) return acc / (float)count;    // IEEE fp32 divide, like sum / L
        return acc;
    }
    template <int K>
    __device__ __forceinline__ void finish4(float4 &a, int64_t count) const {
        a.x = finish<K>(a.x, count);
        a.y = finish<K>(a.y, count);
        a.z = finish<K>(a.z, count);
        a.w = finish<K>(a.w, count);
    }
};

__device__ __forceinline__ float4 ldg_row(const float4 *p) { return __ldg(p); }

// Output of one pooled vector.  ACC = false: plain store.  ACC = true (RF_FIELD_ACCUMULATE): fire-and-forget
// reduction into the destination (red.global.add: local HBM/L2, or a peer GPU's memory over NVLink) -- the
// owners of a row-sharded table add their partial pools straight into the source rank's zeroed buffer, so no
// per-owner partial buffers and no combine pass exist.  The order in which owners land is free: this mode is
// not bit-reproducible (the ordered rf_combine_partials path stays the parity path).
template <bool ACC>
__device__ __forceinline__ void emit4(float4 *dst, const float4 &v, bool nonempty) {
    if (ACC) {
        if (nonempty)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    } else {
        *dst = v;
    }
}
template <bool ACC>
__device__ __forceinline__ void emit1(float *dst, float v, bool nonempty) {
    if (ACC) {
        if (nonempty) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst), "f"(v) : "memory");
    } else {
        *dst = v;
    }
}

// Accumulate `cnt` rows (ids in shared memory) into acc[], in index order.  Every trip issues up
// to U independent 128-bit row loads before the first add, also on the last (partial) trip: a
// bag's time is ceil(cnt / U) DRAM latencies, never one latency per leftover key.
template <int NV, int K>
__device__ __forceinline__ void accumulate_vec(const PoolOp &op, float4 (&acc)[NV], const float4 *__restrict__ W,
                                               uint32_t row_vecs, uint32_t lg, uint32_t G, const uint32_t *sid, int cnt) {
    constexpr int U = NV == 1 ? 8 : (NV == 2 ? 4 : 2);
    for (int i = 0; i < cnt; i += U) {
        float4 r[U][NV];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            // past the end: re-read the trip's first row (an L1 hit) so that every load is
            // unconditional and independent; only the pooling below is predicated
            const float4 *row = W + (size_t)sid[i + u < cnt ? i + u : i] * row_vecs;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const uint32_t c = lg + v * G;
                r[u][v] = ldg_row(row + ((NV == 1 || c < row_vecs) ? c : lg));
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (i + u < cnt) {
#pragma unroll
                for (int v = 0; v < NV; ++v) op.apply4<K>(acc[v], r[u][v]);
            }
        }
    }
}

struct Round {
    int bag0, bag1;        // bags of this round, relative to the tile's first bag
    int64_t item0;         // first key of the round (field-flat index)
    int n_keys;            // keys hashed this round
    bool partial;          // true: one long bag walked in sub-rounds
    bool first_part, last_part;
};

// ------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------
struct Smem {
    uint32_t ids[RF_MAX_TABLES_PER_FIELD * kChunk];
    int32_t soff[kSub + 4];
    uint32_t stage[kStageBytes / 4 + 8];
    int32_t bbeg[kMaxTileBags + 4];   // jagged mode: first key of each bag of the tile
    int32_t bend[kMaxTileBags + 4];   // ... and one past its last key
    int32_t tile_begin[kMaxFieldsSmem + 1];
    DevField field;   // this tile's descriptor, copied once per tile
    // partial pools of a bag longer than one round (only lane groups 0..T-1 are active then):
    // vec path float4[(t*32 + lane)*4 + v]; scalar path float[(t*32 + lane)*16 + slab]
    float4 carry[RF_MAX_TABLES_PER_FIELD * 32 * 4];
    int next_work;    // phase C: lane groups pull (bag, table) work items from here (jagged balance)
};
static_assert(sizeof(Smem) <= 48 * 1024, "static shared memory limit");
static_assert(kMaxDim / 32 == 16 && sizeof(DevField) % 4 == 0, "DevField is copied word-wise");

template <int NV, int K, bool ACC>
__device__ __forceinline__ void pool_round_vec(const PoolOp &op, const DevField &F, Smem &sm, const Round &R,
                                               int tile_bag0) {
    const uint32_t row_vecs = (uint32_t)F.dim >> 2;
    uint32_t G = 1;
    while (G < row_vecs && G < 32) G <<= 1;
    const uint32_t lg = threadIdx.x & (G - 1);
    const uint32_t grp = threadIdx.x / G;
    const uint32_t n_grp = kThreads / G;
    const int T = F.n_tables;
    const bool dense = F.boffs == nullptr;
    const int n_work = (R.bag1 - R.bag0) * T;
    // bags differ in length, so lane groups pull work items instead of striding over them
    const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31u) & ~(G - 1)));
    (void)grp;
    (void)n_grp;
    for (;;) {
        int w = 0;
        if (lg == 0) w = atomicAdd(&sm.next_work, 1);
        w = __shfl_sync(gmask, w, 0, G);
        if (w >= n_work) break;
        const int bl = R.bag0 + (T == 2 ? (w >> 1) : w);
        const int t = T == 2 ? (w & 1) : 0;
        int64_t lo, hi;   // field-flat key range of the bag
        if (dense) {
            lo = (int64_t)(tile_bag0 + bl) * F.bag_len;
            hi = lo + F.bag_len;
        } else {
            lo = sm.bbeg[bl];
            hi = sm.bend[bl];
        }
        int rel = (int)(lo - R.item0), cnt = (int)(hi - lo);
        if (R.partial) {   // the round holds a slice [item0, item0 + n_keys) of this one bag
            rel = 0;
            cnt = R.n_keys;
        }
        float4 acc[NV];
        float4 *carry = sm.carry + (t * 32 + lg) * 4;
        if (R.partial && !R.first_part) {
#pragma unroll
            for (int v = 0; v < NV; ++v) acc[v] = carry[v];
        } else {
            const float z = op.init<K>();
#pragma unroll
            for (int v = 0; v < NV; ++v) acc[v] = make_float4(z, z, z, z);
        }
        if (lg < row_vecs)
            accumulate_vec<NV, K>(op, acc, reinterpret_cast<const float4 *>(F.t[t].w), row_vecs, lg, G,
                                  sm.ids + t * kChunk + rel, cnt);
        if (R.partial && !R.last_part) {
#pragma unroll
            for (int v = 0; v < NV; ++v) carry[v] = acc[v];
            continue;
        }
        float4 *o = reinterpret_cast<float4 *>(F.out + (int64_t)(tile_bag0 + bl) * F.out_stride + (int64_t)t * F.dim);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const uint32_t c = lg + v * G;
            op.finish4<K>(acc[v], hi - lo);
            if (c < row_vecs) emit4<ACC>(o + c, acc[v], hi > lo);
        }
    }
}

// Dense bags of exactly L keys (L in {1,2,3,4}), D <= 128: a lane group fetches NB bags together,
// keeping NB * L (6..8) independent 128-bit row loads in flight per lane instead of L, because
// this phase is bound by DRAM latency x dependent rounds.  Same in-order pooling per bag.
// With T == 2 the table index of a lane group never changes (its work items keep their parity).
template <int K, int NB, int L, bool ACC>
__device__ __forceinline__ void pool_round_short(const PoolOp &op, const DevField &F, const Smem &sm, const Round &R,
                                                 int tile_bag0) {
    const uint32_t row_vecs = (uint32_t)F.dim >> 2;
    uint32_t G = 1;
    while (G < row_vecs) G <<= 1;
    const uint32_t lg = threadIdx.x & (G - 1);
    if (lg >= row_vecs) return;
    const int grp = threadIdx.x / G;
    const int n_grp = kThreads / G;                       // even: the parity of w is the parity of grp
    const int T = F.n_tables;
    const int n_work = (R.bag1 - R.bag0) * T;
    const int t = T == 2 ? (grp & 1) : 0;
    const int wshift = T == 2 ? 1 : 0;
    const float4 *__restrict__ W = reinterpret_cast<const float4 *>(F.t[t].w) + lg;
    const uint32_t *sid_base = sm.ids + t * kChunk;
    float *out_base = F.out + (int64_t)(tile_bag0 + R.bag0) * F.out_stride + (int64_t)t * F.dim + lg * 4;
    const int64_t ostride = F.out_stride;
    int w = grp;
    for (; w + (NB - 1) * n_grp < n_work; w += NB * n_grp) {
        float4 r[NB][L];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const uint32_t *sid = sid_base + ((w + b * n_grp) >> wshift) * L;
#pragma unroll
            for (int u = 0; u < L; ++u) r[b][u] = ldg_row(W + (size_t)sid[u] * row_vecs);
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const float z = op.init<K>();
            float4 acc = make_float4(z, z, z, z);
#pragma unroll
            for (int u = 0; u < L; ++u) op.apply4<K>(acc, r[b][u]);
            op.finish4<K>(acc, L);
            emit4<ACC>(reinterpret_cast<float4 *>(out_base + (int64_t)((w + b * n_grp) >> wshift) * ostride), acc, true);
        }
    }
    for (; w < n_work; w += n_grp) {
        const uint32_t *sid = sid_base + (w >> wshift) * L;
        float4 r[L];
#pragma unroll
        for (int u = 0; u < L; ++u) r[u] = ldg_row(W + (size_t)sid[u] * row_vecs);
        const float z = op.init<K>();
        float4 acc = make_float4(z, z, z, z);
#pragma unroll
        for (int u = 0; u < L; ++u) op.apply4<K>(acc, r[u]);
        op.finish4<K>(acc, L);
        emit4<ACC>(reinterpret_cast<float4 *>(out_base + (int64_t)(w >> wshift) * ostride), acc, true);
    }
}

// Generic-D path (dim % 4 != 0 or unaligned): one warp per (bag, table), one float per lane per
// 32-column slab, same in-order accumulation.  Long bags carry kMaxDim/32 partial slabs.
template <int K, bool ACC>
__device__ __forceinline__ void pool_round_scalar(const PoolOp &op, const DevField &F, Smem &sm, const Round &R,
                                                  int tile_bag0) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warp = kThreads / 32;
    const int T = F.n_tables;
    const bool dense = F.boffs == nullptr;
    const int n_work = (R.bag1 - R.bag0) * T;
    const int D = F.dim;
    for (int w = warp; w < n_work; w += n_warp) {
        const int bl = R.bag0 + (T == 2 ? (w >> 1) : w);
        const int t = T == 2 ? (w & 1) : 0;
        int64_t lo, hi;
        if (dense) {
            lo = (int64_t)(tile_bag0 + bl) * F.bag_len;
            hi = lo + F.bag_len;
        } else {
            lo = sm.bbeg[bl];
            hi = sm.bend[bl];
        }
        int rel = (int)(lo - R.item0), cnt = (int)(hi - lo);
        if (R.partial) {
            rel = 0;
            cnt = R.n_keys;
        }
        const float *W = F.t[t].w;
        const uint32_t *sid = sm.ids + t * kChunk + rel;
        float *o = F.out + (int64_t)(tile_bag0 + bl) * F.out_stride + (int64_t)t * D;
        float *carry = reinterpret_cast<float *>(sm.carry) + (t * 32 + lane) * (kMaxDim / 32);
        for (int s = 0; s * 32 < D; ++s) {
            const int d = s * 32 + (int)lane;
            float acc = (R.partial && !R.first_part) ? carry[s] : op.init<K>();
            if (d < D)
                for (int i = 0; i < cnt; ++i) acc = op.apply<K>(acc, __ldg(W + (size_t)sid[i] * D + d));
            if (R.partial && !R.last_part) {
                carry[s] = acc;
                continue;
            }
            if (d < D) emit1<ACC>(o + d, op.finish<K>(acc, hi - lo), hi > lo);
        }
    }
}

template <int K, bool ACC>
__device__ __noinline__ void pool_round(const PoolOp &op, const DevField &F, Smem &sm, const Round &R, int tile_bag0) {
    if (F.vec_ok) {
        const int row_vecs = F.dim >> 2;
        if (row_vecs <= 32 && !R.partial && F.boffs == nullptr && F.bag_len >= 1 && F.bag_len <= 4) {
            if (F.bag_len == 1) pool_round_short<K, 8, 1, ACC>(op, F, sm, R, tile_bag0);
            else if (F.bag_len == 2) pool_round_short<K, 4, 2, ACC>(op, F, sm, R, tile_bag0);
            else if (F.bag_len == 3) pool_round_short<K, 2, 3, ACC>(op, F, sm, R, tile_bag0);
            else pool_round_short<K, 2, 4, ACC>(op, F, sm, R, tile_bag0);
        } else if (row_vecs <= 32) {
            pool_round_vec<1, K, ACC>(op, F, sm, R, tile_bag0);
        } else if (row_vecs <= 64) {
            pool_round_vec<2, K, ACC>(op, F, sm, R, tile_bag0);
        } else {
            pool_round_vec<4, K, ACC>(op, F, sm, R, tile_bag0);
        }
    } else {
        pool_round_scalar<K, ACC>(op, F, sm, R, tile_bag0);
    }
}

template <bool ACC>
__device__ __forceinline__ void bag_forward_body(const DevField *__restrict__ fields, int n_fields, int total_tiles) {
    __shared__ Smem sm;      // 46.4 KB static: 3 CTAs/SM leave most of the 228 KB carve-out to L1
    const int tid = threadIdx.x;

    const bool prefix_in_smem = n_fields <= kMaxFieldsSmem;
    if (prefix_in_smem) {
        for (int i = tid; i < n_fields; i += kThreads) sm.tile_begin[i] = fields[i].tile_begin;
        if (tid == 0) sm.tile_begin[n_fields] = total_tiles;
        __syncthreads();
    }

    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        // ---- which field does this tile belong to (largest f with tile_begin[f] <= tile) -----
        int f_lo = 0, f_hi = n_fields - 1;
        while (f_lo < f_hi) {
            const int mid = (f_lo + f_hi + 1) >> 1;
            const int tb = prefix_in_smem ? sm.tile_begin[mid] : fields[mid].tile_begin;
            if (tb <= tile) f_lo = mid; else f_hi = mid - 1;
        }
        // One round trip: every thread reads the few words it needs for the tile geometry
        // straight from the (L2-resident) descriptor while the descriptor itself and the tile's
        // CSR slice are copied into shared memory for the phases below.
        const DevField &G = fields[f_lo];
        const int tile_bag0 = (tile - G.tile_begin) * G.bags_per_tile;
        const int tile_nbags = min(G.bags_per_tile, G.batch - tile_bag0);
        const int32_t *g_boffs = G.boffs;
        const int32_t *g_bends = G.bends;
        const bool dense = g_boffs == nullptr;
        {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(&G);
            uint32_t *dst = reinterpret_cast<uint32_t *>(&sm.field);
            for (int i = tid; i < (int)(sizeof(DevField) / 4); i += kThreads) dst[i] = src[i];
            if (!dense)
                for (int i = tid; i < tile_nbags; i += kThreads) {
                    sm.bbeg[i] = g_boffs[tile_bag0 + i];
                    sm.bend[i] = g_bends ? g_bends[tile_bag0 + i] : g_boffs[tile_bag0 + i + 1];
                }
        }
        __syncthreads();
        const DevField &F = sm.field;
        const int T = F.n_tables;
        // Bags with explicit ends may have gaps between them (sharded "tile" routing).  Re-base them onto a
        // virtual gap-free key space -- bbeg / bend become running counts -- so that rounds are carved by
        // real key counts; the true begins move to `gbeg` (aliases soff, unused by the ids path) and are
        // only needed when the ids are staged.
        int32_t *gbeg = sm.soff;
        const bool gapped = !dense && g_bends != nullptr;
        if (gapped) {
            if (tid < 32) {
                constexpr int kPer = (kMaxTileBags + 31) / 32;
                int sum = 0;
                for (int i = 0; i < kPer; ++i) {
                    const int b = tid * kPer + i;
                    sum += b < tile_nbags ? sm.bend[b] - sm.bbeg[b] : 0;
                }
                int incl = sum;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int y = __shfl_up_sync(0xffffffffu, incl, d);
                    if (tid >= d) incl += y;
                }
                int run = incl - sum;
                for (int i = 0; i < kPer; ++i) {
                    const int b = tid * kPer + i;
                    if (b < tile_nbags) {
                        const int beg = sm.bbeg[b], cnt = sm.bend[b] - beg;
                        gbeg[b] = beg;
                        sm.bbeg[b] = run;
                        sm.bend[b] = run + cnt;
                        run += cnt;
                    }
                }
            }
            __syncthreads();
        }
        const int64_t tile_item0 = dense ? (int64_t)tile_bag0 * F.bag_len : (int64_t)sm.bbeg[0];

        int bag = 0;            // next bag of the tile, relative
        int64_t part_done = 0;  // keys of a long bag already consumed
        while (bag < tile_nbags) {
            // ---- carve the next round ------------------------------------------------------
            Round R;
            const int64_t b_lo = dense ? tile_item0 + (int64_t)bag * F.bag_len : (int64_t)sm.bbeg[bag];
            const int64_t b_hi = dense ? b_lo + F.bag_len : (int64_t)sm.bend[bag];
            if (b_hi - b_lo > kChunk) {
                R.partial = true;
                R.bag0 = bag;
                R.bag1 = bag + 1;
                R.item0 = b_lo + part_done;
                R.n_keys = (int)min((int64_t)kChunk, b_hi - R.item0);
                R.first_part = part_done == 0;
                R.last_part = R.item0 + R.n_keys == b_hi;
            } else {
                R.partial = false;
                R.first_part = R.last_part = true;
                R.bag0 = bag;
                R.item0 = b_lo;
                int end;
                if (dense) {
                    const int per = F.bag_len > 0 ? kChunk / F.bag_len : tile_nbags;
                    end = min(tile_nbags, bag + max(per, 1));
                } else {
                    // largest end with bend[end - 1] - b_lo <= kChunk (ends are non-decreasing; with explicit
                    // ends there may be unused gaps between bags: they just count towards the round's span)
                    int lo = bag + 1, hi = tile_nbags;
                    while (lo < hi) {
                        const int mid = (lo + hi + 1) >> 1;
                        if ((int64_t)sm.bend[mid - 1] - b_lo <= kChunk) lo = mid; else hi = mid - 1;
                    }
                    end = lo;
                }
                R.bag1 = end;
                const int64_t e_hi = dense ? tile_item0 + (int64_t)end * F.bag_len : (int64_t)sm.bend[end - 1];
                R.n_keys = (int)(e_hi - b_lo);
            }

            // ---- phase A+B: keys -> bucket ids in shared memory -----------------------------
            if (F.ids_in != nullptr && gapped) {
                // virtual key -> its bag (binary search over the round's bags) -> true position
                for (int j = tid; j < R.n_keys; j += kThreads) {
                    const int vk = (int)R.item0 + j;
                    int lo = R.bag0, hi = R.bag1 - 1;
                    while (lo < hi) {
                        const int mid = (lo + hi + 1) >> 1;
                        if (sm.bbeg[mid] <= vk) lo = mid; else hi = mid - 1;
                    }
                    sm.ids[j] = (uint32_t)F.ids_in[(int64_t)gbeg[lo] + (vk - sm.bbeg[lo])];
                }
            } else if (F.ids_in != nullptr) {
                for (int j = tid; j < R.n_keys; j += kThreads)
                    for (int t = 0; t < T; ++t)
                        sm.ids[t * kChunk + j] = (uint32_t)F.ids_in[(int64_t)t * F.n_items + R.item0 + j];
            } else if (F.ints != nullptr) {
                uint32_t *scratch = sm.stage + tid * 6;
                for (int j = tid; j < R.n_keys; j += kThreads) {
                    const int64_t v = F.ints[R.item0 + j];
                    const uint32_t len = format_int64(v, scratch);
                    const WordSrcShared src{scratch, 0u};
                    const bool is_mask = F.mask_mode == RF_MASK_INT_VALUE && v == F.int_mask;
                    for (int t = 0; t < T; ++t) {
                        const uint32_t id = bucket_of(src, len, F.t[t].h, is_mask);
                        sm.ids[t * kChunk + j] = id;
                        if (F.ids_out) F.ids_out[(int64_t)t * F.n_items + R.item0 + j] = (int64_t)id;
                    }
                }
            } else {
                for (int sub0 = 0; sub0 < R.n_keys; sub0 += kSub) {
                    const int n_sub = min(kSub, R.n_keys - sub0);
                    const int64_t key0 = R.item0 + sub0;
                    if (sub0) __syncthreads();          // previous sub-chunk's offsets / bytes are consumed
                    for (int j = tid; j <= n_sub; j += kThreads) sm.soff[j] = F.soffs[key0 + j];
                    __syncthreads();
                    const int32_t byte0 = sm.soff[0];
                    const uint32_t n_bytes = (uint32_t)(sm.soff[n_sub] - byte0);
                    const uintptr_t addr0 = reinterpret_cast<uintptr_t>(F.bytes) + (uintptr_t)byte0;
                    const uint32_t shift = (uint32_t)(addr0 & 15u);
                    const bool staged = shift + n_bytes + 8u <= (uint32_t)kStageBytes;
                    if (staged) {
                        const uint4 *g = reinterpret_cast<const uint4 *>(addr0 - shift);
                        uint4 *s = reinterpret_cast<uint4 *>(sm.stage);
                        const uint32_t n_vec = (shift + n_bytes + 8u + 15u) >> 4;
                        for (uint32_t i = tid; i < n_vec; i += kThreads) s[i] = __ldg(g + i);
                        __syncthreads();
                    }
                    for (int j = tid; j < n_sub; j += kThreads) {
                        const int32_t o = sm.soff[j];
                        const uint32_t len = (uint32_t)(sm.soff[j + 1] - o);
                        bool is_mask = F.mask_mode == RF_MASK_EMPTY_STRING && len == 0;
                        if (F.mask_mode == RF_MASK_STRING_VALUE && len == (uint32_t)F.mask_len) {
                            if (staged) {
                                is_mask = equals_mask(WordSrcShared{sm.stage, shift + (uint32_t)(o - byte0)}, len, F.mask_words);
                            } else {
                                const uintptr_t a = reinterpret_cast<uintptr_t>(F.bytes) + (uintptr_t)o;
                                is_mask = equals_mask(WordSrcGlobal{reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3), (uint32_t)(a & 3u)},
                                                      len, F.mask_words);
                            }
                        }
                        if (staged && T == 2 && F.t[0].h.use_strong && F.t[1].h.use_strong) {
                            // both SipHashes of the key in lockstep (twice the instruction-level parallelism)
                            const WordSrcShared src{sm.stage, shift + (uint32_t)(o - byte0)};
                            uint32_t ida, idb;
                            bucket_of_x2(src, len, F.t[0].h, F.t[1].h, is_mask, ida, idb);
                            sm.ids[sub0 + j] = ida;
                            sm.ids[kChunk + sub0 + j] = idb;
                            if (F.ids_out) {
                                F.ids_out[key0 + j] = (int64_t)ida;
                                F.ids_out[F.n_items + key0 + j] = (int64_t)idb;
                            }
                            continue;
                        }
                        for (int t = 0; t < T; ++t) {
                            uint32_t id;
                            if (staged) {
                                const WordSrcShared src{sm.stage, shift + (uint32_t)(o - byte0)};
                                id = bucket_of(src, len, F.t[t].h, is_mask);
                            } else {
                                const uintptr_t a = reinterpret_cast<uintptr_t>(F.bytes) + (uintptr_t)o;
                                const WordSrcGlobal src{reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3), (uint32_t)(a & 3u)};
                                id = bucket_of(src, len, F.t[t].h, is_mask);
                            }
                            sm.ids[t * kChunk + sub0 + j] = id;
                            if (F.ids_out) F.ids_out[(int64_t)t * F.n_items + key0 + j] = (int64_t)id;
                        }
                    }
                }
            }
            if (tid == 0) sm.next_work = 0;
            __syncthreads();

            // ---- phase C: gather + pool -----------------------------------------------------
            if (F.dim > 0) {
                const PoolOp op{F.combiner, F.flags & RF_FIELD_PARTIAL};
                if (F.combiner <= RF_COMBINER_AVG) pool_round<kAdd, ACC>(op, F, sm, R, tile_bag0);
                else if (!ACC) pool_round<kSelect, false>(op, F, sm, R, tile_bag0);
            }
            __syncthreads();

            if (R.partial) {
                part_done += R.n_keys;
                if (R.last_part) {
                    part_done = 0;
                    bag += 1;
                }
            } else {
                bag = R.bag1;
            }
        }
    }
}

template <int MINB>
__global__ void __launch_bounds__(kThreads, MINB) bag_forward_kernel(const DevField *__restrict__ fields, int n_fields,
                                                                     int total_tiles) {
    bag_forward_body<false>(fields, n_fields, total_tiles);
}

// RF_FIELD_ACCUMULATE launches (every field of the launch carries the flag): same body, outputs reduced into
// the destination.  A separate instance so that the store path above stays byte-identical.
template <int MINB>
__global__ void __launch_bounds__(kThreads, MINB) bag_forward_acc_kernel(const DevField *__restrict__ fields, int n_fields,
                                                                         int total_tiles) {
    bag_forward_body<true>(fields, n_fields, total_tiles);
}

// ------------------------------------------------------------------------------------------
// host side: descriptor ring + launch
// ------------------------------------------------------------------------------------------
std::atomic<int64_t> g_launches{0};

struct DescSlot {
    void *host = nullptr;
    void *dev = nullptr;
    size_t cap = 0;
    cudaEvent_t done = nullptr;
    bool used = false;
};

constexpr int kGraphSlots = 256;          // descriptor slots reserved for launches recorded into CUDA graphs
constexpr size_t kGraphSlotBytes = 32 * 1024;

struct DeviceState {
    DescSlot slots[8];
    int next = 0;
    // A launch captured into a CUDA graph replays its descriptor upload from the same pinned host
    // bytes every time, so it gets a slot of its own that is never rewritten.  The pool is
    // allocated outside capture (cudaMalloc is not capturable).
    char *graph_host = nullptr;
    char *graph_dev = nullptr;
    int graph_used = 0;
    int sm_count = 0;
    std::mutex mu;
};

static DeviceState *device_state(int dev) {
    static std::mutex mu;
    static std::vector<DeviceState *> states;
    std::lock_guard<std::mutex> lk(mu);
    if ((int)states.size() <= dev) states.resize(dev + 1, nullptr);
    if (!states[dev]) states[dev] = new DeviceState();
    return states[dev];
}

// A run of descriptor slots for ONE captured launch: pinned host bytes that are never rewritten (the graph's memcpy node
// re-reads them on every replay) and the matching device bytes.  Shared with the optimizer (rf_bag_adam.cu).
static int graph_slots_locked(DeviceState *st, size_t bytes, char **host, char **device) {
    if (!st->graph_host) return set_error(RF_ERR_CUDA, "run the call once outside stream capture before capturing it");
    const int need = (int)((bytes + kGraphSlotBytes - 1) / kGraphSlotBytes);
    if (need < 1 || st->graph_used + need > kGraphSlots)
        return set_error(RF_ERR_UNSUPPORTED, "captured launches need more than the %d descriptor slots of %zu bytes (rf_release_captured_launches frees them)",
                         kGraphSlots, kGraphSlotBytes);
    *host = st->graph_host + (size_t)st->graph_used * kGraphSlotBytes;
    *device = st->graph_dev + (size_t)st->graph_used * kGraphSlotBytes;
    st->graph_used += need;
    return RF_OK;
}

static int graph_pool_ready_locked(DeviceState *st) {
    if (!st->graph_host) {
        RF_CUDA(cudaMallocHost(reinterpret_cast<void **>(&st->graph_host), kGraphSlots * kGraphSlotBytes));
        RF_CUDA(cudaMalloc(reinterpret_cast<void **>(&st->graph_dev), kGraphSlots * kGraphSlotBytes));
    }
    return RF_OK;
}

// Upload of a captured launch's descriptors: a KERNEL node that reads the pinned host slot through its device mapping (UVA).
// A cudaMemcpyAsync node would sit in the copy engine's queue, behind whatever bulk H2D the application prefetches for the
// next step on another stream -- measured: a graph-replayed forward then waits for the next batch's 35 MB input copy.
__global__ void __launch_bounds__(256) desc_copy_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, int n16) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

int graph_desc_upload(char *device, const char *host, size_t bytes, cudaStream_t stream) {
    const int n16 = (int)((bytes + 15) / 16);                 // slots are 32 KiB granular: rounding up stays inside the slot run
    int blocks = (n16 + 255) / 256;
    if (blocks > 16) blocks = 16;
    desc_copy_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const uint4 *>(host), reinterpret_cast<uint4 *>(device), n16);
    RF_CUDA(cudaGetLastError());
    return RF_OK;
}

int graph_desc_slots(int dev, size_t bytes, char **host, char **device) {
    DeviceState *st = device_state(dev);
    std::lock_guard<std::mutex> lk(st->mu);
    return graph_slots_locked(st, bytes, host, device);
}

int graph_desc_pool_ready(int dev) {
    DeviceState *st = device_state(dev);
    std::lock_guard<std::mutex> lk(st->mu);
    return graph_pool_ready_locked(st);
}

// Optional cap on resident CTAs per SM (0 = none): the kernel walks its tiles grid-stride, so a
// smaller grid leaves registers / CTA slots on every SM for kernels of OTHER streams -- used by the
// sharded pipeline so that the routing of the next step really runs under the pooling of this one.
static std::atomic<int> g_grid_ctas_per_sm{0};

static int launch_kernel(const DevField *dptr, int nf, int tiles_i, int ctas_per_sm, bool acc, cudaStream_t stream) {
    int64_t tiles = tiles_i;
    const int cap = ctas_per_sm > 0 ? ctas_per_sm : g_grid_ctas_per_sm.load();
    if (cap > 0) {
        int dev = 0, sms = 0;
        RF_CUDA(cudaGetDevice(&dev));
        RF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (tiles > (int64_t)cap * sms) tiles = (int64_t)cap * sms;
    }
    const int total_tiles = tiles_i;
    // CTAs per SM the kernel is compiled for (register cap): 4 (64 registers) measured best on
    // B200 for C2 (0.383 ms vs 0.401 @3); RF_BAG_MINB overrides for experiments.
    static const int cfg = getenv("RF_BAG_MINB") ? atoi(getenv("RF_BAG_MINB")) : 4;
    if (acc) {
        bag_forward_acc_kernel<4><<<(unsigned)tiles, kThreads, 0, stream>>>(dptr, nf, total_tiles);
        RF_CUDA(cudaGetLastError());
        g_launches.fetch_add(1);
        return RF_OK;
    }
    switch (cfg) {
        case 2: bag_forward_kernel<2><<<(unsigned)tiles, kThreads, 0, stream>>>(dptr, nf, total_tiles); break;
        case 3: bag_forward_kernel<3><<<(unsigned)tiles, kThreads, 0, stream>>>(dptr, nf, total_tiles); break;
        case 5: bag_forward_kernel<5><<<(unsigned)tiles, kThreads, 0, stream>>>(dptr, nf, total_tiles); break;
        default: bag_forward_kernel<4><<<(unsigned)tiles, kThreads, 0, stream>>>(dptr, nf, total_tiles); break;
    }
    RF_CUDA(cudaGetLastError());
    g_launches.fetch_add(1);
    return RF_OK;
}

static int launch_fields(std::vector<DevField> &dev_fields, int ctas_per_sm, cudaStream_t stream) {
    int dev = 0;
    RF_CUDA(cudaGetDevice(&dev));
    DeviceState *st = device_state(dev);
    std::lock_guard<std::mutex> lk(st->mu);
    if (st->sm_count == 0) RF_CUDA(cudaDeviceGetAttribute(&st->sm_count, cudaDevAttrMultiProcessorCount, dev));

    // tile sizing: ~1024 keys per tile for short bags, a full round (kChunk keys, i.e. many bags
    // to balance over the CTA's lane groups) for long ones; smaller when the launch would not
    // fill the GPU
    int64_t total_keys = 0;
    for (auto &f : dev_fields) total_keys += f.n_items;
    int64_t fill = total_keys / ((int64_t)st->sm_count * 8);
    if (fill < 128) fill = 128;
    int64_t tiles = 0;
    for (auto &f : dev_fields) {
        int64_t avg = f.batch > 0 ? (f.n_items + f.batch - 1) / f.batch : 1;
        if (avg < 1) avg = 1;
        int64_t target = avg >= 32 ? kChunk : kSub;
        if (target > fill) target = fill;
        int64_t bpt = target / avg;
        if (bpt < 1) bpt = 1;
        if (bpt > kMaxTileBags) bpt = kMaxTileBags;
        f.bags_per_tile = (int32_t)bpt;
        if (tiles > INT32_MAX) return set_error(RF_ERR_UNSUPPORTED, "too many tiles in one launch");
        f.tile_begin = (int32_t)tiles;
        tiles += (f.batch + bpt - 1) / bpt;
    }
    if (tiles == 0) return RF_OK;
    if (tiles > INT32_MAX) return set_error(RF_ERR_UNSUPPORTED, "too many tiles in one launch");
    int n_acc = 0;
    for (auto &f : dev_fields) n_acc += (f.flags & RF_FIELD_ACCUMULATE) ? 1 : 0;
    if (n_acc != 0 && n_acc != (int)dev_fields.size())
        return set_error(RF_ERR_INVALID, "RF_FIELD_ACCUMULATE must be set on every field of a launch or on none");
    const bool acc = n_acc != 0;

    const size_t bytes = dev_fields.size() * sizeof(DevField);
    cudaStreamCaptureStatus capture = cudaStreamCaptureStatusNone;
    RF_CUDA(cudaStreamIsCapturing(stream, &capture));
    if (capture != cudaStreamCaptureStatusNone) {
        char *h = nullptr, *d = nullptr;
        int rc = graph_slots_locked(st, bytes, &h, &d);
        if (rc != RF_OK) return rc;
        memcpy(h, dev_fields.data(), bytes);
        rc = graph_desc_upload(d, h, bytes, stream);
        if (rc != RF_OK) return rc;
        return launch_kernel(reinterpret_cast<const DevField *>(d), (int)dev_fields.size(), (int)tiles, ctas_per_sm, acc, stream);
    }
    {
        int rc = graph_pool_ready_locked(st);
        if (rc != RF_OK) return rc;
    }
    DescSlot &slot = st->slots[st->next];
    st->next = (st->next + 1) % 8;
    if (slot.used) RF_CUDA(cudaEventSynchronize(slot.done));
    if (slot.cap < bytes) {
        if (slot.host) cudaFreeHost(slot.host);
        if (slot.dev) cudaFree(slot.dev);
        slot.host = slot.dev = nullptr;
        size_t cap = bytes < 65536 ? 65536 : bytes * 2;
        RF_CUDA(cudaMallocHost(&slot.host, cap));
        RF_CUDA(cudaMalloc(&slot.dev, cap));
        slot.cap = cap;
    }
    if (!slot.done) RF_CUDA(cudaEventCreateWithFlags(&slot.done, cudaEventDisableTiming));
    memcpy(slot.host, dev_fields.data(), bytes);
    RF_CUDA(cudaMemcpyAsync(slot.dev, slot.host, bytes, cudaMemcpyHostToDevice, stream));
    {
        int rc = launch_kernel(static_cast<const DevField *>(slot.dev), (int)dev_fields.size(), (int)tiles, ctas_per_sm, acc, stream);
        if (rc != RF_OK) return rc;
    }
    RF_CUDA(cudaEventRecord(slot.done, stream));
    slot.used = true;
    return RF_OK;
}

static int fill_table(DevTable &dt, const rf_table_desc &t, int mask_mode, bool need_weights, int fi, int ti) {
    if (t.num_bins <= 0) return set_error(RF_ERR_INVALID, "`num_bins` cannot be `None` or non-positive values.");
    if (t.num_bins > 0xffffffffLL) return set_error(RF_ERR_UNSUPPORTED, "num_bins above 2^32-1 is not supported");
    if (need_weights && t.weights == nullptr)
        return set_error(RF_ERR_INVALID, "field %d table %d: weights pointer is NULL", fi, ti);
    dt.w = t.weights;
    dt.h = make_hash_spec(t.num_bins, mask_mode != RF_MASK_NONE, t.use_strong, t.key0, t.key1);
    return RF_OK;
}

static int build_field(DevField &d, const rf_field_desc &f, int64_t batch, int fi) {
    memset(&d, 0, sizeof(d));
    const int n_src = (f.bytes != nullptr) + (f.int_values != nullptr) + (f.ids != nullptr);
    if (n_src != 1) return set_error(RF_ERR_INVALID, "field %d: exactly one of bytes / int_values / ids must be set", fi);
    if (f.bytes && !f.str_offsets) return set_error(RF_ERR_INVALID, "field %d: bytes given without str_offsets", fi);
    if (f.n_tables < 1 || f.n_tables > RF_MAX_TABLES_PER_FIELD)
        return set_error(RF_ERR_INVALID, "field %d: n_tables must be 1..%d", fi, RF_MAX_TABLES_PER_FIELD);
    if (batch < 0 || batch > INT32_MAX) return set_error(RF_ERR_INVALID, "batch out of range");
    if (f.dim < 0 || f.dim > kMaxDim) return set_error(RF_ERR_UNSUPPORTED, "field %d: dim must be in [0, %d]", fi, kMaxDim);
    if (f.dim == 0 && !f.ids_out) return set_error(RF_ERR_INVALID, "field %d: dim == 0 (hash only) needs ids_out", fi);
    if (f.dim > 0 && !f.out) return set_error(RF_ERR_INVALID, "field %d: out pointer is NULL", fi);
    if (f.combiner < RF_COMBINER_SUM || f.combiner > RF_COMBINER_MAX)
        return set_error(RF_ERR_INVALID, "field %d: Do not support combiner = %d", fi, f.combiner);
    if (f.mask_mode < RF_MASK_NONE || f.mask_mode > RF_MASK_STRING_VALUE)
        return set_error(RF_ERR_INVALID, "field %d: bad mask_mode", fi);
    if (f.bytes && f.mask_mode == RF_MASK_INT_VALUE) return set_error(RF_ERR_INVALID, "field %d: integer mask on string keys", fi);
    if (f.int_values && (f.mask_mode == RF_MASK_EMPTY_STRING || f.mask_mode == RF_MASK_STRING_VALUE))
        return set_error(RF_ERR_INVALID, "field %d: string mask on integer keys", fi);
    if (f.mask_mode == RF_MASK_STRING_VALUE && (f.mask_len < 0 || f.mask_len > RF_MAX_MASK_BYTES))
        return set_error(RF_ERR_UNSUPPORTED, "field %d: a string mask_value takes at most %d bytes", fi, RF_MAX_MASK_BYTES);
    if (!f.bag_offsets && f.bag_len < 0) return set_error(RF_ERR_INVALID, "field %d: negative bag_len", fi);
    if (f.bag_ends && !f.bag_offsets) return set_error(RF_ERR_INVALID, "field %d: bag_ends given without bag_offsets", fi);
    if (f.bag_ends && (!f.ids || f.n_tables != 1))
        return set_error(RF_ERR_UNSUPPORTED, "field %d: bag_ends (gapped bags) takes pre-hashed ids and one table", fi);
    if ((f.flags & RF_FIELD_ACCUMULATE) && f.combiner > RF_COMBINER_AVG)
        return set_error(RF_ERR_INVALID, "field %d: RF_FIELD_ACCUMULATE takes the sum / avg combiners", fi);
    d.bytes = f.bytes;
    d.soffs = f.str_offsets;
    d.ints = f.int_values;
    d.ids_in = f.ids;
    d.boffs = f.bag_offsets;
    d.bends = f.bag_ends;
    d.out = f.out;
    d.ids_out = f.ids_out;
    d.out_stride = f.out_stride;
    d.batch = (int32_t)batch;
    d.bag_len = f.bag_offsets ? 0 : f.bag_len;
    d.n_items = f.bag_offsets ? f.n_items : batch * (int64_t)f.bag_len;
    if (d.n_items < 0) return set_error(RF_ERR_INVALID, "field %d: negative n_items", fi);
    d.int_mask = f.int_mask_value;
    d.dim = f.dim;
    d.n_tables = f.n_tables;
    d.combiner = f.combiner;
    d.mask_mode = f.mask_mode;
    if (f.mask_mode == RF_MASK_STRING_VALUE) {
        if (f.mask_len == 0) {
            d.mask_mode = RF_MASK_EMPTY_STRING;      // mask_value="" is the length test alone
        } else {
            memcpy(d.mask_words, f.mask_bytes, (size_t)f.mask_len);
            d.mask_len = f.mask_len;
        }
    }
    d.flags = f.flags;
    bool aligned = (f.dim % 4 == 0) && (reinterpret_cast<uintptr_t>(f.out) % 16 == 0) && (f.out_stride % 4 == 0);
    for (int t = 0; t < f.n_tables; ++t) {
        int rc = fill_table(d.t[t], f.tables[t], f.mask_mode, f.dim > 0, fi, t);
        if (rc != RF_OK) return rc;
        aligned = aligned && (reinterpret_cast<uintptr_t>(f.tables[t].weights) % 16 == 0);
    }
    d.vec_ok = aligned ? 1 : 0;
    return RF_OK;
}

}  // namespace rf

using namespace rf;

extern "C" {

int rf_abi_version(void) { return RF_B200_ABI_VERSION; }

int64_t rf_launch_count(void) { return g_launches.load(); }

int rf_set_bag_grid_limit(int ctas_per_sm) {
    if (ctas_per_sm < 0) return set_error(RF_ERR_INVALID, "ctas_per_sm must be >= 0");
    g_grid_ctas_per_sm.store(ctas_per_sm);
    return RF_OK;
}

uint64_t rf_debug_fastmod(uint64_t x, uint64_t d) {
    const FastMod m = make_fastmod(d);
    return fastmod(x, m);
}

int rf_bag_forward_ex(const rf_field_desc *fields, int n_fields, int64_t batch, int max_ctas_per_sm, void *stream) {
    if (n_fields < 0 || (n_fields > 0 && !fields)) return set_error(RF_ERR_INVALID, "bad fields array");
    if (max_ctas_per_sm < 0) return set_error(RF_ERR_INVALID, "max_ctas_per_sm must be >= 0");
    if (n_fields == 0 || batch == 0) return RF_OK;
    std::vector<DevField> dev(n_fields);
    for (int i = 0; i < n_fields; ++i) {
        int rc = build_field(dev[i], fields[i], batch, i);
        if (rc != RF_OK) return rc;
    }
    return launch_fields(dev, max_ctas_per_sm, static_cast<cudaStream_t>(stream));
}

int rf_bag_forward(const rf_field_desc *fields, int n_fields, int64_t batch, void *stream) {
    return rf_bag_forward_ex(fields, n_fields, batch, 0, stream);
}

int rf_release_captured_launches(void) {
    int dev = 0;
    RF_CUDA(cudaGetDevice(&dev));
    DeviceState *st = device_state(dev);
    std::lock_guard<std::mutex> lk(st->mu);
    st->graph_used = 0;
    return RF_OK;
}

static int hash_only(rf_field_desc &f, int64_t n_items, int64_t num_bins, int mask_mode, int use_strong, uint64_t key0,
                     uint64_t key1, int64_t *d_ids_out, void *stream) {
    if (n_items < 0 || n_items > INT32_MAX) return set_error(RF_ERR_INVALID, "n_items out of range");
    if (n_items == 0) return RF_OK;
    if (!d_ids_out) return set_error(RF_ERR_INVALID, "ids_out is NULL");
    f.bag_len = 1;
    f.n_tables = 1;
    f.tables[0].num_bins = num_bins;
    f.tables[0].use_strong = use_strong;
    f.tables[0].key0 = key0;
    f.tables[0].key1 = key1;
    f.dim = 0;
    f.combiner = RF_COMBINER_SUM;
    f.mask_mode = mask_mode;
    f.ids_out = d_ids_out;
    return rf_bag_forward(&f, 1, n_items, stream);
}

int rf_hash_strings(const uint8_t *d_bytes, const int32_t *d_str_offsets, int64_t n_items, int64_t num_bins, int mask_mode,
                    int use_strong, uint64_t key0, uint64_t key1, int64_t *d_ids_out, void *stream) {
    rf_field_desc f;
    memset(&f, 0, sizeof(f));
    if (n_items > 0 && (!d_bytes || !d_str_offsets)) return set_error(RF_ERR_INVALID, "bytes / str_offsets is NULL");
    f.bytes = d_bytes;
    f.str_offsets = d_str_offsets;
    return hash_only(f, n_items, num_bins, mask_mode, use_strong, key0, key1, d_ids_out, stream);
}

int rf_hash_strings_masked(const uint8_t *d_bytes, const int32_t *d_str_offsets, int64_t n_items, int64_t num_bins,
                           const uint8_t *h_mask_value, int32_t mask_len, int use_strong, uint64_t key0, uint64_t key1,
                           int64_t *d_ids_out, void *stream) {
    rf_field_desc f;
    memset(&f, 0, sizeof(f));
    if (n_items > 0 && (!d_bytes || !d_str_offsets)) return set_error(RF_ERR_INVALID, "bytes / str_offsets is NULL");
    if (mask_len < 0 || mask_len > RF_MAX_MASK_BYTES)
        return set_error(RF_ERR_UNSUPPORTED, "a string mask_value takes at most %d bytes", RF_MAX_MASK_BYTES);
    if (mask_len > 0 && !h_mask_value) return set_error(RF_ERR_INVALID, "mask_value is NULL");
    f.bytes = d_bytes;
    f.str_offsets = d_str_offsets;
    if (mask_len > 0) memcpy(f.mask_bytes, h_mask_value, (size_t)mask_len);
    f.mask_len = mask_len;
    return hash_only(f, n_items, num_bins, RF_MASK_STRING_VALUE, use_strong, key0, key1, d_ids_out, stream);
}

int rf_hash_int64(const int64_t *d_values, int64_t n_items, int64_t num_bins, int mask_mode, int64_t int_mask_value,
                  int use_strong, uint64_t key0, uint64_t key1, int64_t *d_ids_out, void *stream) {
    rf_field_desc f;
    memset(&f, 0, sizeof(f));
    if (n_items > 0 && !d_values) return set_error(RF_ERR_INVALID, "values is NULL");
    f.int_values = d_values;
    f.int_mask_value = int_mask_value;
    return hash_only(f, n_items, num_bins, mask_mode, use_strong, key0, key1, d_ids_out, stream);
}

}  // extern "C"
