// rf_shard.cu -- routing and combining kernels for ROW-SHARDED embedding tables.
//
// The reference replicates every table on every GPU (tf.distribute.MirroredStrategy,
// /root/reference/backend/utils/gpu_utils.py:13-14); row sharding is the new design that
// north_star asks for once a table no longer belongs on one GPU.  Row `id` of a table lives on
// rank `id % world` as local row `id / world`.  Per step and per rank:
//
//   1. hash the local keys (rf_hash_strings)                                    -> ids[n]
//   2. rf_shard_route: count / scan / scatter the ids by owner, keeping bag order ->
//        for every owner g: CSR offsets[batch+1] + local rows[], written THROUGH THE POINTERS
//        THE CALLER GIVES -- peer-mapped NVLink pointers in the fused path, local send buffers
//        in the NCCL path.
//   3. on the owner: rf_bag_forward over the received (rows, offsets) with `out` pointing at the
//        source rank's partial buffer (peer pointer => the pooled vector crosses NVLink as the
//        kernel produces it; no separate all-to-all).
//   4. rf_combine_partials on the source: reduce the `world` partial vectors in rank order.
//
// Everything here is integer routing and streaming adds: HBM / NVLink bound, no tensor cores.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "../../include/rf_b200.h"
#include "rf_common.h"
#include "rf_hash.cuh"

namespace rf {

constexpr int kMaxWorld = 16;
extern std::atomic<int64_t> g_launches;

struct PtrTable {
    void *p[kMaxWorld];
};

__device__ __forceinline__ void bag_range(const int32_t *boffs, int bag_len, int64_t b, int64_t &lo, int64_t &hi) {
    if (boffs) {
        lo = boffs[b];
        hi = boffs[b + 1];
    } else {
        lo = b * bag_len;
        hi = lo + bag_len;
    }
}

// counts[g * batch + b] = #keys of bag b owned by rank g.
// A CTA takes a run of consecutive bags (~kCntCap keys).  Pass 1 -- one key per thread, every lane
// busy, coalesced offsets and neighbouring key bytes: hash the key (strings, straight from global
// memory) or fetch its id, keep it in shared memory and (HASH) leave it in ids_ws for the scatter
// pass.  Pass 2 -- one warp per bag counts owners out of shared memory with ballots.  A bag longer
// than the staging capacity is counted by one warp directly from global memory.
constexpr int kCntCap = 4096;
constexpr int kCntThreads = 256;

template <bool HASH>
__device__ __forceinline__ uint32_t key_id(int64_t k, const int64_t *ids, const uint8_t *bytes, const int32_t *soffs,
                                           const HashSpec &spec, int mask_empty) {
    if (!HASH) return (uint32_t)ids[k];
    const int32_t o = soffs[k];
    const uint32_t len = (uint32_t)(soffs[k + 1] - o);
    const uintptr_t a = reinterpret_cast<uintptr_t>(bytes) + (uintptr_t)o;
    const WordSrcGlobal src{reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3), (uint32_t)(a & 3u)};
    return bucket_of(src, len, spec, mask_empty && len == 0);
}

template <bool HASH>
__global__ void __launch_bounds__(kCntThreads) shard_count_kernel(const int64_t *__restrict__ ids, const uint8_t *__restrict__ bytes,
                                                                  const int32_t *__restrict__ soffs, HashSpec spec, int mask_empty,
                                                                  int64_t *__restrict__ ids_ws, const int32_t *__restrict__ boffs,
                                                                  int bag_len, int64_t batch, int world, int tile_bags,
                                                                  int32_t *__restrict__ counts) {
    __shared__ uint32_t sid[kCntCap];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t n_tiles = (batch + tile_bags - 1) / tile_bags;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t t0 = tile * tile_bags, t1 = min(batch, t0 + tile_bags);
        int64_t r0 = t0;
        while (r0 < t1) {
            int64_t k0, dummy;
            bag_range(boffs, bag_len, r0, k0, dummy);
            int64_t r1 = r0, k1 = k0;
            while (r1 < t1) {                           // whole bags, <= kCntCap keys (uniform across the CTA)
                int64_t lo, hi;
                bag_range(boffs, bag_len, r1, lo, hi);
                if (hi - k0 > kCntCap && r1 > r0) break;
                k1 = hi;
                ++r1;
                if (hi - k0 > kCntCap) break;
            }
            const bool staged = k1 - k0 <= kCntCap;
            if (staged) {
                for (int64_t k = k0 + tid; k < k1; k += kCntThreads) {
                    const uint32_t id = key_id<HASH>(k, ids, bytes, soffs, spec, mask_empty);
                    sid[k - k0] = id;
                    if (HASH) ids_ws[k] = (int64_t)id;
                }
                __syncthreads();
            }
            for (int64_t b = r0 + wid; b < r1; b += kCntThreads / 32) {
                int64_t lo, hi;
                bag_range(boffs, bag_len, b, lo, hi);
                int mine = 0;                           // lane g accumulates the count of owner g
                for (int64_t i = lo; i < hi; i += 32) {
                    const bool on = i + lane < hi;
                    uint32_t id = 0;
                    if (on) {
                        if (staged) {
                            id = sid[i + lane - k0];
                        } else {
                            id = key_id<HASH>(i + lane, ids, bytes, soffs, spec, mask_empty);
                            if (HASH) ids_ws[i + lane] = (int64_t)id;
                        }
                    }
                    const int owner = on ? (int)(id % (uint32_t)world) : -1;
                    for (int g = 0; g < world; ++g) {
                        const unsigned m = __ballot_sync(0xffffffffu, owner == g);
                        if (lane == g) mine += __popc(m);
                    }
                }
                if (lane < world) counts[(int64_t)lane * batch + b] = mine;
            }
            __syncthreads();
            r0 = r1;
        }
    }
}

constexpr int kScanChunk = 1024;   // bags per scan chunk

// Grid (chunks, world): exclusive scan of counts[g][chunk*1024 ..] inside the chunk ->
// excl[g][b] (chunk-relative) and chunk_tot[g][chunk].  The chunk bases are summed on the fly by
// the scatter pass, so the whole scan is two fully parallel kernels instead of a serial one.
__global__ void __launch_bounds__(kScanChunk) shard_scan_kernel(const int32_t *__restrict__ counts, int64_t batch,
                                                                int n_chunks, int32_t *__restrict__ excl,
                                                                int32_t *__restrict__ chunk_tot) {
    __shared__ int32_t warp_sums[32];
    const int g = blockIdx.y, chunk = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t i = (int64_t)chunk * kScanChunk + tid;
    const int32_t v = i < batch ? counts[(int64_t)g * batch + i] : 0;
    int32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int32_t y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        int32_t t = warp_sums[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int32_t y = __shfl_up_sync(0xffffffffu, t, d);
            if (lane >= d) t += y;
        }
        warp_sums[lane] = t;
    }
    __syncthreads();
    if (i < batch) excl[(int64_t)g * batch + i] = (wid ? warp_sums[wid - 1] : 0) + x - v;
    if (tid == kScanChunk - 1) chunk_tot[(int64_t)g * n_chunks + chunk] = warp_sums[31];
}

// One CTA per owner: chunk_tot[g][0..n_chunks) -> exclusive chunk bases in place; the grand total
// closes the owner's CSR (offs_dst[g][batch]).
__global__ void __launch_bounds__(1024) shard_chunk_base_kernel(int32_t *__restrict__ chunk_tot, int n_chunks, int64_t batch,
                                                                PtrTable offs_dst) {
    __shared__ int32_t warp_sums[32];
    __shared__ int32_t carry_s;
    const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int32_t *ct = chunk_tot + (int64_t)g * n_chunks;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n_chunks; base += 1024) {
        const int i = base + tid;
        const int32_t v = i < n_chunks ? ct[i] : 0;
        int32_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int32_t y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) warp_sums[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int32_t t = warp_sums[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int32_t y = __shfl_up_sync(0xffffffffu, t, d);
                if (lane >= d) t += y;
            }
            warp_sums[lane] = t;
        }
        __syncthreads();
        const int32_t carry = carry_s;
        if (i < n_chunks) ct[i] = carry + (wid ? warp_sums[wid - 1] : 0) + x - v;
        __syncthreads();
        if (tid == 1023) carry_s = carry + warp_sums[31];
        __syncthreads();
    }
    if (tid == 0 && offs_dst.p[g]) static_cast<int32_t *>(offs_dst.p[g])[batch] = carry_s;
}

// One warp per bag: key k of bag b owned by g goes to rows_dst[g][offs[g][b] + (rank of k among the
// bag's keys owned by g)] as the owner-local row id / world.  Order inside (owner, bag) is kept.
// offs[g][b] = chunk base + excl[g][b]; lane g also publishes it as the owner's CSR (offs_dst[g][b]).
__global__ void __launch_bounds__(256) shard_scatter_warp_kernel(const int64_t *__restrict__ ids, const int32_t *__restrict__ boffs,
                                                            int bag_len, int64_t batch, int world, int n_chunks,
                                                            const int32_t *__restrict__ excl,
                                                            const int32_t *__restrict__ chunk_tot, PtrTable offs_dst,
                                                            PtrTable rows_dst) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const unsigned lt = (1u << lane) - 1u;
    for (int64_t b = warp0; b < batch; b += n_warps) {
        int64_t lo, hi;
        bag_range(boffs, bag_len, b, lo, hi);
        const int chunk = (int)(b / kScanChunk);
        // lane g: base of owner g = chunk totals before this chunk + in-chunk exclusive offset
        int next = 0;
        if (lane < world) {
            next = chunk_tot[(int64_t)lane * n_chunks + chunk] + excl[(int64_t)lane * batch + b];
            int32_t *od = static_cast<int32_t *>(offs_dst.p[lane]);
            if (od) od[b] = next;
        }
        for (int64_t i = lo; i < hi; i += 32) {
            const bool on = i + lane < hi;
            const uint32_t id = on ? (uint32_t)ids[i + lane] : 0u;
            const uint32_t row = id / (uint32_t)world;
            const int owner = on ? (int)(id - row * (uint32_t)world) : -1;
            int pos = 0;
            for (int g = 0; g < world; ++g) {
                const unsigned m = __ballot_sync(0xffffffffu, owner == g);
                const int base = __shfl_sync(0xffffffffu, next, g);
                if (owner == g) pos = base + __popc(m & lt);
                if (lane == g) next += __popc(m);
            }
            if (on) static_cast<int64_t *>(rows_dst.p[owner])[pos] = (int64_t)row;
        }
    }
}

// Scatter, tiled: a CTA takes a run of consecutive bags.  Because every owner's CSR is in bag order,
// the tile's keys for owner g form ONE contiguous range of g's receive buffer; the CTA assembles
// the W ranges in shared memory (warp per bag, ballot ranks keep key order) and then streams each
// range out with fully coalesced stores -- long contiguous writes are what NVLink moves
// efficiently (8-byte scattered remote stores cost 2x the whole routing pass at 8 GPUs).
// The owners' CSR entries of the tile's bags are written the same way.
constexpr int kScatCap = 4096;     // keys staged per round
constexpr int kScatThreads = 256;

__device__ __forceinline__ int owner_offset(const int32_t *excl, const int32_t *chunk_tot, const int32_t *counts, int g,
                                            int64_t b, int64_t batch, int n_chunks) {
    if (b < batch) return chunk_tot[(int64_t)g * n_chunks + (int)(b / kScanChunk)] + excl[(int64_t)g * batch + b];
    const int64_t l = batch - 1;
    return chunk_tot[(int64_t)g * n_chunks + (int)(l / kScanChunk)] + excl[(int64_t)g * batch + l] + counts[(int64_t)g * batch + l];
}

__global__ void __launch_bounds__(kScatThreads) shard_scatter_kernel(const int64_t *__restrict__ ids,
                                                                     const int32_t *__restrict__ boffs, int bag_len,
                                                                     int64_t batch, int world, int n_chunks, int tile_bags,
                                                                     const int32_t *__restrict__ excl,
                                                                     const int32_t *__restrict__ chunk_tot,
                                                                     const int32_t *__restrict__ counts, PtrTable offs_dst,
                                                                     PtrTable rows_dst) {
    __shared__ uint32_t srow[kScatCap];
    __shared__ int seg_first[kMaxWorld], seg_len[kMaxWorld], seg_base[kMaxWorld + 1];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t n_tiles = (batch + tile_bags - 1) / tile_bags;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t t0 = tile * tile_bags, t1 = min(batch, t0 + tile_bags);
        int64_t r0 = t0;
        while (r0 < t1) {
            // ---- carve a round [r0, r1) of whole bags with <= kScatCap keys (uniform across the CTA) ----
            int64_t k0, dummy;
            bag_range(boffs, bag_len, r0, k0, dummy);
            int64_t r1 = r0;
            int64_t k1 = k0;
            while (r1 < t1) {
                int64_t lo, hi;
                bag_range(boffs, bag_len, r1, lo, hi);
                if (hi - k0 > kScatCap && r1 > r0) break;
                k1 = hi;
                ++r1;
                if (hi - k0 > kScatCap) break;      // a single oversized bag: handled unstaged below
            }
            const bool staged = k1 - k0 <= kScatCap;
            if (tid < world) {
                seg_first[tid] = owner_offset(excl, chunk_tot, counts, tid, r0, batch, n_chunks);
                seg_len[tid] = owner_offset(excl, chunk_tot, counts, tid, r1, batch, n_chunks) - seg_first[tid];
            }
            __syncthreads();
            if (tid == 0) {
                int acc = 0;
                for (int g = 0; g < world; ++g) {
                    seg_base[g] = acc;
                    acc += seg_len[g];
                }
                seg_base[world] = acc;
            }
            __syncthreads();
            // ---- warp per bag: rank every key inside its (owner, bag) run ----
            for (int64_t b = r0 + wid; b < r1; b += kScatThreads / 32) {
                int64_t lo, hi;
                bag_range(boffs, bag_len, b, lo, hi);
                int next = lane < world ? owner_offset(excl, chunk_tot, counts, lane, b, batch, n_chunks) : 0;
                for (int64_t i = lo; i < hi; i += 32) {
                    const bool on = i + lane < hi;
                    const uint32_t id = on ? (uint32_t)ids[i + lane] : 0u;
                    const uint32_t row = id / (uint32_t)world;
                    const int owner = on ? (int)(id - row * (uint32_t)world) : -1;
                    int pos = 0;
                    for (int g = 0; g < world; ++g) {
                        const unsigned m = __ballot_sync(0xffffffffu, owner == g);
                        const int base = __shfl_sync(0xffffffffu, next, g);
                        if (owner == g) pos = base + __popc(m & lt);
                        if (lane == g) next += __popc(m);
                    }
                    if (on) {
                        if (staged) srow[seg_base[owner] + (pos - seg_first[owner])] = row;
                        else static_cast<int64_t *>(rows_dst.p[owner])[pos] = (int64_t)row;
                    }
                }
            }
            __syncthreads();
            // ---- stream the W contiguous ranges (and the CSR entries) out, coalesced ----
            for (int g = 0; g < world; ++g) {
                if (staged) {
                    int64_t *dst = static_cast<int64_t *>(rows_dst.p[g]) + seg_first[g];
                    const uint32_t *src = srow + seg_base[g];
                    for (int i = tid; i < seg_len[g]; i += kScatThreads) dst[i] = (int64_t)src[i];
                }
                int32_t *od = static_cast<int32_t *>(offs_dst.p[g]);
                if (od)
                    for (int64_t bb = r0 + tid; bb < r1; bb += kScatThreads)
                        od[bb] = owner_offset(excl, chunk_tot, counts, g, bb, batch, n_chunks);
            }
            __syncthreads();
            r0 = r1;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Single-pass routing ("tile" layout).  The two-pass route above needs a global scan between its
// count and scatter passes only because every owner's CSR must be gap-free.  If the owner accepts
// bags with explicit [begin, end) and gaps between them (rf_field_desc.bag_ends), a CTA can finish
// its run of bags on its own: keys k0 .. k1 of this source go to [k0, k0 + n_g) of owner g's
// buffer (n_g <= k1 - k0 always fits; the rest of that range stays unused).  One pass over the keys,
// no global scan, no second read of the ids, and all remote writes are long contiguous runs:
//   1. one key per thread: hash (or fetch) the id                                   -> smem
//   2. one warp per bag: owner of each key, rank inside its (bag, owner) run with match.any
//      (one instruction instead of a ballot per owner), per-(bag, owner) counts    -> smem
//   3. one warp per owner: exclusive scan of the counts over the round's bags       -> smem
//   4. one key per thread: place the owner-local row at its slot of the staging tile
//   5. stream each owner's run, and the bags' begin / end arrays, out with coalesced stores
// ------------------------------------------------------------------------------------------
constexpr int kTileCap = 4096;        // keys per round
constexpr int kTileBags = 128;        // bags per round
constexpr int kTileThreads = 256;

struct TileSmem {
    uint32_t sid[kTileCap];                   // bucket id of each key of the round
    uint32_t meta[kTileCap + 8];              // rank (12) | owner (4) << 12 | bag (8) << 16; before that: key offsets
    uint32_t srow[kTileCap];                  // staging: owner-local rows, grouped by owner
    uint16_t cnt[kTileBags][kMaxWorld];       // keys of (bag, owner)
    uint16_t beg[kTileBags][kMaxWorld];       // exclusive scan of cnt over the bags, per owner
    int seg_len[kMaxWorld], seg_base[kMaxWorld + 1];
    int32_t tile_boffs[kTileBags + 4];        // STAGE: key offsets of the tile's bags (tile_bags <= kTileBags)
};
constexpr int kStageKeys = 1024;              // keys whose bytes are staged at a time (<= 16 KiB - slack, else unstaged)

// STAGE = true (default, RF_ROUTE_STAGE=0 selects the round-1 kernel): the tile's bag offsets and, 1024 keys at a
// time, the contiguous slice of the string arena those keys occupy are copied into shared memory with coalesced
// 16-byte loads first, so carving rounds and hashing never wait on a dependent global load (round 1: offset -> bytes
// -> hash was one DRAM round trip per key with 16 keys per thread in series; the kernel ran at 11 % of HBM speed).
template <bool HASH, bool STAGE>
__global__ void __launch_bounds__(kTileThreads, 4) shard_route_tile_kernel(
    const int64_t *__restrict__ ids, const uint8_t *__restrict__ bytes, const int32_t *__restrict__ soffs, HashSpec spec,
    int mask_empty, int64_t *__restrict__ ids_ws, const int32_t *__restrict__ boffs, int bag_len, int64_t batch, int world,
    int tile_bags, const __grid_constant__ PtrTable rows_dst, const __grid_constant__ PtrTable begin_dst,
    const __grid_constant__ PtrTable end_dst) {
    extern __shared__ __align__(16) unsigned char tile_raw[];
    TileSmem &sm = *reinterpret_cast<TileSmem *>(tile_raw);
    int32_t *tb = reinterpret_cast<int32_t *>(sm.tile_boffs);      // STAGE: key offsets of the tile's bags
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const bool pow2 = (world & (world - 1)) == 0;
    const int wshift = 31 - __clz(world);
    const int64_t n_tiles = (batch + tile_bags - 1) / tile_bags;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t t0 = tile * tile_bags, t1 = min(batch, t0 + tile_bags);
        if (STAGE) {
            __syncthreads();                        // the previous tile's rounds are done with tb
            for (int i = tid; i <= (int)(t1 - t0); i += kTileThreads)
                tb[i] = boffs ? boffs[t0 + i] : (int32_t)((t0 + i) * bag_len);
            __syncthreads();
        }
        int64_t r0 = t0;
        while (r0 < t1) {
            // ---- carve a round: whole bags, <= kTileCap keys, <= kTileBags bags ----
            int64_t k0, k1, r1;
            if (STAGE) {
                // tb is non-decreasing: binary search for the last bag whose end still fits the round
                k0 = tb[r0 - t0];
                int lo = (int)(r0 - t0), hi = (int)min(t1 - t0, r0 - t0 + kTileBags);
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if ((int64_t)tb[mid] - k0 <= kTileCap) lo = mid; else hi = mid - 1;
                }
                r1 = t0 + lo;
                k1 = tb[lo];
            } else {
                int64_t dummy;
                bag_range(boffs, bag_len, r0, k0, dummy);
                r1 = r0;
                k1 = k0;
                while (r1 < t1 && r1 - r0 < kTileBags) {
                    int64_t lo, hi;
                    bag_range(boffs, bag_len, r1, lo, hi);
                    if (hi - k0 > kTileCap) break;
                    k1 = hi;
                    ++r1;
                }
            }
            const int nb = (int)(r1 - r0);
            if (nb == 0) {
                // one bag longer than a round: walk it in pieces of kTileCap keys; the pieces of one (bag, owner)
                // run stay in key order because every piece starts at its own key index
                int64_t lo, hi;
                bag_range(boffs, bag_len, r0, lo, hi);
                int run[kMaxWorld];                     // thread 0: keys of each owner placed so far
                if (tid == 0)
                    for (int g = 0; g < world; ++g) run[g] = 0;
                // the simple layout cannot express a bag whose owner-run is split by gaps, so such bags are
                // packed serially by one thread (rare: > 4096 keys in one bag)
                if (tid == 0) {
                    for (int64_t k = lo; k < hi; ++k) {
                        const uint32_t id = key_id<HASH>(k, ids, bytes, soffs, spec, mask_empty);
                        if (HASH && ids_ws) ids_ws[k] = (int64_t)id;
                        const uint32_t row = id / (uint32_t)world, g = id - row * (uint32_t)world;
                        static_cast<int64_t *>(rows_dst.p[g])[lo + run[g]] = (int64_t)row;
                        ++run[g];
                    }
                    for (int g = 0; g < world; ++g) {
                        static_cast<int32_t *>(begin_dst.p[g])[r0] = (int32_t)lo;
                        static_cast<int32_t *>(end_dst.p[g])[r0] = (int32_t)(lo + run[g]);
                    }
                }
                __syncthreads();
                r0 += 1;
                continue;
            }
            const int n_keys = (int)(k1 - k0);
            // ---- 1. ids (string keys: the round's offsets are staged first, coalesced, so that hashing
            //         costs one dependent global round trip per key instead of two) ----
            if (HASH && STAGE) {
                int32_t *soff_s = reinterpret_cast<int32_t *>(sm.meta);
                for (int j = tid; j <= n_keys; j += kTileThreads) soff_s[j] = soffs[k0 + j];
                __syncthreads();
                // key bytes, kStageKeys keys at a time, into the staging area (srow is not live until step 4)
                uint32_t *stage = sm.srow;
                for (int sub0 = 0; sub0 < n_keys; sub0 += kStageKeys) {
                    const int n_sub = min(kStageKeys, n_keys - sub0);
                    const int32_t byte0 = soff_s[sub0];
                    const uint32_t n_bytes = (uint32_t)(soff_s[sub0 + n_sub] - byte0);
                    const uintptr_t addr0 = reinterpret_cast<uintptr_t>(bytes) + (uintptr_t)byte0;
                    const uint32_t shift = (uint32_t)(addr0 & 15u);
                    const bool staged = shift + n_bytes + 8u <= (uint32_t)(kTileCap * 4 - 32);
                    if (sub0) __syncthreads();              // the previous slice has been hashed
                    if (staged) {
                        const uint4 *g = reinterpret_cast<const uint4 *>(addr0 - shift);
                        uint4 *s4 = reinterpret_cast<uint4 *>(stage);
                        const uint32_t n_vec = (shift + n_bytes + 8u + 15u) >> 4;
                        for (uint32_t i = tid; i < n_vec; i += kTileThreads) s4[i] = __ldg(g + i);
                        __syncthreads();
                    }
                    for (int j = sub0 + tid; j < sub0 + n_sub; j += kTileThreads) {
                        const int32_t o = soff_s[j];
                        const uint32_t len = (uint32_t)(soff_s[j + 1] - o);
                        uint32_t id;
                        if (staged) {
                            const WordSrcShared src{stage, shift + (uint32_t)(o - byte0)};
                            id = bucket_of(src, len, spec, mask_empty && len == 0);
                        } else {
                            const uintptr_t a = reinterpret_cast<uintptr_t>(bytes) + (uintptr_t)o;
                            const WordSrcGlobal src{reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3), (uint32_t)(a & 3u)};
                            id = bucket_of(src, len, spec, mask_empty && len == 0);
                        }
                        sm.sid[j] = id;
                        if (ids_ws) ids_ws[k0 + j] = (int64_t)id;
                    }
                }
            } else if (HASH) {
                int32_t *soff_s = reinterpret_cast<int32_t *>(sm.meta);
                for (int j = tid; j <= n_keys; j += kTileThreads) soff_s[j] = soffs[k0 + j];
                __syncthreads();
                for (int j = tid; j < n_keys; j += kTileThreads) {
                    const int32_t o = soff_s[j];
                    const uint32_t len = (uint32_t)(soff_s[j + 1] - o);
                    const uintptr_t a = reinterpret_cast<uintptr_t>(bytes) + (uintptr_t)o;
                    const WordSrcGlobal src{reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3), (uint32_t)(a & 3u)};
                    const uint32_t id = bucket_of(src, len, spec, mask_empty && len == 0);
                    sm.sid[j] = id;
                    if (ids_ws) ids_ws[k0 + j] = (int64_t)id;
                }
            } else {
                for (int j = tid; j < n_keys; j += kTileThreads) sm.sid[j] = (uint32_t)ids[k0 + j];
            }
            for (int j = tid; j < nb * kMaxWorld; j += kTileThreads) (&sm.cnt[0][0])[j] = 0;
            __syncthreads();
            // ---- 2. owner, rank inside the (bag, owner) run ----
            for (int bl = wid; bl < nb; bl += kTileThreads / 32) {
                int64_t lo, hi;
                if (STAGE) {
                    lo = tb[r0 - t0 + bl];
                    hi = tb[r0 - t0 + bl + 1];
                } else {
                    bag_range(boffs, bag_len, r0 + bl, lo, hi);
                }
                const int j0 = (int)(lo - k0), j1 = (int)(hi - k0);
                for (int j = j0; j < j1; j += 32) {
                    const bool on = j + lane < j1;
                    const uint32_t id = on ? sm.sid[j + lane] : 0u;
                    const uint32_t owner = on ? (pow2 ? (id & (uint32_t)(world - 1)) : id % (uint32_t)world) : 0xffu;
                    const unsigned grp = __match_any_sync(0xffffffffu, owner);
                    const int leader = __ffs(grp) - 1;
                    int base = 0;
                    if (on && lane == leader) {
                        base = sm.cnt[bl][owner];
                        sm.cnt[bl][owner] = (uint16_t)(base + __popc(grp));
                    }
                    base = __shfl_sync(0xffffffffu, base, leader);
                    if (on) sm.meta[j + lane] = (uint32_t)(base + __popc(grp & lt)) | (owner << 12) | ((uint32_t)bl << 16);
                    __syncwarp();
                }
            }
            __syncthreads();
            // ---- 3. per owner: exclusive scan of the counts over the bags (one warp per owner) ----
            for (int g = wid; g < world; g += kTileThreads / 32) {
                int c[kTileBags / 32], sum = 0;
#pragma unroll
                for (int i = 0; i < kTileBags / 32; ++i) {
                    const int bl = lane * (kTileBags / 32) + i;
                    c[i] = bl < nb ? sm.cnt[bl][g] : 0;
                    sum += c[i];
                }
                int incl = sum;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int y = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += y;
                }
                int run = incl - sum;
#pragma unroll
                for (int i = 0; i < kTileBags / 32; ++i) {
                    const int bl = lane * (kTileBags / 32) + i;
                    if (bl < nb) sm.beg[bl][g] = (uint16_t)run;
                    run += c[i];
                }
                if (lane == 31) sm.seg_len[g] = incl;
            }
            __syncthreads();
            if (tid == 0) {
                int acc = 0;
                for (int g = 0; g < world; ++g) {
                    sm.seg_base[g] = acc;
                    acc += sm.seg_len[g];
                }
                sm.seg_base[world] = acc;
            }
            __syncthreads();
            // ---- 4. place every key's owner-local row ----
            for (int j = tid; j < n_keys; j += kTileThreads) {
                const uint32_t m = sm.meta[j], id = sm.sid[j];
                const uint32_t rank = m & 0xfffu, owner = (m >> 12) & 0xfu, bl = m >> 16;
                const uint32_t row = pow2 ? id >> wshift : id / (uint32_t)world;
                sm.srow[sm.seg_base[owner] + sm.beg[bl][owner] + rank] = row;
            }
            __syncthreads();
            // ---- 5. stream out: owner g's run goes to [k0, k0 + seg_len[g]) of its buffer for this source ----
            // one pass over the staged keys: key i belongs to the owner whose segment [seg_base[g], seg_base[g+1]) holds
            // it; consecutive threads write consecutive slots of that owner's run (all 256 threads busy, instead of one
            // short loop per owner)
            for (int i = tid; i < n_keys; i += kTileThreads) {
                int g = 0;
                while (g + 1 < world && i >= sm.seg_base[g + 1]) ++g;
                static_cast<int64_t *>(rows_dst.p[g])[k0 + (i - sm.seg_base[g])] = (int64_t)sm.srow[i];
            }
            for (int e = tid; e < nb * world; e += kTileThreads) {
                const int g = e / nb, bl = e - g * nb;
                const int32_t b = (int32_t)k0 + sm.beg[bl][g];
                static_cast<int32_t *>(begin_dst.p[g])[r0 + bl] = b;
                static_cast<int32_t *>(end_dst.p[g])[r0 + bl] = b + sm.cnt[bl][g];
            }
            __syncthreads();
            r0 = r1;
        }
    }
}

// out[b] = reduce over g = 0..world-1 (in that order) of partials[g][b]; avg divides by the bag's
// key count.  One float4 (or float) per thread.
template <bool VEC>
__global__ void __launch_bounds__(256) combine_partials_kernel(const float *__restrict__ partials, int world, int64_t batch,
                                                               int dim, int combiner, int bag_len,
                                                               const int32_t *__restrict__ boffs, float *__restrict__ out,
                                                               int64_t out_stride) {
    const int per_row = VEC ? dim >> 2 : dim;
    const int64_t total = batch * per_row;
    const int64_t plane = batch * (int64_t)dim;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = e / per_row;
        const int c = (int)(e - b * per_row);
        const int64_t cnt = boffs ? (int64_t)(boffs[b + 1] - boffs[b]) : (int64_t)bag_len;
        if (VEC) {
            const float4 *p = reinterpret_cast<const float4 *>(partials + b * dim) + c;
            float4 acc = __ldg(p);
            for (int g = 1; g < world; ++g) {
                const float4 x = __ldg(reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(p) + g * plane));
                if (combiner <= RF_COMBINER_AVG) {
                    acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
                } else if (combiner == RF_COMBINER_MIN) {
                    acc.x = x.x < acc.x ? x.x : acc.x; acc.y = x.y < acc.y ? x.y : acc.y;
                    acc.z = x.z < acc.z ? x.z : acc.z; acc.w = x.w < acc.w ? x.w : acc.w;
                } else {
                    acc.x = x.x > acc.x ? x.x : acc.x; acc.y = x.y > acc.y ? x.y : acc.y;
                    acc.z = x.z > acc.z ? x.z : acc.z; acc.w = x.w > acc.w ? x.w : acc.w;
                }
            }
            if (cnt == 0) acc = make_float4(0.f, 0.f, 0.f, 0.f);
            else if (combiner == RF_COMBINER_AVG) {
                const float d = (float)cnt;
                acc.x = acc.x / d; acc.y = acc.y / d; acc.z = acc.z / d; acc.w = acc.w / d;
            }
            reinterpret_cast<float4 *>(out + b * out_stride)[c] = acc;
        } else {
            const float *p = partials + b * dim + c;
            float acc = __ldg(p);
            for (int g = 1; g < world; ++g) {
                const float x = __ldg(p + g * plane);
                if (combiner <= RF_COMBINER_AVG) acc += x;
                else if (combiner == RF_COMBINER_MIN) acc = x < acc ? x : acc;
                else acc = x > acc ? x : acc;
            }
            if (cnt == 0) acc = 0.f;
            else if (combiner == RF_COMBINER_AVG) acc = acc / (float)cnt;
            out[b * out_stride + c] = acc;
        }
    }
}

// Cross-GPU barrier on peer-mapped signal pads (one CTA): thread t tells rank t "rank `me` has reached `value`" and
// waits until rank t has told this rank the same.  Everything this stream wrote to peers before the barrier is
// released with a system-scope fence first, so after the barrier every rank sees every other rank's earlier writes.
// One rank per GPU (ranks sharing a GPU would wait on kernels that cannot be co-scheduled).  Bounded spin: traps.
__global__ void __launch_bounds__(32) shard_barrier_kernel(const __grid_constant__ PtrTable signals, int me, int world, int slot,
                                                           uint32_t value) {
    const int t = threadIdx.x;
    if (t >= world) return;
    __threadfence_system();
    uint32_t *remote = static_cast<uint32_t *>(signals.p[t]) + slot * kMaxWorld + me;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(value) : "memory");
    const uint32_t *mine = static_cast<const uint32_t *>(signals.p[me]) + slot * kMaxWorld + t;
    uint32_t seen = 0;
    for (uint64_t spin = 0;; ++spin) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
        if ((int32_t)(seen - value) >= 0) break;
        if (spin > (1ull << 31)) __trap();
    }
    __threadfence_system();
}

static int grid_for(int64_t work_items, int per_block, int dev_sms) {
    int64_t blocks = (work_items + per_block - 1) / per_block;
    const int64_t cap = (int64_t)dev_sms * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace rf

using namespace rf;

extern "C" {

static int route_impl(const int64_t *d_ids, const uint8_t *d_bytes, const int32_t *d_str_offsets, const HashSpec *spec,
                      int mask_empty, int64_t *d_ids_ws, const int32_t *d_bag_offsets, int32_t bag_len, int64_t batch,
                      int world, int32_t *d_counts_ws, int32_t *d_offsets_local, int32_t *const *h_offsets_dst,
                      int64_t *const *h_rows_dst, void *stream) {
    if (world < 1 || world > kMaxWorld) return set_error(RF_ERR_INVALID, "world must be in [1, %d]", kMaxWorld);
    if (batch < 0 || batch > INT32_MAX) return set_error(RF_ERR_INVALID, "batch out of range");
    if (batch == 0) return RF_OK;
    if (!d_counts_ws || !d_offsets_local || !h_rows_dst || !h_offsets_dst)
        return set_error(RF_ERR_INVALID, "rf_shard_route: NULL buffer");
    if (!d_bag_offsets && bag_len < 0) return set_error(RF_ERR_INVALID, "negative bag_len");
    int dev = 0, sms = 0;
    RF_CUDA(cudaGetDevice(&dev));
    RF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    PtrTable offs{}, rows{};
    for (int g = 0; g < world; ++g) {
        offs.p[g] = h_offsets_dst ? h_offsets_dst[g] : nullptr;
        rows.p[g] = h_rows_dst[g];
        if (!rows.p[g]) return set_error(RF_ERR_INVALID, "rf_shard_route: rows_dst[%d] is NULL", g);
    }
    const int grid = grid_for(batch, 8, sms);   // 8 warps (bags) per 256-thread block
    // tiles of consecutive bags: ~4096 keys per CTA round (jagged: sized for 128 keys/bag, rounds adapt)
    const int64_t per_bag = d_bag_offsets ? 128 : (bag_len > 0 ? bag_len : 1);
    int64_t tile_bags = kScatCap / per_bag;
    if (tile_bags < 8) tile_bags = 8;
    if (tile_bags > 512) tile_bags = 512;
    const int sgrid = grid_for((batch + tile_bags - 1) / tile_bags, 1, sms);
    if (spec)
        shard_count_kernel<true><<<sgrid, kCntThreads, 0, st>>>(nullptr, d_bytes, d_str_offsets, *spec, mask_empty, d_ids_ws,
                                                                d_bag_offsets, bag_len, batch, world, (int)tile_bags, d_counts_ws);
    else
        shard_count_kernel<false><<<sgrid, kCntThreads, 0, st>>>(d_ids, nullptr, nullptr, HashSpec{}, 0, nullptr, d_bag_offsets,
                                                                 bag_len, batch, world, (int)tile_bags, d_counts_ws);
    const int n_chunks = (int)((batch + kScanChunk - 1) / kScanChunk);
    int32_t *excl = d_offsets_local;                         // [world][batch]
    int32_t *chunk_tot = d_offsets_local + (int64_t)world * batch;   // [world][n_chunks]
    shard_scan_kernel<<<dim3(n_chunks, world), kScanChunk, 0, st>>>(d_counts_ws, batch, n_chunks, excl, chunk_tot);
    shard_chunk_base_kernel<<<world, 1024, 0, st>>>(chunk_tot, n_chunks, batch, offs);
    // the scatter pass walks the same tiles
    static const bool tiled = !(getenv("RF_SCATTER_TILED") && atoi(getenv("RF_SCATTER_TILED")) == 0);
    if (tiled)
        shard_scatter_kernel<<<sgrid, kScatThreads, 0, st>>>(spec ? d_ids_ws : d_ids, d_bag_offsets, bag_len, batch, world,
                                                             n_chunks, (int)tile_bags, excl, chunk_tot, d_counts_ws, offs, rows);
    else
        shard_scatter_warp_kernel<<<grid, 256, 0, st>>>(spec ? d_ids_ws : d_ids, d_bag_offsets, bag_len, batch, world,
                                                        n_chunks, excl, chunk_tot, offs, rows);
    RF_CUDA(cudaGetLastError());
    g_launches.fetch_add(4);
    return RF_OK;
}

int rf_shard_route_tiles_ex(const uint8_t *d_bytes, const int32_t *d_str_offsets, const int64_t *d_ids, int64_t num_bins,
                            int mask_mode, int use_strong, uint64_t key0, uint64_t key1, int64_t *d_ids_ws,
                            const int32_t *d_bag_offsets, int32_t bag_len, int64_t batch, int world, int64_t *const *h_rows_dst,
                            int32_t *const *h_begin_dst, int32_t *const *h_end_dst, int max_ctas_per_sm, void *stream) {
    if (max_ctas_per_sm < 0) return set_error(RF_ERR_INVALID, "max_ctas_per_sm must be >= 0");
    if (world < 1 || world > kMaxWorld) return set_error(RF_ERR_INVALID, "world must be in [1, %d]", kMaxWorld);
    if (batch < 0 || batch > INT32_MAX) return set_error(RF_ERR_INVALID, "batch out of range");
    if (batch == 0) return RF_OK;
    const bool hash = d_bytes != nullptr;
    if (hash == (d_ids != nullptr)) return set_error(RF_ERR_INVALID, "rf_shard_route_tiles: give string keys or ids, not both");
    if (hash && !d_str_offsets) return set_error(RF_ERR_INVALID, "rf_shard_route_tiles: NULL key buffer");
    if (!h_rows_dst || !h_begin_dst || !h_end_dst) return set_error(RF_ERR_INVALID, "rf_shard_route_tiles: NULL destination table");
    if (!d_bag_offsets && bag_len < 0) return set_error(RF_ERR_INVALID, "negative bag_len");
    HashSpec spec{};
    if (hash) {
        if (num_bins <= 0) return set_error(RF_ERR_INVALID, "`num_bins` cannot be `None` or non-positive values.");
        if (num_bins > 0xffffffffLL) return set_error(RF_ERR_UNSUPPORTED, "num_bins above 2^32-1 is not supported");
        if (mask_mode != RF_MASK_NONE && mask_mode != RF_MASK_EMPTY_STRING)
            return set_error(RF_ERR_INVALID, "string keys take RF_MASK_NONE or RF_MASK_EMPTY_STRING");
        spec = make_hash_spec(num_bins, mask_mode != RF_MASK_NONE, use_strong, key0, key1);
    }
    PtrTable rows{}, begs{}, ends{};
    for (int g = 0; g < world; ++g) {
        rows.p[g] = h_rows_dst[g];
        begs.p[g] = h_begin_dst[g];
        ends.p[g] = h_end_dst[g];
        if (!rows.p[g] || !begs.p[g] || !ends.p[g]) return set_error(RF_ERR_INVALID, "rf_shard_route_tiles: destination %d is NULL", g);
    }
    int dev = 0, sms = 0;
    RF_CUDA(cudaGetDevice(&dev));
    RF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // a tile is about one round (<= 4096 keys, <= 128 bags); jagged bag lengths are only known on the
    // device, so tiles are sized for 128 keys per bag there and the rounds adapt
    const int64_t per_bag = d_bag_offsets ? 128 : (bag_len > 0 ? bag_len : 1);
    int64_t tile_bags = kTileCap / per_bag;
    if (tile_bags < 8) tile_bags = 8;
    if (tile_bags > kTileBags) tile_bags = kTileBags;
    int grid = grid_for((batch + tile_bags - 1) / tile_bags, 1, sms);
    // experiments: RF_ROUTE_CTAS_PER_SM caps the grid (the kernel walks its tiles grid-stride), e.g. 1 leaves the
    // rest of every SM to a concurrently running pooling kernel
    static const int env_cap = getenv("RF_ROUTE_CTAS_PER_SM") ? atoi(getenv("RF_ROUTE_CTAS_PER_SM")) : -1;
    const int route_cap = env_cap >= 0 ? env_cap : max_ctas_per_sm;
    if (route_cap > 0 && grid > route_cap * sms) grid = route_cap * sms;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t smem = sizeof(TileSmem);
    static const bool stage = !(getenv("RF_ROUTE_STAGE") && atoi(getenv("RF_ROUTE_STAGE")) == 0);
#define RF_ROUTE_LAUNCH(H, S, ...)                                                                                      \
    do {                                                                                                                \
        RF_CUDA(cudaFuncSetAttribute(shard_route_tile_kernel<H, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        shard_route_tile_kernel<H, S><<<grid, kTileThreads, smem, st>>>(__VA_ARGS__);                                    \
    } while (0)
    if (hash && stage)
        RF_ROUTE_LAUNCH(true, true, nullptr, d_bytes, d_str_offsets, spec, mask_mode == RF_MASK_EMPTY_STRING, d_ids_ws, d_bag_offsets,
                        bag_len, batch, world, (int)tile_bags, rows, begs, ends);
    else if (hash)
        RF_ROUTE_LAUNCH(true, false, nullptr, d_bytes, d_str_offsets, spec, mask_mode == RF_MASK_EMPTY_STRING, d_ids_ws, d_bag_offsets,
                        bag_len, batch, world, (int)tile_bags, rows, begs, ends);
    else if (stage)
        RF_ROUTE_LAUNCH(false, true, d_ids, nullptr, nullptr, spec, 0, nullptr, d_bag_offsets, bag_len, batch, world, (int)tile_bags,
                        rows, begs, ends);
    else
        RF_ROUTE_LAUNCH(false, false, d_ids, nullptr, nullptr, spec, 0, nullptr, d_bag_offsets, bag_len, batch, world, (int)tile_bags,
                        rows, begs, ends);
#undef RF_ROUTE_LAUNCH
    RF_CUDA(cudaGetLastError());
    g_launches.fetch_add(1);
    return RF_OK;
}

int rf_shard_route_tiles(const uint8_t *d_bytes, const int32_t *d_str_offsets, const int64_t *d_ids, int64_t num_bins,
                         int mask_mode, int use_strong, uint64_t key0, uint64_t key1, int64_t *d_ids_ws,
                         const int32_t *d_bag_offsets, int32_t bag_len, int64_t batch, int world, int64_t *const *h_rows_dst,
                         int32_t *const *h_begin_dst, int32_t *const *h_end_dst, void *stream) {
    return rf_shard_route_tiles_ex(d_bytes, d_str_offsets, d_ids, num_bins, mask_mode, use_strong, key0, key1, d_ids_ws, d_bag_offsets,
                                   bag_len, batch, world, h_rows_dst, h_begin_dst, h_end_dst, 0, stream);
}

int rf_shard_route(const int64_t *d_ids, const int32_t *d_bag_offsets, int32_t bag_len, int64_t batch, int world,
                   int32_t *d_counts_ws, int32_t *d_offsets_local, int32_t *const *h_offsets_dst,
                   int64_t *const *h_rows_dst, void *stream) {
    if (batch > 0 && !d_ids) return set_error(RF_ERR_INVALID, "rf_shard_route: ids is NULL");
    return route_impl(d_ids, nullptr, nullptr, nullptr, 0, nullptr, d_bag_offsets, bag_len, batch, world, d_counts_ws,
                      d_offsets_local, h_offsets_dst, h_rows_dst, stream);
}

int rf_shard_route_keys(const uint8_t *d_bytes, const int32_t *d_str_offsets, int64_t num_bins, int mask_mode,
                        int use_strong, uint64_t key0, uint64_t key1, int64_t *d_ids_ws, const int32_t *d_bag_offsets,
                        int32_t bag_len, int64_t batch, int world, int32_t *d_counts_ws, int32_t *d_offsets_local,
                        int32_t *const *h_offsets_dst, int64_t *const *h_rows_dst, void *stream) {
    if (batch > 0 && (!d_bytes || !d_str_offsets || !d_ids_ws))
        return set_error(RF_ERR_INVALID, "rf_shard_route_keys: NULL key buffer");
    if (num_bins <= 0) return set_error(RF_ERR_INVALID, "`num_bins` cannot be `None` or non-positive values.");
    if (num_bins > 0xffffffffLL) return set_error(RF_ERR_UNSUPPORTED, "num_bins above 2^32-1 is not supported");
    if (mask_mode != RF_MASK_NONE && mask_mode != RF_MASK_EMPTY_STRING)
        return set_error(RF_ERR_INVALID, "string keys take RF_MASK_NONE or RF_MASK_EMPTY_STRING");
    const HashSpec spec = make_hash_spec(num_bins, mask_mode != RF_MASK_NONE, use_strong, key0, key1);
    return route_impl(nullptr, d_bytes, d_str_offsets, &spec, mask_mode == RF_MASK_EMPTY_STRING, d_ids_ws, d_bag_offsets,
                      bag_len, batch, world, d_counts_ws, d_offsets_local, h_offsets_dst, h_rows_dst, stream);
}

int rf_combine_partials(const float *d_partials, int world, int64_t batch, int32_t dim, int combiner, int32_t bag_len,
                        const int32_t *d_bag_offsets, float *d_out, int64_t out_stride, void *stream) {
    if (world < 1 || world > kMaxWorld) return set_error(RF_ERR_INVALID, "world must be in [1, %d]", kMaxWorld);
    if (batch < 0 || dim <= 0) return set_error(RF_ERR_INVALID, "bad batch / dim");
    if (combiner < RF_COMBINER_SUM || combiner > RF_COMBINER_MAX) return set_error(RF_ERR_INVALID, "bad combiner");
    if (batch == 0) return RF_OK;
    if (!d_partials || !d_out) return set_error(RF_ERR_INVALID, "rf_combine_partials: NULL buffer");
    int dev = 0, sms = 0;
    RF_CUDA(cudaGetDevice(&dev));
    RF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec = dim % 4 == 0 && out_stride % 4 == 0 && reinterpret_cast<uintptr_t>(d_partials) % 16 == 0 &&
                     reinterpret_cast<uintptr_t>(d_out) % 16 == 0;
    const int64_t work = batch * (vec ? dim / 4 : dim);
    const int grid = grid_for(work, 256, sms);
    if (vec)
        combine_partials_kernel<true><<<grid, 256, 0, st>>>(d_partials, world, batch, dim, combiner, bag_len, d_bag_offsets,
                                                           d_out, out_stride);
    else
        combine_partials_kernel<false><<<grid, 256, 0, st>>>(d_partials, world, batch, dim, combiner, bag_len,
                                                            d_bag_offsets, d_out, out_stride);
    RF_CUDA(cudaGetLastError());
    g_launches.fetch_add(1);
    return RF_OK;
}


int64_t rf_shard_exchange_bytes(int world, int64_t max_batch, int64_t max_keys, int32_t dim) {
    if (world < 1 || world > kMaxWorld || max_batch <= 0 || max_keys <= 0 || dim <= 0) return -1;
    const int64_t rows = (int64_t)world * max_keys * 8;
    const int64_t bounds = ((2 * (int64_t)world * max_batch * 4 + 15) / 16) * 16;
    return rows + bounds + (int64_t)world * max_batch * dim * 4;
}

int rf_sharded_bag_forward(const rf_shard_ctx *ctx, const uint8_t *d_bytes, const int32_t *d_str_offsets, const int64_t *d_ids,
                           int64_t num_bins, int mask_mode, int use_strong, uint64_t key0, uint64_t key1,
                           const int32_t *d_bag_offsets, int32_t bag_len, int64_t batch, const float *d_shard, int64_t shard_rows,
                           int combiner, uint64_t step, float *d_out, int64_t out_stride, void *stream) {
    if (!ctx) return set_error(RF_ERR_INVALID, "rf_sharded_bag_forward: NULL context");
    const int W = ctx->world, me = ctx->rank;
    if (W < 1 || W > kMaxWorld || me < 0 || me >= W) return set_error(RF_ERR_INVALID, "bad rank / world");
    if (batch <= 0 || batch > ctx->max_batch) return set_error(RF_ERR_INVALID, "batch must be in [1, max_batch]");
    if (step == 0 || step > 0x7fffffffull) return set_error(RF_ERR_INVALID, "step must count up from 1 (same value on every rank)");
    if (combiner < RF_COMBINER_SUM || combiner > RF_COMBINER_MAX) return set_error(RF_ERR_INVALID, "bad combiner");
    if (!d_shard || !d_out) return set_error(RF_ERR_INVALID, "rf_sharded_bag_forward: NULL shard / out");
    const int64_t B = ctx->max_batch, K = ctx->max_keys, D = ctx->dim;
    const int64_t rows_bytes = (int64_t)W * K * 8;
    const int64_t bounds_bytes = ((2 * (int64_t)W * B * 4 + 15) / 16) * 16;
    PtrTable sig{};
    int64_t *rows_dst[kMaxWorld];
    int32_t *beg_dst[kMaxWorld], *end_dst[kMaxWorld];
    for (int g = 0; g < W; ++g) {
        char *base = static_cast<char *>(ctx->peer_exchange[g]);
        if (!base || !ctx->peer_signals[g]) return set_error(RF_ERR_INVALID, "rf_sharded_bag_forward: peer %d is not mapped", g);
        rows_dst[g] = reinterpret_cast<int64_t *>(base) + (int64_t)me * K;                 // my slice of owner g's buffers
        beg_dst[g] = reinterpret_cast<int32_t *>(base + rows_bytes) + (int64_t)me * B;
        end_dst[g] = reinterpret_cast<int32_t *>(base + rows_bytes) + (int64_t)(W + me) * B;
        sig.p[g] = ctx->peer_signals[g];
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // 1. route my keys to their owners (peer stores)
    int rc = rf_shard_route_tiles(d_bytes, d_str_offsets, d_ids, num_bins, mask_mode, use_strong, key0, key1, nullptr, d_bag_offsets,
                                  bag_len, batch, W, rows_dst, beg_dst, end_dst, stream);
    if (rc != RF_OK) return rc;
    // 2. every source's routing has landed
    shard_barrier_kernel<<<1, 32, 0, st>>>(sig, me, W, 0, (uint32_t)step);
    // 3. pool what each source asked of my shard, straight into that source's partial plane `me`
    char *mine = static_cast<char *>(ctx->peer_exchange[me]);
    rf_field_desc fields[kMaxWorld];
    memset(fields, 0, sizeof(fields));
    for (int k = 0; k < W; ++k) {
        const int s = (me + k) % W;                    // rotated: at any moment the W owners write to W different ranks
        rf_field_desc &f = fields[k];
        f.ids = reinterpret_cast<const int64_t *>(mine) + (int64_t)s * K;
        f.bag_offsets = reinterpret_cast<const int32_t *>(mine + rows_bytes) + (int64_t)s * B;
        f.bag_ends = reinterpret_cast<const int32_t *>(mine + rows_bytes) + (int64_t)(W + s) * B;
        f.n_items = K / W > 0 ? K / W : 1;             // an estimate only steers tile sizing
        f.n_tables = 1;
        f.tables[0].weights = d_shard;
        f.tables[0].num_bins = shard_rows;
        f.dim = (int32_t)D;
        f.combiner = combiner == RF_COMBINER_AVG ? RF_COMBINER_SUM : combiner;
        f.flags = RF_FIELD_PARTIAL;
        f.out = reinterpret_cast<float *>(static_cast<char *>(ctx->peer_exchange[s]) + rows_bytes + bounds_bytes) + (int64_t)me * B * D;
        f.out_stride = D;
    }
    rc = rf_bag_forward(fields, W, batch, stream);
    if (rc != RF_OK) return rc;
    // 4. every owner's partials have landed here
    shard_barrier_kernel<<<1, 32, 0, st>>>(sig, me, W, 1, (uint32_t)step);
    RF_CUDA(cudaGetLastError());
    g_launches.fetch_add(2);
    // 5. reduce the W partials in rank order (avg divides by the bag's key count)
    return rf_combine_partials(reinterpret_cast<const float *>(mine + rows_bytes + bounds_bytes), W, batch, (int32_t)D, combiner,
                               bag_len, d_bag_offsets, d_out, out_stride, stream);
}

}  // extern "C"
