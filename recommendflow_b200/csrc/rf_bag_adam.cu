// rf_bag_adam.cu -- backward of the pooled embedding bag fused with the Adam row update the
// reference trains with.
//
// Reference: `tf.keras.optimizers.Adam(learning_rate)` on the Embedding variables
// (example/ranking_search/train.py:97-104, example/recall_search/train.py:97).  Keras applies a
// sparse (IndexedSlices) gradient like this: duplicates are summed first
// (`_deduplicate_indexed_slices`), then -- non-lazily, for EVERY row of the variable --
//     m = m * b1 (+ g * (1 - b1) on touched rows)
//     v = v * b2 (+ g * g * (1 - b2) on touched rows)
//     w = w - lr_t * m / (sqrt(v) + eps),   lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t)
// so the moments of untouched rows decay and those rows keep moving.  This file does exactly that
// (`lazy = 0`): a sparse pass over the touched rows and a dense streaming pass over the rest of
// the table (HBM-bound: 6 x 4 B per element).  `lazy = 1` updates touched rows only (the LazyAdam
// variant; NOT the reference's semantics, offered because the dense pass costs 6 x the table
// bytes per step).
//
// Sparse pass: the pooled gradient flows to every gathered row (pads included, like the forward).
//   1. ids -> (uint32 key, position) pairs, cub::DeviceRadixSort (library plumbing; stable, so
//      positions inside a run stay ascending and the summation order is fixed),
//   2. cub::DeviceSelect::If compacts the run heads -> unique rows + run lengths, no host sync,
//   3. one lane group (dim / 4 lanes, 128-bit) per unique row sums its run in position order,
//      applies Adam to (w, m, v) in registers and marks the row in a bitmap; runs longer than
//      kHeavyRun (the pad row 0 of dense-padded batches) go to a list and are reduced by a whole
//      CTA each (slot-strided partial sums + ordered combine),
//   4. dense pass over rows whose bit is clear.
// All arithmetic uses the round-to-nearest intrinsics (no FMA contraction) so a plain C restatement
// reproduces it bit for bit wherever the summation order is the same.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <atomic>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "../../include/rf_b200.h"
#include "rf_common.h"

namespace rf {

extern std::atomic<int64_t> g_launches;

namespace {

constexpr int kAdamThreads = 256;
constexpr int kHeavyRun = 128;

struct AdamC {
    float b1, b2, omb1, omb2, lr_t, eps;
};

__device__ __forceinline__ void adam_scalar(float &w, float &m, float &v, float g, bool touched, const AdamC &c) {
    m = __fmul_rn(m, c.b1);
    v = __fmul_rn(v, c.b2);
    if (touched) {
        m = __fadd_rn(m, __fmul_rn(g, c.omb1));
        v = __fadd_rn(v, __fmul_rn(__fmul_rn(g, g), c.omb2));
    }
    w = __fsub_rn(w, __fdiv_rn(__fmul_rn(c.lr_t, m), __fadd_rn(__fsqrt_rn(v), c.eps)));
}

__device__ __forceinline__ void adam_vec(float4 &w, float4 &m, float4 &v, const float4 &g, bool touched, const AdamC &c) {
    adam_scalar(w.x, m.x, v.x, g.x, touched, c);
    adam_scalar(w.y, m.y, v.y, g.y, touched, c);
    adam_scalar(w.z, m.z, v.z, g.z, touched, c);
    adam_scalar(w.w, m.w, v.w, g.w, touched, c);
}

struct GradSrc {
    const float *grad;
    int64_t stride;
    const int32_t *boffs;
    int64_t batch;
    int bag_len;
    int avg;
};

// gradient that key `pos` receives for column group c: grad[bag(pos)] (x 1 / count for avg)
__device__ __forceinline__ float4 key_grad(const GradSrc &s, int pos, int c) {
    int64_t b;
    float scale = 1.0f;
    if (s.boffs) {
        int64_t lo = 0, hi = s.batch;
        while (lo + 1 < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (s.boffs[mid] <= pos) lo = mid; else hi = mid;
        }
        b = lo;
        if (s.avg) scale = __fdiv_rn(1.0f, (float)(s.boffs[b + 1] - s.boffs[b]));
    } else {
        b = pos / s.bag_len;
        if (s.avg) scale = __fdiv_rn(1.0f, (float)s.bag_len);
    }
    float4 g = __ldg(reinterpret_cast<const float4 *>(s.grad + b * s.stride) + c);
    if (s.avg) {
        g.x = __fmul_rn(g.x, scale);
        g.y = __fmul_rn(g.y, scale);
        g.z = __fmul_rn(g.z, scale);
        g.w = __fmul_rn(g.w, scale);
    }
    return g;
}

__device__ __forceinline__ void add4(float4 &a, const float4 &b) {
    a.x = __fadd_rn(a.x, b.x);
    a.y = __fadd_rn(a.y, b.y);
    a.z = __fadd_rn(a.z, b.z);
    a.w = __fadd_rn(a.w, b.w);
}

__global__ void __launch_bounds__(kAdamThreads) adam_prep_kernel(const int64_t *__restrict__ ids, int n, uint32_t *__restrict__ keys,
                                                                 int32_t *__restrict__ pos, int32_t *__restrict__ counters) {
    const int i = blockIdx.x * kAdamThreads + threadIdx.x;
    if (i == 0) counters[0] = counters[1] = 0;
    if (i < n) {
        keys[i] = (uint32_t)ids[i];
        pos[i] = i;
    }
}

struct HeadPred {
    const uint32_t *keys;
    __device__ __forceinline__ bool operator()(int i) const { return i == 0 || keys[i] != keys[i - 1]; }
};

struct Tables {
    float *w, *m, *v;
    uint32_t *bitmap;   // NULL in lazy mode
    int per;            // dim / 4
};

__device__ __forceinline__ void apply_row(const Tables &t, uint32_t row, int c, const float4 &g, const AdamC &k) {
    const int64_t at = (int64_t)row * t.per + c;
    float4 w = reinterpret_cast<float4 *>(t.w)[at], m = reinterpret_cast<float4 *>(t.m)[at], v = reinterpret_cast<float4 *>(t.v)[at];
    adam_vec(w, m, v, g, true, k);
    reinterpret_cast<float4 *>(t.w)[at] = w;
    reinterpret_cast<float4 *>(t.m)[at] = m;
    reinterpret_cast<float4 *>(t.v)[at] = v;
    if (c == 0 && t.bitmap) atomicOr(t.bitmap + (row >> 5), 1u << (row & 31));
}

// one lane group per unique row; *n_unique is read from device memory (written by the select)
__global__ void __launch_bounds__(kAdamThreads)
adam_rows_kernel(const uint32_t *__restrict__ keys, const int32_t *__restrict__ pos, const int32_t *__restrict__ heads,
                 int32_t *__restrict__ counters, int n_keys, int2 *__restrict__ heavy, GradSrc src, Tables t, AdamC k) {
    const int groups = kAdamThreads / t.per;
    const int grp = threadIdx.x / t.per, c = threadIdx.x - grp * t.per;
    if (grp >= groups) return;
    const int n_unique = counters[0];
    for (int u = blockIdx.x * groups + grp; u < n_unique; u += gridDim.x * groups) {
        const int begin = heads[u], end = u + 1 < n_unique ? heads[u + 1] : n_keys;
        if (end - begin > kHeavyRun) {
            if (c == 0) heavy[atomicAdd(counters + 1, 1)] = make_int2(begin, end);
            continue;
        }
        float4 g = key_grad(src, pos[begin], c);
        for (int i = begin + 1; i < end; ++i) add4(g, key_grad(src, pos[i], c));
        apply_row(t, keys[begin], c, g, k);
    }
}

// one CTA per long run: slot s sums keys begin + s, begin + s + slots, ...; slots are combined in order
__global__ void __launch_bounds__(kAdamThreads)
adam_heavy_kernel(const uint32_t *__restrict__ keys, const int32_t *__restrict__ pos, const int32_t *__restrict__ counters,
                  const int2 *__restrict__ heavy, GradSrc src, Tables t, AdamC k) {
    __shared__ float4 part[kAdamThreads];
    const int slots = kAdamThreads / t.per;
    const int s = threadIdx.x / t.per, c = threadIdx.x - s * t.per;
    const int n_heavy = counters[1];
    for (int h = blockIdx.x; h < n_heavy; h += gridDim.x) {
        const int2 run = heavy[h];
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s < slots)
            for (int i = run.x + s; i < run.y; i += slots) add4(g, key_grad(src, pos[i], c));
        part[threadIdx.x] = g;
        __syncthreads();
        if (s == 0) {
            for (int j = 1; j < slots; ++j) add4(g, part[j * t.per + c]);
            apply_row(t, keys[run.x], c, g, k);
        }
        __syncthreads();
    }
}

// every row whose bit is clear: decay the moments and move the row (Keras' non-lazy sparse Adam)
__global__ void __launch_bounds__(kAdamThreads)
adam_dense_kernel(Tables t, int64_t rows, AdamC k) {
    const int64_t total = rows * t.per;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t e = (int64_t)blockIdx.x * kAdamThreads + threadIdx.x; e < total; e += (int64_t)gridDim.x * kAdamThreads) {
        const int64_t row = e / t.per;
        if ((__ldg(t.bitmap + (row >> 5)) >> (row & 31)) & 1u) continue;
        float4 w = reinterpret_cast<float4 *>(t.w)[e], m = reinterpret_cast<float4 *>(t.m)[e], v = reinterpret_cast<float4 *>(t.v)[e];
        adam_vec(w, m, v, zero, false, k);
        reinterpret_cast<float4 *>(t.w)[e] = w;
        reinterpret_cast<float4 *>(t.m)[e] = m;
        reinterpret_cast<float4 *>(t.v)[e] = v;
    }
}

inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

struct Workspace {
    uint32_t *keys_in, *keys_out;
    int32_t *pos_in, *pos_out, *heads, *counters;
    int2 *heavy;
    uint32_t *bitmap;
    void *cub_temp;
    size_t cub_bytes, bitmap_bytes, total;
};

int key_bits(int64_t rows) {
    int bits = 1;
    while (bits < 32 && ((int64_t)1 << bits) < rows) ++bits;
    return bits;
}

int carve(Workspace &ws, char *base, int64_t n_keys, int64_t rows) {
    const int n = (int)n_keys;
    size_t sort_bytes = 0, select_bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                                    (const int32_t *)nullptr, (int32_t *)nullptr, n, 0, key_bits(rows));
    if (e != cudaSuccess) return set_error(RF_ERR_CUDA, "cub sort size query failed: %s", cudaGetErrorString(e));
    thrust::counting_iterator<int> iota(0);
    e = cub::DeviceSelect::If(nullptr, select_bytes, iota, (int32_t *)nullptr, (int32_t *)nullptr, n, HeadPred{nullptr});
    if (e != cudaSuccess) return set_error(RF_ERR_CUDA, "cub select size query failed: %s", cudaGetErrorString(e));
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char *p = base ? base + off : nullptr;
        off += up256(bytes);
        return p;
    };
    ws.keys_in = reinterpret_cast<uint32_t *>(take(sizeof(uint32_t) * n));
    ws.keys_out = reinterpret_cast<uint32_t *>(take(sizeof(uint32_t) * n));
    ws.pos_in = reinterpret_cast<int32_t *>(take(sizeof(int32_t) * n));
    ws.pos_out = reinterpret_cast<int32_t *>(take(sizeof(int32_t) * n));
    ws.heads = reinterpret_cast<int32_t *>(take(sizeof(int32_t) * ((size_t)n + 1)));
    ws.counters = reinterpret_cast<int32_t *>(take(sizeof(int32_t) * 4));
    ws.heavy = reinterpret_cast<int2 *>(take(sizeof(int2) * ((size_t)n / kHeavyRun + 1)));
    ws.bitmap_bytes = sizeof(uint32_t) * (size_t)((rows + 31) / 32);
    ws.bitmap = reinterpret_cast<uint32_t *>(take(ws.bitmap_bytes));
    ws.cub_bytes = sort_bytes > select_bytes ? sort_bytes : select_bytes;
    ws.cub_temp = take(ws.cub_bytes);
    ws.total = off;
    return RF_OK;
}

}  // namespace
}  // namespace rf

using namespace rf;

extern "C" {

int64_t rf_bag_adam_workspace_bytes(int64_t n_keys, int64_t table_rows) {
    if (n_keys < 0 || n_keys > INT32_MAX || table_rows <= 0 || table_rows > (int64_t)UINT32_MAX) {
        set_error(RF_ERR_INVALID, "rf_bag_adam_workspace_bytes: n_keys / table_rows out of range");
        return -1;
    }
    Workspace ws;
    if (carve(ws, nullptr, n_keys, table_rows) != RF_OK) return -1;
    return (int64_t)ws.total;
}

int rf_bag_backward_adam(const int64_t *d_ids, int64_t n_keys, const int32_t *d_bag_offsets, int32_t bag_len, int64_t batch,
                         const float *d_grad_out, int64_t grad_stride, int32_t dim, int combiner, const rf_adam_params *params,
                         float *d_table, float *d_m, float *d_v, int64_t table_rows, void *d_workspace, int64_t workspace_bytes,
                         void *stream) {
    if (!params) return set_error(RF_ERR_INVALID, "rf_bag_backward_adam: params is NULL");
    if (n_keys < 0 || n_keys > INT32_MAX || batch < 0 || dim <= 0 || table_rows <= 0 || table_rows > (int64_t)UINT32_MAX)
        return set_error(RF_ERR_INVALID, "bad backward shape");
    if (combiner != RF_COMBINER_SUM && combiner != RF_COMBINER_AVG)
        return set_error(RF_ERR_UNSUPPORTED, "backward is implemented for the sum and avg combiners");
    if (params->step < 1) return set_error(RF_ERR_INVALID, "Adam step must be >= 1, got %lld", (long long)params->step);
    if (dim % 4 != 0 || dim > 4 * kAdamThreads || grad_stride % 4 != 0)
        return set_error(RF_ERR_UNSUPPORTED, "rf_bag_backward_adam needs dim %% 4 == 0, dim <= %d and grad_stride %% 4 == 0", 4 * kAdamThreads);
    if (!d_table || !d_m || !d_v) return set_error(RF_ERR_INVALID, "rf_bag_backward_adam: NULL table / moment buffer");
    if (n_keys > 0 && (!d_ids || !d_grad_out)) return set_error(RF_ERR_INVALID, "rf_bag_backward_adam: NULL ids / grad_out");
    if (n_keys > 0 && !d_bag_offsets && (bag_len <= 0 || batch * (int64_t)bag_len != n_keys))
        return set_error(RF_ERR_INVALID, "dense backward: batch x bag_len != n_keys");
    for (const void *p : {(const void *)d_table, (const void *)d_m, (const void *)d_v, (const void *)d_grad_out})
        if (reinterpret_cast<uintptr_t>(p) % 16 != 0) return set_error(RF_ERR_INVALID, "rf_bag_backward_adam: buffers must be 16-byte aligned");
    Workspace ws;
    int rc = carve(ws, static_cast<char *>(d_workspace), n_keys, table_rows);
    if (rc != RF_OK) return rc;
    if (!d_workspace || workspace_bytes < (int64_t)ws.total)
        return set_error(RF_ERR_INVALID, "workspace too small: %lld < %lld bytes (rf_bag_adam_workspace_bytes)",
                         (long long)workspace_bytes, (long long)ws.total);

    // Keras: lr_t = lr * sqrt(1 - beta_2^t) / (1 - beta_1^t), all in fp32
    AdamC k;
    k.b1 = params->beta1;
    k.b2 = params->beta2;
    k.omb1 = 1.0f - params->beta1;
    k.omb2 = 1.0f - params->beta2;
    k.eps = params->epsilon;
    const float b1p = powf(params->beta1, (float)params->step), b2p = powf(params->beta2, (float)params->step);
    k.lr_t = params->lr * sqrtf(1.0f - b2p) / (1.0f - b1p);

    int dev = 0, sms = 0;
    RF_CUDA(cudaGetDevice(&dev));
    RF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool lazy = params->lazy != 0;
    Tables t{d_table, d_m, d_v, lazy ? nullptr : ws.bitmap, dim / 4};
    if (!lazy) RF_CUDA(cudaMemsetAsync(ws.bitmap, 0, ws.bitmap_bytes, st));
    int launches = 0;
    if (n_keys > 0) {
        const int n = (int)n_keys;
        adam_prep_kernel<<<(n + kAdamThreads - 1) / kAdamThreads, kAdamThreads, 0, st>>>(d_ids, n, ws.keys_in, ws.pos_in, ws.counters);
        RF_CUDA(cudaGetLastError());
        size_t tmp = ws.cub_bytes;
        RF_CUDA(cub::DeviceRadixSort::SortPairs(ws.cub_temp, tmp, ws.keys_in, ws.keys_out, ws.pos_in, ws.pos_out, n, 0,
                                                key_bits(table_rows), st));
        thrust::counting_iterator<int> iota(0);
        tmp = ws.cub_bytes;
        RF_CUDA(cub::DeviceSelect::If(ws.cub_temp, tmp, iota, ws.heads, ws.counters, n, HeadPred{ws.keys_out}, st));
        GradSrc src{d_grad_out, grad_stride, d_bag_offsets, batch, bag_len, combiner == RF_COMBINER_AVG};
        const int groups = kAdamThreads / t.per;
        int64_t blocks = ((int64_t)n + groups - 1) / groups;
        if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
        adam_rows_kernel<<<(unsigned)blocks, kAdamThreads, 0, st>>>(ws.keys_out, ws.pos_out, ws.heads, ws.counters, n, ws.heavy, src, t, k);
        RF_CUDA(cudaGetLastError());
        int64_t hblocks = (int64_t)n / kHeavyRun + 1;
        if (hblocks > (int64_t)sms * 4) hblocks = (int64_t)sms * 4;
        adam_heavy_kernel<<<(unsigned)hblocks, kAdamThreads, 0, st>>>(ws.keys_out, ws.pos_out, ws.counters, ws.heavy, src, t, k);
        RF_CUDA(cudaGetLastError());
        launches += 3;
    }
    if (!lazy) {
        int64_t blocks = (table_rows * t.per + kAdamThreads - 1) / kAdamThreads;
        if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
        adam_dense_kernel<<<(unsigned)blocks, kAdamThreads, 0, st>>>(t, table_rows, k);
        RF_CUDA(cudaGetLastError());
        ++launches;
    }
    g_launches.fetch_add(launches);
    return RF_OK;
}

}  // extern "C"
