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
// Any number of tables that share `dim` go through ONE such pass (rf_bag_backward_adam_multi): rows and keys
// are numbered globally across the tables, so the sort, the select and the four kernels are launched once
// per step instead of once per table (456 tables in the C3 plan).
// All arithmetic uses the round-to-nearest intrinsics (no FMA contraction) so a plain C restatement
// reproduces it bit for bit wherever the summation order is the same.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <string.h>

#include <atomic>
#include <vector>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "../../include/rf_b200.h"
#include "rf_common.h"

namespace rf {

extern std::atomic<int64_t> g_launches;
int graph_desc_slots(int dev, size_t bytes, char **host, char **device);       // rf_bag.cu
int graph_desc_pool_ready(int dev);
int graph_desc_upload(char *device, const char *host, size_t bytes, cudaStream_t stream);

namespace {

constexpr int kAdamThreads = 256;
constexpr int kHeavyRun = 128;

struct AdamC {
    float b1, b2, omb1, omb2, lr_t, eps;
    const float *d_lr_t;          // when set, lr_t is read from device memory (a step captured into a CUDA graph)
};

__device__ __forceinline__ AdamC resolved(AdamC k) {
    if (k.d_lr_t) k.lr_t = __ldg(k.d_lr_t);
    return k;
}

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

// One table of a (possibly multi-table) update, as the device sees it.  Tables of one call share `dim`; their
// rows are numbered globally (rbase + row) so that ONE sort / select / update pass covers all of them, and
// their keys are numbered globally (kbase + k) so that a sorted position leads back to its bag.
struct DevAdamField {
    const int64_t *ids;
    const int32_t *boffs;
    const float *grad;
    int64_t grad_stride;
    float *w, *m, *v;
    int64_t rows;
    uint32_t rbase;
    int32_t kbase;
    int32_t n_keys;
    int32_t bag_len;
    int32_t avg;
    int32_t pad;
};

struct Job {
    const DevAdamField *fields;
    int n_fields;
    int per;              // dim / 4
    int64_t batch;
    uint32_t *bitmap;     // over the global rows; NULL in lazy mode
    uint32_t *live;       // persistent across steps (caller-owned, rf_adam_params.d_live_rows): rows that EVER received a
                          // gradient; a row whose bit is clear has m == v == 0 and is skipped by the decay pass unread.  NULL: none
};

// last field whose first global row (BY_ROW) / first global key position is <= x
template <bool BY_ROW>
__device__ __forceinline__ int field_of(const Job &j, uint32_t x) {
    int lo = 0, hi = j.n_fields - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        const uint32_t first = BY_ROW ? j.fields[mid].rbase : (uint32_t)j.fields[mid].kbase;
        if (first <= x) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// gradient that (field-local) key `k` receives for column group c: grad[bag(k)] (x 1 / count for avg)
__device__ __forceinline__ float4 key_grad(const DevAdamField &f, int64_t batch, int k, int c) {
    int64_t b;
    float scale = 1.0f;
    if (f.boffs) {
        int64_t lo = 0, hi = batch;
        while (lo + 1 < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (f.boffs[mid] <= k) lo = mid; else hi = mid;
        }
        b = lo;
        if (f.avg) scale = __fdiv_rn(1.0f, (float)(f.boffs[b + 1] - f.boffs[b]));
    } else {
        b = k / f.bag_len;
        if (f.avg) scale = __fdiv_rn(1.0f, (float)f.bag_len);
    }
    float4 g = __ldg(reinterpret_cast<const float4 *>(f.grad + b * f.grad_stride) + c);
    if (f.avg) {
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

__global__ void __launch_bounds__(kAdamThreads) adam_prep_kernel(Job j, int n, uint32_t *__restrict__ keys, int32_t *__restrict__ pos,
                                                                 int32_t *__restrict__ counters) {
    const int i = blockIdx.x * kAdamThreads + threadIdx.x;
    if (i == 0) counters[0] = counters[1] = 0;
    if (i < n) {
        const DevAdamField &f = j.fields[field_of<false>(j, (uint32_t)i)];
        keys[i] = f.rbase + (uint32_t)f.ids[i - f.kbase];
        pos[i] = i;
    }
}

struct HeadPred {
    const uint32_t *keys;
    __device__ __forceinline__ bool operator()(int i) const { return i == 0 || keys[i] != keys[i - 1]; }
};

__device__ __forceinline__ void apply_row(const Job &j, const DevAdamField &f, uint32_t grow, int c, const float4 &g, const AdamC &k) {
    const int64_t at = (int64_t)(grow - f.rbase) * j.per + c;
    float4 w = reinterpret_cast<float4 *>(f.w)[at], m = reinterpret_cast<float4 *>(f.m)[at], v = reinterpret_cast<float4 *>(f.v)[at];
    adam_vec(w, m, v, g, true, k);
    reinterpret_cast<float4 *>(f.w)[at] = w;
    reinterpret_cast<float4 *>(f.m)[at] = m;
    reinterpret_cast<float4 *>(f.v)[at] = v;
    if (c == 0 && j.bitmap) atomicOr(j.bitmap + (grow >> 5), 1u << (grow & 31));
    if (c == 0 && j.live && !((j.live[grow >> 5] >> (grow & 31)) & 1u)) atomicOr(j.live + (grow >> 5), 1u << (grow & 31));
}

// one lane group per unique row; *n_unique is read from device memory (written by the select)
__global__ void __launch_bounds__(kAdamThreads)
adam_rows_kernel(const uint32_t *__restrict__ keys, const int32_t *__restrict__ pos, const int32_t *__restrict__ heads,
                 int32_t *__restrict__ counters, int n_keys, int2 *__restrict__ heavy, Job j, AdamC k_in) {
    const AdamC k = resolved(k_in);
    const int groups = kAdamThreads / j.per;
    const int grp = threadIdx.x / j.per, c = threadIdx.x - grp * j.per;
    if (grp >= groups) return;
    const int n_unique = counters[0];
    for (int u = blockIdx.x * groups + grp; u < n_unique; u += gridDim.x * groups) {
        const int begin = heads[u], end = u + 1 < n_unique ? heads[u + 1] : n_keys;
        if (end - begin > kHeavyRun) {
            if (c == 0) heavy[atomicAdd(counters + 1, 1)] = make_int2(begin, end);
            continue;
        }
        const uint32_t grow = keys[begin];
        const DevAdamField &f = j.fields[field_of<true>(j, grow)];
        float4 g = key_grad(f, j.batch, pos[begin] - f.kbase, c);
        for (int i = begin + 1; i < end; ++i) add4(g, key_grad(f, j.batch, pos[i] - f.kbase, c));
        apply_row(j, f, grow, c, g, k);
    }
}

// one CTA per long run: slot s sums keys begin + s, begin + s + slots, ...; slots are combined in order
__global__ void __launch_bounds__(kAdamThreads)
adam_heavy_kernel(const uint32_t *__restrict__ keys, const int32_t *__restrict__ pos, const int32_t *__restrict__ counters,
                  const int2 *__restrict__ heavy, Job j, AdamC k_in) {
    const AdamC k = resolved(k_in);
    __shared__ float4 part[kAdamThreads];
    const int slots = kAdamThreads / j.per;
    const int s = threadIdx.x / j.per, c = threadIdx.x - s * j.per;
    const int n_heavy = counters[1];
    for (int h = blockIdx.x; h < n_heavy; h += gridDim.x) {
        const int2 run = heavy[h];
        const uint32_t grow = keys[run.x];
        const DevAdamField &f = j.fields[field_of<true>(j, grow)];
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s < slots)
            for (int i = run.x + s; i < run.y; i += slots) add4(g, key_grad(f, j.batch, pos[i] - f.kbase, c));
        part[threadIdx.x] = g;
        __syncthreads();
        if (s == 0) {
            for (int q = 1; q < slots; ++q) add4(g, part[q * j.per + c]);
            apply_row(j, f, grow, c, g, k);
        }
        __syncthreads();
    }
}

// every row whose bit is clear: decay the moments and move the row (Keras' non-lazy sparse Adam); blockIdx.y = table.
// A warp takes 32 consecutive rows: the touched / live bits of those rows are two (broadcast) word loads and a funnel shift,
// and only the lanes whose row is due look at m and v (the per-element version paid a divide and two dependent loads for
// each of the 91 M float4 of C3's tables, 0.3 ms with 8 % of the rows live).  Both bitmaps carry one spare word at the end.
__global__ void __launch_bounds__(kAdamThreads) adam_dense_kernel(Job j, AdamC k_in) {
    const AdamC k = resolved(k_in);
    const DevAdamField &f = j.fields[blockIdx.y];
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const int lane = threadIdx.x & 31;
    const int64_t n_blocks = (f.rows + 31) >> 5;
    const int64_t warp0 = ((int64_t)blockIdx.x * kAdamThreads + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * kAdamThreads) >> 5;
    const int per = j.per, span = 32 * per;
    for (int64_t blk = warp0; blk < n_blocks; blk += n_warps) {
        const int64_t row0 = blk << 5;
        const uint32_t g = f.rbase + (uint32_t)row0, wi = g >> 5, sh = g & 31;
        uint32_t due = ~__funnelshift_r(__ldg(j.bitmap + wi), __ldg(j.bitmap + wi + 1), sh);         // not updated this step
        if (j.live) due &= __funnelshift_r(j.live[wi], j.live[wi + 1], sh);                          // and ever touched
        if (f.rows - row0 < 32) due &= (1u << (int)(f.rows - row0)) - 1u;
        if (due == 0) continue;
        for (int idx = lane; idx < span; idx += 32) {
            const int r = idx / per;
            if (!((due >> r) & 1u)) continue;
            const int64_t e = row0 * per + idx;
            float4 m = reinterpret_cast<float4 *>(f.m)[e], v = reinterpret_cast<float4 *>(f.v)[e];
            // moments that are (still, or again after underflow) zero leave the row where it is: w -= 0, m = 0, v = 0 bit for bit
            if (m.x == 0.f && m.y == 0.f && m.z == 0.f && m.w == 0.f && v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f) continue;
            float4 w = reinterpret_cast<float4 *>(f.w)[e];
            adam_vec(w, m, v, zero, false, k);
            reinterpret_cast<float4 *>(f.w)[e] = w;
            reinterpret_cast<float4 *>(f.m)[e] = m;
            reinterpret_cast<float4 *>(f.v)[e] = v;
        }
    }
}

inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

struct Workspace {
    DevAdamField *fields;
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

int carve(Workspace &ws, char *base, int n_fields, int64_t n_keys, int64_t rows) {
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
    ws.fields = reinterpret_cast<DevAdamField *>(take(sizeof(DevAdamField) * (size_t)n_fields));
    ws.keys_in = reinterpret_cast<uint32_t *>(take(sizeof(uint32_t) * n));
    ws.keys_out = reinterpret_cast<uint32_t *>(take(sizeof(uint32_t) * n));
    ws.pos_in = reinterpret_cast<int32_t *>(take(sizeof(int32_t) * n));
    ws.pos_out = reinterpret_cast<int32_t *>(take(sizeof(int32_t) * n));
    ws.heads = reinterpret_cast<int32_t *>(take(sizeof(int32_t) * ((size_t)n + 1)));
    ws.counters = reinterpret_cast<int32_t *>(take(sizeof(int32_t) * 4));
    ws.heavy = reinterpret_cast<int2 *>(take(sizeof(int2) * ((size_t)n / kHeavyRun + 1)));
    ws.bitmap_bytes = sizeof(uint32_t) * (size_t)((rows + 31) / 32 + 1);       // + 1: the funnel shift reads word + 1
    ws.bitmap = reinterpret_cast<uint32_t *>(take(ws.bitmap_bytes));
    ws.cub_bytes = sort_bytes > select_bytes ? sort_bytes : select_bytes;
    ws.cub_temp = take(ws.cub_bytes);
    ws.total = off;
    return RF_OK;
}

// totals over the fields + argument checks shared by the size query and the launch
int totals(const rf_adam_field *fields, int n_fields, int64_t &n_keys, int64_t &rows) {
    if (n_fields <= 0 || !fields) return set_error(RF_ERR_INVALID, "rf_bag_backward_adam: no tables");
    if (n_fields > 65535) return set_error(RF_ERR_INVALID, "rf_bag_backward_adam: at most 65535 tables per call");
    n_keys = rows = 0;
    for (int i = 0; i < n_fields; ++i) {
        const rf_adam_field &f = fields[i];
        if (f.n_keys < 0 || f.table_rows <= 0) return set_error(RF_ERR_INVALID, "table %d: bad n_keys / table_rows", i);
        if (f.dim != fields[0].dim) return set_error(RF_ERR_INVALID, "table %d: all tables of one call share dim (%d != %d)", i, f.dim, fields[0].dim);
        n_keys += f.n_keys;
        rows += f.table_rows;
    }
    if (n_keys > INT32_MAX) return set_error(RF_ERR_INVALID, "rf_bag_backward_adam: %lld keys exceed 2^31 - 1", (long long)n_keys);
    if (rows > (int64_t)UINT32_MAX) return set_error(RF_ERR_INVALID, "rf_bag_backward_adam: %lld rows in total exceed 2^32 - 1", (long long)rows);
    return RF_OK;
}

}  // namespace
}  // namespace rf

using namespace rf;

extern "C" {

int64_t rf_bag_adam_multi_workspace_bytes(const rf_adam_field *fields, int n_fields) {
    int64_t n_keys = 0, rows = 0;
    if (totals(fields, n_fields, n_keys, rows) != RF_OK) return -1;
    Workspace ws;
    if (carve(ws, nullptr, n_fields, n_keys, rows) != RF_OK) return -1;
    return (int64_t)ws.total;
}

int rf_bag_backward_adam_multi(const rf_adam_field *fields, int n_fields, int64_t batch, const rf_adam_params *params,
                               void *d_workspace, int64_t workspace_bytes, void *stream) {
    if (!params) return set_error(RF_ERR_INVALID, "rf_bag_backward_adam: params is NULL");
    int64_t n_keys = 0, rows = 0;
    int rc = totals(fields, n_fields, n_keys, rows);
    if (rc != RF_OK) return rc;
    const int dim = fields[0].dim;
    if (batch < 0 || dim <= 0) return set_error(RF_ERR_INVALID, "bad backward shape");
    if (params->step < 1) return set_error(RF_ERR_INVALID, "Adam step must be >= 1, got %lld", (long long)params->step);
    if (dim % 4 != 0 || dim > 4 * kAdamThreads)
        return set_error(RF_ERR_UNSUPPORTED, "rf_bag_backward_adam needs dim %% 4 == 0 and dim <= %d", 4 * kAdamThreads);
    std::vector<DevAdamField> dev(n_fields);
    int64_t kbase = 0, rbase = 0;
    for (int i = 0; i < n_fields; ++i) {
        const rf_adam_field &f = fields[i];
        if (f.combiner != RF_COMBINER_SUM && f.combiner != RF_COMBINER_AVG)
            return set_error(RF_ERR_UNSUPPORTED, "table %d: backward is implemented for the sum and avg combiners", i);
        if (!f.table || !f.m || !f.v) return set_error(RF_ERR_INVALID, "table %d: NULL table / moment buffer", i);
        if (f.n_keys > 0 && (!f.ids || !f.grad_out)) return set_error(RF_ERR_INVALID, "table %d: NULL ids / grad_out", i);
        if (f.n_keys > 0 && !f.bag_offsets && (f.bag_len <= 0 || batch * (int64_t)f.bag_len != f.n_keys))
            return set_error(RF_ERR_INVALID, "table %d: dense backward needs batch x bag_len == n_keys", i);
        if (f.grad_stride % 4 != 0) return set_error(RF_ERR_UNSUPPORTED, "table %d: grad_stride %% 4 != 0", i);
        for (const void *p : {(const void *)f.table, (const void *)f.m, (const void *)f.v, (const void *)f.grad_out})
            if (reinterpret_cast<uintptr_t>(p) % 16 != 0) return set_error(RF_ERR_INVALID, "table %d: buffers must be 16-byte aligned", i);
        DevAdamField &d = dev[i];
        d.ids = f.ids;
        d.boffs = f.bag_offsets;
        d.grad = f.grad_out;
        d.grad_stride = f.grad_stride;
        d.w = f.table;
        d.m = f.m;
        d.v = f.v;
        d.rows = f.table_rows;
        d.rbase = (uint32_t)rbase;
        d.kbase = (int32_t)kbase;
        d.n_keys = (int32_t)f.n_keys;
        d.bag_len = f.bag_len;
        d.avg = f.combiner == RF_COMBINER_AVG;
        d.pad = 0;
        kbase += f.n_keys;
        rbase += f.table_rows;
    }
    Workspace ws;
    rc = carve(ws, static_cast<char *>(d_workspace), n_fields, n_keys, rows);
    if (rc != RF_OK) return rc;
    if (!d_workspace || workspace_bytes < (int64_t)ws.total)
        return set_error(RF_ERR_INVALID, "workspace too small: %lld < %lld bytes (rf_bag_adam[_multi]_workspace_bytes)",
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
    k.d_lr_t = params->d_lr_t;

    int devid = 0, sms = 0;
    RF_CUDA(cudaGetDevice(&devid));
    RF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, devid));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool lazy = params->lazy != 0;
    cudaStreamCaptureStatus capture = cudaStreamCaptureStatusNone;
    RF_CUDA(cudaStreamIsCapturing(st, &capture));
    if (capture != cudaStreamCaptureStatusNone) {
        // recorded into a CUDA graph: the upload is replayed from pinned bytes that stay put (rf_bag.cu's descriptor slots),
        // and the step-dependent learning rate has to come from device memory
        if (!params->d_lr_t) return set_error(RF_ERR_INVALID, "a captured optimizer step needs rf_adam_params.d_lr_t (the host-side step would be frozen into the graph)");
        char *h = nullptr, *d = nullptr;
        rc = graph_desc_slots(devid, sizeof(DevAdamField) * (size_t)n_fields, &h, &d);
        if (rc != RF_OK) return rc;
        memcpy(h, dev.data(), sizeof(DevAdamField) * (size_t)n_fields);
        rc = graph_desc_upload(reinterpret_cast<char *>(ws.fields), h, sizeof(DevAdamField) * (size_t)n_fields, st);
        if (rc != RF_OK) return rc;
    } else {
        rc = graph_desc_pool_ready(devid);
        if (rc != RF_OK) return rc;
        // the descriptors travel through a pageable staging copy: the driver snapshots `dev` before returning
        RF_CUDA(cudaMemcpyAsync(ws.fields, dev.data(), sizeof(DevAdamField) * (size_t)n_fields, cudaMemcpyHostToDevice, st));
    }
    Job job{ws.fields, n_fields, dim / 4, batch, lazy ? nullptr : ws.bitmap, lazy ? nullptr : params->d_live_rows};
    if (!lazy) RF_CUDA(cudaMemsetAsync(ws.bitmap, 0, ws.bitmap_bytes, st));
    int launches = 0;
    if (n_keys > 0) {
        const int n = (int)n_keys;
        adam_prep_kernel<<<(n + kAdamThreads - 1) / kAdamThreads, kAdamThreads, 0, st>>>(job, n, ws.keys_in, ws.pos_in, ws.counters);
        RF_CUDA(cudaGetLastError());
        size_t tmp = ws.cub_bytes;
        RF_CUDA(cub::DeviceRadixSort::SortPairs(ws.cub_temp, tmp, ws.keys_in, ws.keys_out, ws.pos_in, ws.pos_out, n, 0, key_bits(rows), st));
        thrust::counting_iterator<int> iota(0);
        tmp = ws.cub_bytes;
        RF_CUDA(cub::DeviceSelect::If(ws.cub_temp, tmp, iota, ws.heads, ws.counters, n, HeadPred{ws.keys_out}, st));
        const int groups = kAdamThreads / job.per;
        int64_t blocks = ((int64_t)n + groups - 1) / groups;
        if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
        adam_rows_kernel<<<(unsigned)blocks, kAdamThreads, 0, st>>>(ws.keys_out, ws.pos_out, ws.heads, ws.counters, n, ws.heavy, job, k);
        RF_CUDA(cudaGetLastError());
        int64_t hblocks = (int64_t)n / kHeavyRun + 1;
        if (hblocks > (int64_t)sms * 4) hblocks = (int64_t)sms * 4;
        adam_heavy_kernel<<<(unsigned)hblocks, kAdamThreads, 0, st>>>(ws.keys_out, ws.pos_out, ws.counters, ws.heavy, job, k);
        RF_CUDA(cudaGetLastError());
        launches += 3;
    }
    if (!lazy) {
        int64_t max_rows = 0;
        for (int i = 0; i < n_fields; ++i) max_rows = fields[i].table_rows > max_rows ? fields[i].table_rows : max_rows;
        int64_t blocks = ((max_rows + 31) / 32 + kAdamThreads / 32 - 1) / (kAdamThreads / 32);      // one warp per 32 rows
        const int64_t cap = ((int64_t)sms * 16 + n_fields - 1) / n_fields;
        if (blocks > cap) blocks = cap;
        if (blocks < 1) blocks = 1;
        adam_dense_kernel<<<dim3((unsigned)blocks, (unsigned)n_fields), kAdamThreads, 0, st>>>(job, k);
        RF_CUDA(cudaGetLastError());
        ++launches;
    }
    g_launches.fetch_add(launches);
    return RF_OK;
}

int64_t rf_bag_adam_workspace_bytes(int64_t n_keys, int64_t table_rows) {
    rf_adam_field f;
    memset(&f, 0, sizeof f);
    f.n_keys = n_keys;
    f.table_rows = table_rows;
    f.dim = 4;
    return rf_bag_adam_multi_workspace_bytes(&f, 1);
}

int rf_bag_backward_adam(const int64_t *d_ids, int64_t n_keys, const int32_t *d_bag_offsets, int32_t bag_len, int64_t batch,
                         const float *d_grad_out, int64_t grad_stride, int32_t dim, int combiner, const rf_adam_params *params,
                         float *d_table, float *d_m, float *d_v, int64_t table_rows, void *d_workspace, int64_t workspace_bytes,
                         void *stream) {
    rf_adam_field f;
    memset(&f, 0, sizeof f);
    f.ids = d_ids;
    f.bag_offsets = d_bag_offsets;
    f.n_keys = n_keys;
    f.bag_len = bag_len;
    f.combiner = combiner;
    f.grad_out = d_grad_out;
    f.grad_stride = grad_stride;
    f.table = d_table;
    f.m = d_m;
    f.v = d_v;
    f.table_rows = table_rows;
    f.dim = dim;
    return rf_bag_backward_adam_multi(&f, 1, batch, params, d_workspace, workspace_bytes, stream);
}

}  // extern "C"
