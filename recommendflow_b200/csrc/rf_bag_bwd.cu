// rf_bag_bwd.cu -- backward of the pooled embedding bag, fused with the SGD row update.
//
// "Next" row: the reference trains through Keras autodiff of
// Embedding + reduce_sum / reduce_mean (backend/layers/preprocess_layers.py:43-68): the gradient of
// a pooled vector flows back to every row that was gathered into it (pads included -- row 0 collects
// the pads' gradient, exactly as the reference's unmasked pooling implies).
//
//   W[id[k]] += alpha * g[bag(k)] * (avg ? 1 / count(bag) : 1)        for every key k
// alpha = -lr gives the fused SGD step; alpha = 1 with W = a zeroed buffer accumulates the dense
// gradient of the touched rows.  One lane group (D / 4 lanes, 128-bit) per key, 128-bit vector
// reductions into HBM/L2 (red.global.add.v4.f32): HBM-bound, duplicates are resolved by the
// atomics (order-dependent fp32 rounding: results are reproducible only up to re-association).
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/rf_b200.h"
#include "rf_common.h"

namespace rf {

extern std::atomic<int64_t> g_launches;

__device__ __forceinline__ void red_add_v4(float *addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <bool VEC>
__global__ void __launch_bounds__(256) bag_backward_kernel(const int64_t *__restrict__ ids, const int32_t *__restrict__ boffs,
                                                           int bag_len, int64_t batch, int64_t n_keys,
                                                           const float *__restrict__ grad, int64_t grad_stride, int dim,
                                                           int avg, float alpha, float *__restrict__ table) {
    const int per = VEC ? dim >> 2 : dim;                 // lanes of work per key
    const int64_t total = n_keys * per;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = e / per;
        const int c = (int)(e - k * per);
        int64_t b;
        float scale = alpha;
        if (boffs) {                                      // jagged: binary search the bag of key k
            int64_t lo = 0, hi = batch;
            while (lo + 1 < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (boffs[mid] <= k) lo = mid; else hi = mid;
            }
            b = lo;
            if (avg) scale = alpha / (float)(boffs[b + 1] - boffs[b]);
        } else {
            b = k / bag_len;
            if (avg) scale = alpha / (float)bag_len;
        }
        const int64_t row = ids[k];
        if (VEC) {
            float4 g = __ldg(reinterpret_cast<const float4 *>(grad + b * grad_stride) + c);
            g.x *= scale; g.y *= scale; g.z *= scale; g.w *= scale;
            red_add_v4(table + row * dim + c * 4, g);
        } else {
            atomicAdd(table + row * dim + c, __ldg(grad + b * grad_stride + c) * scale);
        }
    }
}

}  // namespace rf

using namespace rf;

extern "C" int rf_bag_backward(const int64_t *d_ids, int64_t n_keys, const int32_t *d_bag_offsets, int32_t bag_len, int64_t batch,
                               const float *d_grad_out, int64_t grad_stride, int32_t dim, int combiner, float alpha,
                               float *d_table, void *stream) {
    if (n_keys < 0 || batch < 0 || dim <= 0) return set_error(RF_ERR_INVALID, "bad backward shape");
    if (combiner != RF_COMBINER_SUM && combiner != RF_COMBINER_AVG)
        return set_error(RF_ERR_UNSUPPORTED, "backward is implemented for the sum and avg combiners");
    if (n_keys == 0 || batch == 0) return RF_OK;
    if (!d_ids || !d_grad_out || !d_table) return set_error(RF_ERR_INVALID, "rf_bag_backward: NULL buffer");
    if (!d_bag_offsets && (bag_len <= 0 || batch * (int64_t)bag_len != n_keys))
        return set_error(RF_ERR_INVALID, "dense backward: batch x bag_len != n_keys");
    int dev = 0, sms = 0;
    RF_CUDA(cudaGetDevice(&dev));
    RF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const bool vec = dim % 4 == 0 && grad_stride % 4 == 0 && reinterpret_cast<uintptr_t>(d_grad_out) % 16 == 0 &&
                     reinterpret_cast<uintptr_t>(d_table) % 16 == 0;
    const int64_t work = n_keys * (vec ? dim / 4 : dim);
    int64_t blocks = (work + 255) / 256;
    if (blocks > (int64_t)sms * 32) blocks = (int64_t)sms * 32;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (vec)
        bag_backward_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(d_ids, d_bag_offsets, bag_len, batch, n_keys, d_grad_out,
                                                                    grad_stride, dim, combiner == RF_COMBINER_AVG, alpha, d_table);
    else
        bag_backward_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(d_ids, d_bag_offsets, bag_len, batch, n_keys, d_grad_out,
                                                                     grad_stride, dim, combiner == RF_COMBINER_AVG, alpha, d_table);
    RF_CUDA(cudaGetLastError());
    g_launches.fetch_add(1);
    return RF_OK;
}
