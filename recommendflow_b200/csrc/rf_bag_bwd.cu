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

// --------------------------------------------------------------------------------------------
// min / max pooling backward (EmbeddingBag combiner "min" / "max": tf.reduce_min / tf.reduce_max over the bag,
// backend/layers/preprocess_layers.py:43-68).  TensorFlow's gradient (_MinOrMaxGrad) sends the pooled element's gradient to
// the keys whose row element EQUALS the pooled value, split equally among ties:
//     dX[k][d] = g[bag][d] * (W[id_k][d] == y[bag][d]) / #{k' in bag : W[id_k'][d] == y[bag][d]}
// Unlike sum / avg this differs per key, so it is materialised as one gradient row per key; the row update is then the
// ordinary sum-pooling backward with bags of ONE key (rf_bag_backward / rf_bag_backward_adam with bag_len = 1, batch = n_keys)
// -- which also keeps the optimizer's in-place row updates from racing with the equality tests (they all happen here, first).
// One thread per (bag, 4 columns): two passes over the bag's rows (count the ties, write the shares).
// --------------------------------------------------------------------------------------------
namespace rf {

__global__ void __launch_bounds__(256) bag_minmax_key_grads_kernel(const int64_t *__restrict__ ids, const int32_t *__restrict__ boffs,
                                                                   int bag_len, int64_t batch, const float *__restrict__ table, int dim,
                                                                   const float *__restrict__ pooled, int64_t pooled_stride,
                                                                   const float *__restrict__ grad, int64_t grad_stride,
                                                                   float *__restrict__ key_grads) {
    const int per = dim >> 2;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < batch * per; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = e / per;
        const int c = (int)(e - b * per) * 4;
        const int64_t k0 = boffs ? boffs[b] : b * bag_len, k1 = boffs ? boffs[b + 1] : k0 + bag_len;
        const float4 y = *reinterpret_cast<const float4 *>(pooled + b * pooled_stride + c);
        const float4 g = *reinterpret_cast<const float4 *>(grad + b * grad_stride + c);
        float4 ties = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int64_t k = k0; k < k1; ++k) {
            const float4 w = *reinterpret_cast<const float4 *>(table + ids[k] * dim + c);
            ties.x += w.x == y.x ? 1.f : 0.f;
            ties.y += w.y == y.y ? 1.f : 0.f;
            ties.z += w.z == y.z ? 1.f : 0.f;
            ties.w += w.w == y.w ? 1.f : 0.f;
        }
        const float4 share = make_float4(ties.x > 0.f ? g.x / ties.x : 0.f, ties.y > 0.f ? g.y / ties.y : 0.f,
                                         ties.z > 0.f ? g.z / ties.z : 0.f, ties.w > 0.f ? g.w / ties.w : 0.f);
        for (int64_t k = k0; k < k1; ++k) {
            const float4 w = *reinterpret_cast<const float4 *>(table + ids[k] * dim + c);
            *reinterpret_cast<float4 *>(key_grads + k * dim + c) = make_float4(w.x == y.x ? share.x : 0.f, w.y == y.y ? share.y : 0.f,
                                                                              w.z == y.z ? share.z : 0.f, w.w == y.w ? share.w : 0.f);
        }
    }
}

}  // namespace rf

extern "C" int rf_bag_minmax_key_grads(const int64_t *d_ids, int64_t n_keys, const int32_t *d_bag_offsets, int32_t bag_len, int64_t batch,
                                       const float *d_table, int32_t dim, const float *d_pooled, int64_t pooled_stride,
                                       const float *d_grad_out, int64_t grad_stride, float *d_key_grads, void *stream) {
    using namespace rf;
    if (n_keys < 0 || batch < 0 || dim <= 0) return set_error(RF_ERR_INVALID, "rf_bag_minmax_key_grads: bad shape");
    if (n_keys == 0 || batch == 0) return RF_OK;
    if (!d_ids || !d_table || !d_pooled || !d_grad_out || !d_key_grads) return set_error(RF_ERR_INVALID, "rf_bag_minmax_key_grads: NULL buffer");
    if (!d_bag_offsets && (bag_len <= 0 || batch * (int64_t)bag_len != n_keys))
        return set_error(RF_ERR_INVALID, "rf_bag_minmax_key_grads: dense bags need batch * bag_len == n_keys");
    const uintptr_t al = reinterpret_cast<uintptr_t>(d_table) | reinterpret_cast<uintptr_t>(d_pooled) | reinterpret_cast<uintptr_t>(d_grad_out) |
                         reinterpret_cast<uintptr_t>(d_key_grads);
    if (dim % 4 || (al & 15) || pooled_stride % 4 || grad_stride % 4)
        return set_error(RF_ERR_UNSUPPORTED, "rf_bag_minmax_key_grads needs dim and the row pitches to be multiples of 4 floats and 16-byte aligned buffers");
    int64_t blocks = (batch * (dim / 4) + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    bag_minmax_key_grads_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        d_ids, d_bag_offsets, bag_len, batch, d_table, dim, d_pooled, pooled_stride, d_grad_out, grad_stride, d_key_grads);
    RF_CUDA(cudaGetLastError());
    g_launches.fetch_add(1);
    return RF_OK;
}
