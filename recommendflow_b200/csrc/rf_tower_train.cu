// Training-time passes of one tower stage  y = act(BatchNormalization_batch(x) W + b)
// (reference: backend/blocks/mlp.py:4-15 builds [norm, Dense, Dropout] * n; Keras runs the normalisation on batch
// statistics under model.fit).  The three GEMMs of a stage (forward, dX, dW) are rf_dense_forward_tc; what is left is
// HBM-bound column work over [rows, dim] matrices, done here in as few passes as the data dependencies allow:
//
//   rf_column_stats          one read of x:  batch mean / biased variance per column  (+ x^T for the dW GEMM)
//   rf_activation_backward   one read of dY, y:  dZ = dY * act'(y), dZ^T, db = colsum(dZ)
//   rf_batchnorm_backward    one read of dXhat, x:  dbeta = colsum(dXhat), dgamma = colsum(dXhat * xn);
//                            second read:  dX = s * (dXhat - dbeta / B - xn * dgamma / B)
//
// Layout: a CTA owns 32 columns x kSplitRows rows; warp w of 8 reads rows w, w + 8, ... of each 32 x 32 sub-tile (one
// 128-byte line per warp instruction, four in flight per thread); column sums are reduced warp -> CTA in a fixed order and
// written as partials [split][value][column]; a finalize kernel adds the splits in order: results are deterministic.
// Transposes go through a padded 32 x 33 shared tile so both the read and the write are coalesced.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>

#include "../../include/rf_b200.h"
#include "rf_common.h"

namespace rf {

extern std::atomic<int64_t> g_launches;

namespace {

constexpr int kThreads = 256;
constexpr int kSplitRows = 256;      // rows per CTA: many CTAs keep the loads in flight; the finalize spreads the splits over 8 warps
constexpr float kSeluScale = 1.0507009873554805f, kSeluAlpha = 1.6732632423543772f;

__device__ __forceinline__ float act_grad(int act, float y) {
    switch (act) {
        case RF_ACT_RELU: return y > 0.f ? 1.f : 0.f;
        case RF_ACT_SELU: return y > 0.f ? kSeluScale : y + kSeluScale * kSeluAlpha;      // z <= 0: scale * alpha * e^z = y + scale * alpha
        case RF_ACT_TANH: return 1.f - y * y;
        case RF_ACT_SIGMOID: return y * (1.f - y);
        default: return 1.f;
    }
}

// reduce `acc` (one value per thread: lane = column, warp = row class) over the 8 warps, fixed order; warp 0 returns the sum
__device__ __forceinline__ float cta_column_sum(float acc, float (*red)[32]) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    red[w][lane] = acc;
    __syncthreads();
    float s = 0.f;
    if (w == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) s += red[k][lane];
    }
    return s;
}

// MODE 0: column statistics of x (sum and sum of squares around shift = x[0][c]) + optional transpose
// MODE 1: dz = dy * act'(y) (+ transpose), colsum(dz)
// MODE 2: colsum(dxh), colsum(dxh * xn)
struct PassArgs {
    const float *a;        // x | dy | dxh
    const float *b;        // - | y  | x
    int64_t lda, ldb;
    float *out;            // - | dz | -
    float *out_t;          // x^T | dz^T | -      [dim, rows], may be NULL
    const float *mean, *rstd;
    float *partials;       // [splits][2][dim]
    int64_t rows;
    int dim;
    int act;
    int split_rows;
};

// Vector path (dim, both row pitches multiples of 4, 16-byte aligned buffers): a CTA owns 128 columns x split_rows rows, a lane
// owns 4 columns (128-bit loads / stores), warp w takes rows w, w + 8, ...; the transposed output leaves as 128-bit stores of
// four consecutive rows.  (The scalar kernel below moved 32-column tiles one float at a time: 3.2 TB/s on the 630 MB of the
// q|k|v projection's backward pass.)
template <int MODE>
__global__ void __launch_bounds__(kThreads) column_pass_vec_kernel(PassArgs p) {
    __shared__ float tile[32][129];
    __shared__ float red[8][128];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 128, c = c0 + lane * 4;
    const bool col_ok = c < p.dim;                      // dim % 4 == 0: a lane's four columns are all in or all out
    const int64_t r_begin = (int64_t)blockIdx.y * p.split_rows;
    const int64_t r_end = r_begin + p.split_rows < p.rows ? r_begin + p.split_rows : p.rows;
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1, shift = s1, mean = s1, rstd = s1;
    if (MODE == 0 && col_ok) shift = *reinterpret_cast<const float4 *>(p.a + c);
    if (MODE == 2 && col_ok) {
        mean = *reinterpret_cast<const float4 *>(p.mean + c);
        rstd = *reinterpret_cast<const float4 *>(p.rstd + c);
    }
    for (int64_t r0 = r_begin; r0 < r_end; r0 += 32) {
        float4 va[4], vb[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t r = r0 + w + 8 * k;
            const bool ok = col_ok && r < r_end;
            va[k] = ok ? *reinterpret_cast<const float4 *>(p.a + r * p.lda + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            vb[k] = (MODE != 0 && ok) ? *reinterpret_cast<const float4 *>(p.b + r * p.ldb + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t r = r0 + w + 8 * k;
            const bool ok = col_ok && r < r_end;
            float4 t = va[k];
            if (MODE == 0) {
                if (ok) {
                    const float4 d = make_float4(va[k].x - shift.x, va[k].y - shift.y, va[k].z - shift.z, va[k].w - shift.w);
                    s1.x += d.x; s1.y += d.y; s1.z += d.z; s1.w += d.w;
                    s2.x += d.x * d.x; s2.y += d.y * d.y; s2.z += d.z * d.z; s2.w += d.w * d.w;
                }
            } else if (MODE == 1) {
                t = make_float4(va[k].x * act_grad(p.act, vb[k].x), va[k].y * act_grad(p.act, vb[k].y), va[k].z * act_grad(p.act, vb[k].z),
                                va[k].w * act_grad(p.act, vb[k].w));
                if (ok) {
                    s1.x += t.x; s1.y += t.y; s1.z += t.z; s1.w += t.w;
                    if (p.out) *reinterpret_cast<float4 *>(p.out + r * p.dim + c) = t;
                }
            } else {
                if (ok) {
                    s1.x += va[k].x; s1.y += va[k].y; s1.z += va[k].z; s1.w += va[k].w;
                    s2.x += va[k].x * ((vb[k].x - mean.x) * rstd.x);
                    s2.y += va[k].y * ((vb[k].y - mean.y) * rstd.y);
                    s2.z += va[k].z * ((vb[k].z - mean.z) * rstd.z);
                    s2.w += va[k].w * ((vb[k].w - mean.w) * rstd.w);
                }
            }
            if (MODE != 2 && p.out_t) {
                float *row = tile[w + 8 * k] + lane * 4;
                row[0] = ok ? t.x : 0.f;
                row[1] = ok ? t.y : 0.f;
                row[2] = ok ? t.z : 0.f;
                row[3] = ok ? t.w : 0.f;
            }
        }
        if (MODE != 2 && p.out_t) {
            __syncthreads();
            // 128 columns x 8 row quads: thread -> (column cc, quad rq); rows r0 + 4 rq .. + 3 of column c0 + cc
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int idx = threadIdx.x + kThreads * k;            // 0 .. 1023
                const int rq = idx & 7, cc = idx >> 3;
                const int64_t r = r0 + rq * 4;
                if (c0 + cc < p.dim && r < r_end) {
                    float *dst = p.out_t + (int64_t)(c0 + cc) * p.rows + r;
                    if (r + 3 < r_end)
                        *reinterpret_cast<float4 *>(dst) = make_float4(tile[rq * 4][cc], tile[rq * 4 + 1][cc], tile[rq * 4 + 2][cc], tile[rq * 4 + 3][cc]);
                    else
                        for (int e = 0; e < 4 && r + e < r_end; ++e) dst[e] = tile[rq * 4 + e][cc];
                }
            }
            __syncthreads();
        }
    }
    // column sums: 8 warps -> one, fixed order; lane owns columns c .. c + 3
    auto reduce4 = [&](const float4 &v, int which) {
        __syncthreads();
        red[w][lane * 4 + 0] = v.x;
        red[w][lane * 4 + 1] = v.y;
        red[w][lane * 4 + 2] = v.z;
        red[w][lane * 4 + 3] = v.w;
        __syncthreads();
        if (threadIdx.x < 128 && c0 + (int)threadIdx.x < p.dim) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
            p.partials[((int64_t)blockIdx.y * 2 + which) * p.dim + c0 + threadIdx.x] = s;
        }
    };
    reduce4(s1, 0);
    if (MODE != 1) reduce4(s2, 1);
}

template <int MODE>
__global__ void __launch_bounds__(kThreads) column_pass_kernel(PassArgs p) {
    __shared__ float tile[32][33];
    __shared__ float red[8][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c0 = blockIdx.x * 32, c = c0 + lane;
    const bool col_ok = c < p.dim;
    const int64_t r_begin = (int64_t)blockIdx.y * p.split_rows;
    const int64_t r_end = r_begin + p.split_rows < p.rows ? r_begin + p.split_rows : p.rows;
    float s1 = 0.f, s2 = 0.f;
    float shift = 0.f, mean = 0.f, rstd = 0.f;
    if (MODE == 0 && col_ok) shift = __ldg(p.a + c);
    if (MODE == 2 && col_ok) {
        mean = __ldg(p.mean + c);
        rstd = __ldg(p.rstd + c);
    }
    for (int64_t r0 = r_begin; r0 < r_end; r0 += 32) {
        float va[4], vb[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t r = r0 + w + 8 * k;
            const bool ok = col_ok && r < r_end;
            va[k] = ok ? __ldg(p.a + r * p.lda + c) : 0.f;
            vb[k] = (MODE != 0 && ok) ? __ldg(p.b + r * p.ldb + c) : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t r = r0 + w + 8 * k;
            const bool ok = col_ok && r < r_end;
            float t = va[k];
            if (MODE == 0) {
                if (ok) {
                    const float d = va[k] - shift;
                    s1 += d;
                    s2 += d * d;
                }
            } else if (MODE == 1) {
                t = va[k] * act_grad(p.act, vb[k]);
                if (ok) {
                    s1 += t;
                    if (p.out) p.out[r * p.dim + c] = t;
                }
            } else {
                if (ok) {
                    s1 += va[k];
                    s2 += va[k] * ((vb[k] - mean) * rstd);
                }
            }
            if (MODE != 2 && p.out_t) tile[w + 8 * k][lane] = ok ? t : 0.f;
        }
        if (MODE != 2 && p.out_t) {
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int cc = w + 8 * k;
                const int64_t r = r0 + lane;
                if (c0 + cc < p.dim && r < r_end) p.out_t[(int64_t)(c0 + cc) * p.rows + r] = tile[lane][cc];
            }
            __syncthreads();
        }
    }
    const float t1 = cta_column_sum(s1, red);
    if (w == 0 && col_ok) p.partials[((int64_t)blockIdx.y * 2 + 0) * p.dim + c] = t1;
    if (MODE != 1) {
        const float t2 = cta_column_sum(s2, red);
        if (w == 0 && col_ok) p.partials[((int64_t)blockIdx.y * 2 + 1) * p.dim + c] = t2;
    }
}

// MODE 0: mean / biased variance from the shifted sums;  MODE 1: out0 = sum;  MODE 2: out0 = sum0 (dbeta), out1 = sum1 (dgamma)
template <int MODE>
__global__ void __launch_bounds__(kThreads) column_finalize_kernel(const float *__restrict__ partials, int splits, int dim, int64_t rows,
                                                                   const float *__restrict__ x_row0, float *__restrict__ out0,
                                                                   float *__restrict__ out1) {
    // one CTA per 32 columns: warp w adds splits w, w + 8, ... (in order), the 8 warp sums are added in order
    __shared__ float red[8][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    const bool ok = c < dim;
    float a = 0.f, b = 0.f;
    if (ok)
        for (int s = w; s < splits; s += 8) {
            a += partials[((int64_t)s * 2 + 0) * dim + c];
            if (MODE != 1) b += partials[((int64_t)s * 2 + 1) * dim + c];
        }
    a = cta_column_sum(a, red);
    if (MODE != 1) b = cta_column_sum(b, red);
    if (w != 0 || !ok) return;
    if (MODE == 0) {
        const float inv = 1.f / (float)rows, m = a * inv;
        out0[c] = x_row0[c] + m;
        out1[c] = fmaxf(b * inv - m * m, 0.f);
    } else {
        out0[c] = a;
        if (MODE == 2) out1[c] = b;
    }
}

// dx = s * (dxh - dbeta / B - xn * dgamma / B),  xn = (x - mean) * rstd;  4 columns per thread
__global__ void __launch_bounds__(kThreads) batchnorm_dx_kernel(const float *__restrict__ dxh, const float *__restrict__ x, int64_t ldx,
                                                                const float *__restrict__ mean, const float *__restrict__ rstd,
                                                                const float *__restrict__ s, const float *__restrict__ dbeta,
                                                                const float *__restrict__ dgamma, int64_t rows, int dim,
                                                                float *__restrict__ dx) {
    const int q = dim >> 2;
    const float inv = 1.f / (float)rows;
    for (int64_t e = (int64_t)blockIdx.x * kThreads + threadIdx.x; e < rows * q; e += (int64_t)gridDim.x * kThreads) {
        const int64_t r = e / q;
        const int c = (int)(e - r * q) * 4;
        const float4 g = *reinterpret_cast<const float4 *>(dxh + r * dim + c);
        const float4 v = *reinterpret_cast<const float4 *>(x + r * ldx + c);
        const float4 m = *reinterpret_cast<const float4 *>(mean + c), rs = *reinterpret_cast<const float4 *>(rstd + c);
        const float4 sc = *reinterpret_cast<const float4 *>(s + c), db = *reinterpret_cast<const float4 *>(dbeta + c);
        const float4 dg = *reinterpret_cast<const float4 *>(dgamma + c);
        float4 o;
        o.x = sc.x * (g.x - db.x * inv - (v.x - m.x) * rs.x * (dg.x * inv));
        o.y = sc.y * (g.y - db.y * inv - (v.y - m.y) * rs.y * (dg.y * inv));
        o.z = sc.z * (g.z - db.z * inv - (v.z - m.z) * rs.z * (dg.z * inv));
        o.w = sc.w * (g.w - db.w * inv - (v.w - m.w) * rs.w * (dg.w * inv));
        *reinterpret_cast<float4 *>(dx + r * dim + c) = o;
    }
}

inline int split_rows_of(int64_t) { return kSplitRows; }

// the 128-bit kernel needs 4-float granularity everywhere it touches: columns, both row pitches, the transposed rows, pointers
inline bool vec_ok(int dim, int64_t lda, int64_t ldb, int64_t rows, const void *a, const void *b, const void *out, const void *out_t) {
    const uintptr_t al = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out) |
                         reinterpret_cast<uintptr_t>(out_t);
    return dim % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0 && (out_t == nullptr || rows % 4 == 0) && (al & 15) == 0;
}
inline int splits_of(int64_t rows) {
    const int64_t r = split_rows_of(rows);
    return (int)((rows + r - 1) / r);
}

int check_ws(int64_t rows, int dim, const void *ws, int64_t ws_bytes) {
    const int64_t need = rf_tower_train_workspace_bytes(rows, dim);
    if (!ws || ws_bytes < need) return set_error(RF_ERR_INVALID, "workspace too small: %lld < %lld bytes (rf_tower_train_workspace_bytes)",
                                                 (long long)ws_bytes, (long long)need);
    return RF_OK;
}

}  // namespace
}  // namespace rf

using namespace rf;

extern "C" {

int64_t rf_tower_train_workspace_bytes(int64_t rows, int32_t dim) {
    if (rows <= 0 || dim <= 0) return 0;
    return (int64_t)splits_of(rows) * 2 * dim * (int64_t)sizeof(float);
}

int rf_column_stats(const float *d_x, int64_t rows, int32_t dim, int64_t ldx, float *d_mean, float *d_var, float *d_x_t,
                    void *d_workspace, int64_t workspace_bytes, void *stream) {
    if (rows <= 0 || dim <= 0 || ldx < dim) return set_error(RF_ERR_INVALID, "rf_column_stats: bad shape");
    if (!d_x || !d_mean || !d_var) return set_error(RF_ERR_INVALID, "rf_column_stats: NULL buffer");
    int rc = check_ws(rows, dim, d_workspace, workspace_bytes);
    if (rc != RF_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    PassArgs p{d_x, nullptr, ldx, 0, nullptr, d_x_t, nullptr, nullptr, static_cast<float *>(d_workspace), rows, dim, 0, split_rows_of(rows)};
    const int splits = splits_of(rows);
    if (vec_ok(dim, ldx, 0, rows, d_x, nullptr, nullptr, d_x_t))
        column_pass_vec_kernel<0><<<dim3((unsigned)((dim + 127) / 128), (unsigned)splits), kThreads, 0, st>>>(p);
    else
        column_pass_kernel<0><<<dim3((unsigned)((dim + 31) / 32), (unsigned)splits), kThreads, 0, st>>>(p);
    column_finalize_kernel<0><<<(dim + 31) / 32, kThreads, 0, st>>>(p.partials, splits, dim, rows, d_x, d_mean, d_var);
    RF_CUDA(cudaGetLastError());
    g_launches.fetch_add(2);
    return RF_OK;
}

int rf_activation_backward(const float *d_grad_out, const float *d_out, int64_t rows, int32_t units, int activation, float *d_grad_pre,
                           float *d_grad_pre_t, float *d_grad_bias, void *d_workspace, int64_t workspace_bytes, void *stream) {
    if (rows <= 0 || units <= 0) return set_error(RF_ERR_INVALID, "rf_activation_backward: bad shape");
    if (activation < RF_ACT_NONE || activation > RF_ACT_SIGMOID)
        return set_error(RF_ERR_UNSUPPORTED, "rf_activation_backward: the derivative is taken from the output; activation %d is not covered", activation);
    if (!d_grad_out || !d_grad_bias || (activation != RF_ACT_NONE && (!d_out || !d_grad_pre)))
        return set_error(RF_ERR_INVALID, "rf_activation_backward: NULL buffer");
    int rc = check_ws(rows, units, d_workspace, workspace_bytes);
    if (rc != RF_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    PassArgs p{d_grad_out, d_out ? d_out : d_grad_out, units, units, d_grad_pre, d_grad_pre_t, nullptr, nullptr,
               static_cast<float *>(d_workspace), rows, units, activation, split_rows_of(rows)};
    const int splits = splits_of(rows);
    if (vec_ok(units, units, units, rows, d_grad_out, p.b, d_grad_pre, d_grad_pre_t))
        column_pass_vec_kernel<1><<<dim3((unsigned)((units + 127) / 128), (unsigned)splits), kThreads, 0, st>>>(p);
    else
        column_pass_kernel<1><<<dim3((unsigned)((units + 31) / 32), (unsigned)splits), kThreads, 0, st>>>(p);
    column_finalize_kernel<1><<<(units + 31) / 32, kThreads, 0, st>>>(p.partials, splits, units, rows, nullptr, d_grad_bias, nullptr);
    RF_CUDA(cudaGetLastError());
    g_launches.fetch_add(2);
    return RF_OK;
}

int rf_batchnorm_backward(const float *d_grad_normed, const float *d_x, int64_t ldx, const float *d_mean, const float *d_rstd,
                          const float *d_scale, int64_t rows, int32_t dim, float *d_grad_gamma, float *d_grad_beta, float *d_grad_x,
                          void *d_workspace, int64_t workspace_bytes, void *stream) {
    if (rows <= 0 || dim <= 0 || dim % 4 || ldx < dim || ldx % 4) return set_error(RF_ERR_INVALID, "rf_batchnorm_backward: bad shape (dim and ldx multiples of 4)");
    if (!d_grad_normed || !d_x || !d_mean || !d_rstd || !d_scale || !d_grad_gamma || !d_grad_beta || !d_grad_x)
        return set_error(RF_ERR_INVALID, "rf_batchnorm_backward: NULL buffer");
    int rc = check_ws(rows, dim, d_workspace, workspace_bytes);
    if (rc != RF_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    PassArgs p{d_grad_normed, d_x, dim, ldx, nullptr, nullptr, d_mean, d_rstd, static_cast<float *>(d_workspace), rows, dim, 0, split_rows_of(rows)};
    const int splits = splits_of(rows);
    if (vec_ok(dim, dim, ldx, rows, d_grad_normed, d_x, nullptr, nullptr))
        column_pass_vec_kernel<2><<<dim3((unsigned)((dim + 127) / 128), (unsigned)splits), kThreads, 0, st>>>(p);
    else
        column_pass_kernel<2><<<dim3((unsigned)((dim + 31) / 32), (unsigned)splits), kThreads, 0, st>>>(p);
    column_finalize_kernel<2><<<(dim + 31) / 32, kThreads, 0, st>>>(p.partials, splits, dim, rows, nullptr, d_grad_beta, d_grad_gamma);
    int64_t blocks = (rows * (dim / 4) + kThreads - 1) / kThreads;
    if (blocks > 148 * 16) blocks = 148 * 16;
    batchnorm_dx_kernel<<<(unsigned)blocks, kThreads, 0, st>>>(d_grad_normed, d_x, ldx, d_mean, d_rstd, d_scale, d_grad_beta, d_grad_gamma,
                                                               rows, dim, d_grad_x);
    RF_CUDA(cudaGetLastError());
    g_launches.fetch_add(3);
    return RF_OK;
}

}  // extern "C"
