// rf_dense.cu -- the two dense contractions of the path in exact fp32 (CUDA-core) form:
//
//   * scaled_dot_product_attention         /root/reference/backend/layers/layer_utils.py:4-24
//   * the B x B in-batch logits q . d^T and the per-row statistics every two-tower loss of
//     /root/reference/backend/lossess/match_losses.py needs (:119-226), reduced on the fly:
//     the B x B matrix (268 MB at B = 8192, 17 GB at 65536) is never written to memory.
//
// These kernels are the fp32 reference-precision mode (the reference computes in fp32, its zipped
// wrappers even in fp64).  rf_logits_tc.cu holds the tcgen05 tensor-core version of the logits
// statistics for the throughput mode.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <atomic>

#include "../../include/rf_b200.h"
#include "rf_common.h"

namespace rf {

extern std::atomic<int64_t> g_launches;

bool sdpa_tc_supported(int64_t n_seq, int S, int dh, const float *q, const float *k, const float *v, const float *out);
int launch_sdpa_tc(const float *q, const float *k, const float *v, const float *mask, int64_t n_seq, int S, int dh, float *out,
                   cudaStream_t st, int64_t ld = 0);
struct RowStat;
int launch_logits_tc(const float *q, const float *d, const float *diag, const float *colw, int B, int Dt, float scale,
                     float margin, bool full_stats, RowStat *part, int max_splits, int *splits, cudaStream_t st);
int launch_logits_bf16(const void *q16, const void *d16, const float *diag, const float *colw, int B, int Dt, float scale,
                       float margin, bool full_stats, RowStat *part, int max_splits, int *splits, cudaStream_t st);

// ------------------------------------------------------------------------------------------
// In-batch row statistics.  For S = q d^T (q, d: [B, Dt] fp32 row-major), per row i:
//   m_i, l_i : running max and sum of exp(scale * S_ij - m_i) over this CTA's column range
//   hinge_i  : sum_j clip(S_ij - S_ii + margin, 0, 1e14)            (batch_neg_sample_margin_rank_loss)
//   maxoff_i : max_j (j == i ? 0 : S_ij)                            (batch_hard_neg_sample_margin_rank_loss)
// Grid: (row tiles of 64, column splits).  64 x 64 output tile per step, 256 threads, 4 x 4 per
// thread, K staged through shared memory in slabs of 16.
// ------------------------------------------------------------------------------------------
constexpr int kTM = 64, kTN = 64, kTK = 16;

struct RowStat {
    float m, l, hinge, maxoff;
};

__global__ void __launch_bounds__(256) inbatch_rowstats_kernel(const float *__restrict__ q, const float *__restrict__ d,
                                                               const float *__restrict__ diag,
                                                               const float *__restrict__ colw, int B, int Dt, float scale,
                                                               float margin, int cols_per_split,
                                                               RowStat *__restrict__ part /* [splits][B] */) {
    __shared__ float qs[kTK][kTM + 4];
    __shared__ float ds[kTK][kTN + 4];
    __shared__ RowStat red[kTM][17];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;           // thread (ty, tx): rows ty*4.., cols tx*4..
    const int row0 = blockIdx.x * kTM;
    const int col_begin = blockIdx.y * cols_per_split;
    const int col_end = min(B, col_begin + cols_per_split);

    float rm[4], rl[4], rh[4], rx[4], rdiag[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        rm[i] = -INFINITY;
        rl[i] = 0.f;
        rh[i] = 0.f;
        rx[i] = -INFINITY;
        const int r = row0 + ty * 4 + i;
        rdiag[i] = r < B ? diag[r] : 0.f;
    }
    for (int col0 = col_begin; col0 < col_end; col0 += kTN) {
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        for (int k0 = 0; k0 < Dt; k0 += kTK) {
            // 64 x 16 slabs of q and d, transposed into [k][row]; 256 threads x 4 elements each
            {
                const int r = tid >> 2, kk = (tid & 3) * 4;
                const int gr = row0 + r, gc = col0 + r;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int k = k0 + kk + e;
                    qs[kk + e][r] = (gr < B && k < Dt) ? q[(size_t)gr * Dt + k] : 0.f;
                    ds[kk + e][r] = (gc < col_end && k < Dt) ? d[(size_t)gc * Dt + k] : 0.f;
                }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < kTK; ++k) {
                float a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = qs[k][ty * 4 + i];
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = ds[k][tx * 4 + j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = row0 + ty * 4 + i;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = col0 + tx * 4 + j;
                if (r < B && c < col_end) {
                    const float s = acc[i][j];
                    const float x = scale * s;
                    if (x > rm[i]) {
                        rl[i] = rl[i] * expf(rm[i] - x) + 1.f;
                        rm[i] = x;
                    } else {
                        rl[i] += expf(x - rm[i]);
                    }
                    const float h = s - rdiag[i] + margin;
                    rh[i] += fminf(fmaxf(h, 0.f), 1e14f) * (colw ? colw[c] : 1.f);
                    const float off = (c == r) ? 0.f : s;
                    rx[i] = fmaxf(rx[i], off);
                }
            }
        }
    }
    // combine the 16 threads that share a row
#pragma unroll
    for (int i = 0; i < 4; ++i) red[ty * 4 + i][tx] = RowStat{rm[i], rl[i], rh[i], rx[i]};
    __syncthreads();
    if (tid < kTM) {
        const int r = row0 + tid;
        float m = -INFINITY, l = 0.f, h = 0.f, x = -INFINITY;
        for (int t = 0; t < 16; ++t) {
            const RowStat s = red[tid][t];
            if (s.m > m) {
                l = l * expf(m - s.m) + s.l;
                m = s.m;
            } else if (s.m > -INFINITY) {
                l += s.l * expf(s.m - m);
            }
            h += s.hinge;
            x = fmaxf(x, s.maxoff);
        }
        if (r < B) part[(size_t)blockIdx.y * B + r] = RowStat{m, l, h, x};
    }
}

// Round fp32 to the nearest TF32 value (cvt.rna), kept in fp32 storage.  The tensor core truncates
// whatever it is given to TF32; rounding first makes the operand error unbiased (2^-11 relative,
// random sign) instead of a systematic 2^-10 shrink of every product.
__global__ void __launch_bounds__(256) round_tf32_kernel(const float *__restrict__ in, float *__restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t r;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(in[i]));
        out[i] = __uint_as_float(r);
    }
}

// The bf16 variant's whole pre-pass in ONE launch (it was three: two conversions and the diagonal; at B = 8192 the launches
// around the 36 us tensor-core kernel cost almost as much as the kernel): one warp per row reads q_i and d_i once, writes both
// bf16 copies and leaves diag[i] = q_i . d_i in exact fp32, summed exactly as rowdot_kernel does.
__global__ void __launch_bounds__(256) prep_bf16_kernel(const float *__restrict__ q, const float *__restrict__ d, int B, int Dt,
                                                        unsigned short *__restrict__ q16, unsigned short *__restrict__ d16,
                                                        float *__restrict__ diag) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= B) return;
    float acc = 0.f;
    for (int k = lane; k < Dt; k += 32) {
        const float a = q[(size_t)row * Dt + k], b = d[(size_t)row * Dt + k];
        unsigned short ra, rb;
        asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(ra) : "f"(a));
        asm("cvt.rn.bf16.f32 %0, %1;" : "=h"(rb) : "f"(b));
        q16[(size_t)row * Dt + k] = ra;
        d16[(size_t)row * Dt + k] = rb;
        acc = fmaf(a, b, acc);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) diag[row] = acc;
}

// diag[i] = q_i . d_i (one warp per row)
__global__ void __launch_bounds__(256) rowdot_kernel(const float *__restrict__ q, const float *__restrict__ d, int B, int Dt,
                                                     float *__restrict__ diag) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= B) return;
    float acc = 0.f;
    for (int k = lane; k < Dt; k += 32) acc = fmaf(q[(size_t)row * Dt + k], d[(size_t)row * Dt + k], acc);
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) diag[row] = acc;
}

// merge the column splits; lse_i = m + log(l); loss = mean_i( -(scale * diag_i - lse_i) * y_i ).
// Many CTAs; each leaves a float64 partial of the loss, and the last one to finish (ticket counter)
// adds the partials in block order, so the scalar is deterministic.
constexpr int kFinThreads = 256;

__global__ void __launch_bounds__(kFinThreads) inbatch_finalize_kernel(const RowStat *__restrict__ part, int splits, int B,
                                                                       const float *__restrict__ diag,
                                                                       const float *__restrict__ y, float scale,
                                                                       float *__restrict__ lse_out, float *__restrict__ hinge_out,
                                                                       float *__restrict__ maxoff_out, float *__restrict__ loss_out,
                                                                       double *__restrict__ block_sums, unsigned int *ticket) {
    __shared__ double wsum[kFinThreads / 32];
    __shared__ bool last;
    double local = 0.0;
    const int r = blockIdx.x * kFinThreads + threadIdx.x;
    if (r < B) {
        float m = -INFINITY, l = 0.f, h = 0.f, x = -INFINITY;
        for (int s = 0; s < splits; ++s) {
            const RowStat p = part[(size_t)s * B + r];
            if (p.m > m) {
                l = l * expf(m - p.m) + p.l;
                m = p.m;
            } else if (p.m > -INFINITY) {
                l += p.l * expf(p.m - m);
            }
            h += p.hinge;
            x = fmaxf(x, p.maxoff);
        }
        const float lse = m + logf(l);
        if (lse_out) lse_out[r] = lse;
        if (hinge_out) hinge_out[r] = h;
        if (maxoff_out) maxoff_out[r] = x;
        if (y) local = (double)(-(scale * diag[r] - lse) * y[r]);
    }
    if (!loss_out) return;
    for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < kFinThreads / 32; ++w) t += wsum[w];
        block_sums[blockIdx.x] = t;
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double t = 0.0;
        for (unsigned int i = 0; i < gridDim.x; ++i) t += block_sums[i];
        *loss_out = (float)(t / (double)B);
        *ticket = 0;                                 // ready for the next call
    }
}

static int launch_finalize(const RowStat *part, int splits, int B, const float *diag, const float *y, float scale, float *lse,
                           float *hinge, float *maxoff, float *loss, void *fin_ws, cudaStream_t st) {
    const int blocks = (B + kFinThreads - 1) / kFinThreads;
    double *block_sums = static_cast<double *>(fin_ws);
    unsigned int *ticket = reinterpret_cast<unsigned int *>(block_sums + blocks);
    if (loss) RF_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned int), st));
    inbatch_finalize_kernel<<<blocks, kFinThreads, 0, st>>>(part, splits, B, diag, y, scale, lse, hinge, maxoff, loss,
                                                           block_sums, ticket);
    return RF_OK;
}

static int64_t finalize_ws_bytes(int64_t batch) { return ((batch + kFinThreads - 1) / kFinThreads) * 8 + 64; }

// ------------------------------------------------------------------------------------------
// scaled_dot_product_attention.  q, k, v: [NB, S, dh] fp32; mask: [NB, S] or NULL (the reference's
// [..., S, 1] mask broadcasts over keys: mask[i] == 0 replaces the whole QUERY row i of the logits
// by -4294967295 -> uniform attention).  One CTA per (batch x head); everything in shared memory.
// ------------------------------------------------------------------------------------------
constexpr int kSdpaRows = 32;     // query rows per pass: the S x S score matrix never has to fit at once

__global__ void __launch_bounds__(128) sdpa_kernel(const float *__restrict__ q, const float *__restrict__ k,
                                                   const float *__restrict__ v, const float *__restrict__ mask, int S, int dh,
                                                   float inv_sqrt_dk, float *__restrict__ out) {
    extern __shared__ float sm[];
    const int dp = dh + 1;                      // padded row stride: conflict-free column walks
    const int sp = S + 1;
    float *ks = sm;                             // [S][dp]
    float *vs = ks + S * dp;                    // [S][dp]
    float *qs = vs + S * dp;                    // [kSdpaRows][dp]
    float *ps = qs + kSdpaRows * dp;            // [kSdpaRows][sp]
    const size_t base = (size_t)blockIdx.x * S * dh;
    for (int e = threadIdx.x; e < S * dh; e += blockDim.x) {
        const int r = e / dh, c = e - r * dh;
        ks[r * dp + c] = k[base + e];
        vs[r * dp + c] = v[base + e];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warp = blockDim.x >> 5;
    for (int i0 = 0; i0 < S; i0 += kSdpaRows) {
        const int rows = min(kSdpaRows, S - i0);
        __syncthreads();                        // K/V staged; previous pass's qs/ps consumed
        for (int e = threadIdx.x; e < rows * dh; e += blockDim.x) {
            const int r = e / dh, c = e - r * dh;
            qs[r * dp + c] = q[base + (size_t)(i0 + r) * dh + c];
        }
        __syncthreads();
        for (int e = threadIdx.x; e < rows * S; e += blockDim.x) {
            const int i = e / S, j = e - i * S;
            float acc = 0.f;
            for (int c = 0; c < dh; ++c) acc = fmaf(qs[i * dp + c], ks[j * dp + c], acc);
            float logit = acc * inv_sqrt_dk;
            if (mask && mask[(size_t)blockIdx.x * S + i0 + i] == 0.f) logit = -4294967295.0f;
            ps[i * sp + j] = logit;
        }
        __syncthreads();
        for (int i = warp; i < rows; i += n_warp) {          // softmax per row: one warp per row
            float mx = -INFINITY;
            for (int j = lane; j < S; j += 32) mx = fmaxf(mx, ps[i * sp + j]);
            for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float sum = 0.f;
            for (int j = lane; j < S; j += 32) {
                const float p = expf(ps[i * sp + j] - mx);
                ps[i * sp + j] = p;
                sum += p;
            }
            for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float inv = 1.f / sum;
            for (int j = lane; j < S; j += 32) ps[i * sp + j] *= inv;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < rows * dh; e += blockDim.x) {
            const int i = e / dh, c = e - i * dh;
            float acc = 0.f;
            for (int j = 0; j < S; ++j) acc = fmaf(ps[i * sp + j], vs[j * dp + c], acc);
            out[base + (size_t)(i0 + i) * dh + c] = acc;
        }
    }
}

}  // namespace rf

using namespace rf;

extern "C" {

int rf_sdpa_forward(const float *d_q, const float *d_k, const float *d_v, const float *d_mask, int64_t n_batch_heads,
                    int32_t seq_len, int32_t head_dim, float *d_out, void *stream) {
    if (n_batch_heads < 0 || seq_len <= 0 || head_dim <= 0) return set_error(RF_ERR_INVALID, "bad SDPA shape");
    if (n_batch_heads == 0) return RF_OK;
    if (!d_q || !d_k || !d_v || !d_out) return set_error(RF_ERR_INVALID, "rf_sdpa_forward: NULL buffer");
    const size_t smem = sizeof(float) * ((size_t)(2 * seq_len + kSdpaRows) * (head_dim + 1) + (size_t)kSdpaRows * (seq_len + 1));
    if (smem > 200 * 1024) return set_error(RF_ERR_UNSUPPORTED, "SDPA tile (S=%d, dh=%d) exceeds shared memory", seq_len, head_dim);
    if (n_batch_heads > INT32_MAX) return set_error(RF_ERR_INVALID, "too many batch x heads");
    RF_CUDA(cudaFuncSetAttribute(sdpa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sdpa_kernel<<<(unsigned)n_batch_heads, 128, smem, static_cast<cudaStream_t>(stream)>>>(
        d_q, d_k, d_v, d_mask, seq_len, head_dim, 1.0f / sqrtf((float)head_dim), d_out);
    RF_CUDA(cudaGetLastError());
    g_launches.fetch_add(1);
    return RF_OK;
}

int rf_sdpa_forward_tc(const float *d_q, const float *d_k, const float *d_v, const float *d_mask, int64_t n_batch_heads,
                       int32_t seq_len, int32_t head_dim, float *d_out, void *stream) {
    if (n_batch_heads < 0 || seq_len <= 0 || head_dim <= 0) return set_error(RF_ERR_INVALID, "bad SDPA shape");
    if (n_batch_heads == 0) return RF_OK;
    if (!d_q || !d_k || !d_v || !d_out) return set_error(RF_ERR_INVALID, "rf_sdpa_forward_tc: NULL buffer");
    if (!sdpa_tc_supported(n_batch_heads, seq_len, head_dim, d_q, d_k, d_v, d_out))
        return set_error(RF_ERR_UNSUPPORTED, "tensor-core SDPA takes seq_len <= 64 and head_dim in {32, 64, 96}, 16-byte aligned");
    return launch_sdpa_tc(d_q, d_k, d_v, d_mask, n_batch_heads, seq_len, head_dim, d_out, static_cast<cudaStream_t>(stream));
}

int rf_sdpa_forward_tc_strided(const float *d_q, const float *d_k, const float *d_v, int64_t row_pitch, const float *d_mask,
                               int64_t n_batch_heads, int32_t seq_len, int32_t head_dim, float *d_out, void *stream) {
    if (n_batch_heads < 0 || seq_len <= 0 || head_dim <= 0) return set_error(RF_ERR_INVALID, "bad SDPA shape");
    if (row_pitch < head_dim || row_pitch % 4) return set_error(RF_ERR_INVALID, "row_pitch must be >= head_dim and a multiple of 4 floats");
    if (n_batch_heads == 0) return RF_OK;
    if (!d_q || !d_k || !d_v || !d_out) return set_error(RF_ERR_INVALID, "rf_sdpa_forward_tc_strided: NULL buffer");
    if (!sdpa_tc_supported(n_batch_heads, seq_len, head_dim, d_q, d_k, d_v, d_out))
        return set_error(RF_ERR_UNSUPPORTED, "tensor-core SDPA takes seq_len <= 64 and head_dim in {32, 64, 96}, 16-byte aligned");
    return launch_sdpa_tc(d_q, d_k, d_v, d_mask, n_batch_heads, seq_len, head_dim, d_out, static_cast<cudaStream_t>(stream), row_pitch);
}

int64_t rf_inbatch_workspace_bytes(int64_t batch) {
    if (batch <= 0) return 0;
    const int64_t row_tiles = (batch + kTM - 1) / kTM;
    int64_t splits = (148 * 4 + row_tiles - 1) / row_tiles;
    if (splits < 1) splits = 1;
    if (splits > 64) splits = 64;
    return splits * batch * (int64_t)sizeof(RowStat) + ((batch * (int64_t)sizeof(float) + 15) & ~(int64_t)15) +
           finalize_ws_bytes(batch);
}

int64_t rf_inbatch_workspace_bytes_tc(int64_t batch, int32_t dim) {
    if (batch <= 0 || dim <= 0) return 0;
    return rf_inbatch_workspace_bytes(batch) + 512 + 2 * (((int64_t)batch * dim + 63) & ~(int64_t)63) * (int64_t)sizeof(float);
}

int rf_inbatch_rowstats_tc(const float *d_query, const float *d_doc, const float *d_y, const float *d_col_weight, int64_t batch,
                           int32_t dim, float scale, float margin, void *d_workspace, float *d_lse, float *d_diag,
                           float *d_hinge, float *d_maxoff, float *d_loss, void *stream) {
    if (batch < 0 || dim <= 0) return set_error(RF_ERR_INVALID, "bad in-batch shape");
    if (batch == 0) return RF_OK;
    if (batch > (1 << 24)) return set_error(RF_ERR_UNSUPPORTED, "batch too large");
    if (!d_query || !d_doc || !d_workspace) return set_error(RF_ERR_INVALID, "rf_inbatch_rowstats_tc: NULL buffer");
    if (d_loss && !d_y) return set_error(RF_ERR_INVALID, "loss requested without y_true");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int B = (int)batch;
    // workspace layout as in rf_inbatch_rowstats: [splits <= 64][B] RowStat, then B floats for the diagonal
    const int64_t row_tiles = (batch + kTM - 1) / kTM;
    int64_t max_splits = (148 * 4 + row_tiles - 1) / row_tiles;
    if (max_splits < 1) max_splits = 1;
    if (max_splits > 64) max_splits = 64;
    RowStat *part = static_cast<RowStat *>(d_workspace);
    float *diag_ws = reinterpret_cast<float *>(part + (size_t)max_splits * B);
    float *diag = d_diag ? d_diag : diag_ws;
    void *fin_ws = reinterpret_cast<char *>(diag_ws) + (((size_t)B * sizeof(float) + 15) & ~(size_t)15);
    // TF32-rounded operand copies live after the statistics in the workspace, 256-byte aligned
    uintptr_t p = (reinterpret_cast<uintptr_t>(fin_ws) + finalize_ws_bytes(B) + 255) & ~(uintptr_t)255;
    float *q32 = reinterpret_cast<float *>(p);
    float *d32 = q32 + (((size_t)B * dim + 63) & ~(size_t)63);
    const int64_t n = (int64_t)B * dim;
    const int rgrid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
    round_tf32_kernel<<<rgrid, 256, 0, st>>>(d_query, q32, n);
    round_tf32_kernel<<<rgrid, 256, 0, st>>>(d_doc, d32, n);
    rowdot_kernel<<<(B * 32 + 255) / 256, 256, 0, st>>>(d_query, d_doc, B, dim, diag);   // exact fp32 diagonal
    int splits = 0;
    const bool full = d_hinge != nullptr || d_maxoff != nullptr;
    int rc = launch_logits_tc(q32, d32, diag, d_col_weight, B, dim, scale, margin, full, part, (int)max_splits, &splits, st);
    if (rc != RF_OK) return rc;
    rc = launch_finalize(part, splits, B, diag, d_y, scale, d_lse, d_hinge, d_maxoff, d_loss, fin_ws, st);
    if (rc != RF_OK) return rc;
    RF_CUDA(cudaGetLastError());
    g_launches.fetch_add(4);
    return RF_OK;
}

int rf_inbatch_rowstats_bf16(const float *d_query, const float *d_doc, const float *d_y, const float *d_col_weight, int64_t batch,
                             int32_t dim, float scale, float margin, void *d_workspace, float *d_lse, float *d_diag,
                             float *d_hinge, float *d_maxoff, float *d_loss, void *stream) {
    if (batch < 0 || dim <= 0) return set_error(RF_ERR_INVALID, "bad in-batch shape");
    if (batch == 0) return RF_OK;
    if (batch > (1 << 24)) return set_error(RF_ERR_UNSUPPORTED, "batch too large");
    if (dim % 8 != 0 || dim > 256) return set_error(RF_ERR_UNSUPPORTED, "bf16 tensor-core logits need dim %% 8 == 0 and dim <= 256");
    if (!d_query || !d_doc || !d_workspace) return set_error(RF_ERR_INVALID, "rf_inbatch_rowstats_bf16: NULL buffer");
    if (d_loss && !d_y) return set_error(RF_ERR_INVALID, "loss requested without y_true");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int B = (int)batch;
    // same workspace layout as rf_inbatch_rowstats_tc; the operand copies are bf16 here (half the reserved space)
    const int64_t row_tiles = (batch + kTM - 1) / kTM;
    int64_t max_splits = (148 * 4 + row_tiles - 1) / row_tiles;
    if (max_splits < 1) max_splits = 1;
    if (max_splits > 64) max_splits = 64;
    RowStat *part = static_cast<RowStat *>(d_workspace);
    float *diag_ws = reinterpret_cast<float *>(part + (size_t)max_splits * B);
    float *diag = d_diag ? d_diag : diag_ws;
    void *fin_ws = reinterpret_cast<char *>(diag_ws) + (((size_t)B * sizeof(float) + 15) & ~(size_t)15);
    uintptr_t p = (reinterpret_cast<uintptr_t>(fin_ws) + finalize_ws_bytes(B) + 255) & ~(uintptr_t)255;
    unsigned short *q16 = reinterpret_cast<unsigned short *>(p);
    unsigned short *d16 = q16 + (((size_t)B * dim + 127) & ~(size_t)127);
    prep_bf16_kernel<<<(B * 32 + 255) / 256, 256, 0, st>>>(d_query, d_doc, B, dim, q16, d16, diag);   // bf16 copies + exact fp32 diagonal
    int splits = 0;
    const bool full = d_hinge != nullptr || d_maxoff != nullptr;
    int rc = launch_logits_bf16(q16, d16, diag, d_col_weight, B, dim, scale, margin, full, part, (int)max_splits, &splits, st);
    if (rc != RF_OK) return rc;
    rc = launch_finalize(part, splits, B, diag, d_y, scale, d_lse, d_hinge, d_maxoff, d_loss, fin_ws, st);
    if (rc != RF_OK) return rc;
    RF_CUDA(cudaGetLastError());
    g_launches.fetch_add(2);
    return RF_OK;
}

int rf_inbatch_rowstats(const float *d_query, const float *d_doc, const float *d_y, const float *d_col_weight, int64_t batch,
                        int32_t dim, float scale, float margin, void *d_workspace, float *d_lse, float *d_diag, float *d_hinge, float *d_maxoff,
                        float *d_loss, void *stream) {
    if (batch < 0 || dim <= 0) return set_error(RF_ERR_INVALID, "bad in-batch shape");
    if (batch == 0) return RF_OK;
    if (batch > (1 << 24)) return set_error(RF_ERR_UNSUPPORTED, "batch too large");
    if (!d_query || !d_doc || !d_workspace) return set_error(RF_ERR_INVALID, "rf_inbatch_rowstats: NULL buffer");
    if (d_loss && !d_y) return set_error(RF_ERR_INVALID, "loss requested without y_true");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int B = (int)batch;
    const int row_tiles = (B + kTM - 1) / kTM;
    int splits = (148 * 4 + row_tiles - 1) / row_tiles;
    if (splits < 1) splits = 1;
    if (splits > 64) splits = 64;
    const int max_splits = splits;              // the workspace is laid out for this many
    int cols = (B + splits - 1) / splits;
    cols = (cols + kTN - 1) / kTN * kTN;
    splits = (B + cols - 1) / cols;
    RowStat *part = static_cast<RowStat *>(d_workspace);
    float *diag_ws = reinterpret_cast<float *>(part + (size_t)max_splits * B);
    float *diag = d_diag ? d_diag : diag_ws;
    void *fin_ws = reinterpret_cast<char *>(diag_ws) + (((size_t)B * sizeof(float) + 15) & ~(size_t)15);
    rowdot_kernel<<<(B * 32 + 255) / 256, 256, 0, st>>>(d_query, d_doc, B, dim, diag);
    inbatch_rowstats_kernel<<<dim3(row_tiles, splits), 256, 0, st>>>(d_query, d_doc, diag, d_col_weight, B, dim, scale, margin,
                                                                     cols, part);
    {
        int rc = launch_finalize(part, splits, B, diag, d_y, scale, d_lse, d_hinge, d_maxoff, d_loss, fin_ws, st);
        if (rc != RF_OK) return rc;
    }
    RF_CUDA(cudaGetLastError());
    g_launches.fetch_add(3);
    return RF_OK;
}

}  // extern "C"
