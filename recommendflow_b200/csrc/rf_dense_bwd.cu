// rf_dense_bwd.cu -- backward of the two dense contractions of the path (SURVEY.md §8f rank 1:
// what `model.fit` differentiates; first correct CUDA-core fp32 versions, tensor-core versions are
// the follow-up):
//
//  * rf_sdpa_backward: gradient of scaled_dot_product_attention (backend/layers/layer_utils.py:4-24)
//    w.r.t. q, k, v.  One CTA per (batch, head) slice, everything in shared memory: P is recomputed
//    from q, k and the mask exactly like the forward (a masked QUERY row is filled with the constant
//    -4294967295, so it becomes uniform attention and passes no gradient to q / k, only to v).
//        dV = P^T dO;  dP = dO V^T;  dS = P o (dP - rowsum(P o dP));  dQ = dS K / sqrt(dh);
//        dK = dS^T Q / sqrt(dh)
//  * rf_inbatch_softmax_ce_backward: gradient of batch_neg_sample_scaled_multi_class_ce_loss
//    (backend/lossess/match_losses.py:150-165) w.r.t. query and doc, from the per-row log-sum-exp
//    the forward (rf_inbatch_rowstats*) already produced; the B x B matrix is never stored:
//        C_ij = upstream * y_i * scale / B * (exp(scale * S_ij - lse_i) - [i == j])
//        dQ = C D,   dD = C^T Q
//    Two passes of one kernel (owner = query rows / owner = doc rows): a CTA owns 32 rows of its
//    side, keeps their gradient in registers, streams 64-row tiles of the other side through
//    shared memory, recomputes the S tile, forms C and accumulates.  No atomics: deterministic.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <atomic>

#include "../../include/rf_b200.h"
#include "rf_common.h"

namespace rf {

extern std::atomic<int64_t> g_launches;

namespace {

constexpr int kThreads = 256;

// --------------------------------------------------------------------------------------------
// SDPA backward
// --------------------------------------------------------------------------------------------
constexpr int kBwdMaxSeq = 64;
constexpr int kBwdMaxHeadDim = 128;

__global__ void __launch_bounds__(kThreads)
sdpa_backward_kernel(const float *__restrict__ q, const float *__restrict__ k, const float *__restrict__ v,
                     const float *__restrict__ mask, const float *__restrict__ grad_out, int S, int dh,
                     float *__restrict__ dq, float *__restrict__ dk, float *__restrict__ dv) {
    extern __shared__ float smem[];
    const int ld = dh + 1, lp = S + 1;
    float *Q = smem, *K = Q + S * ld, *V = K + S * ld, *G = V + S * ld;      // [S][dh + 1]
    float *P = G + S * ld, *dS = P + S * lp;                                // [S][S + 1]
    float *rowm = dS + S * lp;                                              // [S] mask flag per query row
    const int64_t base = (int64_t)blockIdx.x * S * dh;
    const int tid = threadIdx.x;
    for (int e = tid; e < S * dh; e += kThreads) {
        const int i = e / dh, c = e - i * dh;
        Q[i * ld + c] = q[base + e];
        K[i * ld + c] = k[base + e];
        V[i * ld + c] = v[base + e];
        G[i * ld + c] = grad_out[base + e];
    }
    for (int i = tid; i < S; i += kThreads) rowm[i] = mask ? mask[(int64_t)blockIdx.x * S + i] : 1.0f;
    __syncthreads();
    const float scale = 1.0f / sqrtf((float)dh);
    // logits (mask fill) and dP = dO V^T
    for (int e = tid; e < S * S; e += kThreads) {
        const int i = e / S, j = e - i * S;
        float s = 0.f, dp = 0.f;
        for (int c = 0; c < dh; ++c) {
            s = fmaf(Q[i * ld + c], K[j * ld + c], s);
            dp = fmaf(G[i * ld + c], V[j * ld + c], dp);
        }
        P[i * lp + j] = rowm[i] == 0.0f ? -4294967295.0f : s * scale;
        dS[i * lp + j] = dp;
    }
    __syncthreads();
    // row softmax, delta_i = sum_j P_ij dP_ij, dS = P o (dP - delta); one warp per row
    const int warp = tid >> 5, lane = tid & 31;
    for (int i = warp; i < S; i += kThreads / 32) {
        float mx = -INFINITY;
        for (int j = lane; j < S; j += 32) mx = fmaxf(mx, P[i * lp + j]);
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float den = 0.f;
        for (int j = lane; j < S; j += 32) {
            const float p = expf(P[i * lp + j] - mx);
            P[i * lp + j] = p;
            den += p;
        }
        for (int o = 16; o; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
        const float inv = 1.0f / den;
        float delta = 0.f;
        for (int j = lane; j < S; j += 32) {
            const float p = P[i * lp + j] * inv;
            P[i * lp + j] = p;
            delta = fmaf(p, dS[i * lp + j], delta);
        }
        for (int o = 16; o; o >>= 1) delta += __shfl_xor_sync(0xffffffffu, delta, o);
        const bool masked = rowm[i] == 0.0f;      // constant logits: nothing flows to q / k from this row
        for (int j = lane; j < S; j += 32) dS[i * lp + j] = masked ? 0.f : P[i * lp + j] * (dS[i * lp + j] - delta);
    }
    __syncthreads();
    for (int e = tid; e < S * dh; e += kThreads) {
        const int i = e / dh, c = e - i * dh;
        float aq = 0.f, ak = 0.f, av = 0.f;
        for (int j = 0; j < S; ++j) {
            aq = fmaf(dS[i * lp + j], K[j * ld + c], aq);       // dQ_i = sum_j dS_ij K_j
            ak = fmaf(dS[j * lp + i], Q[j * ld + c], ak);       // dK_i = sum_j dS_ji Q_j
            av = fmaf(P[j * lp + i], G[j * ld + c], av);        // dV_i = sum_j P_ji dO_j
        }
        dq[base + e] = aq * scale;
        dk[base + e] = ak * scale;
        dv[base + e] = av;
    }
}

// Register-tiled version for head_dim % 4 == 0: the kernel above does one FMA per two shared-memory loads (it runs at the
// LDS rate, 1.74 ms for [8192, 50, 64]); here every thread owns a 4 x 4 tile of each product and reads its operands as
// 128-bit rows, ~8 FMAs per load instruction.  Same arithmetic (fp32 FMA, same summation order over the head dimension
// within a tile), rows padded to 64 with zeros so the tiles need no guards.
//   phase 1  S = Q K^T, dP = dO V^T            thread (ty, tx): rows ty + 16 a, columns tx + 16 b
//   phase 2  row softmax, delta, dS            one warp per row (as above)
//   phase 3  dQ = dS K, dK = dS^T Q, dV = P^T dO   thread (ty, tx): rows 4 ty + a, columns 4 tx .. 4 tx + 3 (+ 64 per pass)
constexpr int kTileSeq = 64;

__global__ void __launch_bounds__(kThreads, 2)
sdpa_backward_tiled_kernel(const float *__restrict__ q, const float *__restrict__ k, const float *__restrict__ v,
                           const float *__restrict__ mask, const float *__restrict__ grad_out, int S, int dh,
                           float *__restrict__ dq, float *__restrict__ dk, float *__restrict__ dv, int64_t pitch, int64_t out_pitch) {
    // pitch: row pitch (floats) of q, k, v; out_pitch: of dq, dk, dv (q | k | v may be column windows of one projection output);
    // grad_out is dense [n, S, dh]
    extern __shared__ __align__(16) float smem[];
    const int ld = dh + 4, lp = kTileSeq + 4;                               // row pitches: 16-byte rows, conflict-free float4 reads
    float *Q = smem, *K = Q + kTileSeq * ld, *V = K + kTileSeq * ld, *G = V + kTileSeq * ld;      // [64][dh + 4]
    float *P = G + kTileSeq * ld, *dS = P + kTileSeq * lp;                  // [64][68]
    float *rowm = dS + kTileSeq * lp;
    const int64_t base = (int64_t)blockIdx.x * S * dh, in_base = (int64_t)blockIdx.x * S * pitch, out_base = (int64_t)blockIdx.x * S * out_pitch;
    const int tid = threadIdx.x, q4 = dh >> 2;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = tid; e < kTileSeq * q4; e += kThreads) {
        const int i = e / q4, c = (e - i * q4) * 4;
        const bool ok = i < S;
        const int64_t at = in_base + (int64_t)i * pitch + c;
        *reinterpret_cast<float4 *>(Q + i * ld + c) = ok ? *reinterpret_cast<const float4 *>(q + at) : zero4;
        *reinterpret_cast<float4 *>(K + i * ld + c) = ok ? *reinterpret_cast<const float4 *>(k + at) : zero4;
        *reinterpret_cast<float4 *>(V + i * ld + c) = ok ? *reinterpret_cast<const float4 *>(v + at) : zero4;
        *reinterpret_cast<float4 *>(G + i * ld + c) = ok ? *reinterpret_cast<const float4 *>(grad_out + base + (int64_t)i * dh + c) : zero4;
    }
    for (int i = tid; i < kTileSeq; i += kThreads) rowm[i] = (i < S && mask) ? mask[(int64_t)blockIdx.x * S + i] : 1.0f;
    __syncthreads();
    const float scale = 1.0f / sqrtf((float)dh);
    const int ty = tid >> 4, tx = tid & 15;
    {
        float s[4][4], dp[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) s[a][b] = dp[a][b] = 0.f;
        for (int c = 0; c < dh; c += 4) {
            float4 qa[4], ga[4], kb[4], vb[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                qa[a] = *reinterpret_cast<const float4 *>(Q + (ty + 16 * a) * ld + c);
                ga[a] = *reinterpret_cast<const float4 *>(G + (ty + 16 * a) * ld + c);
                kb[a] = *reinterpret_cast<const float4 *>(K + (tx + 16 * a) * ld + c);
                vb[a] = *reinterpret_cast<const float4 *>(V + (tx + 16 * a) * ld + c);
            }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    s[a][b] = fmaf(qa[a].x, kb[b].x, s[a][b]);
                    s[a][b] = fmaf(qa[a].y, kb[b].y, s[a][b]);
                    s[a][b] = fmaf(qa[a].z, kb[b].z, s[a][b]);
                    s[a][b] = fmaf(qa[a].w, kb[b].w, s[a][b]);
                    dp[a][b] = fmaf(ga[a].x, vb[b].x, dp[a][b]);
                    dp[a][b] = fmaf(ga[a].y, vb[b].y, dp[a][b]);
                    dp[a][b] = fmaf(ga[a].z, vb[b].z, dp[a][b]);
                    dp[a][b] = fmaf(ga[a].w, vb[b].w, dp[a][b]);
                }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int i = ty + 16 * a, j = tx + 16 * b;
                P[i * lp + j] = rowm[i] == 0.0f ? -4294967295.0f : s[a][b] * scale;
                dS[i * lp + j] = dp[a][b];
            }
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    for (int i = warp; i < kTileSeq; i += kThreads / 32) {
        if (i >= S) {                                   // padding rows contribute nothing to dK / dV
            for (int j = lane; j < kTileSeq; j += 32) P[i * lp + j] = dS[i * lp + j] = 0.f;
            continue;
        }
        float mx = -INFINITY;
        for (int j = lane; j < S; j += 32) mx = fmaxf(mx, P[i * lp + j]);
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float den = 0.f;
        for (int j = lane; j < S; j += 32) {
            const float p = expf(P[i * lp + j] - mx);
            P[i * lp + j] = p;
            den += p;
        }
        for (int o = 16; o; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
        const float inv = 1.0f / den;
        float delta = 0.f;
        for (int j = lane; j < S; j += 32) {
            const float p = P[i * lp + j] * inv;
            P[i * lp + j] = p;
            delta = fmaf(p, dS[i * lp + j], delta);
        }
        for (int o = 16; o; o >>= 1) delta += __shfl_xor_sync(0xffffffffu, delta, o);
        const bool masked = rowm[i] == 0.0f;
        for (int j = lane; j < kTileSeq; j += 32) {
            const bool in = j < S;
            dS[i * lp + j] = (masked || !in) ? 0.f : P[i * lp + j] * (dS[i * lp + j] - delta);
            if (!in) P[i * lp + j] = 0.f;
        }
    }
    __syncthreads();
    for (int c0 = 0; c0 < dh; c0 += 64) {
        const int c = c0 + tx * 4;
        if (c >= dh) continue;
        float aq[4][4], ak[4][4], av[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) aq[a][b] = ak[a][b] = av[a][b] = 0.f;
        const int i0 = ty * 4;
        for (int j0 = 0; j0 < S; j0 += 4) {
            float4 dsr[4];                              // dS[i0 + a][j0 .. j0 + 3]
#pragma unroll
            for (int a = 0; a < 4; ++a) dsr[a] = *reinterpret_cast<const float4 *>(dS + (i0 + a) * lp + j0);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + u;                   // rows S .. 63 of K, Q, dO are zero and dS / P are zero there
                const float4 kj = *reinterpret_cast<const float4 *>(K + j * ld + c);
                const float4 qj = *reinterpret_cast<const float4 *>(Q + j * ld + c);
                const float4 gj = *reinterpret_cast<const float4 *>(G + j * ld + c);
                const float4 dst = *reinterpret_cast<const float4 *>(dS + j * lp + i0);     // dS[j][i0 .. i0 + 3]
                const float4 pt = *reinterpret_cast<const float4 *>(P + j * lp + i0);
                const float dsa[4] = {u == 0 ? dsr[0].x : u == 1 ? dsr[0].y : u == 2 ? dsr[0].z : dsr[0].w,
                                      u == 0 ? dsr[1].x : u == 1 ? dsr[1].y : u == 2 ? dsr[1].z : dsr[1].w,
                                      u == 0 ? dsr[2].x : u == 1 ? dsr[2].y : u == 2 ? dsr[2].z : dsr[2].w,
                                      u == 0 ? dsr[3].x : u == 1 ? dsr[3].y : u == 2 ? dsr[3].z : dsr[3].w};
                const float dta[4] = {dst.x, dst.y, dst.z, dst.w}, pta[4] = {pt.x, pt.y, pt.z, pt.w};
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    aq[a][0] = fmaf(dsa[a], kj.x, aq[a][0]);
                    aq[a][1] = fmaf(dsa[a], kj.y, aq[a][1]);
                    aq[a][2] = fmaf(dsa[a], kj.z, aq[a][2]);
                    aq[a][3] = fmaf(dsa[a], kj.w, aq[a][3]);
                    ak[a][0] = fmaf(dta[a], qj.x, ak[a][0]);
                    ak[a][1] = fmaf(dta[a], qj.y, ak[a][1]);
                    ak[a][2] = fmaf(dta[a], qj.z, ak[a][2]);
                    ak[a][3] = fmaf(dta[a], qj.w, ak[a][3]);
                    av[a][0] = fmaf(pta[a], gj.x, av[a][0]);
                    av[a][1] = fmaf(pta[a], gj.y, av[a][1]);
                    av[a][2] = fmaf(pta[a], gj.z, av[a][2]);
                    av[a][3] = fmaf(pta[a], gj.w, av[a][3]);
                }
            }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int i = i0 + a;
            if (i >= S) continue;
            const int64_t at = out_base + (int64_t)i * out_pitch + c;
            *reinterpret_cast<float4 *>(dq + at) = make_float4(aq[a][0] * scale, aq[a][1] * scale, aq[a][2] * scale, aq[a][3] * scale);
            *reinterpret_cast<float4 *>(dk + at) = make_float4(ak[a][0] * scale, ak[a][1] * scale, ak[a][2] * scale, ak[a][3] * scale);
            *reinterpret_cast<float4 *>(dv + at) = make_float4(av[a][0], av[a][1], av[a][2], av[a][3]);
        }
    }
}

// Tensor-core version (TF32 operands, fp32 accumulate) for the default "tf32" precision mode: the five 64 x 64 x dh products
// of a sequence run on warp-level mma.sync.m16n8k8 (four warps, a 16-row strip each; operands come from shared memory, so the
// transposed products dK = dS^T Q and dV = P^T dO are index arithmetic, not data movement).  At S = 50, dh = 64 the arithmetic
// (21 GF for 8192 sequences) stops being the bound: the kernel approaches the time its 8 tensors take through HBM.  The
// softmax / delta phase is the fp32 code of the kernels above; inputs, P and dS are rounded to nearest TF32 (cvt.rna) where
// they become operands.  A tcgen05 version would need 128-row tiles (two sequences per tile as in the forward) and five
// operand layouts per pair for a kernel that is already memory-bound -- not built.
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const float (&a)[4], float b0, float b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
                   "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

// acc[nt] (16 rows x 8 columns each, nt < NT) += A[16 rows starting at m0][K] * B[K][8 * NT columns].
// A(m, k) = a[m * a_m + k * a_k], B(k, n) = b[k * b_k + n * b_n]: either operand may be read transposed.
template <int NT>
__device__ __forceinline__ void warp_gemm(float (&acc)[NT][4], const float *__restrict__ a, int a_m, int a_k, const float *__restrict__ b,
                                          int b_k, int b_n, int m0, int K, int lane) {
    const int g = lane >> 2, t = lane & 3;
    for (int k0 = 0; k0 < K; k0 += 8) {
        float af[4];
        af[0] = a[(m0 + g) * a_m + (k0 + t) * a_k];
        af[1] = a[(m0 + g + 8) * a_m + (k0 + t) * a_k];
        af[2] = a[(m0 + g) * a_m + (k0 + t + 4) * a_k];
        af[3] = a[(m0 + g + 8) * a_m + (k0 + t + 4) * a_k];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const float b0 = b[(k0 + t) * b_k + (nt * 8 + g) * b_n], b1 = b[(k0 + t + 4) * b_k + (nt * 8 + g) * b_n];
            mma_tf32(acc[nt], af, b0, b1);
        }
    }
}

constexpr int kMmaThreads = 128;

template <int NTD>      // head_dim = 8 * NTD
__global__ void __launch_bounds__(kMmaThreads, 2)
sdpa_backward_mma_kernel(const float *__restrict__ q, const float *__restrict__ k, const float *__restrict__ v,
                         const float *__restrict__ mask, const float *__restrict__ grad_out, int S,
                         float *__restrict__ dq, float *__restrict__ dk, float *__restrict__ dv, int64_t pitch, int64_t out_pitch) {
    constexpr int dh = 8 * NTD;
    constexpr int ld = dh + 4, lp = kTileSeq + 4;
    extern __shared__ __align__(16) float smem[];
    float *Q = smem, *K = Q + kTileSeq * ld, *V = K + kTileSeq * ld, *G = V + kTileSeq * ld;      // [64][dh + 4], TF32-rounded
    float *P = G + kTileSeq * ld, *dS = P + kTileSeq * lp;                                        // [64][68]
    float *rowm = dS + kTileSeq * lp;
    const int64_t base = (int64_t)blockIdx.x * S * dh, in_base = (int64_t)blockIdx.x * S * pitch, out_base = (int64_t)blockIdx.x * S * out_pitch;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    constexpr int q4 = dh >> 2;
    for (int e = tid; e < kTileSeq * q4; e += kMmaThreads) {
        const int i = e / q4, c = (e - i * q4) * 4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a, cc = a, d = a;
        if (i < S) {
            const int64_t at = in_base + (int64_t)i * pitch + c;
            a = *reinterpret_cast<const float4 *>(q + at);
            b = *reinterpret_cast<const float4 *>(k + at);
            cc = *reinterpret_cast<const float4 *>(v + at);
            d = *reinterpret_cast<const float4 *>(grad_out + base + (int64_t)i * dh + c);
        }
        *reinterpret_cast<float4 *>(Q + i * ld + c) = make_float4(to_tf32(a.x), to_tf32(a.y), to_tf32(a.z), to_tf32(a.w));
        *reinterpret_cast<float4 *>(K + i * ld + c) = make_float4(to_tf32(b.x), to_tf32(b.y), to_tf32(b.z), to_tf32(b.w));
        *reinterpret_cast<float4 *>(V + i * ld + c) = make_float4(to_tf32(cc.x), to_tf32(cc.y), to_tf32(cc.z), to_tf32(cc.w));
        *reinterpret_cast<float4 *>(G + i * ld + c) = make_float4(to_tf32(d.x), to_tf32(d.y), to_tf32(d.z), to_tf32(d.w));
    }
    for (int i = tid; i < kTileSeq; i += kMmaThreads) rowm[i] = (i < S && mask) ? mask[(int64_t)blockIdx.x * S + i] : 1.0f;
    __syncthreads();
    const float scale = rsqrtf((float)dh);
    const int m0 = warp * 16;
    {   // S = Q K^T and dP = dO V^T: rows m0 .. m0 + 15, all 64 key columns
        float acc[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
        warp_gemm<8>(acc, Q, ld, 1, K, 1, ld, m0, dh, lane);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int j = nt * 8 + 2 * t;
            const bool m_lo = rowm[m0 + g] == 0.0f, m_hi = rowm[m0 + g + 8] == 0.0f;
            P[(m0 + g) * lp + j] = m_lo ? -4294967295.0f : acc[nt][0] * scale;
            P[(m0 + g) * lp + j + 1] = m_lo ? -4294967295.0f : acc[nt][1] * scale;
            P[(m0 + g + 8) * lp + j] = m_hi ? -4294967295.0f : acc[nt][2] * scale;
            P[(m0 + g + 8) * lp + j + 1] = m_hi ? -4294967295.0f : acc[nt][3] * scale;
            acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
        }
        warp_gemm<8>(acc, G, ld, 1, V, 1, ld, m0, dh, lane);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int j = nt * 8 + 2 * t;
            dS[(m0 + g) * lp + j] = acc[nt][0];
            dS[(m0 + g) * lp + j + 1] = acc[nt][1];
            dS[(m0 + g + 8) * lp + j] = acc[nt][2];
            dS[(m0 + g + 8) * lp + j + 1] = acc[nt][3];
        }
    }
    __syncwarp();
    // softmax rows, delta, dS: a warp owns exactly the 16 rows it just produced (no CTA barrier needed yet)
    for (int i = m0; i < m0 + 16; ++i) {
        if (i >= S) {
            for (int j = lane; j < kTileSeq; j += 32) P[i * lp + j] = dS[i * lp + j] = 0.f;
            continue;
        }
        float mx = -INFINITY;
        for (int j = lane; j < S; j += 32) mx = fmaxf(mx, P[i * lp + j]);
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float den = 0.f;
        for (int j = lane; j < S; j += 32) {
            const float p = expf(P[i * lp + j] - mx);
            P[i * lp + j] = p;
            den += p;
        }
        for (int o = 16; o; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
        const float inv = 1.0f / den;
        float delta = 0.f;
        for (int j = lane; j < S; j += 32) {
            const float p = P[i * lp + j] * inv;
            P[i * lp + j] = p;
            delta = fmaf(p, dS[i * lp + j], delta);
        }
        for (int o = 16; o; o >>= 1) delta += __shfl_xor_sync(0xffffffffu, delta, o);
        const bool masked = rowm[i] == 0.0f;
        for (int j = lane; j < kTileSeq; j += 32) {
            const bool in = j < S;
            const float p = in ? P[i * lp + j] : 0.f;
            dS[i * lp + j] = (masked || !in) ? 0.f : to_tf32(p * (dS[i * lp + j] - delta));
            P[i * lp + j] = to_tf32(p);
        }
    }
    __syncthreads();
    // dQ = dS K, dK = dS^T Q, dV = P^T dO: rows m0 .. m0 + 15, all dh columns; C fragment rows g / g + 8, columns 2t, 2t + 1
    float acc[NTD][4];
    auto store = [&](float *dst, float mul) {
#pragma unroll
        for (int nt = 0; nt < NTD; ++nt) {
            const int c = nt * 8 + 2 * t;
            if (m0 + g < S) *reinterpret_cast<float2 *>(dst + out_base + (int64_t)(m0 + g) * out_pitch + c) = make_float2(acc[nt][0] * mul, acc[nt][1] * mul);
            if (m0 + g + 8 < S)
                *reinterpret_cast<float2 *>(dst + out_base + (int64_t)(m0 + g + 8) * out_pitch + c) = make_float2(acc[nt][2] * mul, acc[nt][3] * mul);
            acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
        }
    };
#pragma unroll
    for (int nt = 0; nt < NTD; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
    warp_gemm<NTD>(acc, dS, lp, 1, K, ld, 1, m0, kTileSeq, lane);          // A(m, k) = dS[m][k], B(k, n) = K[k][n]
    store(dq, scale);
    warp_gemm<NTD>(acc, dS, 1, lp, Q, ld, 1, m0, kTileSeq, lane);          // A(m, k) = dS[k][m]
    store(dk, scale);
    warp_gemm<NTD>(acc, P, 1, lp, G, ld, 1, m0, kTileSeq, lane);           // A(m, k) = P[k][m]
    store(dv, 1.0f);
}

// --------------------------------------------------------------------------------------------
// in-batch softmax cross-entropy backward
// --------------------------------------------------------------------------------------------
constexpr int kOwn = 32;     // owner rows per CTA
constexpr int kOth = 64;     // rows of the other side per tile

// OWNER_IS_QUERY: own = query rows i, other = doc rows j, C_ij indexed (own, other);
// else          : own = doc rows j,  other = query rows i, C_ij indexed (other, own).
template <bool OWNER_IS_QUERY, int NV>
__global__ void __launch_bounds__(kThreads)
ce_backward_kernel(const float *__restrict__ own, const float *__restrict__ oth, const float *__restrict__ y,
                   const float *__restrict__ lse, int64_t B, int dim, float scale, float coef, float diag_on,
                   float *__restrict__ grad_own) {
    extern __shared__ float smem[];
    const int ld = dim + 1;
    float *A = smem;                   // [kOwn][dim + 1] owner rows
    float *T = A + kOwn * ld;          // [kOth][dim + 1] tile of the other side
    float *C = T + kOth * ld;          // [kOwn][kOth + 1]
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int64_t own0 = (int64_t)blockIdx.x * kOwn;
    for (int e = tid; e < kOwn * dim; e += kThreads) {
        const int r = e / dim, c = e - r * dim;
        A[r * ld + c] = own0 + r < B ? own[(own0 + r) * dim + c] : 0.f;
    }
    float acc[2][NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[0][v] = acc[1][v] = 0.f;
    const int o0 = ty * 2, o1 = o0 + 1;
    for (int64_t t0 = 0; t0 < B; t0 += kOth) {
        __syncthreads();               // previous tile fully consumed (and A loaded, first time round)
        for (int e = tid; e < kOth * dim; e += kThreads) {
            const int r = e / dim, c = e - r * dim;
            T[r * ld + c] = t0 + r < B ? oth[(t0 + r) * dim + c] : 0.f;
        }
        __syncthreads();
        // S tile: this thread's 2 owner rows x 4 other rows (tx, tx + 16, tx + 32, tx + 48)
        float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        for (int c = 0; c < dim; ++c) {
            const float a0 = A[o0 * ld + c], a1 = A[o1 * ld + c];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float b = T[(tx + 16 * u) * ld + c];
                s[0][u] = fmaf(a0, b, s[0][u]);
                s[1][u] = fmaf(a1, b, s[1][u]);
            }
        }
#pragma unroll
        for (int w = 0; w < 2; ++w)
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t go = own0 + o0 + w, gt = t0 + tx + 16 * u;
                float cv = 0.f;
                if (go < B && gt < B) {
                    const int64_t row = OWNER_IS_QUERY ? go : gt;       // the query index owns lse and y
                    cv = coef * y[row] * (expf(scale * s[w][u] - lse[row]) - (go == gt ? diag_on : 0.0f));
                }
                C[(o0 + w) * (kOth + 1) + tx + 16 * u] = cv;
            }
        __syncthreads();
        // grad_own[o][c] += sum_t C[o][t] * T[t][c]; this thread: rows o0, o1, columns tx + 16 v
        for (int t = 0; t < kOth; ++t) {
            const float c0 = C[o0 * (kOth + 1) + t], c1 = C[o1 * (kOth + 1) + t];
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int col = tx + 16 * v;
                if (col < dim) {
                    const float b = T[t * ld + col];
                    acc[0][v] = fmaf(c0, b, acc[0][v]);
                    acc[1][v] = fmaf(c1, b, acc[1][v]);
                }
            }
        }
    }
#pragma unroll
    for (int w = 0; w < 2; ++w)
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int col = tx + 16 * v;
            const int64_t go = own0 + o0 + w;
            if (col < dim && go < B) grad_own[go * dim + col] = acc[w][v];
        }
}

template <bool OWNER_IS_QUERY>
int launch_ce_pass(const float *own, const float *oth, const float *y, const float *lse, int64_t B, int dim, float scale,
                   float coef, float diag_on, float *grad_own, cudaStream_t st) {
    const size_t smem = sizeof(float) * ((size_t)(kOwn + kOth) * (dim + 1) + (size_t)kOwn * (kOth + 1));
    const unsigned grid = (unsigned)((B + kOwn - 1) / kOwn);
#define RF_CE_LAUNCH(NV)                                                                                              \
    do {                                                                                                              \
        auto fn = ce_backward_kernel<OWNER_IS_QUERY, NV>;                                                              \
        RF_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                    \
        fn<<<grid, kThreads, smem, st>>>(own, oth, y, lse, B, dim, scale, coef, diag_on, grad_own);                   \
    } while (0)
    if (dim <= 128) RF_CE_LAUNCH(8);
    else if (dim <= 256) RF_CE_LAUNCH(16);
    else RF_CE_LAUNCH(32);
#undef RF_CE_LAUNCH
    RF_CUDA(cudaGetLastError());
    return RF_OK;
}


// --------------------------------------------------------------------------------------------
// tensor-core path of the in-batch softmax CE backward: the three contractions (S = Q D^T, dQ = C D, dD = C^T Q) run on
// rf_dense_forward_tc (tcgen05, TF32 operands) over SLABS of query rows; only a [slab x B] piece of the coefficient
// matrix exists at a time (at most 512 MiB + its transpose; the whole 8192 x 8192 matrix at C3's batch, 2048 rows at B = 65536).
//   coef_kernel     C_ij = coef * y_i * (exp(scale * S_ij - lse_i) - [i == j]) in place, plus its transpose (32 x 32 smem tiles)
//   transpose_kernel [R, C] -> [C, R]
//   add_kernel      dD += partial
// --------------------------------------------------------------------------------------------
// 64 x 64 tiles, 128-bit accesses both ways (rows and B are multiples of 4): the first version moved 32 x 32 tiles with scalar
// accesses and ran at 3.5 TB/s (0.227 ms for the 0.8 GB of an 8192 x 8192 slab), the largest piece of the 0.5 ms backward
__global__ void __launch_bounds__(256) ce_coef_kernel(float *__restrict__ S, float *__restrict__ CT, const float *__restrict__ y,
                                                      const float *__restrict__ lse, int rows, int64_t B, int64_t row0, float scale,
                                                      float coef, float diag_on) {
    __shared__ float tile[64][65];
    const int c4 = threadIdx.x & 15, r = threadIdx.x >> 4;             // 16 float4 per tile row, 16 rows per pass
    const int64_t j0 = (int64_t)blockIdx.x * 64;
    const int i0 = blockIdx.y * 64;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int il = r + 16 * k, i = i0 + il;
        const int64_t j = j0 + c4 * 4;
        float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < rows && j < B) {
            const int64_t gi = row0 + i;
            const float cy = coef * y[gi], l = lse[gi];
            float4 s = *reinterpret_cast<const float4 *>(S + (int64_t)i * B + j);
            c.x = cy * (__expf(scale * s.x - l) - (gi == j ? diag_on : 0.f));
            c.y = cy * (__expf(scale * s.y - l) - (gi == j + 1 ? diag_on : 0.f));
            c.z = cy * (__expf(scale * s.z - l) - (gi == j + 2 ? diag_on : 0.f));
            c.w = cy * (__expf(scale * s.w - l) - (gi == j + 3 ? diag_on : 0.f));
            *reinterpret_cast<float4 *>(S + (int64_t)i * B + j) = c;
        }
        tile[il][c4 * 4 + 0] = c.x;
        tile[il][c4 * 4 + 1] = c.y;
        tile[il][c4 * 4 + 2] = c.z;
        tile[il][c4 * 4 + 3] = c.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int jl = r + 16 * k;                                     // column of the tile = row of the transpose
        const int64_t j = j0 + jl;
        const int i = i0 + c4 * 4;
        if (j < B && i < rows)
            *reinterpret_cast<float4 *>(CT + j * rows + i) = make_float4(tile[c4 * 4 + 0][jl], tile[c4 * 4 + 1][jl], tile[c4 * 4 + 2][jl],
                                                                         tile[c4 * 4 + 3][jl]);
    }
}

__global__ void __launch_bounds__(256) transpose_kernel(const float *__restrict__ in, float *__restrict__ out, int64_t R, int64_t Cn) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t r = r0 + ty + 8 * k, c = c0 + tx;
        tile[ty + 8 * k][tx] = (r < R && c < Cn) ? in[r * Cn + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t c = c0 + ty + 8 * k, r = r0 + tx;
        if (r < R && c < Cn) out[c * R + r] = tile[tx][ty + 8 * k];
    }
}

__global__ void __launch_bounds__(256) add_kernel(float *__restrict__ dst, const float *__restrict__ src, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] += src[i];
}

// rows of the coefficient matrix alive at a time: as many as 1 GiB holds (slab + its transpose), at least 512: a slab is the
// M of two of the three GEMMs, and 2048 rows x 256 columns are only 64 tiles for 148 SMs
inline int64_t slab_rows(int64_t batch) {
    int64_t r = ((int64_t)1 << 27) / (batch > 0 ? batch : 1);
    r = r / 128 * 128;
    if (r < 512) r = 512;
    return r < batch ? r : batch;
}

template <int NTD>
int launch_sdpa_backward_mma_t(const float *q, const float *k, const float *v, int64_t pitch, const float *mask, const float *grad_out,
                               int64_t n, int S, float *dq, float *dk, float *dv, int64_t out_pitch, cudaStream_t st) {
    constexpr int dh = 8 * NTD;
    const size_t bytes = sizeof(float) * ((size_t)4 * kTileSeq * (dh + 4) + (size_t)2 * kTileSeq * (kTileSeq + 4) + kTileSeq);
    RF_CUDA(cudaFuncSetAttribute(sdpa_backward_mma_kernel<NTD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    sdpa_backward_mma_kernel<NTD><<<(unsigned)n, kMmaThreads, bytes, st>>>(q, k, v, mask, grad_out, S, dq, dk, dv, pitch, out_pitch);
    RF_CUDA(cudaGetLastError());
    g_launches.fetch_add(1);
    return RF_OK;
}

// head_dim in {32, 64, 96, 128}, pitches multiples of 2 floats (float2 stores) and 16-byte aligned loads
bool sdpa_backward_mma_ok(int dh) { return dh == 32 || dh == 64 || dh == 96 || dh == 128; }

int launch_sdpa_backward_mma(const float *q, const float *k, const float *v, int64_t pitch, const float *mask, const float *grad_out,
                             int64_t n, int S, int dh, float *dq, float *dk, float *dv, int64_t out_pitch, cudaStream_t st) {
    switch (dh) {
        case 32: return launch_sdpa_backward_mma_t<4>(q, k, v, pitch, mask, grad_out, n, S, dq, dk, dv, out_pitch, st);
        case 64: return launch_sdpa_backward_mma_t<8>(q, k, v, pitch, mask, grad_out, n, S, dq, dk, dv, out_pitch, st);
        case 96: return launch_sdpa_backward_mma_t<12>(q, k, v, pitch, mask, grad_out, n, S, dq, dk, dv, out_pitch, st);
        default: return launch_sdpa_backward_mma_t<16>(q, k, v, pitch, mask, grad_out, n, S, dq, dk, dv, out_pitch, st);
    }
}

int launch_sdpa_backward_tiled(const float *q, const float *k, const float *v, int64_t pitch, const float *mask, const float *grad_out,
                               int64_t n, int S, int dh, float *dq, float *dk, float *dv, int64_t out_pitch, cudaStream_t st) {
    const size_t tiled = sizeof(float) * ((size_t)4 * kTileSeq * (dh + 4) + (size_t)2 * kTileSeq * (kTileSeq + 4) + kTileSeq);
    RF_CUDA(cudaFuncSetAttribute(sdpa_backward_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tiled));
    sdpa_backward_tiled_kernel<<<(unsigned)n, kThreads, tiled, st>>>(q, k, v, mask, grad_out, S, dh, dq, dk, dv, pitch, out_pitch);
    RF_CUDA(cudaGetLastError());
    g_launches.fetch_add(1);
    return RF_OK;
}

}  // namespace
}  // namespace rf

using namespace rf;

extern "C" {

int rf_sdpa_backward(const float *d_q, const float *d_k, const float *d_v, const float *d_mask, const float *d_grad_out,
                     int64_t n_batch_heads, int32_t seq_len, int32_t head_dim, float *d_dq, float *d_dk, float *d_dv,
                     void *stream) {
    if (n_batch_heads < 0 || seq_len <= 0 || head_dim <= 0) return set_error(RF_ERR_INVALID, "bad sdpa shape");
    if (seq_len > kBwdMaxSeq || head_dim > kBwdMaxHeadDim)
        return set_error(RF_ERR_UNSUPPORTED, "rf_sdpa_backward handles seq_len <= %d and head_dim <= %d (got %d, %d)", kBwdMaxSeq,
                         kBwdMaxHeadDim, seq_len, head_dim);
    if (n_batch_heads == 0) return RF_OK;
    if (n_batch_heads > INT32_MAX) return set_error(RF_ERR_INVALID, "too many (batch, head) slices");
    if (!d_q || !d_k || !d_v || !d_grad_out || !d_dq || !d_dk || !d_dv) return set_error(RF_ERR_INVALID, "rf_sdpa_backward: NULL buffer");
    const uintptr_t al = reinterpret_cast<uintptr_t>(d_q) | reinterpret_cast<uintptr_t>(d_k) | reinterpret_cast<uintptr_t>(d_v) |
                         reinterpret_cast<uintptr_t>(d_grad_out) | reinterpret_cast<uintptr_t>(d_dq) | reinterpret_cast<uintptr_t>(d_dk) |
                         reinterpret_cast<uintptr_t>(d_dv);
    if (head_dim % 4 == 0 && (al & 15) == 0)          // register-tiled kernel (128-bit rows)
        return launch_sdpa_backward_tiled(d_q, d_k, d_v, head_dim, d_mask, d_grad_out, n_batch_heads, seq_len, head_dim, d_dq, d_dk, d_dv,
                                          head_dim, static_cast<cudaStream_t>(stream));
    const size_t smem = sizeof(float) * ((size_t)4 * seq_len * (head_dim + 1) + (size_t)2 * seq_len * (seq_len + 1) + seq_len);
    RF_CUDA(cudaFuncSetAttribute(sdpa_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sdpa_backward_kernel<<<(unsigned)n_batch_heads, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(
        d_q, d_k, d_v, d_mask, d_grad_out, seq_len, head_dim, d_dq, d_dk, d_dv);
    RF_CUDA(cudaGetLastError());
    g_launches.fetch_add(1);
    return RF_OK;
}

int rf_sdpa_backward_strided(const float *d_q, const float *d_k, const float *d_v, int64_t row_pitch, const float *d_mask,
                             const float *d_grad_out, int64_t n_batch_heads, int32_t seq_len, int32_t head_dim, float *d_dq, float *d_dk,
                             float *d_dv, int64_t grad_row_pitch, void *stream) {
    if (n_batch_heads < 0 || seq_len <= 0 || head_dim <= 0) return set_error(RF_ERR_INVALID, "bad SDPA shape");
    if (seq_len > kBwdMaxSeq || head_dim > kBwdMaxHeadDim)
        return set_error(RF_ERR_UNSUPPORTED, "rf_sdpa_backward handles seq_len <= %d and head_dim <= %d", kBwdMaxSeq, kBwdMaxHeadDim);
    if (n_batch_heads == 0) return RF_OK;
    if (n_batch_heads > INT32_MAX) return set_error(RF_ERR_INVALID, "too many (batch, head) slices");
    if (!d_q || !d_k || !d_v || !d_grad_out || !d_dq || !d_dk || !d_dv) return set_error(RF_ERR_INVALID, "rf_sdpa_backward_strided: NULL buffer");
    const uintptr_t al = reinterpret_cast<uintptr_t>(d_q) | reinterpret_cast<uintptr_t>(d_k) | reinterpret_cast<uintptr_t>(d_v) |
                         reinterpret_cast<uintptr_t>(d_grad_out) | reinterpret_cast<uintptr_t>(d_dq) | reinterpret_cast<uintptr_t>(d_dk) |
                         reinterpret_cast<uintptr_t>(d_dv);
    if (head_dim % 4 || (al & 15) || row_pitch % 4 || grad_row_pitch % 4 || row_pitch < head_dim || grad_row_pitch < head_dim)
        return set_error(RF_ERR_UNSUPPORTED, "strided SDPA backward needs head_dim and both row pitches to be multiples of 4 floats and 16-byte aligned buffers");
    return launch_sdpa_backward_tiled(d_q, d_k, d_v, row_pitch, d_mask, d_grad_out, n_batch_heads, seq_len, head_dim, d_dq, d_dk, d_dv,
                                      grad_row_pitch, static_cast<cudaStream_t>(stream));
}

int rf_sdpa_backward_tc(const float *d_q, const float *d_k, const float *d_v, int64_t row_pitch, const float *d_mask,
                        const float *d_grad_out, int64_t n_batch_heads, int32_t seq_len, int32_t head_dim, float *d_dq, float *d_dk,
                        float *d_dv, int64_t grad_row_pitch, void *stream) {
    if (n_batch_heads < 0 || seq_len <= 0 || head_dim <= 0) return set_error(RF_ERR_INVALID, "bad SDPA shape");
    if (seq_len > kBwdMaxSeq || !sdpa_backward_mma_ok(head_dim))
        return set_error(RF_ERR_UNSUPPORTED, "rf_sdpa_backward_tc handles seq_len <= %d and head_dim in {32, 64, 96, 128} (got %d, %d)",
                         kBwdMaxSeq, seq_len, head_dim);
    if (n_batch_heads == 0) return RF_OK;
    if (n_batch_heads > INT32_MAX) return set_error(RF_ERR_INVALID, "too many (batch, head) slices");
    if (!d_q || !d_k || !d_v || !d_grad_out || !d_dq || !d_dk || !d_dv) return set_error(RF_ERR_INVALID, "rf_sdpa_backward_tc: NULL buffer");
    const uintptr_t al = reinterpret_cast<uintptr_t>(d_q) | reinterpret_cast<uintptr_t>(d_k) | reinterpret_cast<uintptr_t>(d_v) |
                         reinterpret_cast<uintptr_t>(d_grad_out) | reinterpret_cast<uintptr_t>(d_dq) | reinterpret_cast<uintptr_t>(d_dk) |
                         reinterpret_cast<uintptr_t>(d_dv);
    if ((al & 15) || row_pitch % 4 || grad_row_pitch % 4 || row_pitch < head_dim || grad_row_pitch < head_dim)
        return set_error(RF_ERR_UNSUPPORTED, "rf_sdpa_backward_tc needs row pitches that are multiples of 4 floats and 16-byte aligned buffers");
    return launch_sdpa_backward_mma(d_q, d_k, d_v, row_pitch, d_mask, d_grad_out, n_batch_heads, seq_len, head_dim, d_dq, d_dk, d_dv,
                                    grad_row_pitch, static_cast<cudaStream_t>(stream));
}

int rf_inbatch_softmax_ce_backward_block(const float *d_query, const float *d_doc, const float *d_y, const float *d_lse,
                                         int64_t batch, int32_t dim, float scale, float upstream, int positives_on_diagonal,
                                         float *d_grad_query, float *d_grad_doc, void *stream) {
    if (batch < 0 || dim <= 0) return set_error(RF_ERR_INVALID, "bad in-batch shape");
    if (dim > 512) return set_error(RF_ERR_UNSUPPORTED, "rf_inbatch_softmax_ce_backward handles dim <= 512 (got %d)", dim);
    if (batch == 0) return RF_OK;
    if (!d_query || !d_doc || !d_y || !d_lse) return set_error(RF_ERR_INVALID, "rf_inbatch_softmax_ce_backward: NULL input");
    if (!d_grad_query && !d_grad_doc) return set_error(RF_ERR_INVALID, "rf_inbatch_softmax_ce_backward: no output requested");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float coef = upstream * scale / (float)batch;
    const float diag_on = positives_on_diagonal ? 1.0f : 0.0f;
    int launches = 0;
    if (d_grad_query) {
        int rc = launch_ce_pass<true>(d_query, d_doc, d_y, d_lse, batch, dim, scale, coef, diag_on, d_grad_query, st);
        if (rc != RF_OK) return rc;
        ++launches;
    }
    if (d_grad_doc) {
        int rc = launch_ce_pass<false>(d_doc, d_query, d_y, d_lse, batch, dim, scale, coef, diag_on, d_grad_doc, st);
        if (rc != RF_OK) return rc;
        ++launches;
    }
    g_launches.fetch_add(launches);
    return RF_OK;
}

int64_t rf_inbatch_ce_backward_tc_workspace_bytes(int64_t batch, int32_t dim) {
    if (batch <= 0 || dim <= 0) return 0;
    const int64_t R = slab_rows(batch);
    // S / C slab [R, B], its transpose [B, R], doc^T [dim, B], query-slab^T [dim, R], one partial of dD [B, dim]
    return (2 * R * batch + (int64_t)dim * batch + (int64_t)dim * R + batch * (int64_t)dim) * (int64_t)sizeof(float) + 1024;
}

int rf_inbatch_softmax_ce_backward_tc(const float *d_query, const float *d_doc, const float *d_y, const float *d_lse, int64_t batch,
                                      int32_t dim, float scale, float upstream, int positives_on_diagonal, void *d_workspace,
                                      int64_t workspace_bytes, float *d_grad_query, float *d_grad_doc, void *stream) {
    if (batch < 0 || dim <= 0) return set_error(RF_ERR_INVALID, "bad in-batch shape");
    if (batch == 0) return RF_OK;
    if (batch % 4 || dim % 4) return set_error(RF_ERR_UNSUPPORTED, "tensor-core CE backward needs batch %% 4 == 0 and dim %% 4 == 0");
    if (!d_query || !d_doc || !d_y || !d_lse || !d_grad_query || !d_grad_doc)
        return set_error(RF_ERR_INVALID, "rf_inbatch_softmax_ce_backward_tc: NULL buffer (both gradients are produced)");
    if (!d_workspace || workspace_bytes < rf_inbatch_ce_backward_tc_workspace_bytes(batch, dim))
        return set_error(RF_ERR_INVALID, "workspace too small (rf_inbatch_ce_backward_tc_workspace_bytes)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t B = batch, R = slab_rows(B);
    float *ws = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(d_workspace) + 255) & ~(uintptr_t)255);
    float *slab = ws, *slab_t = slab + R * B, *doc_t = slab_t + R * B, *q_t = doc_t + (int64_t)dim * B, *part = q_t + (int64_t)dim * R;
    const float coef = upstream * scale / (float)B;
    const dim3 tb(256);
    transpose_kernel<<<dim3((unsigned)((dim + 31) / 32), (unsigned)((B + 31) / 32)), tb, 0, st>>>(d_doc, doc_t, B, dim);   // [B, dim] -> [dim, B]
    int launches = 1;
    for (int64_t r0 = 0; r0 < B; r0 += R) {
        const int64_t rows = B - r0 < R ? B - r0 : R;
        // S slab = Q[r0 : r0 + rows] . D^T
        int rc = rf_dense_forward_tc(d_query + r0 * dim, rows, dim, dim, d_doc, nullptr, (int32_t)B, RF_ACT_NONE, 0, slab, B, stream);
        if (rc != RF_OK) return rc;
        ce_coef_kernel<<<dim3((unsigned)((B + 63) / 64), (unsigned)((rows + 63) / 64)), tb, 0, st>>>(
            slab, slab_t, d_y, d_lse, (int)rows, B, r0, scale, coef, positives_on_diagonal ? 1.0f : 0.0f);
        // dQ slab = C . D        (weight_t = D^T [dim, B])
        rc = rf_dense_forward_tc(slab, rows, (int32_t)B, B, doc_t, nullptr, dim, RF_ACT_NONE, 0, d_grad_query + r0 * dim, dim, stream);
        if (rc != RF_OK) return rc;
        // dD (+)= C^T . Q slab   (x = C^T [B, rows], weight_t = Q_slab^T [dim, rows])
        transpose_kernel<<<dim3((unsigned)((dim + 31) / 32), (unsigned)((rows + 31) / 32)), tb, 0, st>>>(d_query + r0 * dim, q_t, rows, dim);
        float *dst = r0 == 0 ? d_grad_doc : part;
        rc = rf_dense_forward_tc(slab_t, B, (int32_t)rows, rows, q_t, nullptr, dim, RF_ACT_NONE, 0, dst, dim, stream);
        if (rc != RF_OK) return rc;
        launches += 2;
        if (r0 != 0) {
            add_kernel<<<(unsigned)std::min<int64_t>((B * dim + 255) / 256, 148 * 8), tb, 0, st>>>(d_grad_doc, part, B * dim);
            ++launches;
        }
    }
    RF_CUDA(cudaGetLastError());
    g_launches.fetch_add(launches);
    return RF_OK;
}

int rf_inbatch_softmax_ce_backward(const float *d_query, const float *d_doc, const float *d_y, const float *d_lse, int64_t batch,
                                   int32_t dim, float scale, float upstream, float *d_grad_query, float *d_grad_doc,
                                   void *stream) {
    return rf_inbatch_softmax_ce_backward_block(d_query, d_doc, d_y, d_lse, batch, dim, scale, upstream, 1, d_grad_query, d_grad_doc,
                                                stream);
}

}  // extern "C"
