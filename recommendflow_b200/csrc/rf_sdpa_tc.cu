// rf_sdpa_tc.cu -- scaled_dot_product_attention on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces the two BatchMatMul + Softmax + Select of /root/reference/backend/layers/layer_utils.py:4-24
// for the behaviour-sequence shape (S <= 64 keys, head_dim 32, 64 or 96 -- what fits one SM's smem).
//
// One CTA works on PAIRS of (batch x head) sequences: the two sequences fill the 128 rows of one
// UMMA tile (rows 0-63 / 64-127; S is padded to 64 by simply letting the 64-row TMA box run into
// the next sequence -- those rows/keys are masked out below).  Per pair:
//   TMA     Q, K, V of both sequences -> smem (128-byte swizzle), head_dim in blocks of 32 floats
//   GEMM 1  logits[128 x 128] = Qpair . Kpair^T      tcgen05.mma kind::tf32, K-major A and B
//           (only the two diagonal 64 x 64 blocks are meaningful)
//   softmax thread = query row = TMEM lane: tcgen05.ld its 64 logits, scale 1/sqrt(dh), the
//           reference's QUERY-row mask (mask == 0 -> every logit = -4294967295 -> uniform), keys
//           >= S dropped, exp / sum in registers, probabilities rounded to TF32 and written into a
//           K-major swizzled smem tile P[128 x 128] that is block-diagonal (off-diagonal blocks
//           stay zero), fence.proxy.async
//   GEMM 2  out[128 x dh] = P . Vpair               A = P (K-major), B = V (MN-major: V is [keys][dh])
//   store   thread = row: tcgen05.ld dh columns -> global
// fp32 operands are read by the tensor core as TF32 (truncated), accumulation is fp32.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include <atomic>

#include "../../include/rf_b200.h"
#include "rf_common.h"

namespace rf {

extern std::atomic<int64_t> g_launches;

namespace sdpa_tc {

constexpr int kRows = 128;            // UMMA M: two sequences of up to 64 rows
constexpr int kSeqPad = 64;
constexpr int kKB = 32;               // fp32 elements per 128-byte swizzle row
constexpr int kBlkBytes = kRows * 128;   // one [128 rows x 128 B] swizzled block = 16 KiB

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (spin > (1u << 26)) __trap();            // fail instead of hanging the GPU
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 128-byte-swizzled operand descriptors (version 1).  K-major: rows at 128 B, 8-row atoms 1024 B apart
// (SBO), LBO unused (=1), layout SWIZZLE_128B.  MN-major TF32 operands only exist in the 32-byte-base
// flavour (SWIZZLE_128B_BASE32B = TMA's 128B_ATOM_32B: 32-byte chunks XOR-ed with the row index mod 4):
// 32 contiguous MN elements per 128-B row, successive K at +128 B, 4-K atoms 512 B apart (SBO), next
// 32 MN elements `lbo_bytes` further.
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t addr) {
    return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t addr, uint32_t lbo_bytes) {
    return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}

__device__ __forceinline__ float ex2(float x) {       // MUFU.EX2: 2^x, ex2(-inf) = +0
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

struct Params {
    const float *mask;     // [n_seq, S] or NULL
    float *out;            // [n_seq, S, dh]
    int n_seq, S, dh;
    float inv_sqrt_dk;
};

// ------------------------------------------------------------------------------------------------------------
// Round-2 kernel: the same arithmetic, software-pipelined across pairs, softmax on 8 warps.
// Round 1 ran load -> GEMM 1 -> softmax -> GEMM 2 -> store strictly in series with one CTA per SM (ncu: 33 % of the
// op's HBM floor).  A first pipelined version (loads and GEMM 1 of pair n+1 under the softmax of pair n) only gained
// 16 %: its profile showed the 4 softmax warps -- one per scheduler, ~1500 dependent instructions per row of 64
// logits -- pacing the kernel at ~10 k cycles per pair (profiles/r2a_ncu_sdpa_pipelined_summary.csv).  Now
//   warp 8 (one lane)  TMA of Q, K and GEMM 1.  Q/K smem is refilled for pair n+1 the moment GEMM 1 of pair n
//                      retires, and GEMM 1 writes alternating TMEM logit buffers;
//   warp 9 (one lane)  TMA of V (double-buffered when it fits) and GEMM 2;
//   warps 0-7          softmax: TWO threads per query row (warps w and w+4 own the same TMEM lanes), 32 logits each:
//                      local max / exp2 / sum, one (max, sum) exchange through shared memory, then P = e * factor;
//                      2 warps per scheduler and ~400 instructions per thread.
// TMEM: logits[0] cols 0-127, logits[1] cols 128-255, output cols 256-(256+dh).  smem (dh = 64): Q 32 + K 32 +
// V 2 x 32 + P 64 = 192 KiB, one CTA per SM.
// ------------------------------------------------------------------------------------------------------------
constexpr int kThreads2 = 320;
constexpr int kSoftThreads = 256;
constexpr int kTmemCols2 = 512;

__global__ void __launch_bounds__(kThreads2, 1)
sdpa_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
               const __grid_constant__ CUtensorMap map_v, Params p, int n_vbuf, int n_qkbuf) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int n_db = p.dh / kKB;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t q_s = base;                                       // n_qkbuf buffers of n_db blocks
    const uint32_t k_s = q_s + n_qkbuf * n_db * kBlkBytes;           // n_qkbuf buffers of n_db blocks
    const uint32_t v_s = k_s + n_qkbuf * n_db * kBlkBytes;           // n_vbuf buffers of n_db blocks
    const uint32_t p_s = v_s + n_vbuf * n_db * kBlkBytes;            // 4 blocks
    const uint32_t xch_s = p_s + 4 * kBlkBytes;                      // float2 [2][128]: (local max, local sum) per half row
    const uint32_t bars = xch_s + 2 * 2 * 128 * 8;                   // two exchange buffers (pairs n, n+1)
    const uint32_t qk_full0 = bars, mma1_done0 = bars + 16, s_free0 = bars + 32, v_full0 = bars + 48, p_ready = bars + 64,
                   mma2_done = bars + 72, tmem_slot = bars + 80;
    uint8_t *smem_gen = smem_raw + (base - smem_u32(smem_raw));
    float2 *xch = reinterpret_cast<float2 *>(smem_gen + (xch_s - base));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    {   // the carve-up must fit the launch's dynamic shared memory (the host sizes it without alignment slack)
        uint32_t dyn;
        asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
        if (tmem_slot + 4 > smem_u32(smem_raw) + dyn) __trap();
    }

    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(qk_full0 + 8 * b, 1);
            mbar_init(mma1_done0 + 8 * b, 1);
            mbar_init(s_free0 + 8 * b, kSoftThreads);
            mbar_init(v_full0 + 8 * b, 1);
        }
        mbar_init(p_ready, kSoftThreads);
        mbar_init(mma2_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        uint4 *pz = reinterpret_cast<uint4 *>(smem_gen + (p_s - base));
        for (int i = threadIdx.x; i < 4 * kBlkBytes / 16; i += kThreads2) pz[i] = make_uint4(0, 0, 0, 0);
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols2) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int n_pairs = (p.n_seq + 1) / 2;
    const uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(p.dh >> 3) << 17) |
                            ((uint32_t)(kRows >> 4) << 24);
    const uint32_t qk_bytes = 2u * n_db * kBlkBytes, v_bytes = (uint32_t)n_db * kBlkBytes;
    const int stride = gridDim.x;

    if (warp == 8) {
        // ===== Q / K loader + GEMM 1 =====
        if (lane == 0) {
            auto load_qk = [&](int pair, uint32_t qb) {
                mbar_expect_tx(qk_full0 + 8 * qb, qk_bytes);
                for (int db = 0; db < n_db; ++db)
                    for (int h = 0; h < 2; ++h) {
                        const int row = (2 * pair + h) * p.S;
                        const uint32_t off = (qb * n_db + db) * kBlkBytes + h * (kSeqPad * 128);
                        tma_load_2d(q_s + off, &map_q, qk_full0 + 8 * qb, db * kKB, row);
                        tma_load_2d(k_s + off, &map_k, qk_full0 + 8 * qb, db * kKB, row);
                    }
            };
            // with two Q/K buffers the loads of pair n+2 start when GEMM 1 of pair n retires: a full pair ahead
            for (int j = 0; j < n_qkbuf; ++j)
                if ((int)blockIdx.x + j * stride < n_pairs) load_qk(blockIdx.x + j * stride, (uint32_t)j);
            uint32_t n = 0;
            for (int pair = blockIdx.x; pair < n_pairs; pair += stride, ++n) {
                const uint32_t b = n & 1u, use = (n >> 1) & 1u;
                const uint32_t qb = n_qkbuf == 2 ? b : 0u, quse = n_qkbuf == 2 ? use : (n & 1u);
                mbar_wait(qk_full0 + 8 * qb, quse);
                mbar_wait(s_free0 + 8 * b, use ^ 1u);               // softmax of pair n-2 has read this logit buffer
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int db = 0; db < n_db; ++db) {
                    const uint64_t a = desc_kmajor(q_s + (qb * n_db + db) * kBlkBytes), bd = desc_kmajor(k_s + (qb * n_db + db) * kBlkBytes);
#pragma unroll
                    for (int k = 0; k < kKB / 8; ++k)
                        umma_tf32(tmem_base + b * 128u, a + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc1, (db | k) ? 1u : 0u);
                }
                umma_commit(mma1_done0 + 8 * b);
                const int nxt = pair + n_qkbuf * stride;
                if (nxt < n_pairs) {
                    mbar_wait(mma1_done0 + 8 * b, use);             // this Q / K buffer is free again
                    load_qk(nxt, qb);
                }
            }
        }
        __syncwarp();
    } else if (warp == 9) {
        // ===== V loader + GEMM 2 =====
        if (lane == 0) {
            auto load_v = [&](int pair, int vb) {
                mbar_expect_tx(v_full0 + 8 * vb, v_bytes);
                for (int db = 0; db < n_db; ++db)
                    for (int h = 0; h < 2; ++h)
                        tma_load_2d(v_s + vb * n_db * kBlkBytes + db * kBlkBytes + h * (kSeqPad * 128), &map_v, v_full0 + 8 * vb,
                                    db * kKB, (2 * pair + h) * p.S);
            };
            for (int j = 0; j < n_vbuf; ++j)
                if ((int)blockIdx.x + j * stride < n_pairs) load_v(blockIdx.x + j * stride, j);
            uint32_t n = 0;
            for (int pair = blockIdx.x; pair < n_pairs; pair += stride, ++n) {
                const uint32_t vb = n_vbuf == 2 ? (n & 1u) : 0u;
                const uint32_t vuse = n_vbuf == 2 ? ((n >> 1) & 1u) : (n & 1u);
                mbar_wait(p_ready, n & 1u);
                mbar_wait(v_full0 + 8 * vb, vuse);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t vbase = v_s + vb * n_db * kBlkBytes;
                for (int kb = 0; kb < 4; ++kb) {
                    const uint64_t a = desc_kmajor(p_s + kb * kBlkBytes);
#pragma unroll
                    for (int k = 0; k < kKB / 8; ++k) {
                        const uint64_t bd = desc_mnmajor(vbase + (uint32_t)(kb * 4 + k) * 1024u, (uint32_t)kBlkBytes);
                        umma_tf32(tmem_base + 256u, a + (uint64_t)(2 * k), bd, idesc2, (kb | k) ? 1u : 0u);
                    }
                }
                umma_commit(mma2_done);
                const int nxt = pair + n_vbuf * stride;
                if (nxt < n_pairs) {
                    mbar_wait(mma2_done, n & 1u);                   // this V buffer has been consumed
                    load_v(nxt, (int)vb);
                }
            }
        }
        __syncwarp();
    } else {
        // ===== softmax: two threads per row of the pair tile (warps w and w + 4 share TMEM lanes 32 (w % 4) ..) =====
        const int quarter = warp & 3, ch = warp >> 2;                // ch: which 32 of the row's 64 logits
        const int r = quarter * 32 + lane;                          // row of the pair tile = TMEM lane
        const int half = r >> 6, i = r & 63;                        // sequence of the pair, query index
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const float c2 = p.inv_sqrt_dk * 1.4426950408889634f;       // logits enter exp2 pre-multiplied by log2(e) / sqrt(dh)
        const int nv = min(32, max(0, p.S - ch * 32));              // valid keys among this thread's 32
        const int n32 = p.dh >> 5, o_lo = ch == 0 ? 0 : (n32 + 1) / 2, o_hi = ch == 0 ? (n32 + 1) / 2 : n32;
        // One pair's softmax as two steps, software-pipelined across pairs so that the exponentials of pair n+1 run while
        // the tensor core does GEMM 2 of pair n (the softmax warps used to idle through it):
        //   stage(n):  wait logits(n), read them, exp2 / partial sums / exchange with the partner thread -> x[], fac
        //   emit(n):   P(n) = x * fac into the smem tile, release GEMM 2(n)
        //   loop:      emit(n); stage(n+1); wait GEMM 2(n); store out(n)
        float x[32];
        float fac = 0.f;
        bool row_ok = false;
        auto stage = [&](int pair, uint32_t n) {
            const uint32_t b = n & 1u, use = (n >> 1) & 1u;
            const int seq = 2 * pair + half;
            mbar_wait(mma1_done0 + 8 * b, use);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            tmem_ld32(lane_addr + b * 128u + (uint32_t)(half * 64 + ch * 32), x);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(s_free0 + 8 * b);                           // GEMM 1 of pair n+2 may overwrite this buffer
            row_ok = seq < p.n_seq && i < p.S;
            // the reference's QUERY-row mask fills the whole row with one constant: uniform attention over the S keys
            const bool masked = row_ok && p.mask && p.mask[(size_t)seq * p.S + i] == 0.f;
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float l = masked ? 0.f : x[j];
                x[j] = j < nv ? l : -INFINITY;                      // padded keys: exp2(-inf) = 0
                mx = fmaxf(mx, x[j]);
            }
            const float mc = mx == -INFINITY ? 0.f : mx * c2;
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                x[j] = ex2(fmaf(x[j], c2, -mc));
                sum += x[j];
            }
            float2 *xb = xch + (n & 1u) * 256;                      // double-buffered: a fast thread never laps its partner
            xb[ch * 128 + r] = make_float2(mx, sum);
            asm volatile("bar.sync 1, %0;" ::"n"(kSoftThreads) : "memory");
            const float2 other = xb[(ch ^ 1) * 128 + r];
            const float big = fmaxf(mx, other.x);                   // finite: every row has >= 1 valid key in one half
            const float f_me = mx == -INFINITY ? 0.f : ex2((mx - big) * c2);
            const float f_ot = other.x == -INFINITY ? 0.f : ex2((other.x - big) * c2);
            const float total = sum * f_me + other.y * f_ot;
            fac = row_ok ? f_me / total : 0.f;                      // rows that are padding produce zeros
        };
        uint32_t n = 0;
        if ((int)blockIdx.x < n_pairs) stage(blockIdx.x, 0);
        for (int pair = blockIdx.x; pair < n_pairs; pair += stride, ++n) {
            const int seq = 2 * pair + half;
            const bool ok_n = row_ok;
            // emit(n): GEMM 2 of pair n-1 has been waited for below, so the P tile is free
            uint8_t *prow = smem_gen + (p_s - base) + (r >> 3) * 1024 + (r & 7) * 128 + (half * 2 + ch) * kBlkBytes;
#pragma unroll
            for (int c = 0; c < 8; ++c) {                             // 8 chunks of 4 keys
                uint4 w;
                w.x = to_tf32(x[c * 4 + 0] * fac);
                w.y = to_tf32(x[c * 4 + 1] * fac);
                w.z = to_tf32(x[c * 4 + 2] * fac);
                w.w = to_tf32(x[c * 4 + 3] * fac);
                *reinterpret_cast<uint4 *>(prow + ((c ^ (r & 7)) << 4)) = w;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(p_ready);
            if (pair + stride < n_pairs) stage(pair + stride, n + 1);   // under GEMM 2 of pair n
            mbar_wait(mma2_done, n & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float *orow = p.out + ((size_t)seq * p.S + i) * p.dh;
            for (int cb = o_lo; cb < o_hi; ++cb) {
                float o[32];
                tmem_ld32(lane_addr + 256u + (uint32_t)(cb * 32), o);
                if (ok_n) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4 *>(orow + cb * 32 + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols2) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(CUtensorMap *map, const float *ptr, int64_t rows, int64_t cols, int64_t ld, CUtensorMapSwizzle swizzle) {
    static EncodeTiledFn fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) return (EncodeTiledFn) nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    if (!fn) return set_error(RF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kKB, (cuuint32_t)kSeqPad};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(RF_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return RF_OK;
}

}  // namespace sdpa_tc

bool sdpa_tc_supported(int64_t n_seq, int S, int dh, const float *q, const float *k, const float *v, const float *out) {
    const uintptr_t al = reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                         reinterpret_cast<uintptr_t>(out);
    return n_seq > 0 && S >= 1 && S <= sdpa_tc::kSeqPad && dh >= 32 && dh <= 96 && dh % 32 == 0 && (al & 15) == 0 &&
           n_seq * (int64_t)S < INT32_MAX;
}

int launch_sdpa_tc(const float *q, const float *k, const float *v, const float *mask, int64_t n_seq, int S, int dh, float *out,
                   cudaStream_t st, int64_t ld) {
    using namespace sdpa_tc;
    CUtensorMap mq, mk, mv;
    int rc;
    if (ld <= 0) ld = dh;                      // rows of q / k / v are `ld` floats apart (contiguous by default)
    if ((rc = make_map(&mq, q, n_seq * S, dh, ld, CU_TENSOR_MAP_SWIZZLE_128B)) != RF_OK) return rc;
    if ((rc = make_map(&mk, k, n_seq * S, dh, ld, CU_TENSOR_MAP_SWIZZLE_128B)) != RF_OK) return rc;
    if ((rc = make_map(&mv, v, n_seq * S, dh, ld, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) != RF_OK) return rc;   // MN-major TF32 operand
    const int n_db = dh / kKB;
    int dev = 0, sms = 0;
    RF_CUDA(cudaGetDevice(&dev));
    RF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int n_pairs = (int)((n_seq + 1) / 2);
    Params p{mask, out, (int)n_seq, S, dh, 1.0f / sqrtf((float)dh)};
    {
        // buffering that fits 227 KiB: dh = 32: Q/K x 2, V x 2 (160 KiB); dh = 64: Q/K x 1, V x 2 (192 KiB; Q/K x 2 with
        // V x 1 measured no faster: 0.114 vs 0.110 ms); dh = 96: x 1, x 1
        const int n_qkbuf = n_db <= 1 ? 2 : 1, n_vbuf = n_db <= 2 ? 2 : 1;
        // no alignment slack: the dynamic shared memory of a kernel without static shared memory starts 1 KiB aligned
        // (the kernel traps if its carve-up does not fit)
        const size_t smem = (size_t)((2 * n_qkbuf + n_vbuf) * n_db + 4) * kBlkBytes + 4096 + 128;
        RF_CUDA(cudaFuncSetAttribute(sdpa_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int grid = n_pairs < sms ? n_pairs : sms;
        sdpa_tc_kernel<<<grid, kThreads2, smem, st>>>(mq, mk, mv, p, n_vbuf, n_qkbuf);
    }
    RF_CUDA(cudaGetLastError());
    g_launches.fetch_add(1);
    return RF_OK;
}

}  // namespace rf
