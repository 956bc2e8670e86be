// rf_logits_tc.cu -- the B x B in-batch logits S = query . doc^T on the 5th-gen tensor cores
// (tcgen05 + TMEM + TMA, sm_100a), with the row reductions of the two-tower losses fused into the
// epilogue so that S never exists in memory.
//
// Replaces tf.matmul(query, tf.transpose(doc)) + tf.exp + tf.reduce_sum + tf.linalg.diag_part of
// /root/reference/backend/lossess/match_losses.py:160-165 (and the same contraction of :119-226).
//
//   operands   fp32 in HBM, read as TF32 (kind::tf32: the tensor core uses the top 19 bits; this is
//              what TensorFlow itself does for fp32 matmuls on Ampere+ GPUs), fp32 accumulate.
//   tile       128 (query rows) x 256 (docs) x 32 (K, one 128-byte swizzle atom) per stage; for
//              dim <= 256 the CTA's query tile is loaded ONCE and stays resident, only doc tiles stream
//              (3-stage TMA -> smem ring); 4 x tcgen05.mma (K = 8) per stage; TWO 128-lane x 256-column
//              TMEM accumulators, so the MMAs of tile i+1 run under the epilogue of tile i.
//   warps      0: TMA producer (one elected lane)   1: TMEM alloc + MMA issue (one elected lane)
//              2-5: epilogue -- each thread owns one query row (= one TMEM lane), pulls 32 columns at
//              a time with tcgen05.ld and keeps running (max, sum exp, hinge, max-off-diagonal).
//   grid       (row tiles, column groups); a CTA walks `tiles_per_cta` column tiles and writes one
//              partial RowStat per row; rf_dense.cu's finalize kernel merges the column groups.
//
// Every mbarrier wait is bounded and traps instead of hanging if the pipeline is ever wedged.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <atomic>

#include "../../include/rf_b200.h"
#include "rf_common.h"

namespace rf {

extern std::atomic<int64_t> g_launches;

struct RowStat {
    float m, l, hinge, maxoff;
};

constexpr int kBM = 128, kBN = 256, kBK = 32;       // tile; kBK fp32 = 128 bytes = one SW128 atom row
constexpr int kUmmaK = 8;                           // tf32: 32 bytes per MMA along K
constexpr int kABytes = kBM * kBK * 4;              // 16 KiB
constexpr int kBBytes = kBN * kBK * 4;              // 32 KiB
// A_RES = true (dim <= 256): the CTA's query tile stays in shared memory for all of its column tiles
// (8 x 16 KiB) and only doc tiles stream through a 3-stage ring -- L2 -> SM traffic per tile drops from
// 384 KiB to 256 KiB, which is what bounds this fp32-operand kernel.  A_RES = false: both stream, 4 stages.
constexpr int kMaxResidentKB = 8;
constexpr int kTmemCols = 512;                      // two 128 x 256 fp32 accumulators: MMA(i+1) overlaps epilogue(i)
constexpr int kTcThreads = 192;
template <bool A_RES> struct TcCfg {
    static constexpr int kStages = A_RES ? 3 : 4;
    static constexpr int kStageBytes = A_RES ? kBBytes : kABytes + kBBytes;
    static constexpr size_t smem_bytes(int n_kb) {
        return (size_t)(A_RES ? n_kb * kABytes : 0) + (size_t)kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
    }
};

// ---- PTX helpers --------------------------------------------------------------------------------
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
        if (spin > (1u << 26)) __trap();            // a wedged pipeline must fail, not hang the GPU
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

// K-major operand tile in shared memory, 128-byte swizzle: rows at a 128-byte pitch, 8-row groups
// 1024 bytes apart (SBO), LBO = 1 (unused for swizzled K-major), descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);        // start address  [0,14)
    d |= (uint64_t)1 << 16;                              // leading byte offset (>>4) [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset (>>4)  [32,46)
    d |= (uint64_t)1 << 46;                              // version                   [46,48)
    d |= (uint64_t)2 << 61;                              // layout: SWIZZLE_128B      [61,64)
    return d;
}

// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 256
constexpr uint32_t kInstrDesc = (1u << 4)                 // c_format = F32
                                | (2u << 7)               // a_format = TF32
                                | (2u << 10)              // b_format = TF32
                                | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);

template <bool FULL_STATS, bool A_RES>
__global__ void __launch_bounds__(kTcThreads, 1)
logits_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_d,
                 const float *__restrict__ diag, const float *__restrict__ colw, int B, int Dt, float scale, float margin,
                 int n_tiles, int tiles_per_cta, RowStat *__restrict__ part) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int kStages = TcCfg<A_RES>::kStages;
    constexpr int kStageBytes = TcCfg<A_RES>::kStageBytes;
    const int n_kb = (Dt + kBK - 1) / kBK;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;          // SW128 atoms need 1024-byte alignment
    const uint32_t a_res = smem_base;                                           // A_RES: n_kb resident query blocks
    const uint32_t ring = smem_base + (A_RES ? n_kb * kABytes : 0);
    const uint32_t bars = ring + kStages * kStageBytes;
    const uint32_t full0 = bars, empty0 = bars + 8 * kStages;                   // full[s], empty[s]
    const uint32_t tmem_full0 = bars + 16 * kStages, tmem_empty0 = tmem_full0 + 16;   // [2] each
    const uint32_t a_full = tmem_empty0 + 16, tmem_slot = a_full + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tile = blockIdx.x;
    const int tile0 = blockIdx.y * tiles_per_cta;
    const int tile1 = min(n_tiles, tile0 + tiles_per_cta);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full0 + 8 * a, 1);
            mbar_init(tmem_empty0 + 8 * a, 4);     // one arrive per epilogue warp
        }
        mbar_init(a_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {                                // whole warp: allocate the accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            if (A_RES && tile0 < tile1) {
                mbar_expect_tx(a_full, (uint32_t)(n_kb * kABytes));
                for (int kb = 0; kb < n_kb; ++kb) tma_load_2d(a_res + kb * kABytes, &map_q, a_full, kb * kBK, m_tile * kBM);
            }
            for (int tile = tile0; tile < tile1; ++tile) {
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    const uint32_t dst = ring + stage * kStageBytes;
                    mbar_expect_tx(full0 + 8 * stage, kStageBytes);
                    if (!A_RES) tma_load_2d(dst, &map_q, full0 + 8 * stage, kb * kBK, m_tile * kBM);
                    tma_load_2d(dst + (A_RES ? 0 : kABytes), &map_d, full0 + 8 * stage, kb * kBK, tile * kBN);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            int local = 0;
            if (A_RES && tile0 < tile1) mbar_wait(a_full, 0);
            for (int tile = tile0; tile < tile1; ++tile, ++local) {
                const uint32_t acc = (uint32_t)(local & 1);
                mbar_wait(tmem_empty0 + 8 * acc, ((uint32_t)(local >> 1) & 1u) ^ 1u);   // epilogue drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(full0 + 8 * stage, phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st_addr = ring + stage * kStageBytes;
                    const uint64_t adesc = umma_desc_sw128(A_RES ? a_res + kb * kABytes : st_addr);
                    const uint64_t bdesc = umma_desc_sw128(A_RES ? st_addr : st_addr + kABytes);
#pragma unroll
                    for (int k = 0; k < kBK / kUmmaK; ++k) {
                        // step 32 bytes along K inside the swizzle atom: +2 in the (>>4) address field
                        umma_tf32(tmem_base + acc * (uint32_t)kBN, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), kInstrDesc,
                                  (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(empty0 + 8 * stage);            // frees the smem stage when these MMAs retire
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(tmem_full0 + 8 * acc);              // accumulator complete
            }
        }
    } else {
        // ===== epilogue: thread <-> query row (TMEM lane); warp w may touch lanes 32*(w % 4) .. +31 =====
        const int quarter = warp & 3;
        const int row_in_tile = quarter * 32 + lane;
        const int row = m_tile * kBM + row_in_tile;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const float rdiag = (FULL_STATS && row < B) ? diag[row] : 0.f;
        float rm = -INFINITY, rl = 0.f, rh = 0.f, rx = -INFINITY;
        int local = 0;
        for (int tile = tile0; tile < tile1; ++tile, ++local) {
            const uint32_t acc = (uint32_t)(local & 1);
            mbar_wait(tmem_full0 + 8 * acc, (uint32_t)(local >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int col_tile0 = tile * kBN;
#pragma unroll 1
            for (int c0 = 0; c0 < kBN; c0 += 32) {
                float v[32];
                tmem_ld32(lane_addr + acc * (uint32_t)kBN + (uint32_t)c0, v);
                const int col0 = col_tile0 + c0;
                if (col0 >= B) continue;
                const int n_valid = min(32, B - col0);
                float cmax = -INFINITY;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float x = scale * v[j];
                    if (j < n_valid) cmax = fmaxf(cmax, x);
                }
                if (cmax > rm) {
                    rl *= __expf(rm - cmax);
                    rm = cmax;
                }
                float add = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j < n_valid) add += __expf(scale * v[j] - rm);
                rl += add;
                if (FULL_STATS) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (j < n_valid) {
                            const int c = col0 + j;
                            const float h = v[j] - rdiag + margin;
                            rh += fminf(fmaxf(h, 0.f), 1e14f) * (colw ? __ldg(colw + c) : 1.f);
                            rx = fmaxf(rx, c == row ? 0.f : v[j]);
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty0 + 8 * acc);
        }
        if (row < B) part[(size_t)blockIdx.y * B + row] = RowStat{rm, rl, rh, rx};
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// =================================================================================================
// bf16-operand variant (round 2).  The TF32 kernel above is bound by L2 -> SM operand traffic, not by the
// tensor pipe: every 128 x 256 tile streams 256 KiB of fp32 doc rows for 16.8 MFLOP -- 537 MB per loss at
// B = 8192, ~5 TB/s of L2 bandwidth at the measured 0.147 ms (tensor pipe 21-24 % active).  Here
//   * operands are rounded to bf16 once (2 bytes / element, kind::f16, UMMA K = 16, fp32 accumulate),
//   * a CTA owns 256 query rows (two 128-row UMMA tiles, resident in shared memory) and streams 128-doc
//     tiles, so every doc tile fetched from L2 feeds two MMAs: 4x less L2 traffic per flop than above,
//   * FOUR 128 x 128 fp32 accumulators in TMEM = (2 row tiles) x (2 buffers): the MMAs of doc tile i+1 run
//     under the epilogue of doc tile i; 8 epilogue warps, thread = query row.
// The diagonal S_ii stays the exact fp32 dot product.  Tolerance vs the float64 oracle: bf16 rounding is 2^-9
// relative per operand; for l2-normalised embeddings |S_ij| error <~ 2^-8, times the temperature inside the
// exponent (tests/test_dense_gpu.py states the bounds).
// =================================================================================================
namespace bf {

constexpr int kRowsCta = 256;                       // query rows per CTA (2 UMMA M = 128 tiles)
constexpr int kDocs = 128;                          // docs per step (UMMA N)
constexpr int kK = 64;                              // bf16 elements per 128-byte swizzle row
constexpr int kUK = 16;                             // UMMA K for 16-bit operands
constexpr int kABlk = 128 * kK * 2;                 // one [128 rows x 64 k] block: 16 KiB
constexpr int kBBlk = kDocs * kK * 2;               // one [128 docs x 64 k] block: 16 KiB
constexpr int kMaxKB = 4;                           // dim <= 256
constexpr int kStagesB = 6;                         // doc k-blocks in flight
constexpr int kThreadsB = 64 + 256;                 // warp 0: TMA, warp 1: MMA, warps 2-9: epilogue
constexpr size_t smem_bytes(int n_kb) { return (size_t)2 * n_kb * kABlk + (size_t)kStagesB * kBBlk + 1024 + 256; }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// kind::f16, bf16 x bf16 -> fp32, A and B K-major, M = 128, N = 128
constexpr uint32_t kIdescBf = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kDocs >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

template <bool FULL_STATS>
__global__ void __launch_bounds__(kThreadsB, 1)
logits_bf16_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_d,
                   const float *__restrict__ diag, const float *__restrict__ colw, int B, int Dt, float scale, float margin,
                   int n_tiles, int tiles_per_cta, RowStat *__restrict__ part) {
    extern __shared__ uint8_t smem_raw[];
    const int n_kb = (Dt + kK - 1) / kK;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_res = smem_base;                                   // [2 row tiles][n_kb] blocks of 16 KiB
    const uint32_t ring = a_res + 2 * n_kb * kABlk;
    const uint32_t bars = ring + kStagesB * kBBlk;
    const uint32_t full0 = bars, empty0 = bars + 8 * kStagesB;
    const uint32_t tmem_full0 = bars + 16 * kStagesB, tmem_empty0 = tmem_full0 + 16;
    const uint32_t a_full = tmem_empty0 + 16, tmem_slot = a_full + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_blk = blockIdx.x;                                       // 256-row block of queries
    const int tile0 = blockIdx.y * tiles_per_cta;
    const int tile1 = min(n_tiles, tile0 + tiles_per_cta);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStagesB; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full0 + 8 * a, 1);
            mbar_init(tmem_empty0 + 8 * a, 8);     // one arrive per epilogue warp
        }
        mbar_init(a_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    // TMEM columns: accumulator (row tile t, buffer b) at column (b * 2 + t) * 128

    if (warp == 0) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            if (tile0 < tile1) {
                mbar_expect_tx(a_full, (uint32_t)(2 * n_kb * kABlk));
                for (int t = 0; t < 2; ++t)
                    for (int kb = 0; kb < n_kb; ++kb)
                        tma_load_2d(a_res + (t * n_kb + kb) * kABlk, &map_q, a_full, kb * kK, m_blk * kRowsCta + t * 128);
            }
            for (int tile = tile0; tile < tile1; ++tile) {
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    mbar_expect_tx(full0 + 8 * stage, kBBlk);
                    tma_load_2d(ring + stage * kBBlk, &map_d, full0 + 8 * stage, kb * kK, tile * kDocs);
                    if (++stage == kStagesB) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            int local = 0;
            if (tile0 < tile1) mbar_wait(a_full, 0);
            for (int tile = tile0; tile < tile1; ++tile, ++local) {
                const uint32_t buf = (uint32_t)(local & 1);
                mbar_wait(tmem_empty0 + 8 * buf, ((uint32_t)(local >> 1) & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(full0 + 8 * stage, phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t bdesc = umma_desc_sw128(ring + stage * kBBlk);
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        const uint64_t adesc = umma_desc_sw128(a_res + (t * n_kb + kb) * kABlk);
#pragma unroll
                        for (int k = 0; k < kK / kUK; ++k)
                            umma_bf16(tmem_base + (buf * 2u + (uint32_t)t) * 128u, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2),
                                      kIdescBf, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(empty0 + 8 * stage);
                    if (++stage == kStagesB) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(tmem_full0 + 8 * buf);
            }
        }
    } else {
        // ===== epilogue: 8 warps; warp w: row tile (w - 2) / 4, TMEM lanes 32 * (w % 4) .. + 31 =====
        const int e = warp - 2;
        const int t = e >> 2, quarter = warp & 3;
        const int row = m_blk * kRowsCta + t * 128 + quarter * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const float rdiag = (FULL_STATS && row < B) ? diag[row] : 0.f;
        // running (max, sum exp) in the base-2 domain: x2 = S * scale * log2(e); 4 instructions per logit in the
        // common case (FMNMX, FFMA, MUFU.EX2, FADD): with two warps per scheduler the epilogue of a 256 x 128 step
        // issues in ~1300 cycles against 2048 cycles of MMA
        const float c2 = scale * 1.4426950408889634f;
        float rm2 = -INFINITY, rl = 0.f, rh = 0.f, rx = -INFINITY;
        int local = 0;
        for (int tile = tile0; tile < tile1; ++tile, ++local) {
            const uint32_t buf = (uint32_t)(local & 1);
            mbar_wait(tmem_full0 + 8 * buf, (uint32_t)(local >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int col_tile0 = tile * kDocs;
#pragma unroll 1
            for (int c0 = 0; c0 < kDocs; c0 += 32) {
                float v[32];
                tmem_ld32(lane_addr + (buf * 2u + (uint32_t)t) * 128u + (uint32_t)c0, v);
                const int col0 = col_tile0 + c0;
                if (col0 >= B) continue;
                const int n_valid = min(32, B - col0);
                if (n_valid < 32) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j >= n_valid) v[j] = -INFINITY;           // columns past the batch: exp2(-inf) = 0
                }
                float cmax = v[0];
#pragma unroll
                for (int j = 1; j < 32; ++j) cmax = fmaxf(cmax, v[j]);
                cmax *= c2;
                if (cmax > rm2) {
                    rl *= exp2f(rm2 - cmax);
                    rm2 = cmax;
                }
                float add0 = 0.f, add1 = 0.f;
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    float e0, e1;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(v[j], c2, -rm2)));
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(v[j + 1], c2, -rm2)));
                    add0 += e0;
                    add1 += e1;
                }
                rl += add0 + add1;
                if (FULL_STATS) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (j < n_valid) {
                            const int c = col0 + j;
                            const float h = v[j] - rdiag + margin;
                            rh += fminf(fmaxf(h, 0.f), 1e14f) * (colw ? __ldg(colw + c) : 1.f);
                            rx = fmaxf(rx, c == row ? 0.f : v[j]);
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty0 + 8 * buf);
        }
        // back to the natural-log domain the finalize kernel merges in: m = rm2 * ln 2, l unchanged (sum of e^(x - m))
        if (row < B) part[(size_t)blockIdx.y * B + row] = RowStat{rm2 * 0.6931471805599453f, rl, rh, rx};
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

}  // namespace bf

// ---- host ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return (EncodeTiledFn) nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

static int make_map(CUtensorMap *map, const float *base, int64_t rows, int64_t cols, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error(RF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(RF_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return RF_OK;
}

// Launches the tensor-core partial-statistics kernel.  Returns the number of column groups in *splits.
int launch_logits_tc(const float *q, const float *d, const float *diag, const float *colw, int B, int Dt, float scale,
                     float margin, bool full_stats, RowStat *part, int max_splits, int *splits, cudaStream_t st) {
    if (Dt % 4 != 0 || (reinterpret_cast<uintptr_t>(q) & 15) || (reinterpret_cast<uintptr_t>(d) & 15))
        return set_error(RF_ERR_UNSUPPORTED, "tensor-core logits need dim %% 4 == 0 and 16-byte aligned operands");
    CUtensorMap mq, md;
    int rc = make_map(&mq, q, B, Dt, kBM);
    if (rc != RF_OK) return rc;
    rc = make_map(&md, d, B, Dt, kBN);
    if (rc != RF_OK) return rc;
    const int m_tiles = (B + kBM - 1) / kBM;
    const int n_tiles = (B + kBN - 1) / kBN;
    const int n_kb = (Dt + kBK - 1) / kBK;
    const bool a_res = n_kb <= kMaxResidentKB;
    // column groups: one CTA per SM is resident, so choose the split that minimises
    // waves x (tiles per CTA + the one-off query-tile load)
    int dev = 0, sms = 148;
    RF_CUDA(cudaGetDevice(&dev));
    RF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int groups = 1;
    double best = 1e30;
    for (int g = 1; g <= n_tiles && g <= max_splits; ++g) {
        const int tpc = (n_tiles + g - 1) / g;
        const int g_eff = (n_tiles + tpc - 1) / tpc;
        const double waves = (double)(((int64_t)m_tiles * g_eff + sms - 1) / sms);
        const double cost = waves * (tpc + 0.5);
        if (cost < best - 1e-9) {
            best = cost;
            groups = g_eff;
        }
    }
    const int tiles_per_cta = (n_tiles + groups - 1) / groups;
    groups = (n_tiles + tiles_per_cta - 1) / tiles_per_cta;
    *splits = groups;
    const dim3 grid(m_tiles, groups);
#define RF_LAUNCH_TC(FULL, ARES)                                                                                          \
    do {                                                                                                                  \
        const size_t smem = TcCfg<ARES>::smem_bytes(n_kb);                                                                \
        RF_CUDA(cudaFuncSetAttribute(logits_tc_kernel<FULL, ARES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        logits_tc_kernel<FULL, ARES><<<grid, kTcThreads, smem, st>>>(mq, md, diag, colw, B, Dt, scale, margin, n_tiles,     \
                                                                    tiles_per_cta, part);                                 \
    } while (0)
    if (full_stats && a_res) RF_LAUNCH_TC(true, true);
    else if (full_stats) RF_LAUNCH_TC(true, false);
    else if (a_res) RF_LAUNCH_TC(false, true);
    else RF_LAUNCH_TC(false, false);
#undef RF_LAUNCH_TC
    RF_CUDA(cudaGetLastError());
    g_launches.fetch_add(1);
    return RF_OK;
}

static int make_map_bf16(CUtensorMap *map, const void *base, int64_t rows, int64_t cols, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return set_error(RF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    const cuuint32_t box[2] = {(cuuint32_t)bf::kK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(RF_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return RF_OK;
}

// bf16 operands q16, d16: [B, Dt] bf16 row-major (already rounded).  Dt % 8 == 0, Dt <= 256.
int launch_logits_bf16(const void *q16, const void *d16, const float *diag, const float *colw, int B, int Dt, float scale,
                       float margin, bool full_stats, RowStat *part, int max_splits, int *splits, cudaStream_t st) {
    if (Dt % 8 != 0 || Dt > bf::kMaxKB * bf::kK) return set_error(RF_ERR_UNSUPPORTED, "bf16 tensor-core logits need dim %% 8 == 0 and dim <= 256");
    CUtensorMap mq, md;
    int rc = make_map_bf16(&mq, q16, B, Dt, 128);
    if (rc != RF_OK) return rc;
    rc = make_map_bf16(&md, d16, B, Dt, bf::kDocs);
    if (rc != RF_OK) return rc;
    const int m_blks = (B + bf::kRowsCta - 1) / bf::kRowsCta;
    const int n_tiles = (B + bf::kDocs - 1) / bf::kDocs;
    const int n_kb = (Dt + bf::kK - 1) / bf::kK;
    int dev = 0, sms = 148;
    RF_CUDA(cudaGetDevice(&dev));
    RF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int groups = 1;
    double best = 1e30;
    for (int g = 1; g <= n_tiles && g <= max_splits; ++g) {
        const int tpc = (n_tiles + g - 1) / g;
        const int g_eff = (n_tiles + tpc - 1) / tpc;
        const double waves = (double)(((int64_t)m_blks * g_eff + sms - 1) / sms);
        const double cost = waves * (tpc + 1.0);              // + the one-off query block load
        if (cost < best - 1e-9) {
            best = cost;
            groups = g_eff;
        }
    }
    const int tiles_per_cta = (n_tiles + groups - 1) / groups;
    groups = (n_tiles + tiles_per_cta - 1) / tiles_per_cta;
    *splits = groups;
    const dim3 grid(m_blks, groups);
    const size_t smem = bf::smem_bytes(n_kb);
    if (full_stats) {
        RF_CUDA(cudaFuncSetAttribute(bf::logits_bf16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        bf::logits_bf16_kernel<true><<<grid, bf::kThreadsB, smem, st>>>(mq, md, diag, colw, B, Dt, scale, margin, n_tiles, tiles_per_cta, part);
    } else {
        RF_CUDA(cudaFuncSetAttribute(bf::logits_bf16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        bf::logits_bf16_kernel<false><<<grid, bf::kThreadsB, smem, st>>>(mq, md, diag, colw, B, Dt, scale, margin, n_tiles, tiles_per_cta, part);
    }
    RF_CUDA(cudaGetLastError());
    g_launches.fetch_add(1);
    return RF_OK;
}

}  // namespace rf
