// rf_gemm_tc.cu -- Keras Dense (y = activation(x W + b)) on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces the tower MLP of /root/reference/backend/blocks/mlp.py:4-15 (create_mlp: [norm, Dense(units,
// activation), Dropout] * n; models/matching/dssm.py:25-26 builds [1024, 512, 256], selu, BatchNormalization(1e-6))
// and the Dense q / k / v projections of backend/layers/attention_layers.py:141-155.  At inference the
// BatchNormalization in front of a Dense is an affine map per input column and is folded into W and b by the
// host layer, so one launch of this kernel is a whole [norm, Dense, activation] stage; the last stage can also
// l2-normalise its rows (dssm.py:36 / que2search.py embedding_norm) in the same epilogue.
//
//   operands   x [M, K] fp32 (row pitch ldx) and W^T [N, K] fp32, both K-major, read by TMA straight from HBM as
//              TF32 (kind::tf32 -- what TensorFlow itself does for fp32 matmuls on Ampere and later), fp32 accumulate
//   tile       128 x BN x 32 per stage (BN = 64 / 128 / 256 picked per shape), 4-6 stage TMA -> smem ring,
//              4 x tcgen05.mma (K = 8) per stage, TWO 128-lane x BN-column TMEM accumulators so that the MMAs of
//              tile i+1 run under the epilogue of tile i; persistent CTAs walk the tiles grid-stride
//   warps      0: TMA producer   1: TMEM alloc + MMA issue   2-9: epilogue, thread = output row = TMEM lane, EIGHT
//              warps (two per scheduler: warps w and w + 4 own the same TMEM lanes and take alternate 32-column
//              chunks) -- the first version had four and its profile (profiles/r2b_ncu_dense_tc_summary.csv) showed
//              the kernel paced by their instruction issue (tensor pipe 9 % active): a branchy activation with a
//              bias load per element cost ~55 dependent instructions per output on one warp per scheduler.
//              tcgen05.ld 32 columns, + bias (one coalesced load per chunk, broadcast by shuffles), branch-free
//              activation chosen once per chunk, (optional row l2-norm), then the warp's 32 x 32 block goes
//              through a 128-byte-swizzled smem tile and leaves as ONE TMA store (UTMASTG): full 128-byte lines per
//              row instead of 32 scattered 16-byte stores per instruction (first version: 0.2 ms for a 64 -> 64
//              projection of 409 600 rows whose HBM floor is 32 us); TMA clips rows / columns past the edges
// Every mbarrier wait is bounded and traps instead of hanging.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <atomic>

#include "../../include/rf_b200.h"
#include "rf_common.h"

namespace rf {

extern std::atomic<int64_t> g_launches;

namespace gemm_tc {

constexpr int kBM = 128, kBK = 32, kUmmaK = 8;
constexpr int kABytes = kBM * kBK * 4;              // 16 KiB
constexpr int kThreads = 64 + 256;
constexpr int kEpiWarps = 8;
constexpr int kOutStage = kEpiWarps * 2 * 4096;     // per epilogue warp: two 32-row x 128-byte staging tiles for the TMA store

// PAIR: two CTAs of a cluster (one TPC) compute a 256 x BN tile with tcgen05.mma.cta_group::2: each CTA stages ITS 128 rows
// of x and only HALF of the BN weight rows, the tensor core reads the other half from the partner's shared memory -- per CTA
// (128 + BN / 2) * K * 4 bytes cross L2 -> smem per tile instead of (128 + BN) * K * 4, which is what bounds this kernel.
template <int BN, bool PAIR = false>
struct Cfg {
    static constexpr int kBBytes = (PAIR ? BN / 2 : BN) * kBK * 4;          // weight rows this CTA stages
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = PAIR ? 4 : (BN == 256 ? 3 : (BN == 128 ? 4 : 6));     // ring + 64 KiB of store staging <= 227 KiB
    static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
    static constexpr size_t kSmem = (size_t)kStages * kStageBytes + kOutStage + 2048 + 1024 + 256;
};

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
        if (spin > (1u << 26)) __trap();
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
// ---- CTA-pair (cta_group::2) forms -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address of this CTA's window) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t local, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into THIS CTA's shared memory whose completion is counted on a barrier that may live in the partner CTA
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// completion of all prior MMAs of the pair, delivered to the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
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
// K-major operand tile, 128-byte swizzle: rows at a 128-byte pitch, 8-row groups 1024 bytes apart (SBO)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

template <int ACT>
__device__ __forceinline__ float activate(float x) {
    if (ACT == RF_ACT_RELU) return fmaxf(x, 0.f);
    if (ACT == RF_ACT_SELU) {          // scale * (x > 0 ? x : alpha * (e^x - 1)), both sides evaluated, no branch
        const float neg = 1.7580993408473766f * (__expf(fminf(x, 0.f)) - 1.f);
        return x > 0.f ? 1.0507009873554805f * x : neg;
    }
    if (ACT == RF_ACT_TANH) return tanhf(x);
    if (ACT == RF_ACT_SIGMOID) return __fdividef(1.f, 1.f + __expf(-x));
    if (ACT == RF_ACT_GELU) return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
    return x;
}

// v[j] = activation(v[j] + bias[col0 + j]) for the 32 columns of one chunk; `bias_lane` = bias[col0 + lane] (or 0)
template <int ACT>
__device__ __forceinline__ void bias_act_chunk(float (&v)[32], float bias_lane) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = activate<ACT>(v[j] + __shfl_sync(0xffffffffu, bias_lane, j));
}
__device__ __forceinline__ void bias_act(float (&v)[32], float bias_lane, int act) {
    switch (act) {
        case RF_ACT_RELU: bias_act_chunk<RF_ACT_RELU>(v, bias_lane); break;
        case RF_ACT_SELU: bias_act_chunk<RF_ACT_SELU>(v, bias_lane); break;
        case RF_ACT_TANH: bias_act_chunk<RF_ACT_TANH>(v, bias_lane); break;
        case RF_ACT_SIGMOID: bias_act_chunk<RF_ACT_SIGMOID>(v, bias_lane); break;
        case RF_ACT_GELU: bias_act_chunk<RF_ACT_GELU>(v, bias_lane); break;
        default: bias_act_chunk<RF_ACT_NONE>(v, bias_lane); break;
    }
}

struct Params {
    const float *bias;      // [N] or NULL
    float *out;             // [M, N], row pitch ldo
    int64_t ldo;
    int M, N, K;
    int act, l2norm;
    int n_m_tiles, n_n_tiles;
    float l2_eps;
    // split-K: tile t = (split, m_tile, n_tile); split s contracts k-blocks [s * kb_per, (s + 1) * kb_per) and stores its
    // partial product at rows s * m_pad + ... of a [k_splits * m_pad, N] buffer (summed in order by splitk_sum_kernel)
    int k_splits, kb_per, m_pad;
};

template <int BN, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
dense_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const __grid_constant__ CUtensorMap map_o, Params p) {
    extern __shared__ uint8_t smem_raw[];
    using C = Cfg<BN, PAIR>;
    // PAIR: launched as clusters of 2; rank 0 (the leader) issues the MMAs of the pair and owns the barriers the MMA issuer
    // waits on (full: the TMA bytes of BOTH CTAs are counted there; tmem_empty: the epilogue warps of BOTH CTAs arrive there);
    // the barriers the MMA completion signals (empty, tmem_full) exist in both CTAs and are hit by a multicast commit
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    const int cta = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;          // tile walker index: a pair walks as one
    const int n_ctas = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    constexpr int kStages = C::kStages;
    constexpr int kStageBytes = C::kStageBytes;
    constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                ((uint32_t)((PAIR ? 2 * kBM : kBM) >> 4) << 24);
    const int n_kb = (p.K + kBK - 1) / kBK;
    const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t out_stage = ring + kStages * kStageBytes;      // 1024-byte aligned (stage sizes are multiples of 1 KiB)
    const uint32_t xch_s = out_stage + kOutStage;                 // float2 [2][128]: l2-norm partial sums of the two column halves
    const uint32_t bars = xch_s + 2048;
    const uint32_t full0 = bars, empty0 = bars + 8 * kStages;
    const uint32_t tmem_full0 = bars + 16 * kStages, tmem_empty0 = tmem_full0 + 16, tmem_slot = tmem_empty0 + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mn_tiles = p.n_m_tiles * p.n_n_tiles;
    const int n_tiles = mn_tiles * p.k_splits;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tmem_full0 + 8 * a, 1);
            mbar_init(tmem_empty0 + 8 * a, PAIR ? 2 * kEpiWarps : kEpiWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (PAIR) {               // one warp of EACH CTA of the pair performs the (symmetric) allocation
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(C::kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(C::kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (PAIR) cluster_sync_all();          // the partner's barriers are initialised before anything signals them
    else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int t = cta; t < n_tiles; t += n_ctas) {
                const int split = t / mn_tiles, mn = t - split * mn_tiles;
                const int m_tile = (mn / p.n_n_tiles) * (PAIR ? 2 : 1) + (int)rank, n_tile = mn - (mn / p.n_n_tiles) * p.n_n_tiles;
                const int kb0 = split * p.kb_per, kb1 = min(n_kb, kb0 + p.kb_per);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    const uint32_t dst = ring + stage * kStageBytes;
                    if (PAIR) {
                        // both CTAs' bytes are counted on the LEADER's full barrier (the peer's complete_tx may land before the
                        // leader's expect_tx of the phase: the pending arrival keeps the phase open)
                        const uint32_t bar = mapa_rank(full0 + 8 * stage, 0);
                        if (leader) mbar_expect_tx(full0 + 8 * stage, 2 * kStageBytes);
                        tma_load_2d_pair(dst, &map_a, bar, kb * kBK, m_tile * kBM);
                        tma_load_2d_pair(dst + kABytes, &map_b, bar, kb * kBK, n_tile * BN + (int)rank * (BN / 2));
                    } else {
                        mbar_expect_tx(full0 + 8 * stage, kStageBytes);
                        tma_load_2d(dst, &map_a, full0 + 8 * stage, kb * kBK, m_tile * kBM);
                        tma_load_2d(dst + kABytes, &map_b, full0 + 8 * stage, kb * kBK, n_tile * BN);
                    }
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0 && leader) {
            uint32_t stage = 0, phase = 0;
            int local = 0;
            for (int t = cta; t < n_tiles; t += n_ctas, ++local) {
                const uint32_t acc = (uint32_t)(local & 1);
                mbar_wait(tmem_empty0 + 8 * acc, ((uint32_t)(local >> 1) & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int kb0 = (t / mn_tiles) * p.kb_per, kb1 = min(n_kb, kb0 + p.kb_per);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(full0 + 8 * stage, phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st_addr = ring + stage * kStageBytes;
                    const uint64_t adesc = desc_sw128(st_addr), bdesc = desc_sw128(st_addr + kABytes);
#pragma unroll
                    for (int k = 0; k < kBK / kUmmaK; ++k) {
                        if (PAIR)
                            umma_tf32_pair(tmem_base + acc * (uint32_t)BN, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), kIdesc,
                                           (kb != kb0 || k != 0) ? 1u : 0u);
                        else
                            umma_tf32(tmem_base + acc * (uint32_t)BN, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), kIdesc,
                                      (kb != kb0 || k != 0) ? 1u : 0u);
                    }
                    if (PAIR) umma_commit_pair(empty0 + 8 * stage);
                    else umma_commit(empty0 + 8 * stage);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (PAIR) umma_commit_pair(tmem_full0 + 8 * acc);
                else umma_commit(tmem_full0 + 8 * acc);
            }
        }
    } else {
        // ===== epilogue: thread <-> output row (TMEM lane); warp w may touch lanes 32 * (w % 4) .. + 31; warps w and
        // w + 4 share those lanes and take alternate 32-column chunks =====
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        float2 *xch = reinterpret_cast<float2 *>(smem_raw + (xch_s - smem_u32(smem_raw)));
        int local = 0;
        uint32_t n_store = 0;
        for (int t = cta; t < n_tiles; t += n_ctas, ++local) {
            const int split = t / mn_tiles, mn = t - split * mn_tiles;
            const int m_tile = (mn / p.n_n_tiles) * (PAIR ? 2 : 1) + (int)rank, n_tile = mn - (mn / p.n_n_tiles) * p.n_n_tiles;
            const uint32_t acc = (uint32_t)(local & 1);
            mbar_wait(tmem_full0 + 8 * acc, (uint32_t)(local >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float inv = 1.f;
            if (p.l2norm) {                     // pass 1: ||activation(x W + b)||^2 of this row (one column tile holds the row)
                float ss = 0.f;
#pragma unroll 1
                for (int c0 = half * 32; c0 < BN; c0 += 64) {
                    const int col0 = n_tile * BN + c0;
                    if (col0 >= p.N) break;
                    float v[32];
                    tmem_ld32(lane_addr + acc * (uint32_t)BN + (uint32_t)c0, v);
                    const float bl = (p.bias && col0 + lane < p.N) ? __ldg(p.bias + col0 + lane) : 0.f;
                    bias_act(v, bl, p.act);
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (col0 + j < p.N) ss = fmaf(v[j], v[j], ss);
                }
                // the row's other half lives in the partner warp (w +- 4): exchange the partial sums through smem
                const int r = quarter * 32 + lane;
                xch[half * 128 + r].x = ss;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                ss += xch[(half ^ 1) * 128 + r].x;
                asm volatile("bar.sync 1, 256;" ::: "memory");                       // before the next tile overwrites xch
                inv = 1.f / fmaxf(sqrtf(ss), p.l2_eps);
            }
            const int row0 = split * p.m_pad + m_tile * kBM + quarter * 32;
#pragma unroll 1
            for (int c0 = half * 32; c0 < BN; c0 += 64) {
                const int col0 = n_tile * BN + c0;
                if (col0 >= p.N) break;
                float v[32];
                tmem_ld32(lane_addr + acc * (uint32_t)BN + (uint32_t)c0, v);
                const float bl = (p.bias && col0 + lane < p.N) ? __ldg(p.bias + col0 + lane) : 0.f;
                bias_act(v, bl, p.act);
                // stage the warp's 32 rows x 32 columns (128 bytes per row, 16-byte chunks XOR-ed with row & 7 = the
                // TMA 128-byte swizzle) and hand the tile to the TMA store engine
                const uint32_t buf = out_stage + (uint32_t)(warp - 2) * 8192u + (uint32_t)(n_store & 1) * 4096u;
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // this buffer's previous store has read it
                __syncwarp();
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(buf + (uint32_t)lane * 128u + (uint32_t)((c ^ (lane & 7)) << 4)),
                                 "f"(v[4 * c] * inv), "f"(v[4 * c + 1] * inv), "f"(v[4 * c + 2] * inv), "f"(v[4 * c + 3] * inv)
                                 : "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) tma_store_2d(&map_o, buf, col0, row0);
                ++n_store;
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                if (PAIR) mbar_arrive_cluster(mapa_rank(tmem_empty0 + 8 * acc, 0));      // the leader's MMA issuer waits for both CTAs
                else mbar_arrive(tmem_empty0 + 8 * acc);
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // every store has landed before the CTA exits
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (PAIR) cluster_sync_all();          // the leader's MMAs read the partner's shared memory: nobody leaves early
    else __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::kTmemCols) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::kTmemCols) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(CUtensorMap *map, const float *base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                    CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B) {
    static EncodeTiledFn fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) return (EncodeTiledFn) nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    if (!fn) return set_error(RF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(RF_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return RF_OK;
}

__global__ void __launch_bounds__(256) splitk_sum_kernel(const float *__restrict__ part, int splits, int64_t m_pad, int M, int N,
                                                          float *__restrict__ out, int64_t ldo) {
    const int q = N >> 2;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < (int64_t)M * q; e += (int64_t)gridDim.x * 256) {
        const int64_t m = e / q;
        const int n = (int)(e - m * q) * 4;
        float4 a = *reinterpret_cast<const float4 *>(part + m * N + n);
        for (int s = 1; s < splits; ++s) {
            const float4 b = *reinterpret_cast<const float4 *>(part + ((int64_t)s * m_pad + m) * N + n);
            a.x += b.x;
            a.y += b.y;
            a.z += b.z;
            a.w += b.w;
        }
        *reinterpret_cast<float4 *>(out + m * ldo + n) = a;
    }
}

template <int BN, bool PAIR = false>
static int launch(const float *x, int64_t ldx, const float *wt, int64_t ldw, Params p, int sms, cudaStream_t st, float *partials) {
    CUtensorMap ma, mb, mo;
    constexpr int kRowsPerTile = PAIR ? 2 * kBM : kBM;      // a pair of CTAs walks 256-row tiles
    int rc = make_map(&ma, x, p.M, p.K, ldx, kBM);
    if (rc != RF_OK) return rc;
    rc = make_map(&mb, wt, p.N, p.K, ldw, PAIR ? BN / 2 : BN);
    if (rc != RF_OK) return rc;
    p.n_m_tiles = (p.M + kRowsPerTile - 1) / kRowsPerTile;
    p.n_n_tiles = (p.N + BN - 1) / BN;
    p.m_pad = p.n_m_tiles * kRowsPerTile;
    if (p.k_splits > 1)
        rc = make_map(&mo, partials, (int64_t)p.k_splits * p.m_pad, p.N, p.N, 32, CU_TENSOR_MAP_L2_PROMOTION_NONE);
    else
        rc = make_map(&mo, p.out, p.M, p.N, p.ldo, 32, CU_TENSOR_MAP_L2_PROMOTION_NONE);     // 32 x 32 store boxes
    if (rc != RF_OK) return rc;
    const int tiles = p.n_m_tiles * p.n_n_tiles * p.k_splits;
    const int walkers = PAIR ? sms / 2 : sms;
    const int grid = (tiles < walkers ? tiles : walkers) * (PAIR ? 2 : 1);
    RF_CUDA(cudaFuncSetAttribute(dense_tc_kernel<BN, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<BN, PAIR>::kSmem));
    if (PAIR) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = Cfg<BN, PAIR>::kSmem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        RF_CUDA(cudaLaunchKernelEx(&cfg, dense_tc_kernel<BN, PAIR>, ma, mb, mo, p));
    } else {
        dense_tc_kernel<BN, PAIR><<<grid, kThreads, Cfg<BN, PAIR>::kSmem, st>>>(ma, mb, mo, p);
    }
    if (p.k_splits > 1) {
        int64_t blocks = ((int64_t)p.M * (p.N / 4) + 255) / 256;
        if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
        splitk_sum_kernel<<<(unsigned)blocks, 256, 0, st>>>(partials, p.k_splits, p.m_pad, p.M, p.N, p.out, p.ldo);
    }
    RF_CUDA(cudaGetLastError());
    g_launches.fetch_add(p.k_splits > 1 ? 2 : 1);
    return RF_OK;
}

// Column-tile width and K splits.  The kernel is bound by L2 -> shared-memory traffic (fp32 operands: a 128 x BN tile pulls
// (128 + BN) * K * 4 bytes), with a per-tile epilogue cost ~ 2 bytes-equivalents per output element, and the persistent grid
// runs ceil(CTAs / SMs) waves -- measured on B200 (profiles/r2d_gemm_tile_sweep.txt), time ~ waves * ((128 + BN) * K + 256 * BN):
// 8192 x 1888 x 1024 takes 0.065 / 0.083 / 0.123 ms at BN = 256 / 128 / 64 (model 1 : 1.30 : 1.68).  A product whose output
// has few tiles and a long contraction (dW = X^T dZ: K = batch) additionally splits K over CTAs (partials summed in order by
// splitk_sum_kernel, ~5 us of extra launch + traffic).
struct TileChoice {
    int bn, splits;
    bool pair;
};

static TileChoice pick_config(int64_t rows, int units, int in_dim, int sms, bool allow_split, bool l2norm) {
    const int64_t m_tiles = (rows + kBM - 1) / kBM;
    const int n_kb = (in_dim + kBK - 1) / kBK;
    TileChoice best{64, 1, false};
    double best_cost = 1e300;
    static const int kSplits[] = {1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64};
    for (int bn : {256, 128, 64}) {
        if (l2norm && units > bn) continue;
        if (bn > 64 && units <= bn / 2) continue;                      // a mostly empty column tile
        const int64_t tiles = m_tiles * ((units + bn - 1) / bn);
        for (int s : kSplits) {
            if (s > 1 && (!allow_split || n_kb / s < 8)) break;
            const int kb_per = (n_kb + s - 1) / s;
            const int real = (n_kb + kb_per - 1) / kb_per;
            if (real != s) continue;
            const double waves = (double)((tiles * s + sms - 1) / sms);
            double cost = waves * ((double)(kBM + bn) * kb_per * kBK + 2.0 * kBM * bn);
            if (s > 1) cost += 120000.0 + 2.0 * (double)m_tiles * kBM * units * s / sms;      // extra launch + partials written and re-read (whole GPU)
            if (cost < best_cost) {
                best_cost = cost;
                best = TileChoice{bn, s, false};
            }
        }
    }
    // CTA pairs (cta_group::2): 256-row tiles, each CTA stages its 128 rows of x and half of the weight tile.  Measured
    // (profiles/r2e_gemm_pair_sweep.txt): a gain where the single-CTA kernel is L2-bound -- at least a full wave of pair tiles
    // and a contraction of 512 or more (8192 x 1888 x 1024: 0.067 -> 0.059 ms, 65536 rows: 0.425 -> 0.374 ms) -- and a loss
    // for few tiles (half the walkers) or short contractions (epilogue-bound).  RF_DENSE_PAIR=0 turns pairs off, =2 forces
    // them wherever they are legal (tests).
    const char *env = getenv("RF_DENSE_PAIR");
    const int mode = env ? atoi(env) : 1;
    if (mode != 0 && rows >= 2 * kBM && !l2norm) {
        const int64_t m_pairs = (rows + 2 * kBM - 1) / (2 * kBM);
        for (int bn : {256, 128}) {
            if (units <= bn / 2) continue;
            const int64_t tiles = m_pairs * ((units + bn - 1) / bn);
            if (mode != 2 && (tiles < sms / 2 || in_dim < 512)) continue;
            const double waves = (double)((tiles + sms / 2 - 1) / (sms / 2));
            const double cost = waves * ((double)(kBM + bn / 2) * n_kb * kBK + 2.0 * kBM * bn);
            if (cost < best_cost || (mode == 2 && !best.pair)) {
                best_cost = cost;
                best = TileChoice{bn, 1, true};
            }
        }
    }
    return best;
}

}  // namespace gemm_tc
}  // namespace rf

using namespace rf;

extern "C" int64_t rf_dense_tc_workspace_bytes(int64_t rows, int32_t in_dim, int32_t units) {
    using namespace gemm_tc;
    if (rows <= 0 || in_dim <= 0 || units <= 0) return 0;
    const TileChoice c = pick_config(rows, units, in_dim, 148, true, false);
    if (c.splits <= 1) return 0;
    return (int64_t)c.splits * ((rows + kBM - 1) / kBM) * kBM * units * (int64_t)sizeof(float);
}

extern "C" int rf_dense_forward_tc_ex(const float *d_x, int64_t rows, int32_t in_dim, int64_t ldx, const float *d_weight_t,
                                      const float *d_bias, int32_t units, int activation, int l2_normalize, float *d_out,
                                      int64_t ldo, void *d_workspace, int64_t workspace_bytes, void *stream) {
    using namespace gemm_tc;
    if (rows < 0 || in_dim <= 0 || units <= 0) return set_error(RF_ERR_INVALID, "bad Dense shape");
    if (activation < RF_ACT_NONE || activation > RF_ACT_GELU) return set_error(RF_ERR_INVALID, "Unknown activation function: %d", activation);
    if (rows == 0) return RF_OK;
    if (!d_x || !d_weight_t || !d_out) return set_error(RF_ERR_INVALID, "rf_dense_forward_tc: NULL buffer");
    if (rows > INT32_MAX) return set_error(RF_ERR_UNSUPPORTED, "more than 2^31 - 1 rows");
    const uintptr_t al = reinterpret_cast<uintptr_t>(d_x) | reinterpret_cast<uintptr_t>(d_weight_t) | reinterpret_cast<uintptr_t>(d_out);
    if ((al & 15) || in_dim % 4 || units % 4 || ldx % 4 || ldo % 4 || ldx < in_dim || ldo < units)
        return set_error(RF_ERR_UNSUPPORTED, "tensor-core Dense needs in_dim, units and both row pitches to be multiples of 4 "
                                             "floats and 16-byte aligned buffers");
    if (l2_normalize && units > 256) return set_error(RF_ERR_UNSUPPORTED, "fused row l2-normalisation handles units <= 256");
    int dev = 0, sms = 148;
    RF_CUDA(cudaGetDevice(&dev));
    RF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    Params p{d_bias, d_out, ldo, (int)rows, units, in_dim, activation, l2_normalize ? 1 : 0, 0, 0, 1e-12f, 1, 0, 0};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int n_kb = (in_dim + kBK - 1) / kBK;
    p.kb_per = n_kb;
    // split-K only for a plain product (no bias / activation / normalisation to apply to a partial sum) and only when the
    // caller brought the workspace for it
    const bool may_split = d_workspace && !d_bias && activation == RF_ACT_NONE && !l2_normalize &&
                           (reinterpret_cast<uintptr_t>(d_workspace) & 15) == 0 &&
                           workspace_bytes >= rf_dense_tc_workspace_bytes(rows, in_dim, units);
    TileChoice c = pick_config(rows, units, in_dim, 148, may_split, l2_normalize != 0);
    if (const char *force = getenv("RF_DENSE_BN")) {          // experiments: force the column tile (no split)
        const int bn = atoi(force);
        if ((bn == 64 || bn == 128 || bn == 256) && !(l2_normalize && units > bn)) c = TileChoice{bn, 1, false};
    }
    if (c.splits > 1) {
        p.kb_per = (n_kb + c.splits - 1) / c.splits;
        p.k_splits = (n_kb + p.kb_per - 1) / p.kb_per;
    }
    float *partials = static_cast<float *>(d_workspace);
    if (c.pair && c.bn == 256) return launch<256, true>(d_x, ldx, d_weight_t, in_dim, p, sms, st, partials);
    if (c.pair) return launch<128, true>(d_x, ldx, d_weight_t, in_dim, p, sms, st, partials);
    if (c.bn == 256) return launch<256>(d_x, ldx, d_weight_t, in_dim, p, sms, st, partials);
    if (c.bn == 128) return launch<128>(d_x, ldx, d_weight_t, in_dim, p, sms, st, partials);
    return launch<64>(d_x, ldx, d_weight_t, in_dim, p, sms, st, partials);
}

extern "C" int rf_dense_forward_tc(const float *d_x, int64_t rows, int32_t in_dim, int64_t ldx, const float *d_weight_t,
                                   const float *d_bias, int32_t units, int activation, int l2_normalize, float *d_out,
                                   int64_t ldo, void *stream) {
    return rf_dense_forward_tc_ex(d_x, rows, in_dim, ldx, d_weight_t, d_bias, units, activation, l2_normalize, d_out, ldo, nullptr, 0,
                                  stream);
}
