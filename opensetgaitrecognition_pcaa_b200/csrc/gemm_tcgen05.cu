// Blackwell tensor-core GEMMs of the PCAA hot path: tcgen05.mma (bf16 x bf16 -> fp32 in TMEM), operands staged
// by TMA into 128B-swizzled shared memory, mbarrier producer/consumer pipeline, persistent warp-specialised CTAs.
//
//   warp 0      : TMA producer (one elected lane)
//   warp 1      : TMEM allocator + MMA issuer (one elected lane issues tcgen05.mma / tcgen05.commit)
//   warps 2..9  : epilogue (TMEM -> registers -> fused epilogue -> global): two warps per TMEM lane quarter, each
//                 owning half of the tile's columns
//
// Two 128 x BN fp32 accumulators live in TMEM (2*BN <= 512 columns) so the epilogue of tile i overlaps the MMAs of
// tile i+1.  Work items are (m-tile, n-tile, k-split) triples strided over the persistent grid.
//
// Replaces the per-point 1x1 Conv2d contractions of the reference (models.py:21-28, 86-98) and their autograd
// backward (data gradient with fused ELU'/BatchNorm-statistics epilogue, weight gradient with split-K).
#include <cuda.h>

#include "common.cuh"

namespace pcaa {

constexpr int BM = 128;
constexpr int BK = 64;            // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 320;
constexpr int EPI_THREADS = 256;

enum { MODE_BIAS_STATS = 0, MODE_BIAS_ELU = 1, MODE_PLAIN = 2, MODE_DGRAD_ELUBN = 3, MODE_WGRAD = 4, MODE_DGRAD_ELUOUT = 5,
       // "channel-major" modes: the output ROW (TMEM lane) is the channel, the columns are points; per-channel
       // parameters are per-thread scalars and BatchNorm statistics are plain in-thread sums (no cross-lane reduction)
       MODE_T_BIAS_STATS = 7, MODE_T_AFFINE_ELU = 8, MODE_T_DGRAD_ELUBN = 9, MODE_T_AFFINE_ELU_POOL = 10 };

template <int MODE> constexpr bool is_t_mode() { return MODE == MODE_T_BIAS_STATS || MODE == MODE_T_AFFINE_ELU || MODE == MODE_T_DGRAD_ELUBN || MODE == MODE_T_AFFINE_ELU_POOL; }
template <int MODE> constexpr bool has_row_stats() { return MODE == MODE_T_BIAS_STATS || MODE == MODE_T_DGRAD_ELUBN; }

struct GemmParams {
    int64_t M, N;                 // output extent (rows, cols)
    int m_tiles, n_tiles, k_splits;
    int kb_total, kb_per_split;   // K blocks of BK
    void* out;                    // bf16 [M, ldo]  or fp32 (MODE_WGRAD, or out_f32)
    int64_t ldo;
    int out_f32;                  // store fp32 instead of bf16 (modes 1, 2)
    int wgrad_store;              // MODE_WGRAD: overwrite instead of atomicAdd (requires k_splits == 1)
    int out_scalar;               // fp32 rows are not 16-byte aligned (ldo % 4 != 0): scalar stores
    int out_tma;                  // MODE_WGRAD: fp32 rows are 16-byte aligned -> staged in shared memory, TMA store / reduce-add
    const float* bias;
    double* stats;                // [2*N]
    const __nv_bfloat16* yprev;   // [M, ldy] (MODE_DGRAD_ELUBN)
    int64_t ldy;
    const float *scale, *shift, *mean, *invstd;
    int a_tiled, b_tiled;         // operand stored as 256-point tiles [n_tiles][C][256] (3-D tensor map), see pcaa.h
    int sched_mfixed;             // tile order: 0 = items strided over the grid; 1 = CTA keeps one m block (T modes)
    int pool_n;                   // MODE_T_AFFINE_ELU_POOL: points per group (>= 32); out = fp32 pooled [N / pool_n, M], zeroed
    float pool_inv_n;
};

// tile `it` of this CTA -> (m block, n block, k split); false when the CTA has no more work
__device__ __forceinline__ bool tile_at(const GemmParams& p, int it, int& m_blk, int& n_blk, int& ks) {
    if (p.sched_mfixed) {
        // channel-major modes, CTA pairs: cluster c = blockIdx.x / 2 owns the 256-row block c % m_pairs (this CTA its
        // upper or lower 128 rows) for the n blocks c / m_pairs + it * per.  The m_pairs clusters of one n block run at
        // the same time (the activation tile is fetched from DRAM once and hit in L2 by the others) and a CTA never
        // changes its rows (its per-row epilogue state lives in registers for the whole kernel)
        const int m_pairs = (p.m_tiles + 1) >> 1;
        const int cid = blockIdx.x >> 1;
        const int per = (gridDim.x >> 1) / m_pairs;
        m_blk = (cid % m_pairs) * 2 + (int)(blockIdx.x & 1);
        n_blk = cid / m_pairs + it * per;
        ks = 0;
        return cid < per * m_pairs && n_blk < p.n_tiles;
    }
    const int item = blockIdx.x + it * gridDim.x;
    if (item >= p.m_tiles * p.n_tiles * p.k_splits) return false;
    ks = item % p.k_splits;
    const int tile = item / p.k_splits;
    n_blk = tile % p.n_tiles;
    m_blk = tile / p.n_tiles;
    return true;
}

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug traps (kernel error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            printf("pcaa gemm_tc: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* smem, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* smem, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// ---- CTA pair (cta_group::2): two CTAs of a cluster on the SMs of one TPC execute ONE 256 x 256 UMMA tile; each
// stages its own 128 rows of A and HALF of B (so a k block costs 32 KB of shared memory per SM instead of 48 KB: more
// stages in flight), the leader (cluster rank 0) issues the MMAs and both epilogues read their own TMEM
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // default semantics (no .release.cluster): that form costs a cluster-scope memory barrier (MEMBAR + ERRBAR, 28 % of the
    // data-gradient kernel's stall samples); the TMEM reads are ordered by tcgen05.fence::before_thread_sync
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: data lands in THIS CTA's shared memory, the transaction bytes are counted on the mbarrier at
// cluster address `bar` (the leader's)
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t bar, void* smem, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem)), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(const CUtensorMap* map, uint32_t bar, void* smem, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem)), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives (once the MMAs issued so far retire) on the mbarrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}

// smem -> global tensor store (bulk async group) and its completion waits
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(smem)), "r"(c0), "r"(c1)
                 : "memory");
}
// global += smem (element-wise fp32 add performed by the memory system: split-K accumulation without per-thread atomics)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* smem, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(smem)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// one lane of a converged warp (the same one on every call): the issuing lane of TMA / tcgen05 instructions.  The loops
// around it run on all 32 lanes so that addresses / descriptors / phases stay warp-uniform (uniform registers) -- a
// `lane == 0` wrapper makes the compiler treat them as divergent and emit a waterfall loop around every instruction.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
// registers of an issued load are only valid after wait::ld: tie them to the wait so no consumer is scheduled above it
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]),
                   "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]),
                   "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    // the registers are only valid after wait::ld: tie them to the wait so no consumer is scheduled above it
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]),
                   "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]),
                   "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

// shared-memory matrix descriptor (sm_100 UMMA), 128-byte swizzle
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= 1ull << 46;   // descriptor version (Blackwell)
    d |= 2ull << 61;   // SWIZZLE_128B
    return d;
}

// butterfly reduce-scatter: v[j] (j = column within a 32-column chunk) summed over the 32 lanes (rows);
// afterwards lane l holds the total of column l in v[0].
__device__ __forceinline__ float warp_col_reduce32(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool upper = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            float send = upper ? v[i] : v[i + s];
            float keep = upper ? v[i + s] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

// Channel-major modes stage their output (and the dgrad mode its yprev tile) through shared memory, one
// 32-row x 64-column 128B-swizzled sub-tile (4 KB) per epilogue warp, moved by TMA: the global side of the epilogue
// costs no LSU wavefronts (a row-per-lane direct access touches 32 cache lines per instruction).
constexpr int SUB_BYTES = 32 * 64 * 2;
template <int BN, int MODE>
struct SmemLayout {
    static constexpr bool TMODE = is_t_mode<MODE>();
    static constexpr bool PAIR = TMODE;                             // channel-major modes run as CTA pairs
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * BK * 2;   // a pair splits the B tile between its two CTAs
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = PAIR ? ((MODE == MODE_T_DGRAD_ELUBN) ? 5 : 6) : ((BN == 256) ? 4 : 6);
    static constexpr bool STAGED = TMODE || MODE == MODE_WGRAD;     // epilogue output goes through shared memory + TMA
    static constexpr int COLP_FLOATS = STAGED ? 0 : 5 * BN;         // bias | scale, shift, mean, invstd
    static constexpr int STAT_FLOATS = STAGED ? 0 : 4 * 2 * BN;     // per epilogue warp: sum, sum2
    static constexpr int OBUF_BYTES = STAGED ? 8 * SUB_BYTES : 0;   // output staging, one sub-tile per epilogue warp
    static constexpr int YBUF_BYTES = (MODE == MODE_T_DGRAD_ELUBN) ? 8 * SUB_BYTES : 0;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + OBUF_BYTES + YBUF_BYTES + (COLP_FLOATS + STAT_FLOATS) * 4 + 256 + 1024;
};


// epilogue arithmetic of one 32-column chunk of a channel-major tile (thread = row).  FULL: every column is a real
// point; otherwise columns >= valid are pad points and are stored as zeros / kept out of the statistics.
struct TRow { float bias, scale, shift, scale_l2, shift_l2, invstd, nmean_invstd; };
template <int MODE, bool FULL>
__device__ __forceinline__ void t_chunk(float (&v)[32], const uint4* yraw4, const TRow& r, int valid, float& t1, float& t2) {
    if constexpr (MODE == MODE_T_BIAS_STATS) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            v[j] = (FULL || j < valid) ? v[j] + r.bias : 0.f;
            t1 += v[j];
            t2 = fmaf(v[j], v[j], t2);
        }
    } else if constexpr (MODE == MODE_T_AFFINE_ELU || MODE == MODE_T_AFFINE_ELU_POOL) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float z = fmaf(v[j], r.scale, r.shift), zl = fmaf(v[j], r.scale_l2, r.shift_l2);
            v[j] = (FULL || j < valid) ? elu_l2(z, zl) : 0.f;
        }
    } else {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&yraw4[g]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(h[e]);
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int j = g * 8 + 2 * e + u;
                    const float yv = u ? f.y : f.x;
                    const float z = fmaf(yv, r.scale, r.shift), zl = fmaf(yv, r.scale_l2, r.shift_l2);
                    const float gq = (FULL || j < valid) ? v[j] * elu_grad_l2(z, zl) : 0.f;
                    const float xh = fmaf(yv, r.invstd, r.nmean_invstd);
                    v[j] = gq;
                    t1 += gq;
                    t2 = fmaf(gq, xh, t2);
                }
            }
        }
    }
}

template <int BN, bool A_MN, bool B_MN, int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmY, const GemmParams p) {
    using L = SmemLayout<BN, MODE>;
    constexpr int STAGES = L::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* tiles = smem;
    uint8_t* obuf = smem + STAGES * L::STAGE_BYTES;                       // [8][SUB_BYTES], 1024-byte aligned
    uint8_t* ybuf = obuf + L::OBUF_BYTES;                                 // [8][SUB_BYTES]
    float* colp = reinterpret_cast<float*>(ybuf + L::YBUF_BYTES);
    float* wstat = colp + L::COLP_FLOATS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(wstat + L::STAT_FLOATS);
    uint64_t* full = bars;                   // [STAGES]
    uint64_t* empty = bars + STAGES;         // [STAGES]
    uint64_t* tfull = bars + 2 * STAGES;     // [2]
    uint64_t* tempty = bars + 2 * STAGES + 2;  // [2]
    uint64_t* ybar = bars + 2 * STAGES + 4;  // [8] yprev sub-tile landed (one per epilogue warp)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 12);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr bool PAIR = L::PAIR;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;       // 0 = leader of the CTA pair (issues the MMAs)

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], (PAIR ? 2 : 1) * (EPI_THREADS / 32));   // pair: the epilogue warps of both CTAs
        }
        for (int i = 0; i < 8; ++i) mbar_init(&ybar[i], 1);
        if constexpr (is_t_mode<MODE>() || MODE == MODE_WGRAD) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO) : "memory");
            if constexpr (MODE == MODE_T_DGRAD_ELUBN) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        if constexpr (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"(2 * BN)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"(2 * BN)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all();      // the peer's mbarriers must be initialised before anything signals them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================================================== TMA producer (warp-converged, one issuing lane)
        {
            int stage = 0;
            uint32_t phase = 0;
            int m_blk, n_blk, ks;
            for (int it = 0; tile_at(p, it, m_blk, n_blk, ks); ++it) {
                const int kb0 = ks * p.kb_per_split;
                const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
                if constexpr (MODE == MODE_T_DGRAD_ELUBN) {
                    // the epilogue reads this CTA's [128 x 256] block of yprev (contiguous in T256) PF_DIST tiles from now:
                    // pull it into L2 so its per-warp sub-tile loads do not wait on DRAM.  (Two tiles ahead was too
                    // early: ncu showed yprev fetched from DRAM twice -- the line was evicted before its use.)
                    constexpr int PF_DIST = 1;
                    for (int t = (it == 0 ? 0 : it + PF_DIST); t <= it + PF_DIST; ++t) {
                        int pm, pn, pk;
                        if (tile_at(p, t, pm, pn, pk) && lane == 0) {
                            const int64_t rows = min((int64_t)BM, p.M - (int64_t)pm * BM);
                            const __nv_bfloat16* src = p.yprev + ((int64_t)pn * p.M + (int64_t)pm * BM) * 256;
                            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"((uint32_t)(rows * 512)) : "memory");
                        }
                    }
                }
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sa = tiles + stage * L::STAGE_BYTES;
                    uint8_t* sb = sa + L::A_BYTES;
                    if constexpr (PAIR) {
                        // each CTA loads its 128 rows of A and its half of the B tile (64-point groups 2*rank, 2*rank+1);
                        // all bytes of the pair are counted on the LEADER's full barrier
                        static_assert(!PAIR || (B_MN && BN == 256), "CTA pairs: B is a 256-point T256 tile read MN-major");
                        if (elect_one()) {
                            const uint32_t bar = mapa_u32(smem_u32(&full[stage]), 0);
                            if (rank == 0) mbar_expect_tx(&full[stage], 2 * L::STAGE_BYTES);
                            if constexpr (!A_MN) {
                                tma_load_2d_pair(&tmA, bar, sa, kb * BK, m_blk * BM);
                            } else {
#pragma unroll
                                for (int i = 0; i < BM / 64; ++i)
                                    tma_load_2d_pair(&tmA, bar, sa + i * (BK * 128), m_blk * BM + i * 64, kb * BK);
                            }
#pragma unroll
                            for (int i = 0; i < BN / 128; ++i)
                                tma_load_3d_pair(&tmB, bar, sb + i * (BK * 128), ((int)rank * (BN / 128) + i) * 64, kb * BK, n_blk);
                        }
                    } else if (elect_one()) {
                    mbar_expect_tx(&full[stage], L::STAGE_BYTES);
                    if constexpr (!A_MN) {
                        // tiled: k = points, 4 k blocks per 256-point tile, rows = channels
                        if (p.a_tiled) tma_load_3d(&tmA, &full[stage], sa, (kb & 3) * BK, m_blk * BM, kb >> 2);
                        else tma_load_2d(&tmA, &full[stage], sa, kb * BK, m_blk * BM);
                    } else {
#pragma unroll
                        for (int i = 0; i < BM / 64; ++i)
                            tma_load_2d(&tmA, &full[stage], sa + i * (BK * 128), m_blk * BM + i * 64, kb * BK);
                    }
                    if constexpr (!B_MN) {
                        if (p.b_tiled) tma_load_3d(&tmB, &full[stage], sb, (kb & 3) * BK, n_blk * BN, kb >> 2);
                        else tma_load_2d(&tmB, &full[stage], sb, kb * BK, n_blk * BN);
                    } else if (p.b_tiled) {
                        // n = points: the n block IS the 256-point tile (BN == 256), k rows = channels
#pragma unroll
                        for (int i = 0; i < BN / 64; ++i)
                            tma_load_3d(&tmB, &full[stage], sb + i * (BK * 128), i * 64, kb * BK, n_blk);
                    } else {
#pragma unroll
                        for (int i = 0; i < BN / 64; ++i)
                            tma_load_2d(&tmB, &full[stage], sb + i * (BK * 128), n_blk * BN + i * 64, kb * BK);
                    }
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer (warp-converged, one issuing lane)
        if (rank == 0) {
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) |
                                       ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) |
                                       ((uint32_t)(((PAIR ? 2 : 1) * BM) >> 4) << 24);     // pair: one 256-row UMMA
            // K-major: 8-row groups 1024 B apart (SBO), LBO unused.  MN-major: 64-element column groups BK*128 B apart
            // (LBO), 8-k-row groups 1024 B apart (SBO).
            constexpr uint32_t A_LBO = A_MN ? BK * 128 : 16, B_LBO = B_MN ? BK * 128 : 16;
            constexpr uint32_t A_KSTEP = A_MN ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
            constexpr uint32_t B_KSTEP = B_MN ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
            int stage = 0;
            uint32_t phase = 0;
            int m_blk, n_blk, ks;
            for (int it = 0; tile_at(p, it, m_blk, n_blk, ks); ++it) {
                const int kb0 = ks * p.kb_per_split;
                const int kb1 = min(p.kb_total, kb0 + p.kb_per_split);
                const int buf = it & 1;
                mbar_wait(&tempty[buf], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * BN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(tiles + stage * L::STAGE_BYTES);
                    const uint32_t sb = sa + L::A_BYTES;
                    const uint64_t da = make_desc(sa, A_LBO, 1024);
                    const uint64_t db = make_desc(sb, B_LBO, 1024);
                    if (elect_one()) {
                        if constexpr (PAIR) {
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; ++k)
                                tc_mma_bf16_pair(tmem_d, da + (uint64_t)(k * A_KSTEP), db + (uint64_t)(k * B_KSTEP), idesc,
                                                 (kb > kb0 || k > 0) ? 1u : 0u);
                            tc_commit_pair(&empty[stage]);     // frees the slot in both CTAs
                            if (kb + 1 == kb1) tc_commit_pair(&tfull[buf]);
                        } else {
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; ++k)
                                tc_mma_bf16(tmem_d, da + (uint64_t)(k * A_KSTEP), db + (uint64_t)(k * B_KSTEP), idesc,
                                            (kb > kb0 || k > 0) ? 1u : 0u);
                            tc_commit(&empty[stage]);          // smem slot free once these MMAs retire
                            if (kb + 1 == kb1) tc_commit(&tfull[buf]);   // ... and the accumulator is ready for the epilogue
                        }
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if constexpr (is_t_mode<MODE>()) {
        // ===================================================================== epilogue warps, channel-major modes
        // thread = one output row (channel): its bias / BatchNorm coefficients are scalars, its statistics are plain
        // running sums kept in registers over all tiles of the CTA (the scheduler keeps the m block fixed) and
        // flushed with ONE double atomic per row at the end.  A warp owns 32 rows x 128 columns of the tile and moves
        // them as two 64-column sub-tiles: TMEM -> registers -> 128B-swizzled shared memory -> TMA store (and, for the
        // data gradient, yprev sub-tiles arrive by TMA one step ahead).  TMEM loads are software-pipelined.
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const int ew = warp - 2;                         // 0..7
        constexpr int CPW = BN / 64;                     // 32-column chunks per warp per tile (4)
        uint8_t* my_o = obuf + ew * SUB_BYTES;
        uint8_t* my_y = ybuf + ew * SUB_BYTES;
        const uint32_t sw = (uint32_t)(lane & 7);        // 128B swizzle: 16-byte chunk j of row r sits at chunk j ^ (r & 7)
        const uint32_t row_off = (uint32_t)lane * 128u;
        int cur_m = -1;
        int64_t row = 0;
        bool row_ok = false;
        TRow tr{};
        double d1 = 0.0, d2 = 0.0;
        uint32_t yphase = 0;
        int m_blk, n_blk, ks;
        if constexpr (MODE == MODE_T_DGRAD_ELUBN) {
            // first yprev sub-tile of this warp
            if (lane == 0 && tile_at(p, 0, m_blk, n_blk, ks)) {
                mbar_expect_tx(&ybar[ew], SUB_BYTES);
                tma_load_3d(&tmY, &ybar[ew], my_y, half * (BN / 2), m_blk * BM + q * 32, n_blk);
            }
        }
        for (int it = 0; tile_at(p, it, m_blk, n_blk, ks); ++it) {
            if (m_blk != cur_m) {
                if (cur_m >= 0 && row_ok && has_row_stats<MODE>()) {
                    atomicAdd(&p.stats[row], d1);
                    atomicAdd(&p.stats[p.M + row], d2);
                }
                d1 = d2 = 0.0;
                cur_m = m_blk;
                row = (int64_t)m_blk * BM + q * 32 + lane;
                row_ok = row < p.M;
                if (row_ok) {
                    if constexpr (MODE == MODE_T_BIAS_STATS) tr.bias = p.bias ? p.bias[row] : 0.f;
                    if constexpr (MODE == MODE_T_AFFINE_ELU || MODE == MODE_T_AFFINE_ELU_POOL) {
                        tr.scale = p.scale[row];
                        tr.shift = p.bias ? fmaf(p.bias[row], tr.scale, p.shift[row]) : p.shift[row];
                    }
                    if constexpr (MODE == MODE_T_DGRAD_ELUBN) {
                        tr.scale = p.scale[row]; tr.shift = p.shift[row];
                        tr.invstd = p.invstd[row]; tr.nmean_invstd = -p.mean[row] * tr.invstd;
                    }
                    tr.scale_l2 = tr.scale * LOG2E_F;
                    tr.shift_l2 = tr.shift * LOG2E_F;
                }
            }
            const int buf = it & 1;
            const int64_t n0 = (int64_t)n_blk * BN + half * (BN / 2);
            mbar_wait(&tfull[buf], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + half * (BN / 2);
            uint32_t ra[32], rb[32];
            uint4 yraw[8];
            float t1 = 0.f, t2 = 0.f;
            // pooled mode: running sum of the current group of pool_n points (the groups' boundaries depend on the column
            // only: warp-uniform) -- the activation tile is never written, only the [groups, M] means are
            int64_t pool_g = 0;
            int pool_left = 0;
            float pool_acc = 0.f;
            if constexpr (MODE == MODE_T_AFFINE_ELU_POOL) {
                pool_g = n0 / p.pool_n;
                pool_left = (int)((pool_g + 1) * p.pool_n - n0);
            }
            tmem_ld32_issue(taddr, ra);
#pragma unroll
            for (int c = 0; c < CPW; ++c) {
                uint32_t (&r)[32] = (c & 1) ? rb : ra;
                uint32_t (&rn)[32] = (c & 1) ? ra : rb;
                const int64_t col0 = n0 + c * 32;
                if constexpr (MODE == MODE_T_DGRAD_ELUBN) {
                    if ((c & 1) == 0) {
                        // the 64-column yprev sub-tile of this step: shared memory -> registers, then refill the buffer
                        mbar_wait(&ybar[ew], yphase);
                        yphase ^= 1;
#pragma unroll
                        for (int g = 0; g < 8; ++g)
                            yraw[g] = *reinterpret_cast<const uint4*>(my_y + row_off + (((uint32_t)g ^ sw) << 4));
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            int nm = m_blk, nn = n_blk, nk, nsub = (c >> 1) + 1;
                            bool more = true;
                            if (nsub == CPW / 2) { nsub = 0; more = tile_at(p, it + 1, nm, nn, nk); }
                            if (more) {
                                mbar_expect_tx(&ybar[ew], SUB_BYTES);
                                tma_load_3d(&tmY, &ybar[ew], my_y, half * (BN / 2) + nsub * 64, nm * BM + q * 32, nn);
                            }
                        }
                    }
                }
                tmem_ld_wait(r);
                if (c + 1 < CPW) tmem_ld32_issue(taddr + (c + 1) * 32, rn);
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                if (col0 + 32 <= p.N) t_chunk<MODE, true>(v, &yraw[(c & 1) * 4], tr, 32, t1, t2);
                else t_chunk<MODE, false>(v, &yraw[(c & 1) * 4], tr, (int)max((int64_t)0, p.N - col0), t1, t2);
                if constexpr (MODE == MODE_T_AFFINE_ELU_POOL) {
                    if (pool_left > 32) {
                        float s0 = 0.f, s1 = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; j += 2) { s0 += v[j]; s1 += v[j + 1]; }
                        pool_acc += s0 + s1;
                        pool_left -= 32;
                    } else {
                        // pool_n >= 32: one group boundary in this chunk, after its first pool_left columns
                        float s0 = 0.f, s1 = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float a = j < pool_left ? v[j] : 0.f;
                            s0 += a;
                            s1 += v[j] - a;
                        }
                        pool_acc += s0;
                        if (row_ok && (pool_g + 1) * p.pool_n <= p.N)
                            atomicAdd(reinterpret_cast<float*>(p.out) + pool_g * p.M + row, pool_acc * p.pool_inv_n);
                        ++pool_g;
                        pool_acc = s1;
                        pool_left = p.pool_n - (32 - pool_left);
                    }
                    if (c + 1 == CPW && row_ok && (pool_g + 1) * p.pool_n <= p.N)
                        atomicAdd(reinterpret_cast<float*>(p.out) + pool_g * p.M + row, pool_acc * p.pool_inv_n);
                    continue;
                }
                if ((c & 1) == 0) {
                    // the previous TMA store must have finished READING the staging buffer before it is overwritten
                    if (lane == 0) tma_store_wait_read();
                    __syncwarp();
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 u;
                    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                    for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[g * 8 + 2 * e], v[g * 8 + 2 * e + 1]);
                    *reinterpret_cast<uint4*>(my_o + row_off + (((uint32_t)((c & 1) * 4 + g) ^ sw) << 4)) = u;
                }
                if (c & 1) {
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        // rows >= M are clipped by the tensor map
                        tma_store_3d(&tmO, my_o, half * (BN / 2) + (c >> 1) * 64, m_blk * BM + q * 32, n_blk);
                        tma_store_commit();
                    }
                }
            }
            // all TMEM reads of this accumulator are complete -> hand the buffer back to the (leader's) MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[buf]), 0));
            d1 += (double)t1;
            d2 += (double)t2;
        }
        if (cur_m >= 0 && row_ok && has_row_stats<MODE>()) {
            atomicAdd(&p.stats[row], d1);
            atomicAdd(&p.stats[p.M + row], d2);
        }
        if (lane == 0) tma_store_wait_all();
    } else {
        // ===================================================================== epilogue warps
        const int q = warp & 3;                         // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;               // which half of the tile's columns this warp owns
        const int et = threadIdx.x - 64;                // 0..255
        constexpr int CHUNKS = BN / 32;
        const int c_begin = half * (CHUNKS / 2), c_end = c_begin + CHUNKS / 2;
        float* my_s1 = wstat + q * 2 * BN;
        float* my_s2 = my_s1 + BN;
        int m_blk, n_blk, ks;
        for (int it = 0; tile_at(p, it, m_blk, n_blk, ks); ++it) {
            const int64_t n0 = (int64_t)n_blk * BN;
            const int64_t row = (int64_t)m_blk * BM + q * 32 + lane;
            const bool row_ok = row < p.M;
            const int buf = it & 1;
            // stage the per-column parameters of this tile
            if constexpr (MODE == MODE_BIAS_STATS || MODE == MODE_BIAS_ELU || MODE == MODE_PLAIN) {
                for (int c = et; c < BN; c += EPI_THREADS) colp[c] = (n0 + c < p.N && p.bias) ? p.bias[n0 + c] : 0.f;
            }
            if constexpr (MODE == MODE_DGRAD_ELUBN) {
                for (int c = et; c < BN; c += EPI_THREADS) {
                    const bool ok = n0 + c < p.N;
                    colp[BN + c] = ok ? p.scale[n0 + c] : 0.f;
                    colp[2 * BN + c] = ok ? p.shift[n0 + c] : 0.f;
                    colp[3 * BN + c] = ok ? p.mean[n0 + c] : 0.f;
                    colp[4 * BN + c] = ok ? p.invstd[n0 + c] : 0.f;
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            mbar_wait(&tfull[buf], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN;
#pragma unroll 1
            for (int c = c_begin; c < c_end; ++c) {
                const int64_t col0 = n0 + c * 32;
                // previous-layer activation tile (needed by the data-gradient epilogues): issue the loads before
                // the TMEM read so their latency overlaps it
                uint4 yraw[4];
                if constexpr (MODE == MODE_DGRAD_ELUBN || MODE == MODE_DGRAD_ELUOUT) {
                    const bool ok = row_ok && col0 < p.N;
                    const uint4* yp = reinterpret_cast<const uint4*>(p.yprev + (ok ? row * p.ldy + col0 : 0));
#pragma unroll
                    for (int g = 0; g < 4; ++g) yraw[g] = (ok && col0 + g * 8 < p.N) ? __ldg(yp + g) : make_uint4(0, 0, 0, 0);
                }
                uint32_t r[32];
                tmem_ld32(taddr + c * 32, r);
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                if constexpr (MODE == MODE_WGRAD) {
                    if (p.out_tma) {
                        // 32 rows x 32 fp32 columns (128-byte rows) of this warp -> swizzled staging buffer -> one TMA
                        // store (or reduce-add when k is split); rows >= M and columns >= N are clipped by the map
                        uint8_t* my_o = obuf + (warp - 2) * SUB_BYTES;
                        if (lane == 0) tma_store_wait_read();
                        __syncwarp();
#pragma unroll
                        for (int g = 0; g < 8; ++g)
                            *reinterpret_cast<float4*>(my_o + lane * 128 + ((g ^ (lane & 7)) << 4)) =
                                make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0 && col0 < p.N) {
                            if (p.wgrad_store) tma_store_2d(&tmO, my_o, (int)col0, (int)(m_blk * BM + q * 32));
                            else tma_reduce_add_2d(&tmO, my_o, (int)col0, (int)(m_blk * BM + q * 32));
                            tma_store_commit();
                        }
                    } else if (row_ok) {
                        float* o = reinterpret_cast<float*>(p.out) + row * p.ldo + col0;
                        if (p.out_scalar) {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (col0 + j < p.N) {
                                    if (p.wgrad_store) o[j] = v[j];
                                    else atomicAdd(o + j, v[j]);
                                }
                        } else if (p.wgrad_store) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                if (col0 + j < p.N) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                if (col0 + j < p.N) atomicAdd(reinterpret_cast<float4*>(o + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
                        }
                    }
                } else {
                    float s2v[32];
                    if constexpr (MODE == MODE_BIAS_STATS || MODE == MODE_PLAIN) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] += colp[c * 32 + j];
                    } else if constexpr (MODE == MODE_BIAS_ELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = elu_f(v[j] + colp[c * 32 + j]);
                    } else if constexpr (MODE == MODE_DGRAD_ELUBN || MODE == MODE_DGRAD_ELUOUT) {
                        float yv[32];
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&yraw[g]);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                float2 f = __bfloat1622float2(h[e]);
                                yv[g * 8 + 2 * e] = f.x;
                                yv[g * 8 + 2 * e + 1] = f.y;
                            }
                        }
                        if constexpr (MODE == MODE_DGRAD_ELUBN) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const int cc = c * 32 + j;
                                float z = fmaf(yv[j], colp[BN + cc], colp[2 * BN + cc]);
                                float g = v[j] * elu_grad_f(z);
                                float xh = (yv[j] - colp[3 * BN + cc]) * colp[4 * BN + cc];
                                v[j] = g;
                                s2v[j] = g * xh;
                            }
                        } else {
                            // ELU backward from the saved OUTPUT a = ELU(z): ELU'(z) = a > 0 ? 1 : a + 1
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] *= (yv[j] > 0.f ? 1.f : yv[j] + 1.f);
                        }
                    }
                    if (row_ok) {
                        if (p.out_f32) {
                            float* o = reinterpret_cast<float*>(p.out) + row * p.ldo + col0;
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                if (col0 + j < p.N) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        } else {
                            // bf16 store of this thread's 32 consecutive columns (64 B)
                            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ldo + col0;
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                if (col0 + g * 8 < p.N) {
                                    uint4 u;
                                    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                                    for (int e = 0; e < 4; ++e)
                                        h[e] = __floats2bfloat162_rn(v[g * 8 + 2 * e], v[g * 8 + 2 * e + 1]);
                                    *reinterpret_cast<uint4*>(o + g * 8) = u;
                                }
                            }
                        }
                    }
                    if constexpr (MODE == MODE_BIAS_STATS || MODE == MODE_DGRAD_ELUBN) {
                        if constexpr (MODE == MODE_BIAS_STATS) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                v[j] = row_ok ? v[j] : 0.f;
                                s2v[j] = v[j] * v[j];
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                v[j] = row_ok ? v[j] : 0.f;
                                s2v[j] = row_ok ? s2v[j] : 0.f;
                            }
                        }
                        float t1 = warp_col_reduce32(v, lane);
                        float t2 = warp_col_reduce32(s2v, lane);
                        my_s1[c * 32 + lane] = t1;
                        my_s2[c * 32 + lane] = t2;
                    }
                }
            }
            // all TMEM reads of this accumulator are complete -> hand the buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[buf]);
            // every epilogue warp is done with colp / has published its partial statistics
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if constexpr (MODE == MODE_BIAS_STATS || MODE == MODE_DGRAD_ELUBN) {
                for (int c = et; c < 2 * BN; c += EPI_THREADS) {
                    const int which = c / BN, cc = c % BN;
                    if (n0 + cc < p.N) {
                        float s = wstat[(0 * 2 + which) * BN + cc] + wstat[(1 * 2 + which) * BN + cc] +
                                  wstat[(2 * 2 + which) * BN + cc] + wstat[(3 * 2 + which) * BN + cc];
                        atomicAdd(&p.stats[(int64_t)which * p.N + n0 + cc], (double)s);
                    }
                }
            }
        }
    }
    if constexpr (MODE == MODE_WGRAD) {
        if (warp >= 2 && lane == 0) tma_store_wait_all();
    }
    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all();      // neither CTA may exit (or free TMEM) while the other still signals it
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if constexpr (PAIR)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)f;
        else
            cudaGetLastError();
    }
    return fn;
}

// 2-D bf16 tensor map: inner (contiguous) extent d0, outer extent d1, outer stride ld elements; box = b0 x b1
static int make_map(CUtensorMap* m, const void* ptr, int64_t d0, int64_t d1, int64_t ld, int b0, int b1) {
    EncodeTiledFn enc = get_encode();
    PCAA_REQUIRE(enc != nullptr, PCAA_ERR_DRIVER, "cuTensorMapEncodeTiled is unavailable (driver too old?)");
    PCAA_REQUIRE(((uintptr_t)ptr & 15) == 0 && (ld * 2) % 16 == 0, PCAA_ERR_ALIGN,
                 "tensor-core operand must be 16-byte aligned with a leading dimension multiple of 8 (ld=%lld)",
                 (long long)ld);
    cuuint64_t dims[2] = {(cuuint64_t)d0, (cuuint64_t)d1};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)b0, (cuuint32_t)b1};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PCAA_REQUIRE(r == CUDA_SUCCESS, PCAA_ERR_DRIVER, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return PCAA_OK;
}

// 2-D fp32 tensor map of a row-major output [rows, cols] (ld elements): box = 32 columns (128 bytes) x 32 rows
static int make_map_out_f32(CUtensorMap* m, const void* ptr, int64_t cols, int64_t rows, int64_t ld) {
    EncodeTiledFn enc = get_encode();
    PCAA_REQUIRE(enc != nullptr, PCAA_ERR_DRIVER, "cuTensorMapEncodeTiled is unavailable (driver too old?)");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PCAA_REQUIRE(r == CUDA_SUCCESS, PCAA_ERR_DRIVER, "cuTensorMapEncodeTiled (fp32 output) failed (%d)", (int)r);
    return PCAA_OK;
}

// 3-D bf16 tensor map of a tiled operand [n_tiles][C][256]: box = b0 points x b1 channels x 1 tile
static int make_map_tiled(CUtensorMap* m, const void* ptr, int64_t C, int64_t n_tiles, int b0, int b1) {
    PCAA_REQUIRE(ptr != nullptr, PCAA_ERR_SHAPE, "gemm_tc: null T256 tensor");
    EncodeTiledFn enc = get_encode();
    PCAA_REQUIRE(enc != nullptr, PCAA_ERR_DRIVER, "cuTensorMapEncodeTiled is unavailable (driver too old?)");
    PCAA_REQUIRE(((uintptr_t)ptr & 15) == 0, PCAA_ERR_ALIGN, "tiled tensor-core operand must be 16-byte aligned");
    cuuint64_t dims[3] = {256, (cuuint64_t)C, (cuuint64_t)n_tiles};
    cuuint64_t strides[2] = {256 * 2, (cuuint64_t)C * 256 * 2};
    cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PCAA_REQUIRE(r == CUDA_SUCCESS, PCAA_ERR_DRIVER, "cuTensorMapEncodeTiled (3-D) failed (%d)", (int)r);
    return PCAA_OK;
}

static int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <int BN, bool A_MN, bool B_MN, int MODE>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& ty,
                     const GemmParams& p, cudaStream_t st) {
    auto kern = gemm_tc_kernel<BN, A_MN, B_MN, MODE>;
    using L = SmemLayout<BN, MODE>;
    static_assert(L::TOTAL <= 232448, "shared memory budget of one CTA exceeded");
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
        if (e != cudaSuccess) {
            set_error("gemm_tc: cannot reserve %d bytes of shared memory: %s", L::TOTAL,
                      cudaGetErrorString(e));
            return PCAA_ERR_LAUNCH;
        }
        attr = true;
    }
    int items = p.m_tiles * p.n_tiles * p.k_splits;
    int grid = items < num_sms() ? items : num_sms();
    if constexpr (L::PAIR) {
        // CTA pairs (clusters of 2, one per TPC): m_pairs clusters per n block, at most one cluster per SM pair and
        // no more column groups than there are n tiles
        static int max_clusters = 0;
        if (max_clusters == 0) {
            static bool attr2 = false;
            if (!attr2) { cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 0); attr2 = true; }
            cudaLaunchConfig_t q{};
            q.gridDim = dim3(num_sms()); q.blockDim = dim3(NUM_THREADS); q.dynamicSmemBytes = L::TOTAL;
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension; qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
            q.attrs = qa; q.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, kern, &q) != cudaSuccess || n <= 0) { cudaGetLastError(); n = num_sms() / 2; }
            max_clusters = n < num_sms() / 2 ? n : num_sms() / 2;
        }
        const int m_pairs = (p.m_tiles + 1) / 2;
        int per = max_clusters / m_pairs;
        if (per > p.n_tiles) per = p.n_tiles;
        if (per < 1) {
            set_error("gemm_tc: %d row-block pairs exceed the resident clusters (channel-major modes)", m_pairs);
            return PCAA_ERR_UNSUPPORTED;
        }
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(2 * per * m_pairs);
        cfg.blockDim = dim3(NUM_THREADS);
        cfg.dynamicSmemBytes = L::TOTAL;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, to, ty, p);
        if (e != cudaSuccess) {
            set_error("gemm_tc: cluster launch failed: %s", cudaGetErrorString(e));
            cudaGetLastError();
            return PCAA_ERR_LAUNCH;
        }
        return check_launch("gemm_tc");
    }
    kern<<<grid, NUM_THREADS, L::TOTAL, st>>>(ta, tb, to, ty, p);
    return check_launch("gemm_tc");
}

}  // namespace pcaa

using namespace pcaa;

// A(m,k): a_mn == 0 -> stored [M, K] (lda), else stored [K, M] (lda).  B(n,k): b_mn == 0 -> [N, K], else [K, N].
static int gemm_tc_dispatch(const void* A, int64_t lda, int a_layout, const void* B, int64_t ldb, int b_layout, void* out,
                            int64_t ldo, int out_dtype, int64_t M, int64_t N, int64_t K, int mode, const float* bias,
                            double* stats, const void* yprev, int64_t ldy, const float* scale, const float* shift,
                            const float* mean, const float* invstd, cudaStream_t st) {
    if (M == 0 || N == 0) return PCAA_OK;
    PCAA_REQUIRE(a_layout >= 0 && a_layout <= 3 && b_layout >= 0 && b_layout <= 3, PCAA_ERR_UNSUPPORTED, "gemm_tc: bad operand layout");
    const int a_mn = a_layout & 1, b_mn = b_layout & 1;
    const bool a_tiled = a_layout >= 2, b_tiled = b_layout >= 2;
    PCAA_REQUIRE(M > 0 && N > 0 && K > 0, PCAA_ERR_SHAPE, "gemm_tc: bad shape");
    PCAA_REQUIRE(mode >= 0 && mode <= PCAA_TC_T_AFFINE_ELU_POOL, PCAA_ERR_UNSUPPORTED, "gemm_tc: unknown mode %d", mode);
    const bool tmode = mode >= PCAA_TC_T_BIAS_STATS;
    const bool pooled = mode == PCAA_TC_T_AFFINE_ELU_POOL;
    const bool wgrad = (mode == PCAA_TC_WGRAD_ACC || mode == PCAA_TC_WGRAD_STORE);
    const bool f32out = !pooled && (wgrad || out_dtype == PCAA_F32);
    if (pooled) {
        // out = fp32 [N / ldo, M] group means, ldo = points per group
        PCAA_REQUIRE(out_dtype == PCAA_F32 && ldo >= 32 && N % ldo == 0, PCAA_ERR_SHAPE,
                     "gemm_tc: pooled mode writes fp32 [N / n, M] means for groups of n = ldo >= 32 points, n | N (n=%lld, N=%lld)",
                     (long long)ldo, (long long)N);
        if (cudaMemsetAsync(out, 0, sizeof(float) * (size_t)(N / ldo) * (size_t)M, st) != cudaSuccess) return check_launch("gemm_tc memset");
    }
    // bf16 rows are written in 16-byte groups: the buffer must have ldo >= round_up(N, 8) (pad columns receive the
    // epilogue of zero accumulators); fp32 rows use 16-byte stores when aligned, scalar stores otherwise (wgrad only)
    const bool scalar_out = f32out && (((uintptr_t)out & 15) != 0 || ldo % 4 != 0 || N % 4 != 0);
    PCAA_REQUIRE(!scalar_out || wgrad, PCAA_ERR_ALIGN, "gemm_tc: fp32 output needs 16-byte aligned rows (ldo %% 4 == 0, N %% 4 == 0)");
    PCAA_REQUIRE(f32out || tmode || (((uintptr_t)out & 15) == 0 && ldo % 8 == 0 && ldo >= (N + 7) / 8 * 8), PCAA_ERR_ALIGN,
                 "gemm_tc: bf16 output needs 16-byte aligned rows with ldo >= round_up(N, 8)");
    if (tmode)
        PCAA_REQUIRE(((uintptr_t)out & 15) == 0 && b_layout == PCAA_OP_T256_MN && !a_tiled, PCAA_ERR_UNSUPPORTED,
                     "gemm_tc: channel-major modes take B as 256-point tiles (PCAA_OP_T256_MN) and write tiled output");
    PCAA_REQUIRE(!a_tiled || a_layout == PCAA_OP_T256_K, PCAA_ERR_UNSUPPORTED, "gemm_tc: a tiled A operand must be PCAA_OP_T256_K");
    if (mode == PCAA_TC_BIAS_STATS || mode == PCAA_TC_DGRAD_ELUBN || mode == PCAA_TC_T_BIAS_STATS || mode == PCAA_TC_T_DGRAD_ELUBN)
        PCAA_REQUIRE(stats != nullptr, PCAA_ERR_SHAPE, "gemm_tc: stats buffer required for mode %d", mode);
    if (mode == PCAA_TC_DGRAD_ELUBN || mode == PCAA_TC_T_DGRAD_ELUBN)
        PCAA_REQUIRE(yprev && scale && shift && mean && invstd, PCAA_ERR_SHAPE, "gemm_tc: mode %d needs yprev/coefficients", mode);
    if (mode == PCAA_TC_T_AFFINE_ELU || pooled) PCAA_REQUIRE(scale && shift, PCAA_ERR_SHAPE, "gemm_tc: modes 8 / 10 need scale/shift");
    if (tmode && !pooled) PCAA_REQUIRE(out_dtype == PCAA_BF16, PCAA_ERR_UNSUPPORTED, "gemm_tc: channel-major modes store bf16");
    if (mode == PCAA_TC_DGRAD_ELUOUT) PCAA_REQUIRE(yprev != nullptr, PCAA_ERR_SHAPE, "gemm_tc: mode 5 needs the saved activation");
    if (yprev) PCAA_REQUIRE(((uintptr_t)yprev & 15) == 0 && (tmode || ldy % 8 == 0), PCAA_ERR_ALIGN, "gemm_tc: yprev alignment");
    // Output tile width.  256 columns everywhere, except for the row-major epilogue modes (decoder / TCN forward and data
    // gradient) when 256-wide tiles would leave more than half of the SMs without a tile: the decoder's inner layers at
    // batch 256 are M = 256 rows x N = 4500 / 9000 columns = 36 / 72 tiles of 128 x 256 on 148 SMs; 128-wide tiles double
    // the CTAs that stream the weight matrix.
    const int m_tiles_ = ceil_div(M, BM);
    const bool narrow = !tmode && !wgrad && (long long)m_tiles_ * ceil_div(N, 256) * 2 <= num_sms() && N > 128;
    const int BN = narrow ? 128 : 256;
    CUtensorMap ta, tb;
    int rc;
    if (a_tiled) rc = make_map_tiled(&ta, A, M, ceil_div(K, 256), BK, BM);                 // [K/256 tiles][M][256], k = points
    else rc = a_mn ? make_map(&ta, A, M, K, lda, 64, BK) : make_map(&ta, A, K, M, lda, BK, BM);
    if (rc) return rc;
    if (b_tiled && b_mn) rc = make_map_tiled(&tb, B, K, ceil_div(N, 256), 64, BK);         // [N/256 tiles][K][256], n = points
    else if (b_tiled) rc = make_map_tiled(&tb, B, N, ceil_div(K, 256), BK, BN);            // [K/256 tiles][N][256], k = points
    else rc = b_mn ? make_map(&tb, B, N, K, ldb, 64, BK) : make_map(&tb, B, K, N, ldb, BK, BN);
    if (rc) return rc;
    PCAA_REQUIRE(a_tiled == (b_tiled && !b_mn), PCAA_ERR_UNSUPPORTED, "gemm_tc: k = points needs BOTH operands tiled (PCAA_OP_T256_K)");
    // channel-major modes: the epilogue moves 64-point x 32-channel sub-tiles of out (and yprev) by TMA
    CUtensorMap to = ta, ty = ta;
    if (tmode && !pooled) {
        rc = make_map_tiled(&to, out, M, ceil_div(N, 256), 64, 32);
        if (rc) return rc;
        if (mode == PCAA_TC_T_DGRAD_ELUBN) {
            rc = make_map_tiled(&ty, yprev, M, ceil_div(N, 256), 64, 32);
            if (rc) return rc;
        }
    }
    const bool out_tma = wgrad && ((uintptr_t)out & 15) == 0 && (ldo * 4) % 16 == 0;
    if (out_tma) {
        rc = make_map_out_f32(&to, out, N, M, ldo);
        if (rc) return rc;
    }
    GemmParams p{};
    p.out_tma = out_tma ? 1 : 0;
    p.M = M;
    p.N = N;
    p.m_tiles = ceil_div(M, BM);
    p.n_tiles = ceil_div(N, BN);
    p.kb_total = a_tiled ? ceil_div(K, 256) * 4 : ceil_div(K, BK);     // tiled k: whole 256-point tiles (pad points are zeros)
    p.a_tiled = a_tiled ? 1 : 0;
    p.b_tiled = b_tiled ? 1 : 0;
    p.k_splits = 1;
    p.kb_per_split = p.kb_total;
    if (mode == PCAA_TC_WGRAD_ACC) {
        int tiles = p.m_tiles * p.n_tiles;
        int splits = num_sms() / tiles;
        if (splits < 1) splits = 1;
        if (splits > p.kb_total) splits = p.kb_total;
        p.kb_per_split = ceil_div(p.kb_total, splits);
        p.k_splits = ceil_div(p.kb_total, p.kb_per_split);
    }
    p.out = out;
    p.ldo = ldo;
    p.out_f32 = f32out ? 1 : 0;
    p.wgrad_store = mode == PCAA_TC_WGRAD_STORE;
    p.out_scalar = scalar_out ? 1 : 0;
    p.bias = bias;
    p.stats = stats;
    p.yprev = (const __nv_bfloat16*)yprev;
    p.ldy = ldy;
    p.scale = scale;
    p.shift = shift;
    p.mean = mean;
    p.invstd = invstd;
    p.sched_mfixed = tmode ? 1 : 0;
    p.pool_n = pooled ? (int)ldo : 0;
    p.pool_inv_n = pooled ? 1.f / (float)ldo : 0.f;
    const int key = (a_mn ? 2 : 0) | (b_mn ? 1 : 0);
    if (narrow) {
        if (key == 0) {
            switch (mode) {
                case PCAA_TC_BIAS_STATS: return launch_tc<128, false, false, MODE_BIAS_STATS>(ta, tb, to, ty, p, st);
                case PCAA_TC_BIAS_ELU: return launch_tc<128, false, false, MODE_BIAS_ELU>(ta, tb, to, ty, p, st);
                case PCAA_TC_PLAIN: return launch_tc<128, false, false, MODE_PLAIN>(ta, tb, to, ty, p, st);
                case PCAA_TC_DGRAD_ELUBN: return launch_tc<128, false, false, MODE_DGRAD_ELUBN>(ta, tb, to, ty, p, st);
                default: break;
            }
        } else if (key == 1) {
            switch (mode) {
                case PCAA_TC_PLAIN: return launch_tc<128, false, true, MODE_PLAIN>(ta, tb, to, ty, p, st);
                case PCAA_TC_DGRAD_ELUOUT: return launch_tc<128, false, true, MODE_DGRAD_ELUOUT>(ta, tb, to, ty, p, st);
                default: break;
            }
        }
        set_error("gemm_tc: operand layout (a=%d, b=%d) is not instantiated for mode %d (128-wide tiles)", a_layout, b_layout, mode);
        return PCAA_ERR_UNSUPPORTED;
    }
    if (key == 0) {
        if (wgrad) return launch_tc<256, false, false, MODE_WGRAD>(ta, tb, to, ty, p, st);
        switch (mode) {
            case PCAA_TC_BIAS_STATS: return launch_tc<256, false, false, MODE_BIAS_STATS>(ta, tb, to, ty, p, st);
            case PCAA_TC_BIAS_ELU: return launch_tc<256, false, false, MODE_BIAS_ELU>(ta, tb, to, ty, p, st);
            case PCAA_TC_PLAIN: return launch_tc<256, false, false, MODE_PLAIN>(ta, tb, to, ty, p, st);
            case PCAA_TC_DGRAD_ELUBN: return launch_tc<256, false, false, MODE_DGRAD_ELUBN>(ta, tb, to, ty, p, st);
            default: break;
        }
    } else if (key == 1) {
        switch (mode) {
            case PCAA_TC_PLAIN: return launch_tc<256, false, true, MODE_PLAIN>(ta, tb, to, ty, p, st);
            case PCAA_TC_DGRAD_ELUOUT: return launch_tc<256, false, true, MODE_DGRAD_ELUOUT>(ta, tb, to, ty, p, st);
            case PCAA_TC_T_BIAS_STATS: return launch_tc<256, false, true, MODE_T_BIAS_STATS>(ta, tb, to, ty, p, st);
            case PCAA_TC_T_AFFINE_ELU: return launch_tc<256, false, true, MODE_T_AFFINE_ELU>(ta, tb, to, ty, p, st);
            case PCAA_TC_T_AFFINE_ELU_POOL: return launch_tc<256, false, true, MODE_T_AFFINE_ELU_POOL>(ta, tb, to, ty, p, st);
            default: break;
        }
    } else if (key == 3) {
        if (wgrad) return launch_tc<256, true, true, MODE_WGRAD>(ta, tb, to, ty, p, st);
        if (mode == PCAA_TC_T_DGRAD_ELUBN) return launch_tc<256, true, true, MODE_T_DGRAD_ELUBN>(ta, tb, to, ty, p, st);
    }
    set_error("gemm_tc: operand layout (a=%d, b=%d) is not instantiated for mode %d", a_layout, b_layout, mode);
    return PCAA_ERR_UNSUPPORTED;
}

extern "C" int pcaa_gemm_tc(const void* A, int64_t lda, int a_mn, const void* B, int64_t ldb, int b_mn, void* out,
                            int64_t ldo, int out_dtype, int64_t M, int64_t N, int64_t K, int mode, const float* bias,
                            double* stats, const void* yprev, int64_t ldy, const float* scale, const float* shift,
                            const float* mean, const float* invstd, pcaa_stream stream) {
    return gemm_tc_dispatch(A, lda, a_mn, B, ldb, b_mn, out, ldo, out_dtype, M, N, K, mode, bias, stats, yprev, ldy, scale,
                            shift, mean, invstd, (cudaStream_t)stream);
}

extern "C" int pcaa_gemm_tc_tn(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo, int64_t M,
                               int64_t N, int64_t K, int mode, const float* bias, double* stats, const void* yprev,
                               const float* scale, const float* shift, const float* mean, const float* invstd,
                               pcaa_stream stream) {
    PCAA_REQUIRE(mode >= 0 && mode <= 3, PCAA_ERR_UNSUPPORTED, "gemm_tc_tn: unknown mode %d", mode);
    return gemm_tc_dispatch(A, lda, 0, W, ldw, 0, out, ldo, PCAA_BF16, M, N, K, mode, bias, stats, yprev, N, scale, shift,
                            mean, invstd, (cudaStream_t)stream);
}

extern "C" int pcaa_gemm_tc_nt_wgrad(const void* A, int64_t lda, const void* B, int64_t ldb, float* dW, int64_t ldw,
                                     int64_t N1, int64_t N2, int64_t K, pcaa_stream stream) {
    return gemm_tc_dispatch(A, lda, 1, B, ldb, 1, dW, ldw, PCAA_F32, N1, N2, K, PCAA_TC_WGRAD_ACC, nullptr, nullptr, nullptr,
                            0, nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}
