// gemm_tcgen05.cu -- hand-written sm_100a GEMM for the dense side of the CTR models.
//
//   D[M,N] (+)= A[M,K] * B[N,K]^T      A, B bf16 K-major, fp32 accumulation in TMEM
//
// One 128x64 output tile per CTA (optionally one K-split of it), warp specialised:
//   warp 0      TMA producer   cp.async.bulk.tensor.2d (128B swizzle) -> 4-stage smem ring
//   warp 1      MMA issuer     one elected thread, tcgen05.mma.cta_group::1.kind::f16,
//                              128x64x16 per instruction, accumulator = 64 TMEM columns;
//                              tcgen05.commit frees the smem stage / signals the epilogue
//   warps 2..5  epilogue       tcgen05.ld 32x32b.x32 (each warp owns its TMEM lane quadrant)
//                              -> fused epilogue -> global
// Fused epilogues (the reference gets these from cuBLAS/cuDNN via TensorFlow, K6 in SURVEY 2.5):
//   EPI_FWD   relu, "ones" column (bias folded into the next layer's weights), bf16 store
//             plus a transposed bf16 copy (the batch-major operand of the dW GEMMs)
//   EPI_DX    relu mask from the forward activation, bf16 store (+ transposed copy)
//   EPI_DW    split-K partial sums, fp32 red.global.add
//   EPI_DX_FM fp32 store of the embedding gradient with the FM second-order term fused
//
// Every GEMM of the training step (3 forward, 3 dX, 3 dW) is this one kernel: all operands
// are kept K-major by writing transposed copies in the producing epilogue instead of using
// MN-major UMMA descriptors.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "pdl.cuh"
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>
#include <algorithm>

namespace {

constexpr int BM = 128, BK = 64;
constexpr int A_BYTES = BM * BK * 2;
// Tile width. The GEMMs of a step are bound by operand traffic out of L2, not by the tensor cores
// (measured: time ~ bytes each SM pulls / ~42 B/clk): a 128x64 tile re-reads (1/128 + 1/64) operand
// bytes per output element, a 128x128 tile (1/128 + 1/128), a third less.
//   BN = 64 : 4 stages x 24 KB = 96 KB, two CTAs per SM -- the small GEMMs (K = 448, 224 tiles)
//   BN = 128: 3 stages x 32 KB = 96 KB, two CTAs per SM -- the wide ones (fwd1, dX1, dW1)
template <int BN> constexpr int stages_for() { return BN == 64 ? 4 : 3; }
constexpr int NUM_THREADS = 192;

enum EpiMode : int { EPI_FWD = 0, EPI_DX = 1, EPI_DW = 2, EPI_DX_FM = 3 };

struct GemmEpi {
    int mode, relu, ones_col, fm_cols;
    int M, N, D, mn_major;   // mn_major: A is stored [K, M], B is stored [K, N] (MN contiguous)
    void* out; long long ldo;
    __nv_bfloat16* outT; long long ldoT;
    const __nv_bfloat16* mask; long long ldmask;
    const float* dlogit; const float* S; const float* emb; long long ldemb;
    int swap, mc;              // mc: CTAs per cluster along the N-tile axis that share (multicast) the A tile; <= 1: off
    unsigned long long* dbg;   // optional: %globaltimer stamps of CTA (0,0,0) [start, setup, first-full, mainloop, epilogue]
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded: a descriptor bug must surface as a wrong result / error, never as a hung GPU
__device__ __forceinline__ bool mbar_wait(uint64_t* b, uint32_t parity) {
    for (uint32_t it = 0; it < (1u << 24); ++it)
        if (mbar_try(b, parity)) return true;
    return false;
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);   // start address
    d |= (uint64_t)1 << 16;                         // leading byte offset (unused with swizzle)
    d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
    return d;
}
// MN-major, SWIZZLE_128B: the tile is [k rows][64 MN elements = 128 B]; 8 k-rows form a 1024-byte
// swizzle atom (stride byte offset), 64-element MN blocks are `lbo_bytes` apart (leading byte
// offset). Canonical form ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units.
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// A-tile multicast: the tile lands at the same shared-memory offset of every CTA in `mask` and completes the
// transaction bytes on each destination CTA's own barrier (same offset)
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1) : "memory");
}
// commit that arrives on the barrier at the same offset in every CTA of `mask` (stage released cluster-wide)
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 2)
exb_gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmT,
                        GemmEpi E, int num_k_blocks, int k_blocks_per_split) {
    static_assert(BN == 64 || BN == 128, "tile width");
    constexpr int STAGES = stages_for<BN>();
    constexpr int B_BYTES = BN * BK * 2;
    constexpr int WSTAGE = BN * 128, TSTAGE = BN * 64;     // per-warp epilogue staging: out tiles | outT tile
    static_assert(4 * (WSTAGE + TSTAGE) <= STAGES * (A_BYTES + B_BYTES), "epilogue staging aliases the stage memory");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte aligned bases; do not rely on the toolchain for it
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    exb::pdl_trigger();   // the next kernel of the step may be scheduled once all CTAs are resident
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // E.swap: N tiles vary fastest over the launch order (neighbouring CTAs share the A tile, not the B tile)
    const int m_blk = E.swap ? blockIdx.y : blockIdx.x, n_blk = E.swap ? blockIdx.x : blockIdx.y;
    const bool dbg = E.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
#define GSTAMP(i) do { if (dbg) { unsigned long long _t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t)); E.dbg[i] = _t; } } while (0)
    if (threadIdx.x == 0) GSTAMP(0);
    const int kb0 = blockIdx.z * k_blocks_per_split;
    const int kb1 = min(num_k_blocks, kb0 + k_blocks_per_split);
    const int nkb = kb1 - kb0;
    // E.mc > 1: launched as clusters of E.mc CTAs with the same M block and consecutive N blocks. Rank 0 loads the A
    // tile ONCE and multicasts it into every CTA of the cluster (A is 2/3 of the operand bytes of a 128x64 tile and
    // the GEMMs of the step are bound by operand traffic out of L2); each CTA loads its own B tile. A stage is
    // re-filled only when EVERY CTA of the cluster has consumed it: the MMA commits arrive on all CTAs' barriers.
    const int C = E.mc > 1 ? E.mc : 1;
    uint32_t crank = 0;
    if (C > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const uint16_t cmask = (uint16_t)((1u << C) - 1u);

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], (uint32_t)C); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 1) {   // TMEM: BN fp32 accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (C > 1) cluster_sync_all();     // every CTA's barriers exist before a peer multicasts into / arrives on them
    exb::pdl_wait();      // barriers, TMEM and tensor maps are ready; operands come from the previous kernel
    if (threadIdx.x == 0) GSTAMP(1);

    if (warp == 0) {
        if (lane == 0) {   // ===== TMA producer
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);
                mbar_expect_tx(&full[s], A_BYTES + B_BYTES);
                if (E.mn_major) {   // boxes of 64 MN elements x 64 k rows; the A tile is two MN blocks
                    if (C > 1) {
                        if (crank == 0) {
                            tma_load_2d_mc(sA + s * A_BYTES, &tmA, &full[s], m_blk * BM, (kb0 + i) * BK, cmask);
                            tma_load_2d_mc(sA + s * A_BYTES + A_BYTES / 2, &tmA, &full[s], m_blk * BM + 64, (kb0 + i) * BK, cmask);
                        }
                    } else {
                        tma_load_2d(sA + s * A_BYTES, &tmA, &full[s], m_blk * BM, (kb0 + i) * BK);
                        tma_load_2d(sA + s * A_BYTES + A_BYTES / 2, &tmA, &full[s], m_blk * BM + 64, (kb0 + i) * BK);
                    }
#pragma unroll
                    for (int h = 0; h < BN / 64; ++h)
                        tma_load_2d(sB + s * B_BYTES + h * 8192, &tmB, &full[s], n_blk * BN + 64 * h, (kb0 + i) * BK);
                    continue;
                }
                if (C > 1) {
                    if (crank == 0) tma_load_2d_mc(sA + s * A_BYTES, &tmA, &full[s], (kb0 + i) * BK, m_blk * BM, cmask);
                } else {
                    tma_load_2d(sA + s * A_BYTES, &tmA, &full[s], (kb0 + i) * BK, m_blk * BM);
                }
                tma_load_2d(sB + s * B_BYTES, &tmB, &full[s], (kb0 + i) * BK, n_blk * BN);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {   // ===== MMA issuer
            // instruction descriptor: D=f32, A=B=bf16, both K-major, N=64, M=128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
                mbar_wait(&full[s], ph);
                if (i == 0) GSTAMP(2);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a0 = smem_u32(sA + s * A_BYTES), b0 = smem_u32(sB + s * B_BYTES);
                if (E.mn_major) {
                    // both operands MN-major (bits 15/16); one UMMA_K = 16 k rows = two swizzle atoms = 2048 B
                    const uint32_t idesc_mn = idesc | (1u << 15) | (1u << 16);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        umma_bf16(tmem_base, umma_desc_mn(a0 + k * 2048, 8192), umma_desc_mn(b0 + k * 2048, 8192),
                                  idesc_mn, (i | k) ? 1u : 0u);
                } else {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)   // UMMA_K = 16 bf16 = 32 bytes inside the 128B swizzle row
                    umma_bf16(tmem_base, umma_desc(a0 + k * 32), umma_desc(b0 + k * 32), idesc, (i | k) ? 1u : 0u);
                }
                if (C > 1) umma_commit_mc(&empty[s], cmask);   // the stage is free once EVERY CTA of the cluster has read it
                else umma_commit(&empty[s]);       // smem stage reusable when these MMAs retire
            }
            umma_commit(tmem_full);                // accumulator complete
        }
    } else {
        // ===== epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32)
        // TMEM hands every thread one output ROW (32 consecutive columns per tcgen05.ld). The fused
        // math runs in that layout; operands it needs (relu mask / FM terms) are fetched BEFORE the
        // accumulator is awaited, so their latency hides behind the main loop. Results leave through
        // 128B-swizzled shared-memory tiles and ONE TMA store (or TMA reduce-add for split-K) per
        // warp: no per-thread global stores, full-line writes, asynchronous to the issuing warp.
        const int q = warp & 3;
        const int row0 = m_blk * BM + q * 32, row = row0 + lane;
        const bool rv = row < E.M;
        uint8_t* wstage = smem + (warp - 2) * (WSTAGE + TSTAGE);   // per warp: out tile(s) | outT tile
        uint8_t* tstage = wstage + WSTAGE;
        const bool f32out = (E.mode == EPI_DW || E.mode == EPI_DX_FM);
        bool ok = true, waited = false;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            const int n0 = n_blk * BN + c0;
            // ---- issue the global reads of this chunk
            uint4 mk[4];
            float4 e4[8], s4[8];
            float dl = 0.f;
            if (E.mode == EPI_DX && rv) {
                const uint4* mp = reinterpret_cast<const uint4*>(E.mask + (size_t)row * E.ldmask + n0);
#pragma unroll
                for (int j = 0; j < 4; ++j) mk[j] = mp[j];
            }
            const bool fm = (E.mode == EPI_DX_FM) && rv && (n0 + 31 < E.fm_cols);
            if (fm) {
                const float4* ep = reinterpret_cast<const float4*>(E.emb + (size_t)row * E.ldemb + n0);
                const float* sbase = E.S + (size_t)row * E.D;
                dl = E.dlogit[row];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    e4[j] = ep[j];
                    s4[j] = *reinterpret_cast<const float4*>(sbase + ((n0 + 4 * j) % E.D));
                }
            }
            if (!waited) {
                if (nkb > 0) ok = mbar_wait(tmem_full, 0);
                if (warp == 2 && lane == 0) GSTAMP(3);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                waited = true;
            }
            uint32_t v[32];
            if (nkb > 0 && ok) tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
            // ---- fused math, lane == output row
            const __nv_bfloat16* mh = reinterpret_cast<const __nv_bfloat16*>(mk);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float x = __uint_as_float(v[j]);
                const int n = n0 + j;
                if (E.mode == EPI_FWD) {
                    if (E.relu) x = fmaxf(x, 0.f);
                    if (n == E.ones_col) x = 1.f;
                    if (n >= E.N) x = 0.f;
                } else if (E.mode == EPI_DX) {
                    if (!rv || !(__bfloat162float(mh[j]) > 0.f) || n == E.ones_col || n >= E.N) x = 0.f;
                } else if (fm) {
                    const float* ef = reinterpret_cast<const float*>(e4);
                    const float* sf = reinterpret_cast<const float*>(s4);
                    x += dl * (sf[j] - ef[j]);
                }
                v[j] = __float_as_uint(x);
            }
            // ---- registers -> swizzled staging tile (rows of 128 bytes, 16-byte chunk index ^ (row & 7))
            if (f32out) {
                uint8_t* tile = wstage + (c0 >> 5) * 4096;       // [32 rows][32 fp32]
#pragma unroll
                for (int t = 0; t < 8; ++t)
                    *reinterpret_cast<uint4*>(tile + lane * 128 + ((t ^ (lane & 7)) << 4)) =
                        make_uint4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0 && E.fm_cols != -7) {
                    if (E.mode == EPI_DW)
                        asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                                     ::"l"(&tmO), "r"(smem_u32(tile)), "r"(n0), "r"(row0) : "memory");
                    else
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                     ::"l"(&tmO), "r"(smem_u32(tile)), "r"(n0), "r"(row0) : "memory");
                }
            } else {
                uint8_t* tile = wstage + (c0 >> 6) * 4096;        // [32 rows][64 bf16] per 64 columns
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    uint32_t pk[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(v[8 * t + 2 * u]), __uint_as_float(v[8 * t + 2 * u + 1]));
                        pk[u] = *reinterpret_cast<uint32_t*>(&h2);
                    }
                    const int chunk = ((c0 & 63) >> 3) + t;
                    *reinterpret_cast<uint4*>(tile + lane * 128 + ((chunk ^ (lane & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
                if (E.outT) {                                     // [64 n][32 rows] bf16: lanes along the row axis
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        *reinterpret_cast<__nv_bfloat16*>(tstage + (c0 + j) * 64 + lane * 2) = __float2bfloat16_rn(__uint_as_float(v[j]));
                }
            }
        }
        if (!f32out) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0 && E.fm_cols != -7) {
#pragma unroll
                for (int h = 0; h < BN / 64; ++h)
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                 ::"l"(&tmO), "r"(smem_u32(wstage + h * 4096)), "r"(n_blk * BN + 64 * h), "r"(row0) : "memory");
                if (E.outT)
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                 ::"l"(&tmT), "r"(smem_u32(tstage)), "r"(row0), "r"(n_blk * BN) : "memory");
            }
        }
        if (lane == 0) {
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncwarp();
        if (warp == 2 && lane == 0) GSTAMP(4);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) GSTAMP(5);
    if (C > 1) cluster_sync_all();     // no CTA leaves while a peer may still multicast into / arrive on its memory
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN) : "memory");
    }
}

// =====================================================================================================
// Persistent GEMM CHAIN: several dependent GEMMs of the training step in ONE launch.
//
// The dense side of a CTR step is nine small GEMMs (M = batch 4096, N / K in {448, 1728}); launched one by one
// every kernel pays launch + barrier init + TMEM allocation + pipeline fill + epilogue drain + a partial last
// wave for 5-30 us of tensor work, and a layer cannot start before the previous kernel has fully drained.
// Here the tiles of all GEMMs of a chain (forward: fwd1 -> fwd2 -> fwd3; backward: dX3, dW3, dX2, dW2, dX1, dW1)
// form ONE static work list consumed by persistent CTAs (two per SM):
//   * set-up once per CTA; the TMA -> MMA smem ring and its phases run on ACROSS tiles;
//   * TMEM holds TWO accumulators: the epilogue of tile i (tcgen05.ld, fused math, TMA store) overlaps the
//     main loop of tile i+1;
//   * dependencies are per 128-row block, not per kernel: a tile of layer l+1 starts as soon as the row block of
//     layer l it reads is complete (`ready` counters, release / acquire at gpu scope + async-proxy fences) -- the
//     batch-parallel structure of an MLP (row blocks are independent through forward AND backward) pipelines
//     naturally; the split-K weight-gradient tiles wait for the row blocks of their K range only;
//   * the static order puts every producer before its consumers, and each CTA walks its items in increasing order,
//     so the smallest unfinished item can always run: no deadlock.
// Tile code (TMA boxes, UMMA descriptors, fused epilogues) is the one of exb_gemm_tcgen05_kernel<64>.
constexpr int CH_STAGES = 3;                  // 3 x 24 KB ring + 32 KB epilogue staging: two CTAs per SM
constexpr int CH_BN = 64;
constexpr int CH_B_BYTES = CH_BN * BK * 2;
constexpr int CH_MAX_PROB = 8;
constexpr int CH_MAX_MB = 64;                 // row blocks per problem tracked by the ready counters

struct ChainMaps { CUtensorMap tmA, tmB, tmO; };
struct ChainMapsAll { ChainMaps m[CH_MAX_PROB]; };    // passed as a __grid_constant__ parameter (3 KB): descriptors in param space
struct ChainMeta {
    GemmEpi E;
    int m_tiles, n_tiles, splits, nkb, per;
    int item0, items;
    int dep, dep_kind, dep_need;              // dep_kind 0 none | 1 A rows = my row block | 2 my K range (batch rows)
    int signal;
};

__device__ __forceinline__ void ch_decode(const ChainMeta* M, int nprob, int item, int& p, int& m, int& n, int& z) {
    p = 0;
#pragma unroll 1
    for (int i = 1; i < nprob; ++i)
        if (item >= M[i].item0) p = i;
    const int local = item - M[p].item0;
    const int mn = M[p].m_tiles * M[p].n_tiles;
    z = local / mn;
    const int r = local - z * mn;
    m = r / M[p].n_tiles;
    n = r - m * M[p].n_tiles;
}
__device__ __forceinline__ unsigned ch_ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(NUM_THREADS, 2)
exb_gemm_chain_kernel(const __grid_constant__ ChainMapsAll MAPS, const ChainMeta* __restrict__ metas, int nprob, int total_items,
                      unsigned* __restrict__ ready, int* __restrict__ err) {
    const ChainMaps* maps = MAPS.m;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;
    uint8_t* sB = smem + CH_STAGES * A_BYTES;
    uint8_t* sStage = sB + CH_STAGES * CH_B_BYTES;                 // 4 warps x 8 KB (1024-byte aligned)
    uint64_t* full = reinterpret_cast<uint64_t*>(sStage + 4 * 8192);
    uint64_t* empty = full + CH_STAGES;
    uint64_t* tfull = empty + CH_STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    ChainMeta* M = reinterpret_cast<ChainMeta*>(tmem_slot + 4);

    exb::pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < nprob * (int)(sizeof(ChainMeta) / 4); i += blockDim.x)
        reinterpret_cast<uint32_t*>(M)[i] = reinterpret_cast<const uint32_t*>(metas)[i];     // host-written
    if (warp == 0 && lane == 0) {
        for (int i = 0; i < CH_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // two 64-column fp32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    exb::pdl_wait();

    if (warp == 0) {
        if (lane == 0) {   // ===== TMA producer
            uint32_t kiter = 0;
            for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
                int p, m_blk, n_blk, z;
                ch_decode(M, nprob, item, p, m_blk, n_blk, z);
                const ChainMeta& Q = M[p];
                const int kb0 = z * Q.per, kb1 = min(Q.nkb, kb0 + Q.per);
                if (Q.dep_kind) {      // the row blocks this tile reads must be complete
                    int mb0 = m_blk, mb1 = m_blk + 1;
                    if (Q.dep_kind == 2) { mb0 = (kb0 * BK) / BM; mb1 = (kb1 * BK + BM - 1) / BM; }
                    const unsigned* cnt = ready + Q.dep * CH_MAX_MB;
                    for (int mb = mb0; mb < mb1; ++mb) {
                        uint32_t it = 0;
                        while (ch_ld_acquire(cnt + mb) < (unsigned)Q.dep_need) {
                            __nanosleep(64);
                            if (++it > (1u << 22)) { atomicCAS(err, 0, 100 + p); break; }
                        }
                    }
                    asm volatile("fence.proxy.async;" ::: "memory");   // generic acquire -> async-proxy (TMA) reads
                }
                const CUtensorMap* tA = &maps[p].tmA;
                const CUtensorMap* tB = &maps[p].tmB;
                for (int kb = kb0; kb < kb1; ++kb, ++kiter) {
                    const int s = kiter % CH_STAGES;
                    const uint32_t ph = (kiter / CH_STAGES) & 1u;
                    mbar_wait(&empty[s], ph ^ 1u);
                    mbar_expect_tx(&full[s], A_BYTES + CH_B_BYTES);
                    if (Q.E.mn_major) {
                        tma_load_2d(sA + s * A_BYTES, tA, &full[s], m_blk * BM, kb * BK);
                        tma_load_2d(sA + s * A_BYTES + A_BYTES / 2, tA, &full[s], m_blk * BM + 64, kb * BK);
                        tma_load_2d(sB + s * CH_B_BYTES, tB, &full[s], n_blk * CH_BN, kb * BK);
                    } else {
                        tma_load_2d(sA + s * A_BYTES, tA, &full[s], kb * BK, m_blk * BM);
                        tma_load_2d(sB + s * CH_B_BYTES, tB, &full[s], kb * BK, n_blk * CH_BN);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {   // ===== MMA issuer
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(CH_BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            uint32_t kiter = 0, lit = 0;
            for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++lit) {
                int p, m_blk, n_blk, z;
                ch_decode(M, nprob, item, p, m_blk, n_blk, z);
                const ChainMeta& Q = M[p];
                const int kb0 = z * Q.per, kb1 = min(Q.nkb, kb0 + Q.per);
                const uint32_t acc = lit & 1u;
                mbar_wait(&tempty[acc], ((lit >> 1) & 1u) ^ 1u);      // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tacc = tmem_base + acc * CH_BN;
                for (int kb = kb0; kb < kb1; ++kb, ++kiter) {
                    const int s = kiter % CH_STAGES;
                    const uint32_t ph = (kiter / CH_STAGES) & 1u;
                    mbar_wait(&full[s], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a0 = smem_u32(sA + s * A_BYTES), b0 = smem_u32(sB + s * CH_B_BYTES);
                    if (Q.E.mn_major) {
                        const uint32_t idesc_mn = idesc | (1u << 15) | (1u << 16);
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            umma_bf16(tacc, umma_desc_mn(a0 + k * 2048, 8192), umma_desc_mn(b0 + k * 2048, 8192),
                                      idesc_mn, ((kb - kb0) | k) ? 1u : 0u);
                    } else {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            umma_bf16(tacc, umma_desc(a0 + k * 32), umma_desc(b0 + k * 32), idesc, ((kb - kb0) | k) ? 1u : 0u);
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(&tfull[acc]);
            }
        }
    } else {
        // ===== epilogue warps (TMEM lane quadrant q), one 32-row x 64-column quarter of every tile
        const int q = warp & 3;
        uint8_t* wstage = sStage + (warp - 2) * 8192;
        uint32_t lit = 0;
        for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++lit) {
            int p, m_blk, n_blk, z;
            ch_decode(M, nprob, item, p, m_blk, n_blk, z);
            const ChainMeta& Q = M[p];
            const GemmEpi& E = Q.E;
            const uint32_t acc = lit & 1u;
            const int row0 = m_blk * BM + q * 32, row = row0 + lane;
            const bool rv = row < E.M;
            const bool f32out = (E.mode == EPI_DW || E.mode == EPI_DX_FM);
            const CUtensorMap* tO = &maps[p].tmO;
            bool waited = false;
#pragma unroll 1
            for (int c0 = 0; c0 < CH_BN; c0 += 32) {
                const int n0 = n_blk * CH_BN + c0;
                uint4 mk[4];
                float4 e4[8], s4[8];
                float dl = 0.f;
                if (E.mode == EPI_DX && rv) {
                    const uint4* mp = reinterpret_cast<const uint4*>(E.mask + (size_t)row * E.ldmask + n0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) mk[j] = mp[j];
                }
                const bool fm = (E.mode == EPI_DX_FM) && rv && (n0 + 31 < E.fm_cols);
                if (fm) {
                    const float4* ep = reinterpret_cast<const float4*>(E.emb + (size_t)row * E.ldemb + n0);
                    const float* sbase = E.S + (size_t)row * E.D;
                    dl = E.dlogit[row];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        e4[j] = ep[j];
                        s4[j] = *reinterpret_cast<const float4*>(sbase + ((n0 + 4 * j) % E.D));
                    }
                }
                if (!waited) {
                    mbar_wait(&tfull[acc], (lit >> 1) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    waited = true;
                }
                uint32_t v[32];
                tmem_ld32(tmem_base + acc * CH_BN + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                if (c0 + 32 >= CH_BN) {          // last read of this accumulator: hand it back to the MMA warp
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tempty[acc])) : "memory");
                }
                const __nv_bfloat16* mh = reinterpret_cast<const __nv_bfloat16*>(mk);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float x = __uint_as_float(v[j]);
                    const int n = n0 + j;
                    if (E.mode == EPI_FWD) {
                        if (E.relu) x = fmaxf(x, 0.f);
                        if (n == E.ones_col) x = 1.f;
                        if (n >= E.N) x = 0.f;
                    } else if (E.mode == EPI_DX) {
                        if (!rv || !(__bfloat162float(mh[j]) > 0.f) || n == E.ones_col || n >= E.N) x = 0.f;
                    } else if (fm) {
                        const float* ef = reinterpret_cast<const float*>(e4);
                        const float* sf = reinterpret_cast<const float*>(s4);
                        x += dl * (sf[j] - ef[j]);
                    }
                    v[j] = __float_as_uint(x);
                }
                if (f32out) {
                    uint8_t* tile = wstage + (c0 >> 5) * 4096;       // [32 rows][32 fp32]
#pragma unroll
                    for (int t = 0; t < 8; ++t)
                        *reinterpret_cast<uint4*>(tile + lane * 128 + ((t ^ (lane & 7)) << 4)) =
                            make_uint4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        if (E.mode == EPI_DW)
                            asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                                         ::"l"(tO), "r"(smem_u32(tile)), "r"(n0), "r"(row0) : "memory");
                        else
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                         ::"l"(tO), "r"(smem_u32(tile)), "r"(n0), "r"(row0) : "memory");
                    }
                } else {
                    uint8_t* tile = wstage;                            // [32 rows][64 bf16]
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        uint32_t pk[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(v[8 * t + 2 * u]), __uint_as_float(v[8 * t + 2 * u + 1]));
                            pk[u] = *reinterpret_cast<uint32_t*>(&h2);
                        }
                        const int chunk = (c0 >> 3) + t;
                        *reinterpret_cast<uint4*>(tile + lane * 128 + ((chunk ^ (lane & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                }
            }
            if (!f32out) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0)
                    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                 ::"l"(tO), "r"(smem_u32(wstage)), "r"(n_blk * CH_BN), "r"(row0) : "memory");
            }
            if (lane == 0) {
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (Q.signal) {
                    // a later GEMM of the chain reads this row block: its writes must have COMPLETED before the release
                    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                    asm volatile("fence.proxy.async;" ::: "memory");
                    __threadfence();
                    atomicAdd(&ready[p * CH_MAX_MB + m_blk], 1u);
                } else {
                    // nobody inside this launch reads the tile (dW, the last dX): only the staging tile has to be free
                    // again; kernel completion makes the writes visible to the next kernel. (Measured: the step time
                    // does not move, 0.2236 ms either way -- the store round trip is not what holds the chain back.)
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
            }
            __syncwarp();
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128) : "memory");
    }
    // self-cleaning dependency counters: the LAST CTA to leave zeroes them for the next launch (no memset node in
    // front of the kernel, so the launch keeps its programmatic-dependent-launch edge inside a CUDA graph)
    if (warp == 0) {
        unsigned last = 0;
        if (lane == 0) {
            __threadfence();
            last = (atomicAdd(&ready[CH_MAX_PROB * CH_MAX_MB], 1u) == gridDim.x - 1) ? 1u : 0u;
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            for (int i = lane; i < CH_MAX_PROB * CH_MAX_MB; i += 32) ready[i] = 0u;
            __syncwarp();
            if (lane == 0) { __threadfence(); ready[CH_MAX_PROB * CH_MAX_MB] = 0u; }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
thread_local std::string g_gemm_err;

bool make_map_ex(CUtensorMap* map, CUtensorMapDataType dt, int esize, const void* ptr, long long rows, long long cols,
                 long long ld, int box_cols, int box_rows, CUtensorMapSwizzle sw);

bool make_map(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
    return make_map_ex(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, rows, cols, ld, BK, box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
}

bool make_map_ex(CUtensorMap* map, CUtensorMapDataType dt, int esize, const void* ptr, long long rows, long long cols,
                 long long ld, int box_cols, int box_rows, CUtensorMapSwizzle sw) {
    if (!g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
            g_gemm_err = "cuTensorMapEncodeTiled entry point not found";
            return false;
        }
        g_encode = (EncodeTiledFn)fn;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * esize};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        g_gemm_err = "cuTensorMapEncodeTiled failed: " + std::to_string((int)r);
        return false;
    }
    return true;
}

template <int BN>
constexpr size_t gemm_smem() { return stages_for<BN>() * (A_BYTES + BN * BK * 2) + (2 * stages_for<BN>() + 1) * 8 + 16 + 1024; }

// tile width: env EXB_GEMM_BN (64 | 128) overrides; default 128 for the operand-traffic bound shapes
int pick_swap() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("EXB_GEMM_SWAP"); v = e ? atoi(e) : 0; }
    return v;
}
int pick_bn(int M, int N, int K) {
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("EXB_GEMM_BN"); forced = e ? atoi(e) : 0; }
    if (forced == 64 || forced == 128) return forced;
    // measured (profiles/dense_path.md): 128-wide tiles cut operand traffic by a third but leave too few
    // CTAs in flight per SM for the fused step's shapes (fwd1 15.0 -> 17.0 us, dX1 32.9 -> 35.6 us); default 64.
    // Exception: tall, short-K products (the CIN input-gradient GEMM: M 36 864, N 1 728, K 128 -- two k-blocks per
    // tile, 7 776 tiles): a tile is all set-up + epilogue (ncu: 116 us, tensor pipe 6.6 %, DRAM 8 %), so the wider
    // tile halves the number of those
    if (K <= 128 && N >= 512 && M >= 8192) return 128;
    return 64;
}

template <int BN>
cudaError_t launch_gemm(dim3 grid, cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO,
                        const CUtensorMap& tmT, const GemmEpi& E, int nkb, int per) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(exb_gemm_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem<BN>());
        attr_set = true;
    }
    if (E.mc > 1) {       // clusters of E.mc CTAs along grid.y (the N-tile axis when swap == 0)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid; cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = gemm_smem<BN>(); cfg.stream = stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = (unsigned)E.mc; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = exb::pdl_enabled() ? 2 : 1;
        return cudaLaunchKernelEx(&cfg, exb_gemm_tcgen05_kernel<BN>, tmA, tmB, tmO, tmT, E, nkb, per);
    }
    return exb::launch_pdl(exb_gemm_tcgen05_kernel<BN>, grid, dim3(NUM_THREADS), gemm_smem<BN>(), stream, tmA, tmB, tmO, tmT,
                           E, nkb, per);
}

// cluster size for the A-tile multicast: the largest divisor (<= EXB_GEMM_MC, <= 8) of the number of N tiles.
// OFF by default (EXB_GEMM_MC unset / 0): measured on B200 the lock-step clusters cost more than the 2.3x smaller
// L2 read traffic buys for this step's shapes (fwd 32.8 -> 37.8 us, dX 54 -> 67 us; profiles/r2/dense_path.md) --
// the main loop is bound by TMA round-trip latency per ring slot, not by L2 bandwidth.
int pick_mc(int n_tiles) {
    static int lim = -1;
    if (lim < 0) { const char* e = getenv("EXB_GEMM_MC"); lim = e ? atoi(e) : 0; }
    if (lim <= 1 || pick_swap()) return 1;
    int best = 1;
    for (int c = 2; c <= 8 && c <= lim; ++c)
        if (n_tiles % c == 0) best = c;
    return best;
}

}  // namespace

extern "C" {

const char* exb_gemm_last_error() { return g_gemm_err.c_str(); }

// D (+)= A[M,K](lda) * B[N,K](ldb)^T, bf16 in. K must be a multiple of 64 (pad the operands);
// lda/ldb in elements, multiples of 8. epi: see GemmEpi. splits >= 1 (EPI_DW only).
int exb_gemm_bf16_nt(uint64_t A, long long lda, uint64_t B, long long ldb, int M, int N, int K, int mode, int relu,
                     int ones_col, uint64_t out, long long ldo, uint64_t outT, long long ldoT, uint64_t mask,
                     long long ldmask, uint64_t dlogit, uint64_t S, uint64_t emb, long long ldemb, int fm_cols, int D,
                     int splits, uint64_t stream, uint64_t dbg) {
    if (K % BK != 0 || lda % 8 != 0 || ldb % 8 != 0) { g_gemm_err = "gemm: K %% 64 / ld %% 8 violated"; return -1; }
    CUtensorMap tmA, tmB;
    if (!make_map(&tmA, (const void*)A, M, K, lda, BM)) return -1;
    const int BN = pick_bn(M, N, K);
    if (!make_map(&tmB, (const void*)B, N, K, ldb, BN)) return -1;
    CUtensorMap tmO, tmT;
    const bool f32out = (mode == EPI_DW || mode == EPI_DX_FM);
    if (f32out) {   // fp32 [M, N] (ld ldo): 32x32 boxes = 128-byte rows
        if (!make_map_ex(&tmO, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (const void*)out, M, N, ldo, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    } else {        // bf16 [M, ceil64(N)] : 64x32 boxes = 128-byte rows
        if (!make_map_ex(&tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (const void*)out, M, (N + 63) / 64 * 64, ldo, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    }
    tmT = tmO;
    if (outT && !make_map_ex(&tmT, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (const void*)outT, (N + 63) / 64 * 64, M, ldoT, 32, BN, CU_TENSOR_MAP_SWIZZLE_NONE)) return -1;
    GemmEpi E;
    E.mode = mode; E.relu = relu; E.ones_col = ones_col; E.fm_cols = fm_cols; E.M = M; E.N = N; E.D = D > 0 ? D : 1; E.mn_major = 0;
    E.out = (void*)out; E.ldo = ldo; E.outT = (__nv_bfloat16*)outT; E.ldoT = ldoT;
    E.mask = (const __nv_bfloat16*)mask; E.ldmask = ldmask;
    E.dlogit = (const float*)dlogit; E.S = (const float*)S; E.emb = (const float*)emb; E.ldemb = ldemb;
    E.dbg = (unsigned long long*)dbg;
    const int nkb = K / BK;
    if (splits < 1) splits = 1;
    if (splits > nkb) splits = nkb;
    const int per = (nkb + splits - 1) / splits;
    splits = (nkb + per - 1) / per;
    dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN, splits);
    E.swap = pick_swap(); E.mc = pick_mc((int)grid.y);
    if (E.swap) grid = dim3(grid.y, grid.x, grid.z);
    cudaError_t err = BN == 128 ? launch_gemm<128>(grid, (cudaStream_t)stream, tmA, tmB, tmO, tmT, E, nkb, per)
                                : launch_gemm<64>(grid, (cudaStream_t)stream, tmA, tmB, tmO, tmT, E, nkb, per);
    if (err == cudaSuccess) err = cudaGetLastError();
    if (err != cudaSuccess) { g_gemm_err = std::string("gemm launch: ") + cudaGetErrorString(err); return -1; }
    return 0;
}

// dW-style product from batch-major operands, no transposed copies needed:
//   out[M, N] (fp32, += via TMA reduce-add) = A[K, M]^T * B[K, N],  A/B bf16 row-major with K rows
// (e.g. M = features of dZ, N = features of the layer input, K = batch). M, N: any; K % 64 == 0;
// lda/ldb multiples of 8 elements. Both operands reach the tensor core as MN-major tiles.
int exb_gemm_bf16_tn(uint64_t A, long long lda, uint64_t B, long long ldb, int M, int N, int K, uint64_t out,
                     long long ldo, int splits, uint64_t stream) {
    if (K % BK != 0 || lda % 8 != 0 || ldb % 8 != 0) { g_gemm_err = "gemm_tn: K %% 64 / ld %% 8 violated"; return -1; }
    CUtensorMap tmA, tmB, tmO, tmT;
    if (!make_map_ex(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (const void*)A, K, M, lda, 64, BK, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    if (!make_map_ex(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (const void*)B, K, N, ldb, 64, BK, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    if (!make_map_ex(&tmO, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (const void*)out, M, N, ldo, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    tmT = tmO;
    GemmEpi E;
    memset(&E, 0, sizeof(E));
    E.mode = EPI_DW; E.ones_col = -1; E.M = M; E.N = N; E.D = 1; E.mn_major = 1;
    E.out = (void*)out; E.ldo = ldo;
    const int BN = pick_bn(M, N, K);
    const int nkb = K / BK;
    if (splits < 1) splits = 1;
    if (splits > nkb) splits = nkb;
    const int per = (nkb + splits - 1) / splits;
    splits = (nkb + per - 1) / per;
    dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN, splits);
    E.swap = pick_swap(); E.mc = pick_mc((int)grid.y);
    if (E.swap) grid = dim3(grid.y, grid.x, grid.z);
    cudaError_t err = BN == 128 ? launch_gemm<128>(grid, (cudaStream_t)stream, tmA, tmB, tmO, tmT, E, nkb, per)
                                : launch_gemm<64>(grid, (cudaStream_t)stream, tmA, tmB, tmO, tmT, E, nkb, per);
    if (err == cudaSuccess) err = cudaGetLastError();
    if (err != cudaSuccess) { g_gemm_err = std::string("gemm_tn launch: ") + cudaGetErrorString(err); return -1; }
    return 0;
}

// ---------------------------------------------------------------- GEMM chains (persistent, one launch)
struct ChainDesc {      // one GEMM of a chain (python: ops/gemm.py ChainDesc)
    int tn;             // 0: D = A[M,K] B[N,K]^T (K-major operands); 1: D[M,N] += A[K,M]^T B[K,N] (split-K, fp32 reduce-add)
    int M, N, K;
    unsigned long long A, B, out;
    long long lda, ldb, ldo;
    int mode, relu, ones_col, fm_cols, D, splits;
    unsigned long long mask; long long ldmask;
    unsigned long long dlogit, S, emb; long long ldemb;
    int dep, dep_kind;  // index of the GEMM of this chain that produces this one's A operand (-1: none); kind 1 row block, 2 K range
};
struct Chain {
    int nprob = 0, total = 0, grid = 0;
    ChainMapsAll maps;
    ChainMeta* d_meta = nullptr;
    unsigned* d_ready = nullptr;
    int* d_err = nullptr;
    size_t smem = 0;
};

int exb_chain_desc_size() { return (int)sizeof(ChainDesc); }

void* exb_chain_create(const void* descs, int n, int sms) {
    if (n < 1 || n > CH_MAX_PROB) { g_gemm_err = "chain: 1..8 GEMMs"; return nullptr; }
    const ChainDesc* D = reinterpret_cast<const ChainDesc*>(descs);
    std::vector<ChainMaps> maps(n);
    std::vector<ChainMeta> meta(n);
    int item0 = 0;
    for (int i = 0; i < n; ++i) {
        const ChainDesc& d = D[i];
        ChainMeta& Q = meta[i];
        memset(&Q, 0, sizeof(Q));
        if (d.K % BK != 0 || d.lda % 8 != 0 || d.ldb % 8 != 0) { g_gemm_err = "chain: K % 64 / ld % 8 violated"; return nullptr; }
        GemmEpi& E = Q.E;
        E.mode = d.tn ? EPI_DW : d.mode; E.relu = d.relu; E.ones_col = d.tn ? -1 : d.ones_col; E.fm_cols = d.fm_cols;
        E.M = d.M; E.N = d.N; E.D = d.D > 0 ? d.D : 1; E.mn_major = d.tn ? 1 : 0;
        E.out = (void*)d.out; E.ldo = d.ldo; E.outT = nullptr; E.ldoT = 0;
        E.mask = (const __nv_bfloat16*)d.mask; E.ldmask = d.ldmask;
        E.dlogit = (const float*)d.dlogit; E.S = (const float*)d.S; E.emb = (const float*)d.emb; E.ldemb = d.ldemb;
        E.swap = 0; E.mc = 1; E.dbg = nullptr;
        const bool f32out = (E.mode == EPI_DW || E.mode == EPI_DX_FM);
        bool ok;
        if (d.tn) {
            ok = make_map_ex(&maps[i].tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (const void*)d.A, d.K, d.M, d.lda, 64, BK, CU_TENSOR_MAP_SWIZZLE_128B) &&
                 make_map_ex(&maps[i].tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (const void*)d.B, d.K, d.N, d.ldb, 64, BK, CU_TENSOR_MAP_SWIZZLE_128B);
        } else {
            ok = make_map(&maps[i].tmA, (const void*)d.A, d.M, d.K, d.lda, BM) && make_map(&maps[i].tmB, (const void*)d.B, d.N, d.K, d.ldb, CH_BN);
        }
        if (ok) {
            if (f32out) ok = make_map_ex(&maps[i].tmO, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (const void*)d.out, d.M, d.N, d.ldo, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
            else ok = make_map_ex(&maps[i].tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (const void*)d.out, d.M, (d.N + 63) / 64 * 64, d.ldo, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B);
        }
        if (!ok) return nullptr;
        Q.nkb = d.K / BK;
        int splits = d.tn ? std::max(1, d.splits) : 1;
        if (splits > Q.nkb) splits = Q.nkb;
        Q.per = (Q.nkb + splits - 1) / splits;
        Q.splits = (Q.nkb + Q.per - 1) / Q.per;
        Q.m_tiles = (d.M + BM - 1) / BM; Q.n_tiles = (d.N + CH_BN - 1) / CH_BN;
        Q.item0 = item0; Q.items = Q.m_tiles * Q.n_tiles * Q.splits;
        item0 += Q.items;
        Q.dep = d.dep; Q.dep_kind = d.dep >= 0 ? d.dep_kind : 0; Q.dep_need = 0; Q.signal = 0;
        if (Q.dep_kind) {
            if (d.dep >= i) { g_gemm_err = "chain: a GEMM may only depend on an earlier one"; return nullptr; }
            meta[d.dep].signal = 1;
            Q.dep_need = 4 * meta[d.dep].n_tiles;      // four epilogue warps sign off every tile of the row block
            if (meta[d.dep].m_tiles > CH_MAX_MB) { g_gemm_err = "chain: too many row blocks"; return nullptr; }
        }
    }
    Chain* c = new Chain();
    c->nprob = n; c->total = item0;
    c->smem = CH_STAGES * (A_BYTES + CH_B_BYTES) + 4 * 8192 + (2 * CH_STAGES + 4) * 8 + 16 + CH_MAX_PROB * sizeof(ChainMeta) + 1024;
    cudaFuncSetAttribute(exb_gemm_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem);
    cudaFuncSetAttribute(exb_gemm_chain_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    // two CTAs per SM by construction: 2 x (smem + 1 KB reserved) <= 228 KB, 2 x 192 threads x 168 registers <= 64 K
    // (the occupancy API answers 1 here when it is asked before the carve-out is configured)
    int occ = (2 * (c->smem + 1024) <= 232448) ? 2 : 1;
    if (const char* ev = getenv("EXB_CHAIN_CTAS_PER_SM")) occ = std::max(1, std::min(2, atoi(ev)));
    // dependencies are spin-waits: every CTA has to be resident
    c->grid = std::min(c->total, (sms > 0 ? sms : 148) * occ);
    memset(&c->maps, 0, sizeof(c->maps));
    for (int i = 0; i < n; ++i) c->maps.m[i] = maps[i];
    if (cudaMalloc(&c->d_meta, n * sizeof(ChainMeta)) != cudaSuccess ||
        cudaMalloc(&c->d_ready, (CH_MAX_PROB * CH_MAX_MB + 32) * 4) != cudaSuccess || cudaMalloc(&c->d_err, 4) != cudaSuccess) {
        g_gemm_err = "chain: cudaMalloc failed"; delete c; return nullptr;
    }
    cudaMemcpy(c->d_meta, meta.data(), n * sizeof(ChainMeta), cudaMemcpyHostToDevice);
    cudaMemset(c->d_ready, 0, (CH_MAX_PROB * CH_MAX_MB + 32) * 4);
    cudaMemset(c->d_err, 0, 4);
    return c;
}
void exb_chain_destroy(void* h) {
    Chain* c = (Chain*)h;
    cudaFree(c->d_meta); cudaFree(c->d_ready); cudaFree(c->d_err);
    delete c;
}
int exb_chain_launch(void* h, uint64_t stream) {
    Chain* c = (Chain*)h;
    cudaError_t err = exb::launch_pdl(exb_gemm_chain_kernel, dim3(c->grid), dim3(NUM_THREADS), c->smem, (cudaStream_t)stream,
                              c->maps, (const ChainMeta*)c->d_meta, c->nprob, c->total, c->d_ready, c->d_err);
    if (err == cudaSuccess) err = cudaGetLastError();
    if (err != cudaSuccess) { g_gemm_err = std::string("chain launch: ") + cudaGetErrorString(err); return -1; }
    return 0;
}
// device sync; returns the error word (0 ok, 100 + p: GEMM p timed out waiting for its producer)
int exb_chain_status(void* h) {
    Chain* c = (Chain*)h;
    int v = 0;
    cudaDeviceSynchronize();
    cudaMemcpy(&v, c->d_err, 4, cudaMemcpyDeviceToHost);
    return v;
}
int exb_chain_info(void* h, int* out) { Chain* c = (Chain*)h; out[0] = c->total; out[1] = c->grid; out[2] = (int)c->smem; return 0; }

}  // extern "C"
