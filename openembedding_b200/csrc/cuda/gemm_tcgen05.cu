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
#include <stdint.h>
#include <stdio.h>

#include <string>

namespace {

constexpr int BM = 128, BK = 64, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;
// BN = 64: 96 KB of stages, two CTAs per SM (small-N GEMMs); BN = 256: 192 KB, one CTA per SM,
// 4x fewer re-reads of the A operand (the wide dX1 / dW1 GEMMs are L2-bandwidth bound at BN = 64)
constexpr int NUM_THREADS = 192;

enum EpiMode : int { EPI_FWD = 0, EPI_DX = 1, EPI_DW = 2, EPI_DX_FM = 3 };

struct GemmEpi {
    int mode, relu, ones_col, fm_cols;
    int M, N, D, _pad;
    void* out; long long ldo;
    __nv_bfloat16* outT; long long ldoT;
    const __nv_bfloat16* mask; long long ldmask;
    const float* dlogit; const float* S; const float* emb; long long ldemb;
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
__device__ __forceinline__ void umma_bf16(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
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
__global__ void __launch_bounds__(NUM_THREADS, (BN <= 64 ? 2 : 1))
exb_gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        GemmEpi E, int num_k_blocks, int k_blocks_per_split) {
    constexpr int B_BYTES = BN * BK * 2;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte aligned bases; do not rely on the toolchain for it
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * B_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_blk = blockIdx.x, n_blk = blockIdx.y;
    const int kb0 = blockIdx.z * k_blocks_per_split;
    const int kb1 = min(num_k_blocks, kb0 + k_blocks_per_split);
    const int nkb = kb1 - kb0;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
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

    if (warp == 0) {
        if (lane == 0) {   // ===== TMA producer
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);
                mbar_expect_tx(&full[s], A_BYTES + B_BYTES);
                tma_load_2d(sA + s * A_BYTES, &tmA, &full[s], (kb0 + i) * BK, m_blk * BM);
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
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a0 = smem_u32(sA + s * A_BYTES), b0 = smem_u32(sB + s * B_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)   // UMMA_K = 16 bf16 = 32 bytes inside the 128B swizzle row
                    umma_bf16(tmem_base, umma_desc(a0 + k * 32), umma_desc(b0 + k * 32), idesc, (i | k) ? 1u : 0u);
                umma_commit(&empty[s]);            // smem stage reusable when these MMAs retire
            }
            umma_commit(tmem_full);                // accumulator complete
        }
    } else {
        // ===== epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32)
        // TMEM hands every thread one ROW (32 consecutive columns). Stores/loads in that layout
        // touch 32 different lines per instruction, so each warp transposes its 32x32 block
        // through shared memory (the pipeline stages are free once tmem_full fired) and does
        // the fused math + global traffic with lanes along the COLUMNS (128-byte rows).
        const int q = warp & 3;
        const int row0 = m_blk * BM + q * 32;
        bool ok = true;
        if (nkb > 0) ok = mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float* stg = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * 33);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            if (nkb > 0 && ok) tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = __uint_as_float(v[j]);
            __syncwarp();
            const int n = n_blk * BN + c0 + lane;
            // issue every global read of the block first (32 independent loads per lane in flight)
            float aux[32];
            if (E.mode == EPI_DX) {
#pragma unroll
                for (int rr = 0; rr < 32; ++rr)
                    aux[rr] = (row0 + rr < E.M) ? __bfloat162float(E.mask[(size_t)(row0 + rr) * E.ldmask + n]) : 0.f;
            } else if (E.mode == EPI_DX_FM) {
                const bool fm = n < E.fm_cols;
                const int dcol = n % E.D;
#pragma unroll
                for (int rr = 0; rr < 32; ++rr) {
                    const int row = row0 + rr;
                    aux[rr] = (fm && row < E.M) ? (E.S[(size_t)row * E.D + dcol] - E.emb[(size_t)row * E.ldemb + n]) : 0.f;
                }
#pragma unroll
                for (int rr = 0; rr < 32; ++rr)
                    if (fm && row0 + rr < E.M) aux[rr] *= E.dlogit[row0 + rr];
            }
#pragma unroll
            for (int rr = 0; rr < 32; ++rr) {
                const int row = row0 + rr;
                float x = stg[rr * 33 + lane];
                if (row < E.M) {
                    if (E.mode == EPI_FWD) {
                        if (E.relu) x = fmaxf(x, 0.f);
                        if (n == E.ones_col) x = 1.f;
                        if (n >= E.N) x = 0.f;
                        reinterpret_cast<__nv_bfloat16*>(E.out)[(size_t)row * E.ldo + n] = __float2bfloat16_rn(x);
                    } else if (E.mode == EPI_DX) {
                        if (!(aux[rr] > 0.f) || n == E.ones_col || n >= E.N) x = 0.f;
                        reinterpret_cast<__nv_bfloat16*>(E.out)[(size_t)row * E.ldo + n] = __float2bfloat16_rn(x);
                    } else if (E.mode == EPI_DW) {
                        if (n < E.N) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(reinterpret_cast<float*>(E.out) + (size_t)row * E.ldo + n), "f"(x) : "memory");
                    } else {   // EPI_DX_FM
                        x += aux[rr];
                        if (n < E.N) reinterpret_cast<float*>(E.out)[(size_t)row * E.ldo + n] = x;
                    }
                }
                if (E.outT) stg[rr * 33 + lane] = x;
            }
            __syncwarp();
            if (E.outT && row0 + lane < E.M) {   // lanes = consecutive rows: coalesced transposed copy
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    E.outT[(size_t)(n_blk * BN + c0 + j) * E.ldoT + row0 + lane] = __float2bfloat16_rn(stg[lane * 33 + j]);
            }
            __syncwarp();
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BN) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
thread_local std::string g_gemm_err;

bool make_map(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
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
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        g_gemm_err = "cuTensorMapEncodeTiled failed: " + std::to_string((int)r);
        return false;
    }
    return true;
}

template <int BN>
constexpr size_t gemm_smem() { return STAGES * (A_BYTES + BN * BK * 2) + (2 * STAGES + 1) * 8 + 16 + 1024; }

}  // namespace

extern "C" {

const char* exb_gemm_last_error() { return g_gemm_err.c_str(); }

// D (+)= A[M,K](lda) * B[N,K](ldb)^T, bf16 in. K must be a multiple of 64 (pad the operands);
// lda/ldb in elements, multiples of 8. epi: see GemmEpi. splits >= 1 (EPI_DW only).
int exb_gemm_bf16_nt(uint64_t A, long long lda, uint64_t B, long long ldb, int M, int N, int K, int mode, int relu,
                     int ones_col, uint64_t out, long long ldo, uint64_t outT, long long ldoT, uint64_t mask,
                     long long ldmask, uint64_t dlogit, uint64_t S, uint64_t emb, long long ldemb, int fm_cols, int D,
                     int splits, uint64_t stream) {
    if (K % BK != 0 || lda % 8 != 0 || ldb % 8 != 0) { g_gemm_err = "gemm: K %% 64 / ld %% 8 violated"; return -1; }
    CUtensorMap tmA, tmB;
    if (!make_map(&tmA, (const void*)A, M, K, lda, BM)) return -1;
    const int BN = (N >= 1024) ? 256 : 64;
    if (!make_map(&tmB, (const void*)B, N, K, ldb, BN)) return -1;
    GemmEpi E;
    E.mode = mode; E.relu = relu; E.ones_col = ones_col; E.fm_cols = fm_cols; E.M = M; E.N = N; E.D = D > 0 ? D : 1; E._pad = 0;
    E.out = (void*)out; E.ldo = ldo; E.outT = (__nv_bfloat16*)outT; E.ldoT = ldoT;
    E.mask = (const __nv_bfloat16*)mask; E.ldmask = ldmask;
    E.dlogit = (const float*)dlogit; E.S = (const float*)S; E.emb = (const float*)emb; E.ldemb = ldemb;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(exb_gemm_tcgen05_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem<64>());
        cudaFuncSetAttribute(exb_gemm_tcgen05_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem<256>());
        attr_set = true;
    }
    const int nkb = K / BK;
    if (splits < 1) splits = 1;
    if (splits > nkb) splits = nkb;
    const int per = (nkb + splits - 1) / splits;
    splits = (nkb + per - 1) / per;
    dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN, splits);
    if (BN == 64)
        exb_gemm_tcgen05_kernel<64><<<grid, NUM_THREADS, gemm_smem<64>(), (cudaStream_t)stream>>>(tmA, tmB, E, nkb, per);
    else
        exb_gemm_tcgen05_kernel<256><<<grid, NUM_THREADS, gemm_smem<256>(), (cudaStream_t)stream>>>(tmA, tmB, E, nkb, per);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { g_gemm_err = std::string("gemm launch: ") + cudaGetErrorString(err); return -1; }
    return 0;
}

}  // extern "C"
