// bulk_rows.cuh -- row movement with the Blackwell bulk-copy engine (1-D TMA).
//
// Measured on B200 (profiles/sparse_v0.md): a random 256-byte row in a 16-54 GB table
// costs 2.5-5 us to fetch (DRAM + TLB miss), so a register-staged gather is purely
// latency-bound: 8 rows in flight per lane group, 25 % occupancy, 8 % of DRAM bandwidth.
// Here every lane owns one ROW and hands it to the copy engine:
//     cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes   (SASS: UBLKCP)
// No registers are held while the bytes are in flight, a warp keeps 32 rows (8-24 KB) in
// flight and an SM several hundred. Rows leave shared memory the same way
// (cp.async.bulk.global.shared::cta). Works on any global address, i.e. also on
// peer-mapped (NVLink) table slabs.
#pragma once
#include "exb_common.cuh"

namespace exb {

#define EXB_PULL_WARP_BUF 8192    // bytes of row staging per warp in the pull kernel
#define EXB_APPLY_WARP_BUF 12288  // bytes per warp in the apply phase (w | state | acc rows)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* b, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a lost copy becomes an error code instead of a hung GPU
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity, int* status) {
    if (mbar_try_wait(b, parity)) return;
    unsigned long long t0 = globaltimer_ns();
    unsigned it = 0;
    while (!mbar_try_wait(b, parity)) {
        if ((++it & 63u) == 0 && globaltimer_ns() - t0 > EXB_SPIN_TIMEOUT_NS) {
            set_error(status, EXB_ERR_TIMEOUT_GRID);
            break;
        }
    }
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

__device__ __noinline__ float4 init_block_masked(const InitParams* I, unsigned long long id, int c, int dim);

// Pull the 32 rows of a warp task through shared memory.
// Lane l holds (src, id, flag) of row l. buf: EXB_PULL_WARP_BUF bytes owned by this warp.
__device__ __forceinline__ void pull_rows_bulk(const TableDev& T, const float* src, unsigned long long id,
                                               int flag, int b0, int n_rows, float* __restrict__ out,
                                               int io_stride, int off, int lane, unsigned char* buf,
                                               unsigned long long* mbar, unsigned& parity, int* status) {
    const unsigned rowbytes = (unsigned)T.wstride * 4u;
    const int R = min(32, (int)(EXB_PULL_WARP_BUF / rowbytes));   // rows per pass (warp uniform)
    for (int r0 = 0; r0 < 32; r0 += R) {
        const int r = r0 + lane;
        const float* s = (const float*)__shfl_sync(0xffffffffu, (unsigned long long)src, r & 31);
        const unsigned long long idr = __shfl_sync(0xffffffffu, id, r & 31);
        int fl = __shfl_sync(0xffffffffu, flag, r & 31);
        const int b = b0 + r;
        const bool mine = lane < R && r < 32 && b < n_rows;
        if (!mine) fl = -1;
        float* row = reinterpret_cast<float*>(buf + (size_t)lane * rowbytes);
        const unsigned total = __reduce_add_sync(0xffffffffu, fl == 1 ? rowbytes : 0u);
        if (total) {
            if (lane == 0) mbar_expect_tx(mbar, total);
            __syncwarp();
            if (fl == 1) bulk_g2s(row, s, rowbytes, mbar);
        }
        if (fl == 0 || fl == 2) {   // invalid id -> zeros, missing hash row -> initializer value
            for (int c = 0; c < T.wstride; c += 4) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (fl == 2) v = init_block_masked(&T.init, idr, c, T.dim);
                *reinterpret_cast<float4*>(row + c) = v;
            }
        }
        if (total) {
            mbar_wait(mbar, parity, status);
            parity ^= 1u;
        }
        fence_proxy_async();
        __syncwarp();
        if (fl >= 0) bulk_s2g(out + (size_t)b * io_stride + off, row, rowbytes);
        bulk_commit();
        bulk_wait_read();
        __syncwarp();
    }
}

}  // namespace exb

namespace exb {

struct WarpMeta {   // per-warp row metadata of the apply phase (shared memory)
    unsigned long long key[32];
    unsigned long long row[32];
    unsigned h[32];
    unsigned cnt[32];
    int flag[32];
};

// Apply the optimizer to the (up to) 32 unique rows of a warp task; weights, state and
// accumulator rows travel through shared memory with bulk copies.
// Lane l holds (key,row,h,cnt,flag) of row l on entry.
__device__ __forceinline__ void apply_rows_bulk(const TableDev& T, const PlanDev& P, float* accbase,
                                                unsigned long long key, unsigned long long row, unsigned h,
                                                unsigned cnt, int flag, int lane, unsigned char* buf,
                                                WarpMeta* M, unsigned long long* mbar, unsigned& parity) {
    const int wstride = T.wstride, sstride = T.sstride, dim = T.dim, nslots = T.nslots, nsc = T.nscalars;
    const unsigned wb = (unsigned)wstride * 4u, sb = (unsigned)sstride * 4u;
    const int R = min(32, (int)(EXB_APPLY_WARP_BUF / (2u * wb + sb)));
    const int lpr = T.lpr, gl = lane % lpr, RP = 32 / lpr;
    float* wloc = T.w[P.rank];
    const OptParams opt = T.opt;
    const float s0i = opt_slot_init<float>(opt, 0), s1i = opt_slot_init<float>(opt, 1);
    M->key[lane] = key; M->row[lane] = row; M->h[lane] = h; M->cnt[lane] = cnt; M->flag[lane] = flag;
    __syncwarp();
    float* wbuf = reinterpret_cast<float*>(buf);
    float* sbuf = reinterpret_cast<float*>(buf + (size_t)R * wb);
    float* abuf = reinterpret_cast<float*>(buf + (size_t)R * (wb + sb));
    for (int r0 = 0; r0 < 32; r0 += R) {
        const int r = r0 + lane;
        const bool mine = lane < R && r < 32;
        const int fl = mine ? M->flag[r] : 0;
        float* wrow_s = wbuf + (size_t)lane * wstride;
        float* srow_s = sbuf + (size_t)lane * sstride;
        float* arow_s = abuf + (size_t)lane * wstride;
        unsigned long long grow = mine ? M->row[r] : 0ull;
        unsigned gh = mine ? M->h[r] : 0u;
        const unsigned bytes = fl == 1 ? (2u * wb + sb) : (fl == 2 ? wb : 0u);
        const unsigned total = __reduce_add_sync(0xffffffffu, bytes);
        if (total) {
            if (lane == 0) mbar_expect_tx(mbar, total);
            __syncwarp();
            if (fl) bulk_g2s(arow_s, accbase + (unsigned long long)gh * wstride, wb, mbar);
            if (fl == 1) {
                bulk_g2s(wrow_s, wloc + grow * (unsigned long long)wstride, wb, mbar);
                bulk_g2s(srow_s, T.state + grow * (unsigned long long)sstride, sb, mbar);
            }
            mbar_wait(mbar, parity, P.status);
            parity ^= 1u;
        }
        __syncwarp();
        // ---- math out of shared memory: lane group `lane / lpr` walks rows jj, jj+RP, ...
        for (int jj = lane / lpr; jj < R; jj += RP) {
            const int rr = r0 + jj;
            if (rr >= 32) break;
            const int f = M->flag[rr];
            if (!f) continue;
            float* wr = wbuf + (size_t)jj * wstride;
            float* sr = sbuf + (size_t)jj * sstride;
            float* ar = abuf + (size_t)jj * wstride;
            float sc[2] = {0.f, 0.f}, nsc_v[2];
            for (int i = 0; i < nsc; ++i)
                sc[i] = (f == 2) ? opt_scalar_init<float>(opt, i) : sr[(size_t)nslots * wstride + i];
            RowCtx<float> rc = opt_row_prologue_pure<float>(opt, sc, (uint64_t)M->cnt[rr], nsc_v);
            for (int c = gl * 4; c < wstride; c += lpr * 4) {
                float4 g = *reinterpret_cast<float4*>(ar + c);
                float4 w, a = make_float4(s0i, s0i, s0i, s0i), b = make_float4(s1i, s1i, s1i, s1i);
                if (f == 2) {
                    w = init_block_masked(&T.init, M->key[rr], c, dim);
                } else {
                    w = *reinterpret_cast<float4*>(wr + c);
                    if (nslots > 0) a = *reinterpret_cast<float4*>(sr + c);
                    if (nslots > 1) b = *reinterpret_cast<float4*>(sr + wstride + c);
                }
                if (c + 0 < dim) opt_elem<float>(opt, rc, w.x, a.x, b.x, g.x);
                if (c + 1 < dim) opt_elem<float>(opt, rc, w.y, a.y, b.y, g.y);
                if (c + 2 < dim) opt_elem<float>(opt, rc, w.z, a.z, b.z, g.z);
                if (c + 3 < dim) opt_elem<float>(opt, rc, w.w, a.w, b.w, g.w);
                *reinterpret_cast<float4*>(wr + c) = w;
                if (nslots > 0) *reinterpret_cast<float4*>(sr + c) = a;
                if (nslots > 1) *reinterpret_cast<float4*>(sr + wstride + c) = b;
                *reinterpret_cast<float4*>(ar + c) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (gl == 0) {
                for (int i = 0; i < nsc; ++i) sr[(size_t)nslots * wstride + i] = nsc_v[i];
                if (f == 2)   // pad words of a brand-new state row
                    for (int i = nslots * wstride + nsc; i < sstride; ++i) sr[i] = 0.f;
            }
        }
        fence_proxy_async();
        __syncwarp();
        if (fl) {
            bulk_s2g(wloc + grow * (unsigned long long)wstride, wrow_s, wb);
            bulk_s2g(T.state + grow * (unsigned long long)sstride, srow_s, sb);
            bulk_s2g(accbase + (unsigned long long)gh * wstride, arow_s, wb);
        }
        bulk_commit();
        bulk_wait_read();
        __syncwarp();
    }
}

}  // namespace exb
