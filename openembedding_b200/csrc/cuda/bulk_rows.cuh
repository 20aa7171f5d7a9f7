// bulk_rows.cuh -- asynchronous row movement through shared memory.
//
// Measured on B200 (profiles/sparse_path.md): a random 256-byte row in a 16-54 GB table
// costs 2.5-5 us to fetch (DRAM + TLB miss), so a register-staged gather is purely
// latency-bound (8 rows in flight per lane group, 25 % occupancy, 8 % of DRAM bandwidth).
// Two async mechanisms were tried:
//   * 1-D TMA, one cp.async.bulk (UBLKCP) per row: correct, but the per-SM TMA unit
//     retires one small copy every ~50-60 cycles -> ~12 us per 384-copy pass (removed);
//   * cp.async 16-byte (LDGSTS): a warp instruction moves 512 B in ~8 issue cycles, holds
//     no registers while in flight and queues arbitrarily deep -> the path used below.
// A warp keeps a whole task (32 weight rows, or up to 13 x {weights, state, accumulator}) in
// flight; peer-mapped (NVLink) addresses take the same path.
#pragma once
#include "exb_common.cuh"

namespace exb {

#define EXB_PULL_WARP_BUF 8192    // bytes of row staging per warp in the pull kernel
#define EXB_APPLY_WARP_BUF 10240  // bytes per warp in the apply phase (w | state | acc rows)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __noinline__ float4 init_block_masked(const InitParams* I, unsigned long long id, int c, int dim);

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait() {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// Pull the 32 rows of a warp task through shared memory (cp.async gather, coalesced write-out).
// Lane l holds (src, id, flag) of row l. buf: EXB_PULL_WARP_BUF bytes owned by this warp.
__device__ __forceinline__ void pull_rows_bulk(const TableDev& T, const float* src, unsigned long long id,
                                               int flag, int b0, int n_rows, float* __restrict__ out,
                                               int io_stride, int off, int lane, unsigned char* buf) {
    const int wstride = T.wstride;
    const unsigned rowbytes = (unsigned)wstride * 4u;
    const int R = min(32, (int)(EXB_PULL_WARP_BUF / rowbytes));   // rows per pass (warp uniform)
    const int lpr = T.lpr, gl = lane % lpr, RP = 32 / lpr;
    float* rows = reinterpret_cast<float*>(buf);
    for (int r0 = 0; r0 < 32; r0 += R) {
        // ---- issue: lane group g copies rows g, g+RP, ... of this pass, 16 bytes per lane
        for (int jb = 0; jb < R; jb += RP) {             // warp-uniform trip count (body shuffles)
            const int j = jb + lane / lpr, r = r0 + j;
            const float* s = (const float*)__shfl_sync(0xffffffffu, (unsigned long long)src, r & 31);
            const int fl = __shfl_sync(0xffffffffu, flag, r & 31);
            if (j < R && r < 32 && fl == 1)
                for (int c = gl * 4; c < wstride; c += lpr * 4) cp_async16(rows + (size_t)j * wstride + c, s + c);
        }
        cp_async_commit_wait();
        __syncwarp();
        // ---- write out (request order), zeros / initializer values for rows that were not loaded
        for (int jb = 0; jb < R; jb += RP) {
            const int j = jb + lane / lpr, r = r0 + j;
            const unsigned long long idr = __shfl_sync(0xffffffffu, id, r & 31);
            const int fl = __shfl_sync(0xffffffffu, flag, r & 31);
            const int b = b0 + r;
            if (j >= R || r >= 32 || b >= n_rows) continue;
            float* dst = out + (size_t)b * io_stride + off;
            for (int c = gl * 4; c < wstride; c += lpr * 4) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (fl == 1) v = *reinterpret_cast<const float4*>(rows + (size_t)j * wstride + c);
                else if (fl == 2) v = init_block_masked(&T.init, idr, c, T.dim);
                *reinterpret_cast<float4*>(dst + c) = v;
            }
        }
        __syncwarp();
    }
}

struct WarpMeta {   // per-warp row metadata of the apply phase (shared memory)
    unsigned long long key[32];
    unsigned long long row[32];
    unsigned h[32];
    unsigned cnt[32];
    int flag[32];
};

// Apply the optimizer to the (up to) 32 unique rows of a warp task: weights, state and
// accumulator rows are gathered into shared memory with cp.async, the math reads them from
// there and writes results straight back to global memory.
// Lane l holds (key,row,h,cnt,flag) of row l on entry.
// KIND is the optimizer category as a compile-time constant: the per-element switch of
// opt_elem folds away and the four elements of a float4 become straight-line, independent
// instruction streams (the runtime-switch version was issue-latency bound: ~10 us per pass).
template <int KIND>
__device__ __noinline__ void apply_rows_bulk_k(const TableDev& T, const PlanDev& P, float* accbase,
                                               unsigned long long key, unsigned long long row, unsigned h,
                                               unsigned cnt, int flag, int lane, unsigned char* buf,
                                               WarpMeta* M, int nrows, unsigned long long* trt) {
    const int wstride = T.wstride, sstride = T.sstride, dim = T.dim, nslots = T.nslots, nsc = T.nscalars;
    const unsigned wb = (unsigned)wstride * 4u, sb = (unsigned)sstride * 4u;
    const int R = min(32, (int)(EXB_APPLY_WARP_BUF / (2u * wb + sb)));
    const int lpr = T.lpr, gl = lane % lpr, RP = 32 / lpr;
    float* wloc = T.w[P.rank];
    OptParams opt = T.opt;
    opt.kind = KIND;
    const float s0i = opt_slot_init<float>(opt, 0), s1i = opt_slot_init<float>(opt, 1);
    M->key[lane] = key; M->row[lane] = row; M->h[lane] = h; M->cnt[lane] = cnt; M->flag[lane] = flag;
    __syncwarp();
    float* wbuf = reinterpret_cast<float*>(buf);
    float* sbuf = reinterpret_cast<float*>(buf + (size_t)R * wb);
    float* abuf = reinterpret_cast<float*>(buf + (size_t)R * (wb + sb));
    for (int r0 = 0; r0 < nrows; r0 += R) {
        // ---- gather: lane group g fetches rows g, g+RP, ... of the pass
        for (int jj = lane / lpr; jj < R; jj += RP) {
            const int rr = r0 + jj;
            if (rr >= 32) break;
            const int f = M->flag[rr];
            if (!f) continue;
            const float* ga = accbase + (unsigned long long)M->h[rr] * wstride;
            for (int c = gl * 4; c < wstride; c += lpr * 4) cp_async16(abuf + (size_t)jj * wstride + c, ga + c);
            if (f == 1) {
                const float* gw = wloc + M->row[rr] * (unsigned long long)wstride;
                const float* gs = T.state + M->row[rr] * (unsigned long long)sstride;
                for (int c = gl * 4; c < wstride; c += lpr * 4) cp_async16(wbuf + (size_t)jj * wstride + c, gw + c);
                for (int c = gl * 4; c < sstride; c += lpr * 4) cp_async16(sbuf + (size_t)jj * sstride + c, gs + c);
            }
        }
        cp_async_commit_wait();
        __syncwarp();
        if (trt && lane == 0) trt[5] = globaltimer_ns();
        // ---- math out of shared memory, results go straight to global
        for (int jj = lane / lpr; jj < R; jj += RP) {
            const int rr = r0 + jj;
            if (rr >= 32) break;
            const int f = M->flag[rr];
            if (!f) continue;
            const float* wr = wbuf + (size_t)jj * wstride;
            const float* sr = sbuf + (size_t)jj * sstride;
            const float* ar = abuf + (size_t)jj * wstride;
            float* gw = wloc + M->row[rr] * (unsigned long long)wstride;
            float* gs = T.state + M->row[rr] * (unsigned long long)sstride;
            float* ga = accbase + (unsigned long long)M->h[rr] * wstride;
            float sc[2] = {0.f, 0.f}, nsc_v[2];
            for (int i = 0; i < nsc; ++i)
                sc[i] = (f == 2) ? opt_scalar_init<float>(opt, i) : sr[(size_t)nslots * wstride + i];
            RowCtx<float> rc = opt_row_prologue_pure<float>(opt, sc, (uint64_t)M->cnt[rr], nsc_v);
            for (int c = gl * 4; c < wstride; c += lpr * 4) {
                float4 g = *reinterpret_cast<const float4*>(ar + c);
                float4 w, a = make_float4(s0i, s0i, s0i, s0i), b = make_float4(s1i, s1i, s1i, s1i);
                if (f == 2) {
                    w = init_block_masked(&T.init, M->key[rr], c, dim);
                } else {
                    w = *reinterpret_cast<const float4*>(wr + c);
                    if (nslots > 0) a = *reinterpret_cast<const float4*>(sr + c);
                    if (nslots > 1) b = *reinterpret_cast<const float4*>(sr + wstride + c);
                }
                if (c + 0 < dim) opt_elem<float>(opt, rc, w.x, a.x, b.x, g.x);
                if (c + 1 < dim) opt_elem<float>(opt, rc, w.y, a.y, b.y, g.y);
                if (c + 2 < dim) opt_elem<float>(opt, rc, w.z, a.z, b.z, g.z);
                if (c + 3 < dim) opt_elem<float>(opt, rc, w.w, a.w, b.w, g.w);
                *reinterpret_cast<float4*>(gw + c) = w;
                if (nslots > 0) *reinterpret_cast<float4*>(gs + c) = a;
                if (nslots > 1) *reinterpret_cast<float4*>(gs + wstride + c) = b;
                *reinterpret_cast<float4*>(ga + c) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (gl == 0) {
                for (int i = 0; i < nsc; ++i) gs[(size_t)nslots * wstride + i] = nsc_v[i];
                if (f == 2)   // pad words of a brand-new state row
                    for (int i = nslots * wstride + nsc; i < sstride; ++i) gs[i] = 0.f;
            }
        }
        __syncwarp();
    }
}

__device__ __forceinline__ void apply_rows_bulk(const TableDev& T, const PlanDev& P, float* accbase,
                                                unsigned long long key, unsigned long long row, unsigned h,
                                                unsigned cnt, int flag, int lane, unsigned char* buf,
                                                WarpMeta* M, int nrows, unsigned long long* trt) {
    switch (T.opt.kind) {   // warp uniform
        case OPT_ADADELTA: apply_rows_bulk_k<OPT_ADADELTA>(T, P, accbase, key, row, h, cnt, flag, lane, buf, M, nrows, trt); break;
        case OPT_ADAGRAD: apply_rows_bulk_k<OPT_ADAGRAD>(T, P, accbase, key, row, h, cnt, flag, lane, buf, M, nrows, trt); break;
        case OPT_ADAM: apply_rows_bulk_k<OPT_ADAM>(T, P, accbase, key, row, h, cnt, flag, lane, buf, M, nrows, trt); break;
        case OPT_ADAMAX: apply_rows_bulk_k<OPT_ADAMAX>(T, P, accbase, key, row, h, cnt, flag, lane, buf, M, nrows, trt); break;
        case OPT_FTRL: apply_rows_bulk_k<OPT_FTRL>(T, P, accbase, key, row, h, cnt, flag, lane, buf, M, nrows, trt); break;
        case OPT_RMSPROP: apply_rows_bulk_k<OPT_RMSPROP>(T, P, accbase, key, row, h, cnt, flag, lane, buf, M, nrows, trt); break;
        case OPT_SGD: apply_rows_bulk_k<OPT_SGD>(T, P, accbase, key, row, h, cnt, flag, lane, buf, M, nrows, trt); break;
        case OPT_TEST: apply_rows_bulk_k<OPT_TEST>(T, P, accbase, key, row, h, cnt, flag, lane, buf, M, nrows, trt); break;
        default: apply_rows_bulk_k<OPT_DEFAULT>(T, P, accbase, key, row, h, cnt, flag, lane, buf, M, nrows, trt); break;
    }
}

}  // namespace exb
