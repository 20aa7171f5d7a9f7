// sparse_kernels.cuh -- the two fused sparse hot paths of the engine (sm_100a).
//
//  exb_pull_kernel        K1+K2+K3 of SURVEY 2.5: bucketize by owner (id % W), one-sided
//                         peer loads of the rows over NVLink (array: direct address, hash:
//                         probe in the owner's key slab), scatter into request order.
//                         Missing hash rows are answered with the Philox initial value.
//  exb_push_update_kernel K4a+K4b: persistent kernel, phases separated by grid barriers:
//     P1 dispatch   (id, grad) -> owner inbox with P2P vector stores; local ids skip the
//                   inbox and go straight to the combine map
//     B1            publish counts, cross-GPU flag barrier (release/acquire .sys)
//     P3 combine    owner folds inbox entries into a per-step open-addressing map
//                   (atomicCAS on key, red.global.add.v4.f32 on the accumulator row)
//     B2
//     P5 apply      one lane group per unique row: (hash: find-or-insert, Philox init of
//                   new rows) -> optimizer functor -> write back, reset map entry
//     B3            cross-GPU "update done" barrier (next pull may read any shard)
//
// Reference semantics preserved: gradients of duplicate ids are SUMMED, counts are
// summed (MpscGradientReducer.h:30-53); rows are materialised at their first update.
#pragma once
#include "exb_common.cuh"

namespace exb {

__device__ __forceinline__ int find_segment(const int* __restrict__ prefix, int n, int task) {
    int lo = 0, hi = n;  // prefix[lo] <= task < prefix[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (prefix[mid] <= task) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ float4 init_block_masked(const InitParams& I, unsigned long long id,
                                                    int c, int dim) {
    float t[4];
    InitGen<float>::block4(I, id, (uint32_t)(c >> 2), t);
    float4 v;
    v.x = (c + 0 < dim) ? t[0] : 0.f;
    v.y = (c + 1 < dim) ? t[1] : 0.f;
    v.z = (c + 2 < dim) ? t[2] : 0.f;
    v.w = (c + 3 < dim) ? t[3] : 0.f;
    return v;
}

// ------------------------------------------------------------------ pull
// flag: 0 -> zeros (invalid id / padding row), 1 -> load from src, 2 -> initializer value
template <int LPR>
__device__ __forceinline__ void pull_rows(const TableDev& T, const float* src, unsigned long long id,
                                          int flag, int b0, int n_rows, float* __restrict__ out,
                                          int io_stride, int off, int lane) {
    constexpr int RP = 32 / LPR;
    const int gl = lane % LPR;
    const int wstride = T.wstride, dim = T.dim;
#pragma unroll 4
    for (int p = 0; p < LPR; ++p) {
        int r = p * RP + lane / LPR;
        const float* s = (const float*)__shfl_sync(0xffffffffu, (unsigned long long)src, r);
        unsigned long long idr = __shfl_sync(0xffffffffu, id, r);
        int fl = __shfl_sync(0xffffffffu, flag, r);
        int b = b0 + r;
        if (b >= n_rows) continue;
        float* dst = out + (size_t)b * io_stride + off;
        if (T.vec4) {
            for (int c = gl * 4; c < wstride; c += LPR * 4) {
                float4 v;
                if (fl == 1) v = ld_stream_v4(s + c);
                else if (fl == 2) v = init_block_masked(T.init, idr, c, dim);
                else v = make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(dst + c) = v;
            }
        } else {
            for (int c = gl; c < dim; c += LPR) {
                float v = 0.f;
                if (fl == 1) v = s[c];
                else if (fl == 2) {
                    float t[4];
                    InitGen<float>::block4(T.init, idr, 0u, t);
                    v = t[c & 3];
                }
                dst[c] = v;
            }
        }
    }
}

__global__ void __launch_bounds__(256)
exb_pull_kernel(const TableDev* __restrict__ tables, PlanDev P, const long long* __restrict__ ids,
                float* __restrict__ out, int n_rows) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int W = P.W;
    for (int task = warp; task < P.num_tasks; task += nwarps) {
        const int f = find_segment(P.task_prefix, P.F, task);
        const int b0 = (task - P.task_prefix[f]) * 32;
        if (b0 >= n_rows) continue;
        const TableDev& T = tables[P.pt_table[P.feat_pt[f]]];
        const int b = b0 + lane;
        unsigned long long id = 0;
        const float* src = nullptr;
        int flag = 0;
        if (b < n_rows) {
            id = (unsigned long long)__ldg(ids + (size_t)b * P.ncols + P.feat_col[f]);
            if (!T.is_hash) {
                if (id < T.vocab) {
                    int o = owner_of(T, id, W);
                    unsigned long long row = local_row_of(T, id);
                    src = T.w[o] + row * (unsigned long long)T.wstride;
                    flag = 1;
                }
            } else if ((id >> 63) == 0) {
                int o = owner_of(T, id, W);
                const unsigned long long* keys = T.keys[o];
                unsigned long long mask = T.rows - 1, h = exb_hash64(id) & mask;
                flag = 2;
                for (unsigned long long probe = 0; probe <= mask; ++probe) {
                    unsigned long long k = keys[h];
                    if (k == id) {
                        src = T.w[o] + h * (unsigned long long)T.wstride;
                        flag = 1;
                        break;
                    }
                    if (k == EXB_EMPTY_KEY) break;
                    h = (h + 1) & mask;
                }
            }
        }
        const int off = P.feat_off[f];
        switch (T.lpr) {
            case 1: pull_rows<1>(T, src, id, flag, b0, n_rows, out, P.io_stride, off, lane); break;
            case 2: pull_rows<2>(T, src, id, flag, b0, n_rows, out, P.io_stride, off, lane); break;
            case 4: pull_rows<4>(T, src, id, flag, b0, n_rows, out, P.io_stride, off, lane); break;
            case 8: pull_rows<8>(T, src, id, flag, b0, n_rows, out, P.io_stride, off, lane); break;
            case 16: pull_rows<16>(T, src, id, flag, b0, n_rows, out, P.io_stride, off, lane); break;
            default: pull_rows<32>(T, src, id, flag, b0, n_rows, out, P.io_stride, off, lane); break;
        }
    }
    if (threadIdx.x == 0 && blockIdx.x == 0)
        atomicAdd(&P.stats[0], (unsigned long long)n_rows * (unsigned long long)P.F);
}

// ------------------------------------------------------------ push + update
#define EXB_MAX_SEG 1024  // max W*PT segments of the combine phase

// find-or-insert `key` into plan-table pt's combine map; returns map position
__device__ __forceinline__ unsigned cmap_insert(const PlanDev& P, int pt, unsigned long long key) {
    const unsigned mask = P.pt_map_mask[pt];
    unsigned long long* keys = P.cmap_keys + P.pt_map_off[pt];
    unsigned h = (unsigned)(exb_hash64(key) >> 20) & mask;
    for (unsigned probe = 0; probe <= mask; ++probe) {
        unsigned long long k = *(volatile unsigned long long*)&keys[h];
        if (k == key) return h;
        if (k == EXB_EMPTY_KEY) {
            unsigned long long prev = atomicCAS(&keys[h], EXB_EMPTY_KEY, key);
            if (prev == EXB_EMPTY_KEY) {
                unsigned u = atomicAdd(&P.ucount[pt], 1u);
                P.ulist[P.pt_ulist_off[pt] + u] = h;
                return h;
            }
            if (prev == key) return h;
        }
        h = (h + 1) & mask;
    }
    set_error(P.status, EXB_ERR_CMAP_FULL);
    return 0xFFFFFFFFu;
}

// Move / accumulate the 32 rows of a warp task.
// mode: 0 skip, 1 accumulate row into acc[h] (local red.add), 2 store row to dst (peer inbox)
template <int LPR>
__device__ __forceinline__ void move_rows(const TableDev& T, const float* src, float* dst, int mode,
                                          int lane) {
    constexpr int RP = 32 / LPR;
    const int gl = lane % LPR;
    const int wstride = T.wstride, dim = T.dim;
#pragma unroll 4
    for (int p = 0; p < LPR; ++p) {
        int r = p * RP + lane / LPR;
        const float* s = (const float*)__shfl_sync(0xffffffffu, (unsigned long long)src, r);
        float* d = (float*)__shfl_sync(0xffffffffu, (unsigned long long)dst, r);
        int m = __shfl_sync(0xffffffffu, mode, r);
        if (m == 0) continue;
        if (T.vec4) {
            for (int c = gl * 4; c < wstride; c += LPR * 4) {
                float4 v = *reinterpret_cast<const float4*>(s + c);
                if (m == 1) red_add_v4(d + c, v);
                else *reinterpret_cast<float4*>(d + c) = v;
            }
        } else {
            for (int c = gl; c < dim; c += LPR) {
                float v = s[c];
                if (m == 1) red_add_f32(d + c, v);
                else d[c] = v;
            }
        }
    }
}

__device__ __forceinline__ void move_rows_dispatch(const TableDev& T, const float* src, float* dst,
                                                   int mode, int lane) {
    switch (T.lpr) {
        case 1: move_rows<1>(T, src, dst, mode, lane); break;
        case 2: move_rows<2>(T, src, dst, mode, lane); break;
        case 4: move_rows<4>(T, src, dst, mode, lane); break;
        case 8: move_rows<8>(T, src, dst, mode, lane); break;
        case 16: move_rows<16>(T, src, dst, mode, lane); break;
        default: move_rows<32>(T, src, dst, mode, lane); break;
    }
}

// block-wide exclusive prefix of ceil(cnt/32) over n (<= EXB_MAX_SEG) segments -> s_prefix[0..n]
__device__ __forceinline__ void block_task_prefix(const unsigned* cnt, int n, int* s_prefix) {
    // each of the first 32 threads scans a contiguous chunk, then a warp scan of chunk sums
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const int chunk = (n + 31) / 32;
        int beg = lane * chunk, end = min(n, beg + chunk);
        int sum = 0;
        for (int i = beg; i < end; ++i) sum += (int)((cnt[i] + 31u) >> 5);
        int incl = sum;
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        int run = incl - sum;
        for (int i = beg; i < end; ++i) {
            s_prefix[i] = run;
            run += (int)((cnt[i] + 31u) >> 5);
        }
        if (lane == 31) s_prefix[n] = incl;
    }
    __syncthreads();
}

template <int LPR>
__device__ __forceinline__ void apply_rows(const TableDev& T, const PlanDev& P, int pt,
                                           unsigned long long key, unsigned long long row, unsigned h,
                                           unsigned cnt, int flag /*0 skip,1 existing,2 new*/,
                                           int lane) {
    constexpr int RP = 32 / LPR;
    const int gl = lane % LPR;
    const int wstride = T.wstride, dim = T.dim, nslots = T.nslots, nsc = T.nscalars;
    float* accbase = P.acc + P.pt_acc_off[pt];
    float* wloc = T.w[P.rank];
#pragma unroll 2
    for (int p = 0; p < LPR; ++p) {
        int r = p * RP + lane / LPR;
        unsigned long long keyr = __shfl_sync(0xffffffffu, key, r);
        unsigned long long rowr = __shfl_sync(0xffffffffu, row, r);
        unsigned hr = __shfl_sync(0xffffffffu, h, r);
        unsigned cr = __shfl_sync(0xffffffffu, cnt, r);
        int fl = __shfl_sync(0xffffffffu, flag, r);
        if (fl == 0) continue;
        float* wrow = wloc + rowr * (unsigned long long)wstride;
        float* srow = T.state + rowr * (unsigned long long)T.sstride;
        float* arow = accbase + (unsigned long long)hr * wstride;
        // per-row scalar prologue (evaluated redundantly by the group, committed by lane 0)
        float sc[2] = {0.f, 0.f}, nsc_v[2];
        float* scal = srow + (size_t)nslots * wstride;
        for (int i = 0; i < nsc; ++i) sc[i] = (fl == 2) ? opt_scalar_init<float>(T.opt, i) : scal[i];
        RowCtx<float> rc = opt_row_prologue_pure<float>(T.opt, sc, (uint64_t)cr, nsc_v);
        if (T.vec4) {
            for (int c = gl * 4; c < wstride; c += LPR * 4) {
                float4 g4 = *reinterpret_cast<float4*>(arow + c);
                *reinterpret_cast<float4*>(arow + c) = make_float4(0.f, 0.f, 0.f, 0.f);
                float4 w4, a4, b4;
                float s0i = opt_slot_init<float>(T.opt, 0), s1i = opt_slot_init<float>(T.opt, 1);
                if (fl == 2) {
                    w4 = init_block_masked(T.init, keyr, c, dim);
                    a4 = make_float4(s0i, s0i, s0i, s0i);
                    b4 = make_float4(s1i, s1i, s1i, s1i);
                } else {
                    w4 = *reinterpret_cast<float4*>(wrow + c);
                    a4 = nslots > 0 ? *reinterpret_cast<float4*>(srow + c) : make_float4(0, 0, 0, 0);
                    b4 = nslots > 1 ? *reinterpret_cast<float4*>(srow + wstride + c)
                                    : make_float4(0, 0, 0, 0);
                }
                float* w = &w4.x; float* a = &a4.x; float* b = &b4.x; const float* g = &g4.x;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (c + i < dim) opt_elem<float>(T.opt, rc, w[i], a[i], b[i], g[i]);
                *reinterpret_cast<float4*>(wrow + c) = w4;
                if (nslots > 0) *reinterpret_cast<float4*>(srow + c) = a4;
                if (nslots > 1) *reinterpret_cast<float4*>(srow + wstride + c) = b4;
            }
        } else {
            for (int c = gl; c < dim; c += LPR) {
                float g = arow[c];
                arow[c] = 0.f;
                float w, a = 0.f, b = 0.f;
                if (fl == 2) {
                    float t[4];
                    InitGen<float>::block4(T.init, keyr, 0u, t);
                    w = t[c & 3];
                    a = opt_slot_init<float>(T.opt, 0);
                    b = opt_slot_init<float>(T.opt, 1);
                } else {
                    w = wrow[c];
                    if (nslots > 0) a = srow[c];
                    if (nslots > 1) b = srow[wstride + c];
                }
                opt_elem<float>(T.opt, rc, w, a, b, g);
                wrow[c] = w;
                if (nslots > 0) srow[c] = a;
                if (nslots > 1) srow[wstride + c] = b;
            }
        }
        if (gl == 0)
            for (int i = 0; i < nsc; ++i) scal[i] = nsc_v[i];
    }
}

__global__ void __launch_bounds__(256)
exb_push_update_kernel(const TableDev* __restrict__ tables, PlanDev P,
                       const long long* __restrict__ ids, const float* __restrict__ grads,
                       int n_rows) {
    __shared__ int s_prefix[EXB_MAX_SEG + 1];
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int W = P.W, PT = P.PT, rank = P.rank;

    // ---------------- P1: dispatch (remote ids -> owner inbox, local ids -> combine map)
    for (int task = warp; task < P.num_tasks; task += nwarps) {
        const int f = find_segment(P.task_prefix, P.F, task);
        const int b0 = (task - P.task_prefix[f]) * 32;
        if (b0 >= n_rows) continue;
        const int pt = P.feat_pt[f];
        const TableDev& T = tables[P.pt_table[pt]];
        const int b = b0 + lane;
        unsigned long long id = 0;
        int owner = -1;
        if (b < n_rows) {
            id = (unsigned long long)__ldg(ids + (size_t)b * P.ncols + P.feat_col[f]);
            bool ok = T.is_hash ? ((id >> 63) == 0) : (id < T.vocab);
            if (ok) owner = owner_of(T, id, W);
        }
        const float* src = grads + (size_t)b * P.io_stride + P.feat_off[f];
        float* dst = nullptr;
        int mode = 0;
        if (owner == rank) {
            unsigned h = cmap_insert(P, pt, id);
            if (h != 0xFFFFFFFFu) {
                atomicAdd(&P.cmap_cnt[P.pt_map_off[pt] + h], 1u);
                dst = P.acc + P.pt_acc_off[pt] + (unsigned long long)h * T.wstride;
                mode = 1;
            }
        }
        if (W > 1) {
            unsigned m = __match_any_sync(0xffffffffu, owner);
            if (owner >= 0 && owner != rank) {
                int leader = __ffs(m) - 1;
                unsigned base = 0;
                if (lane == leader) base = atomicAdd(&P.send_cnt[owner * PT + pt], (unsigned)__popc(m));
                base = __shfl_sync(m, base, leader);
                unsigned pos = base + (unsigned)__popc(m & ((1u << lane) - 1u));
                if (pos < P.pt_cap[pt]) {
                    P.inbox_keys[owner][(unsigned long long)rank * P.src_key_stride + P.pt_key_off[pt] + pos] = id;
                    dst = P.inbox_grads[owner] + (unsigned long long)rank * P.src_grad_stride +
                          P.pt_grad_off[pt] + (unsigned long long)pos * T.wstride;
                    mode = 2;
                } else {
                    set_error(P.status, EXB_ERR_INBOX_OVERFLOW);
                }
            }
        }
        move_rows_dispatch(T, src, dst, mode, lane);
    }

    if (W > 1) {
        // ---------------- B1: publish counts, cross-GPU barrier
        grid_barrier(P, true, [&]() {
            for (int i = threadIdx.x; i < W * PT; i += blockDim.x) {
                int o = i / PT, pt = i - o * PT;
                unsigned c = P.send_cnt[i];
                if (c > P.pt_cap[pt]) c = P.pt_cap[pt];
                if (o != rank) P.inbox_cnt[o][rank * PT + pt] = c;
                P.send_cnt[i] = 0;
            }
            peer_barrier(P);
        });
        // ---------------- P3: combine inbox entries of every remote source
        const unsigned* mycnt = P.inbox_cnt[rank];
        block_task_prefix(mycnt, W * PT, s_prefix);
        const int ntask3 = s_prefix[W * PT];
        for (int task = warp; task < ntask3; task += nwarps) {
            const int seg = find_segment(s_prefix, W * PT, task);
            const int s = seg / PT, pt = seg - s * PT;
            if (s == rank) continue;  // local ids never travel through the inbox
            const TableDev& T = tables[P.pt_table[pt]];
            const unsigned e = (unsigned)(task - s_prefix[seg]) * 32u + lane;
            const unsigned n = mycnt[seg];
            const float* src = nullptr;
            float* dst = nullptr;
            int mode = 0;
            if (e < n) {
                unsigned long long key =
                    P.inbox_keys[rank][(unsigned long long)s * P.src_key_stride + P.pt_key_off[pt] + e];
                unsigned h = cmap_insert(P, pt, key);
                if (h != 0xFFFFFFFFu) {
                    atomicAdd(&P.cmap_cnt[P.pt_map_off[pt] + h], 1u);
                    src = P.inbox_grads[rank] + (unsigned long long)s * P.src_grad_stride +
                          P.pt_grad_off[pt] + (unsigned long long)e * T.wstride;
                    dst = P.acc + P.pt_acc_off[pt] + (unsigned long long)h * T.wstride;
                    mode = 1;
                }
            }
            move_rows_dispatch(T, src, dst, mode, lane);
        }
    }

    // ---------------- B2: all accumulations visible
    grid_barrier(P, false, [&]() {});

    // ---------------- P5: apply optimizer to every unique row
    block_task_prefix(P.ucount, PT, s_prefix);
    const int ntask5 = s_prefix[PT];
    unsigned long long n_unique_local = 0;
    for (int task = warp; task < ntask5; task += nwarps) {
        const int pt = find_segment(s_prefix, PT, task);
        const TableDev& T = tables[P.pt_table[pt]];
        const unsigned u = (unsigned)(task - s_prefix[pt]) * 32u + lane;
        const unsigned n = P.ucount[pt];
        unsigned long long key = 0, row = 0;
        unsigned h = 0, cnt = 0;
        int flag = 0;
        if (u < n) {
            h = P.ulist[P.pt_ulist_off[pt] + u];
            const unsigned long long mo = P.pt_map_off[pt] + h;
            key = *(volatile unsigned long long*)&P.cmap_keys[mo];
            cnt = *(volatile unsigned*)&P.cmap_cnt[mo];
            P.cmap_keys[mo] = EXB_EMPTY_KEY;
            P.cmap_cnt[mo] = 0;
            if (!T.is_hash) {
                row = local_row_of(T, key);
                atomicOr(&T.touched[row >> 5], 1u << (row & 31));
                flag = 1;
            } else {
                unsigned long long* keys = const_cast<unsigned long long*>(T.keys[rank]);
                unsigned long long mask = T.rows - 1, hh = exb_hash64(key) & mask;
                for (unsigned long long probe = 0; probe <= mask; ++probe) {
                    unsigned long long k = *(volatile unsigned long long*)&keys[hh];
                    if (k == key) { flag = 1; break; }
                    if (k == EXB_EMPTY_KEY) {
                        unsigned long long prev = atomicCAS(&keys[hh], EXB_EMPTY_KEY, key);
                        if (prev == EXB_EMPTY_KEY) { flag = 2; atomicAdd(T.size_ctr, 1ull); break; }
                        if (prev == key) { flag = 1; break; }
                    }
                    hh = (hh + 1) & mask;
                }
                if (flag == 0) set_error(P.status, EXB_ERR_HASH_FULL);
                row = hh;
            }
            if (flag) ++n_unique_local;
        }
        switch (T.lpr) {
            case 1: apply_rows<1>(T, P, pt, key, row, h, cnt, flag, lane); break;
            case 2: apply_rows<2>(T, P, pt, key, row, h, cnt, flag, lane); break;
            case 4: apply_rows<4>(T, P, pt, key, row, h, cnt, flag, lane); break;
            case 8: apply_rows<8>(T, P, pt, key, row, h, cnt, flag, lane); break;
            case 16: apply_rows<16>(T, P, pt, key, row, h, cnt, flag, lane); break;
            default: apply_rows<32>(T, P, pt, key, row, h, cnt, flag, lane); break;
        }
    }
    if (n_unique_local) atomicAdd(&P.stats[2], n_unique_local);

    // ---------------- B3: reset per-step counters; cross-GPU "update done"
    grid_barrier(P, W > 1, [&]() {
        for (int i = threadIdx.x; i < PT; i += blockDim.x) P.ucount[i] = 0;
        if (threadIdx.x == 0) atomicAdd(&P.stats[1], (unsigned long long)n_rows * P.F);
        if (W > 1) peer_barrier(P);
    });
}

}  // namespace exb
