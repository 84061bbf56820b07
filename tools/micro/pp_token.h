// pp_token.h -- measurement harness of the REJECTED ping-pong token experiment, used only by
// tools/micro/pp_bench.cu: two teams of one persistent block run the plain fused kernels' phases and hand a token for the
// L1 / shared-memory data pipe back and forth with bar.sync / bar.arrive.  Not part of libslmgs.so.
#pragma once

#include "slmgs_kernels.h"

namespace slmgs {

// the phases of ColKernel<N, COL_FUSED, VAR, CT> / RowKernel<N, ROW_FUSED, ..> with L1-burst hooks: one release and one
// acquire per phase; the token is held across the team barrier between phases
template <class K> struct PPCol {
    typedef typename K::F F;
    typedef typename K::Args Args;
    typedef typename K::State State;
    static constexpr int NS = K::NS, NPHASE = K::NPHASE;
    static constexpr int PP_THREADS = K::MAXT / 2;  // (2-column tiles at 4096 rows)
    static SLMGS_DEVICE int pp_items(const Args& a) { return a.W / (PP_THREADS / F::TPL); }
    template <int P, class Sy> static SLMGS_DEVICE void phase_sy(State& st, const Args& a, cf* smem, const ThreadId& id, Sy& sy) {
        const typename K::Loc L = K::locate(a, smem, id);
        if constexpr (P == 0) {
            K::load_rows(st, a, L);
            sy.release();
            F::template fwd_stage_sy<0>(st.v, L.lt, a.twA, a.twB, L.s, L.C, sy);
        } else if constexpr (P < NS - 1) {
            F::template fwd_stage_sy<P>(st.v, L.lt, a.twA, a.twB, L.s, L.C, sy);
        } else if constexpr (P == NS - 1) {
            K::prefetch_images_head(a, L);
            F::template fwd_stage_sy<NS - 1>(st.v, L.lt, a.twA, a.twB, L.s, L.C, sy);
            K::template constrain<false>(st, a, id, L);
            F::template inv_stage_sy<NS - 1>(st.v, L.lt, a.twA, a.twB, L.s, L.C, sy);
        } else {
            F::template inv_stage_sy<2 * NS - 2 - P>(st.v, L.lt, a.twA, a.twB, L.s, L.C, sy);
            if constexpr (P == NPHASE - 1) {
                sy.acquire();
                K::store_rows(st, a, L);
            }
        }
    }
};
template <class K, int LI> struct PPRow {
    typedef typename K::F F;
    typedef typename K::Args Args;
    typedef typename K::State State;
    static constexpr int NS = K::NS, NPHASE = K::NPHASE;
    static constexpr int PP_THREADS = K::TEAM;
    static SLMGS_DEVICE int pp_items(const Args& a) { return (a.h + LI - 1) / LI; }
    template <int P, class Sy> static SLMGS_DEVICE void phase_sy(State& st, const Args& a, cf* smem, const ThreadId& id, Sy& sy) {
        const typename K::Loc L = K::locate(a, smem, id);
        if constexpr (P == 0) {
            K::load_spectrum(st, a, L);
            sy.release();
        }
        if constexpr (P < NS - 1) {
            F::template inv_stage_sy<NS - 1 - P>(st.v, L.lt, a.twA, a.twB, L.s, LI, sy);
        } else if constexpr (P == NS - 1) {
            F::template inv_stage_sy<0>(st.v, L.lt, a.twA, a.twB, L.s, LI, sy);
            K::template project<true, false>(st, a, id, L);
            F::template fwd_stage_sy<0>(st.v, L.lt, a.twA, a.twB, L.s, LI, sy);
        } else {
            F::template fwd_stage_sy<P - (NS - 1)>(st.v, L.lt, a.twA, a.twB, L.s, LI, sy);
        }
        if constexpr (P == NPHASE - 1) {
            sy.acquire();
            K::store_spectrum(st, a, L);
        }
    }
};

// ------------------------------------------------------------------------------------------------------------------
// Ping-pong teams.  slmgs_kernel_pp<K> runs TWO independent teams of K::PP_THREADS threads in one persistent block
// (one block per SM); each team walks over its own work items (column tiles / row groups) with its own shared-memory
// slice and its own named barrier.  Two equal blocks that start together on one SM stay in lock step: both queue
// their shared-memory exchange at the same time, both wait for it, then both compete for the FMA pipe -- the L1 /
// shared-memory data pipe and the FMA pipe take turns instead of overlapping (ncu: each ~50 % busy).  Here the bursts on the L1 data
// pipe (global loads / stores, exchange writes + reads) are bracketed by a token that the teams hand back and forth
// (bar.sync / bar.arrive on two named barriers, K::phase_sy): team B's burst always queues behind team A's, so A's
// butterflies run under B's exchange and vice versa.
// Every phase releases and acquires exactly once; a team holds the token at phase boundaries.
// ------------------------------------------------------------------------------------------------------------------
template <int T> struct TeamSync {
    int mine, other;  // named barriers: `mine` = this team waits here for the token, `other` = the partner does
    SLMGS_DEVICE void acquire() {
        SLMGS_PP_STAMP(mine - 3, 0);
        asm volatile("bar.sync %0, %1;" ::"r"(mine), "n"(2 * T) : "memory");
        SLMGS_PP_STAMP(mine - 3, 1);
    }
    SLMGS_DEVICE void release() {
        asm volatile("bar.arrive %0, %1;" ::"r"(other), "n"(2 * T) : "memory");
        SLMGS_PP_STAMP(mine - 3, 2);
    }
};
template <class K, int P, class Sy>
SLMGS_DEVICE void run_phases_pp(typename K::State& st, const typename K::Args& a, cf* smem, const ThreadId& id, Sy& sy,
                                int team, bool active) {
    if (active) {
        K::template phase_sy<P>(st, a, smem, id, sy);
    } else {  // a team without a work item in the last round keeps the token moving
        sy.release();
        sy.acquire();
    }
    if constexpr (P + 1 < K::NPHASE) {
        asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "n"(K::PP_THREADS) : "memory");
        SLMGS_PP_STAMP(team, 3);
        run_phases_pp<K, P + 1>(st, a, smem, id, sy, team, active);
    }
}
template <class K> __global__ void __launch_bounds__(2 * K::PP_THREADS, 1) slmgs_kernel_pp(const typename K::Args a, int team_smem_cf) {
    extern __shared__ __align__(16) unsigned char slmgs_smem_raw[];
    constexpr int T = K::PP_THREADS;
    const int team = threadIdx.x / T;
    SLMGS_PP_STAMP(team, 6);
    cf* smem = reinterpret_cast<cf*>(slmgs_smem_raw) + (size_t)team * team_smem_cf;
    typename K::State st;
    ThreadId id;
    id.tid = threadIdx.x % T;
    id.nthreads = T;
    id.by = blockIdx.y;
    id.it = 0;
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int items = K::pp_items(a);
    id.gx = items;
    const int stride = 2 * gridDim.x;
    const int rounds = (items + stride - 1) / stride;
    TeamSync<T> sy;
    sy.mine = 3 + team;
    sy.other = 4 - team;
    if (team == 1) sy.release();  // team 0 starts with the token
    sy.acquire();
    for (int r = 0; r < rounds; ++r) {
        id.bx = 2 * blockIdx.x + team + r * stride;
        run_phases_pp<K, 0>(st, a, smem, id, sy, team, id.bx < items);
        // the exchange buffer is written again in phase 0 of the next item
        asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "n"(T) : "memory");
    }
    if (team == 0) sy.release();
}

// gx: persistent blocks per hologram (<= number of SMs); team_threads = K::PP_THREADS (host copy of the constant)
template <class K>
int launch_kernel_pp(int gx, int gy, size_t team_smem_bytes, cudaStream_t stream, const typename K::Args& a, bool pdl = false) {
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(slmgs_kernel_pp<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_set[dev & 63] = true;
    }
    const size_t team_cf = (team_smem_bytes + sizeof(cf) - 1) / sizeof(cf);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(gx, gy, 1);
    cfg.blockDim = dim3(2 * K::PP_THREADS, 1, 1);
    cfg.dynamicSmemBytes = 2 * team_cf * sizeof(cf);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return (int)cudaLaunchKernelEx(&cfg, slmgs_kernel_pp<K>, a, (int)team_cf);
}

}  // namespace slmgs
