// slmgs_launch.h -- phase-structured kernels: device driver and host-emulation driver.
//
// A kernel class K provides
//     typedef ... Args;            POD argument block (passed by value)
//     struct State { ... };        per-thread registers that live across barriers
//     static constexpr int NPHASE; number of barrier-delimited phases
//     static constexpr int MAXT;   max threads per block (launch bound)
//     template <int P> static SLMGS_DEVICE void phase(State&, const Args&, cf* smem, const ThreadId&);
// On the device the phases are chained with __syncthreads() in ONE kernel.  With
// -DSLMGS_EMULATE (test infrastructure, host compiler) they are executed by loops over
// (block, phase, thread), which is equivalent because phases only communicate through shared
// memory across barriers.
#pragma once

#include "slmgs_common.h"

#include <string.h>
#ifdef SLMGS_EMULATE
#include <vector>
#endif

namespace slmgs {

struct ThreadId {
    int tid;       // threadIdx.x
    int nthreads;  // blockDim.x
    int bx, by;    // blockIdx.x, blockIdx.y
    int gx;        // gridDim.x
    int it;        // persistent kernels: tile iteration of this block (0 otherwise)
};

// ---- block-level sum into a global double accumulator ------------------------------------
#ifndef SLMGS_EMULATE
SLMGS_DEVICE double warp_sum(double v) {
    SLMGS_UNROLL
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Must be called by every thread of the block (full warps).
SLMGS_DEVICE void accum_add(double* slot, double v) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(slot, v);
}
SLMGS_DEVICE void atomic_add_double(double* slot, double v) { atomicAdd(slot, v); }
SLMGS_DEVICE void atomic_min_float_pos(float* slot, float v) {
    // valid for non-negative floats (ordered like their bit patterns)
    atomicMin(reinterpret_cast<int*>(slot), __float_as_int(v));
}
#else
inline void accum_add(double* slot, double v) { *slot += v; }
inline void atomic_add_double(double* slot, double v) { *slot += v; }
#endif

// optional K::skip(args, id): the whole block has nothing to do (evaluated once, before the first phase)
template <class K, class = void> struct has_skip {
    static constexpr bool value = false;
};
template <class K> struct has_skip<K, decltype((void)&K::skip)> {
    static constexpr bool value = true;
};
// optional K::iterations(args, id): PERSISTENT kernel -- the block runs its phase sequence that many times
// (id.it = 0, 1, ...), with a barrier between iterations; optional K::init(args, smem, id) runs once per thread
// before the first iteration, followed by a barrier (mbarrier set-up).
template <class K, class = void> struct has_iterations {
    static constexpr bool value = false;
};
template <class K> struct has_iterations<K, decltype((void)&K::iterations)> {
    static constexpr bool value = true;
};

#ifndef SLMGS_EMULATE
#ifdef SLMGS_TRACE
// Build-time diagnostic (-DSLMGS_TRACE, never in the product build): per-phase SM clock stamps of a few blocks.
// slot layout: [block 0..7][thread group 0..1][64 stamps]; stamp 2P = phase P done, 2P+1 = barrier behind it passed.
static __device__ long long slmgs_trace_buf[8 * 2 * 64];
static __device__ int slmgs_trace_on;
template <class K, class = void> struct trace_class {
    static constexpr int value = 0;
};
template <class K> struct trace_class<K, decltype((void)K::TRACE_CLASS)> {
    static constexpr int value = K::TRACE_CLASS;
};
template <class K> SLMGS_DEVICE void trace_stamp(const ThreadId& id, int slot) {
    if (slmgs_trace_on == trace_class<K>::value && id.bx < 8 && id.by == 0 && (id.tid == 0 || id.tid == id.nthreads - 1) && slot < 64) {
        long long t;
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
        slmgs_trace_buf[(id.bx * 2 + (id.tid != 0)) * 64 + slot] = t;
    }
}
#define SLMGS_STAMP(id, slot) trace_stamp<K>(id, slot)
#else
#define SLMGS_STAMP(id, slot)
#endif
template <class K, int P>
SLMGS_DEVICE void run_phases(typename K::State& st, const typename K::Args& a, cf* smem, const ThreadId& id) {
    K::template phase<P>(st, a, smem, id);
    SLMGS_STAMP(id, 2 + id.it * 2 * K::NPHASE + 2 * P);
    if constexpr (P + 1 < K::NPHASE) {
        K::barrier(id);
        SLMGS_STAMP(id, 2 + id.it * 2 * K::NPHASE + 2 * P + 1);
        run_phases<K, P + 1>(st, a, smem, id);
    }
}

// optional K::MINB: minimum resident blocks per SM (register cap), default 1
template <class K, class = void> struct min_blocks {
    static constexpr int value = 1;
};
template <class K> struct min_blocks<K, decltype((void)K::MINB)> {
    static constexpr int value = K::MINB;
};

template <class K> __global__ void __launch_bounds__(K::MAXT, min_blocks<K>::value) slmgs_kernel(const typename K::Args a) {
    extern __shared__ __align__(16) unsigned char slmgs_smem_raw[];
    cf* smem = reinterpret_cast<cf*>(slmgs_smem_raw);
    typename K::State st;
    ThreadId id;
    id.tid = threadIdx.x;
    id.nthreads = blockDim.x;
    id.bx = blockIdx.x;
    id.by = blockIdx.y;
    id.gx = gridDim.x;
    // Programmatic dependent launch: let the next kernel of the stream start filling SMs as this grid's last
    // wave drains (launch_dependents), and do not touch memory before the previous grid has completed and
    // flushed (wait).  Both are no-ops when the kernel is launched without the PDL attribute.
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    id.it = 0;
    if constexpr (has_skip<K>::value) {
        if (K::skip(a, id)) return;
    }
    SLMGS_STAMP(id, 0);
    if constexpr (has_iterations<K>::value) {
        const int nit = K::iterations(a, id);
        K::init(a, smem, id);
        K::barrier(id);
        for (int it = 0; it < nit; ++it) {
            id.it = it;
            run_phases<K, 0>(st, a, smem, id);
            K::barrier(id);
            SLMGS_STAMP(id, 2 + it * 2 * K::NPHASE + 2 * K::NPHASE - 1);
        }
    } else {
        run_phases<K, 0>(st, a, smem, id);
    }
}

// returns cudaError_t as int
template <class K>
int launch_kernel(int gx, int gy, int nthreads, size_t smem_bytes, cudaStream_t stream, const typename K::Args& a,
                  bool pdl = false) {
    static bool attr_set[64] = {false};  // per instantiation, per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(slmgs_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_set[dev & 63] = true;
    }
    if (pdl) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3(gx, gy, 1);
        cfg.blockDim = dim3(nthreads, 1, 1);
        cfg.dynamicSmemBytes = smem_bytes;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        return (int)cudaLaunchKernelEx(&cfg, slmgs_kernel<K>, a);
    }
    slmgs_kernel<K><<<dim3(gx, gy, 1), dim3(nthreads, 1, 1), smem_bytes, stream>>>(a);
    return (int)cudaGetLastError();
}

// development trace of the team kernels (tools/micro/pp_bench.cu, -DSLMGS_PP_TRACE): clock stamps of one thread per team
#ifdef SLMGS_PP_TRACE
#ifndef SLMGS_PP_TRACE_TID
#define SLMGS_PP_TRACE_TID 0
#endif
// development diagnostic (tools/micro/pp_bench.cu): clock stamps of one thread per team of block 0
static __device__ long long slmgs_pp_trace[2 * 512];
static __device__ int slmgs_pp_trace_n[2];
// (the stamp counter lives in shared memory: a global counter would put a DRAM round trip into every stamp)
SLMGS_DEVICE void pp_stamp(int team, int kind) {
    if (blockIdx.x == 0 && (threadIdx.x & 511) == SLMGS_PP_TRACE_TID) {
        __shared__ int cnt[2];
        long long t;
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
        int n = cnt[team];
        if (n < 0 || n >= 512) n = 0;   // (uninitialised at the first stamp: the host zeroes the buffer and reads until a 0)
        if (kind == 6) n = 0;
        slmgs_pp_trace[team * 512 + n] = t * 8 + kind;
        cnt[team] = n + 1;
        slmgs_pp_trace_n[team] = n + 1;
    }
}
#define SLMGS_PP_STAMP(team, kind) pp_stamp(team, kind)
#else
#define SLMGS_PP_STAMP(team, kind)
#endif
#else
template <class K, int P>
inline void emu_phases(std::vector<typename K::State>& st, const typename K::Args& a, cf* smem, ThreadId id) {
    for (int t = 0; t < id.nthreads; ++t) {
        id.tid = t;
        K::template phase<P>(st[t], a, smem, id);
    }
    if constexpr (P + 1 < K::NPHASE) emu_phases<K, P + 1>(st, a, smem, id);
}

template <class K>
int launch_kernel(int gx, int gy, int nthreads, size_t smem_bytes, void* /*stream*/, const typename K::Args& a,
                  bool /*pdl*/ = false) {
    std::vector<typename K::State> st(nthreads);
    std::vector<cf> smem(smem_bytes / sizeof(cf) + 2);
    for (int by = 0; by < gy; ++by)
        for (int bx = 0; bx < gx; ++bx) {
            ThreadId id;
            id.tid = 0;
            id.nthreads = nthreads;
            id.bx = bx;
            id.by = by;
            id.gx = gx;
            id.it = 0;
            if constexpr (has_skip<K>::value) {
                if (K::skip(a, id)) continue;
            }
            if constexpr (has_iterations<K>::value) {
                const int nit = K::iterations(a, id);
                for (int it = 0; it < nit; ++it) {
                    id.it = it;
                    emu_phases<K, 0>(st, a, smem.data(), id);
                }
            } else {
                emu_phases<K, 0>(st, a, smem.data(), id);
            }
        }
    return 0;
}

#endif

}  // namespace slmgs
