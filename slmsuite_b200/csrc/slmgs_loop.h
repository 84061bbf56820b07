// slmgs_loop.h -- the whole GS loop of a SMALL square field in ONE cooperative kernel.
//
// Fields up to 1024^2 are launch bound: a fused kernel runs a few microseconds, the launch + drain between two kernels
// costs as much (512^2: 11 us per iteration with programmatic dependent launch, of which the kernels are about half;
// replaying the launch sequence as a CUDA graph was measured slower).  Here every block keeps ONE column
// tile and ONE row group for the whole run and the passes are separated by a grid-wide barrier (one atomic counter in
// L2, all blocks co-resident: the kernel is launched cooperatively): n_iter x (column pass, barrier, row pass, barrier)
// with the phases of the plain kernels (same functions, same arithmetic: bit-identical results).  GS only (no weight
// update: every iteration has the same arguments; the last row pass also stores the phase).
#pragma once

#include "slmgs_kernels.h"

namespace slmgs {

struct LoopArgs {
    ColArgs c;
    RowArgs r;        // iterations 0 .. n_iter-2
    RowArgs r_last;   // last iteration (store_phase = 1)
    int n_iter;
    int col_items, row_items;  // blocks with blockIdx.x below these take part in the column / row pass
    unsigned* gbar;            // grid barrier counter (never reset: the host passes the count it starts from)
    unsigned epoch0;
};

#ifndef SLMGS_EMULATE
// Grid barrier: one release-add per block on a counter in L2, then an acquire-load spin.  The __syncthreads() in front makes
// the block's stores happen-before thread 0's release (cumulativity); two full __threadfence() (MEMBAR.GPU) around a relaxed
// atomic cost about a microsecond more per barrier.
SLMGS_DEVICE void loop_grid_sync(unsigned* counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while ((int)(v - target) < 0);
    }
    __syncthreads();
}

template <int N, int LI> __global__ void __launch_bounds__(16384 / Fft<N>::E, 1) slmgs_loop_kernel(const LoopArgs a) {
    typedef ColKernel<N, COL_FUSED, VAR_GS, 0, false> KC;
    typedef RowKernel<N, ROW_FUSED, false, false, LI, false> KR;
    typedef RowKernel<N, ROW_FUSED, true, false, LI, false> KRS;
    extern __shared__ __align__(16) unsigned char slmgs_smem_raw[];
    cf* smem = reinterpret_cast<cf*>(slmgs_smem_raw);
    ThreadId id;
    id.tid = threadIdx.x;
    id.nthreads = blockDim.x;
    id.bx = blockIdx.x;
    id.by = blockIdx.y;
    id.gx = gridDim.x;
    id.it = 0;
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const unsigned nb = gridDim.x * gridDim.y;
    unsigned target = a.epoch0;
    for (int it = 0; it < a.n_iter; ++it) {
        if ((int)blockIdx.x < a.col_items) {
            typename KC::State st;
            run_phases<KC, 0>(st, a.c, smem, id);
        }
        target += nb;
        loop_grid_sync(a.gbar, target);
        if ((int)blockIdx.x < a.row_items) {
            if (it + 1 < a.n_iter) {
                typename KR::State st;
                run_phases<KR, 0>(st, a.r, smem, id);
            } else {
                typename KRS::State st;
                run_phases<KRS, 0>(st, a.r_last, smem, id);
            }
        }
        if (it + 1 < a.n_iter) {
            target += nb;
            loop_grid_sync(a.gbar, target);
        }
    }
}

// returns cudaError_t as int; nthreads threads per block for both passes
template <int N, int LI> int launch_loop(int gx, int gy, int nthreads, size_t smem_bytes, cudaStream_t stream, const LoopArgs& a) {
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(slmgs_loop_kernel<N, LI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_set[dev & 63] = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(gx, gy, 1);
    cfg.blockDim = dim3(nthreads, 1, 1);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, slmgs_loop_kernel<N, LI>, a);
}
// resident blocks of the loop kernel per SM at this block size / shared memory (0 = cannot run)
template <int N, int LI> int loop_blocks_per_sm(int nthreads, size_t smem_bytes) {
    int n = 0;
    cudaFuncSetAttribute(slmgs_loop_kernel<N, LI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, slmgs_loop_kernel<N, LI>, nthreads, smem_bytes) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
#endif

}  // namespace slmgs
