// slmgs_inst.cu -- instantiates the row / column kernels for one line length.
// Compiled once per size with -DSLMGS_N=<N> (see Makefile) so the sizes build in parallel.
#include "slmgs_dispatch.h"

#ifndef SLMGS_N
#error "compile with -DSLMGS_N=<line length>"
#endif

#define SLMGS_CAT2(a, b) a##b
#define SLMGS_CAT(a, b) SLMGS_CAT2(a, b)

namespace slmgs {

int SLMGS_CAT(launch_row_, SLMGS_N)(int mode, int gx, int gy, int nthreads, rt_stream s, const RowArgs& a) {
    switch (mode) {
        case ROW_FIRST: {
            if (a.colflag) {
                typedef RowKernel<SLMGS_N, ROW_FIRST, false, true> K;
                return launch_kernel<K>(gx, gy, nthreads, K::smem_bytes(nthreads), s, a, a.pdl != 0);
            }
            typedef RowKernel<SLMGS_N, ROW_FIRST> K;
            return launch_kernel<K>(gx, gy, nthreads, K::smem_bytes(nthreads), s, a, a.pdl != 0);
        }
        case ROW_FUSED: {
            if (a.colflag) {  // sparse far field: only the column tiles the column kernel processes are moved
                if (a.store_phase) {
                    typedef RowKernel<SLMGS_N, ROW_FUSED, true, true> K;
                    return launch_kernel<K>(gx, gy, nthreads, K::smem_bytes(nthreads), s, a, a.pdl != 0);
                }
                typedef RowKernel<SLMGS_N, ROW_FUSED, false, true> K;
                return launch_kernel<K>(gx, gy, nthreads, K::smem_bytes(nthreads), s, a, a.pdl != 0);
            }
            if (a.store_phase) {
                typedef RowKernel<SLMGS_N, ROW_FUSED, true> K;
                return launch_kernel<K>(gx, gy, nthreads, K::smem_bytes(nthreads), s, a, a.pdl != 0);
            }
            typedef RowKernel<SLMGS_N, ROW_FUSED, false> K;
            return launch_kernel<K>(gx, gy, nthreads, K::smem_bytes(nthreads), s, a, a.pdl != 0);
        }
        case ROW_LAST: {
            if (a.colflag) {
                typedef RowKernel<SLMGS_N, ROW_LAST, false, true> K;
                return launch_kernel<K>(gx, gy, nthreads, K::smem_bytes(nthreads), s, a, a.pdl != 0);
            }
            typedef RowKernel<SLMGS_N, ROW_LAST> K;
            return launch_kernel<K>(gx, gy, nthreads, K::smem_bytes(nthreads), s, a, a.pdl != 0);
        }
    }
    return -1;
}

// a block of MAXT threads has a compile-time tile width (CT); smaller blocks derive it from blockDim
template <int MODE, int VAR> static int launch_col_ct(int gx, int gy, int nthreads, rt_stream s, const ColArgs& a) {
    typedef Fft<SLMGS_N> F;
    constexpr int MAXT = 16384 / F::E;
    if (nthreads == MAXT) {
        typedef ColKernel<SLMGS_N, MODE, VAR, MAXT / F::TPL> K;
        return launch_kernel<K>(gx, gy, nthreads, K::smem_bytes(nthreads), s, a, a.pdl != 0);
    }
    if constexpr (MAXT / F::TPL >= 2) {
        if (nthreads == MAXT / 2) {  // half-size blocks (two per SM) for zero-padded problems
            typedef ColKernel<SLMGS_N, MODE, VAR, MAXT / F::TPL / 2> K;
            return launch_kernel<K>(gx, gy, nthreads, K::smem_bytes(nthreads), s, a, a.pdl != 0);
        }
    }
    typedef ColKernel<SLMGS_N, MODE, VAR, 0> K;
    return launch_kernel<K>(gx, gy, nthreads, K::smem_bytes(nthreads), s, a, a.pdl != 0);
}
template <int VAR> static int launch_col_fused(int gx, int gy, int nthreads, rt_stream s, const ColArgs& a) {
    return launch_col_ct<COL_FUSED, VAR>(gx, gy, nthreads, s, a);
}

int SLMGS_CAT(launch_col_, SLMGS_N)(int mode, int var, int gx, int gy, int nthreads, rt_stream s, const ColArgs& a) {
    switch (mode) {
        case COL_FWD: return launch_col_ct<COL_FWD, VAR_GENERAL>(gx, gy, nthreads, s, a);
        case COL_FUSED:
            switch (var) {
                case VAR_GS: return launch_col_fused<VAR_GS>(gx, gy, nthreads, s, a);
                case VAR_POW: return launch_col_fused<VAR_POW>(gx, gy, nthreads, s, a);
                case VAR_POW_STORED: return launch_col_fused<VAR_POW_STORED>(gx, gy, nthreads, s, a);
                default: return launch_col_fused<VAR_GENERAL>(gx, gy, nthreads, s, a);
            }
        case COL_INV: {
            typedef ColKernel<SLMGS_N, COL_INV> K;
            return launch_kernel<K>(gx, gy, nthreads, K::smem_bytes(nthreads), s, a, a.pdl != 0);
        }
    }
    return -1;
}

LaunchInfo SLMGS_CAT(launch_info_, SLMGS_N)() {
    typedef Fft<SLMGS_N> F;
    LaunchInfo i;
    i.E = F::E;
    i.tpl = F::TPL;
    i.maxt = 16384 / F::E;
    i.padn = F::PADN;
    i.ns = F::NS;
    i.r0 = F::R0;
    i.r1 = F::R1;
    i.r2 = F::R2;
    return i;
}

}  // namespace slmgs
