// slmgs_inst.cu -- instantiates the row / column kernels for one line length.
// Compiled once per size with -DSLMGS_N=<N> (see Makefile) so the sizes build in parallel.
#include "slmgs_dispatch.h"

#ifndef SLMGS_N
#error "compile with -DSLMGS_N=<line length>"
#endif

#define SLMGS_CAT2(a, b) a##b
#define SLMGS_CAT(a, b) SLMGS_CAT2(a, b)

namespace slmgs {

template <int LI> static int launch_row_li(int mode, int gx, int gy, int nthreads, rt_stream s, const RowArgs& a) {
    switch (mode) {
        case ROW_FIRST: {
            if (a.colflag) {
                typedef RowKernel<SLMGS_N, ROW_FIRST, false, true, LI> K;
                return launch_kernel<K>(gx, gy, nthreads, K::smem_bytes(nthreads), s, a, a.pdl != 0);
            }
            typedef RowKernel<SLMGS_N, ROW_FIRST, false, false, LI> K;
            return launch_kernel<K>(gx, gy, nthreads, K::smem_bytes(nthreads), s, a, a.pdl != 0);
        }
        case ROW_FUSED: {
            if (a.colflag) {  // sparse far field: only the column tiles the column kernel processes are moved
                if (a.store_phase) {
                    typedef RowKernel<SLMGS_N, ROW_FUSED, true, true, LI> K;
                    return launch_kernel<K>(gx, gy, nthreads, K::smem_bytes(nthreads), s, a, a.pdl != 0);
                }
                typedef RowKernel<SLMGS_N, ROW_FUSED, false, true, LI> K;
                return launch_kernel<K>(gx, gy, nthreads, K::smem_bytes(nthreads), s, a, a.pdl != 0);
            }
            if (a.store_phase) {
                typedef RowKernel<SLMGS_N, ROW_FUSED, true, false, LI> K;
                return launch_kernel<K>(gx, gy, nthreads, K::smem_bytes(nthreads), s, a, a.pdl != 0);
            }
            if (a.h == a.H && a.w == a.W && !a.amp) {  // dense field, scalar source amplitude: the hot loop of the metric
                typedef RowKernel<SLMGS_N, ROW_FUSED, false, false, LI, true> K;
                return launch_kernel<K>(gx, gy, nthreads, K::smem_bytes(nthreads), s, a, a.pdl != 0);
            }
            typedef RowKernel<SLMGS_N, ROW_FUSED, false, false, LI> K;
            return launch_kernel<K>(gx, gy, nthreads, K::smem_bytes(nthreads), s, a, a.pdl != 0);
        }
        case ROW_LAST: {
            if (a.colflag) {
                typedef RowKernel<SLMGS_N, ROW_LAST, false, true, LI> K;
                return launch_kernel<K>(gx, gy, nthreads, K::smem_bytes(nthreads), s, a, a.pdl != 0);
            }
            typedef RowKernel<SLMGS_N, ROW_LAST, false, false, LI> K;
            return launch_kernel<K>(gx, gy, nthreads, K::smem_bytes(nthreads), s, a, a.pdl != 0);
        }
    }
    return -1;
}

// the row-pair interleaved field layout needs an even number of lines per block (two lines share every warp)
int SLMGS_CAT(launch_row_, SLMGS_N)(int mode, int gx, int gy, int nthreads, rt_stream s, const RowArgs& a) {
    if (a.pairs) {
        if ((nthreads / Fft<SLMGS_N>::TPL) % 2 != 0) return -1;
        return launch_row_li<2>(mode, gx, gy, nthreads, s, a);
    }
    return launch_row_li<1>(mode, gx, gy, nthreads, s, a);
}

// a block of MAXT threads has a compile-time tile width (CT); smaller blocks derive it from blockDim
template <int MODE, int VAR> static int launch_col_ct(int gx, int gy, int nthreads, rt_stream s, const ColArgs& a) {
    typedef Fft<SLMGS_N> F;
    constexpr int MAXT = 16384 / F::E;
    constexpr bool DENSE_VARIANTS = MODE == COL_FUSED && SLMGS_N >= 1024;  // (dense specialisation of the hot kernels only)
    if (nthreads == MAXT) {
        if constexpr (DENSE_VARIANTS) {
            if (a.h == a.H) {
                typedef ColKernel<SLMGS_N, MODE, VAR, MAXT / F::TPL, true> K;
                return launch_kernel<K>(gx, gy, nthreads, K::smem_bytes(nthreads), s, a, a.pdl != 0);
            }
        }
        typedef ColKernel<SLMGS_N, MODE, VAR, MAXT / F::TPL> K;
        return launch_kernel<K>(gx, gy, nthreads, K::smem_bytes(nthreads), s, a, a.pdl != 0);
    }
    if constexpr (MAXT / F::TPL >= 2) {
        if (nthreads == MAXT / 2) {  // half-size blocks (two per SM)
            if constexpr (DENSE_VARIANTS) {
                if (a.h == a.H) {
                    typedef ColKernel<SLMGS_N, MODE, VAR, MAXT / F::TPL / 2, true> K;
                    return launch_kernel<K>(gx, gy, nthreads, K::smem_bytes(nthreads), s, a, a.pdl != 0);
                }
            }
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

// persistent fused column kernel with TMA-staged tiles: long columns only, full-size blocks
// (superseded by the team kernels below; only built with -DSLMGS_WITH_COLP, for A/B measurements)
#if SLMGS_N >= 2048 && SLMGS_N <= 4096 && defined(SLMGS_WITH_COLP)
#define SLMGS_HAVE_COLP 1
template <int VAR> static int launch_colp_var(int dense, int gx, int gy, int nthreads, rt_stream s, const ColArgs& a) {
    typedef Fft<SLMGS_N> F;
    constexpr int MAXT = 16384 / F::E;
    if (nthreads != MAXT) return -1;
    if (dense) {
        typedef ColKernelP<SLMGS_N, VAR, MAXT / F::TPL, true> K;
        return launch_kernel<K>(gx, gy, nthreads, K::smem_bytes(nthreads), s, a, a.pdl != 0);
    }
    typedef ColKernelP<SLMGS_N, VAR, MAXT / F::TPL, false> K;
    return launch_kernel<K>(gx, gy, nthreads, K::smem_bytes(nthreads), s, a, a.pdl != 0);
}
int SLMGS_CAT(launch_colp_, SLMGS_N)(int var, int dense, int gx, int gy, int nthreads, rt_stream s, const ColArgs& a) {
    switch (var) {
        case VAR_GS: return launch_colp_var<VAR_GS>(dense, gx, gy, nthreads, s, a);
        case VAR_POW: return launch_colp_var<VAR_POW>(dense, gx, gy, nthreads, s, a);
        case VAR_POW_STORED: return launch_colp_var<VAR_POW_STORED>(dense, gx, gy, nthreads, s, a);
        default: return launch_colp_var<VAR_GENERAL>(dense, gx, gy, nthreads, s, a);
    }
}
#else
int SLMGS_CAT(launch_colp_, SLMGS_N)(int, int, int, int, int, rt_stream, const ColArgs&) { return -1; }
#endif

// Team kernels (slmgs_teams.h): TMA-staged tiles, two compute teams per persistent block.  Lines of 2048 / 4096 points
// (three radix stages, 16 points per thread).  `tmap` = host copy of the CUtensorMap over the row-pair interleaved field.
// Under emulation the plain kernels run instead (same arithmetic: the team kernels call the same functions).
#if SLMGS_N >= 2048 && SLMGS_N <= 4096
#define SLMGS_HAVE_TEAMS 1
#ifndef SLMGS_EMULATE
template <int VAR> static int launch_colt_var(int dense, int gx, int gy, rt_stream s, const ColArgs& a, const void* tmap) {
    const CUtensorMap& tm = *reinterpret_cast<const CUtensorMap*>(tmap);
    if (dense) return launch_kernel_teams<ColKernelT<SLMGS_N, VAR, true> >(gx, gy, s, a, tm, a.pdl != 0);
    return launch_kernel_teams<ColKernelT<SLMGS_N, VAR, false> >(gx, gy, s, a, tm, a.pdl != 0);
}
int SLMGS_CAT(launch_colt_, SLMGS_N)(int var, int dense, int gx, int gy, rt_stream s, const ColArgs& a, const void* tmap) {
    switch (var) {
        case VAR_GS: return launch_colt_var<VAR_GS>(dense, gx, gy, s, a, tmap);
        case VAR_POW: return launch_colt_var<VAR_POW>(dense, gx, gy, s, a, tmap);
        case VAR_POW_STORED: return launch_colt_var<VAR_POW_STORED>(dense, gx, gy, s, a, tmap);
        default: return launch_colt_var<VAR_GENERAL>(dense, gx, gy, s, a, tmap);
    }
}
int SLMGS_CAT(launch_rowt_, SLMGS_N)(int store, int dense, int gx, int gy, rt_stream s, const RowArgs& a) {
    if (store) return launch_kernel_teams<RowKernelT<SLMGS_N, true, false> >(gx, gy, s, a, a.pdl != 0);
    if (dense) return launch_kernel_teams<RowKernelT<SLMGS_N, false, true> >(gx, gy, s, a, a.pdl != 0);
    return launch_kernel_teams<RowKernelT<SLMGS_N, false, false> >(gx, gy, s, a, a.pdl != 0);
}
#else
int SLMGS_CAT(launch_colt_, SLMGS_N)(int var, int, int, int gy, rt_stream s, const ColArgs& a, const void*) {
    return SLMGS_CAT(launch_col_, SLMGS_N)(COL_FUSED, var, a.W / 2, gy, 2 * Fft<SLMGS_N>::TPL, s, a);
}
int SLMGS_CAT(launch_rowt_, SLMGS_N)(int, int, int, int gy, rt_stream s, const RowArgs& a) {
    return SLMGS_CAT(launch_row_, SLMGS_N)(ROW_FUSED, (a.h + 1) / 2, gy, 2 * Fft<SLMGS_N>::TPL, s, a);
}
#endif
#else
int SLMGS_CAT(launch_colt_, SLMGS_N)(int, int, int, int, rt_stream, const ColArgs&, const void*) { return -1; }
int SLMGS_CAT(launch_rowt_, SLMGS_N)(int, int, int, int, rt_stream, const RowArgs&) { return -1; }
#endif

#if SLMGS_N == 8192
#ifndef SLMGS_EMULATE
int launch_colt8(int mode, int var, int gx, int gy, rt_stream s, const ColArgs& a, const void* tmap) {
    const CUtensorMap& tm = *reinterpret_cast<const CUtensorMap*>(tmap);
    if (mode == COL_FWD) return launch_kernel_teams<ColKernelT8<COL_FWD, VAR_GENERAL> >(gx, gy, s, a, tm, a.pdl != 0);
    if (mode != COL_FUSED) return -1;
    switch (var) {
        case VAR_GS: return launch_kernel_teams<ColKernelT8<COL_FUSED, VAR_GS> >(gx, gy, s, a, tm, a.pdl != 0);
        case VAR_POW: return launch_kernel_teams<ColKernelT8<COL_FUSED, VAR_POW> >(gx, gy, s, a, tm, a.pdl != 0);
        case VAR_POW_STORED: return launch_kernel_teams<ColKernelT8<COL_FUSED, VAR_POW_STORED> >(gx, gy, s, a, tm, a.pdl != 0);
        default: return launch_kernel_teams<ColKernelT8<COL_FUSED, VAR_GENERAL> >(gx, gy, s, a, tm, a.pdl != 0);
    }
}
#else
int launch_colt8(int, int, int, int, rt_stream, const ColArgs&, const void*) { return -1; }
#endif
#endif

// the whole GS loop of a small square field in one cooperative kernel (slmgs_loop.h): 256 .. 1024 points per line.
// query_blocks_per_sm != 0: return the resident blocks per SM at this geometry instead of launching.
#if SLMGS_N >= 256 && SLMGS_N <= 1024 && !defined(SLMGS_EMULATE)
template <int LI> static int launch_loop_li(int gx, int gy, int nthreads, rt_stream s, const LoopArgs& a, int query) {
    typedef ColKernel<SLMGS_N, COL_FUSED, VAR_GS, 0, false> KC;
    typedef RowKernel<SLMGS_N, ROW_FUSED, true, false, LI, false> KRS;
    size_t smem = KC::smem_bytes(nthreads);
    if (KRS::smem_bytes(nthreads) > smem) smem = KRS::smem_bytes(nthreads);
    if (query) return loop_blocks_per_sm<SLMGS_N, LI>(nthreads, smem);
    return launch_loop<SLMGS_N, LI>(gx, gy, nthreads, smem, s, a);
}
int SLMGS_CAT(launch_loop_, SLMGS_N)(int li, int gx, int gy, int nthreads, rt_stream s, const LoopArgs& a, int query) {
    if (li == 2) return launch_loop_li<2>(gx, gy, nthreads, s, a, query);
    return launch_loop_li<1>(gx, gy, nthreads, s, a, query);
}
#else
int SLMGS_CAT(launch_loop_, SLMGS_N)(int, int, int, int, rt_stream, const LoopArgs&, int query) { return query ? 0 : -1; }
#endif

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
    i.r3 = F::R3;
    i.p_npre = 0;
    i.p_ct = 0;
    i.p_box_rows = 0;
    i.teams = 0;
#ifdef SLMGS_HAVE_TEAMS
    i.teams = 1;
#endif
#ifdef SLMGS_HAVE_COLP
    {
        typedef ColKernelP<SLMGS_N, VAR_GS, (16384 / F::E) / F::TPL, true> K;
        i.p_npre = K::NPRE;
        i.p_ct = (16384 / F::E) / F::TPL;
        i.p_box_rows = K::BOX_ROWS;
    }
#endif
    return i;
}

}  // namespace slmgs

#if defined(SLMGS_TRACE) && !defined(SLMGS_EMULATE)
// diagnostic build only (one size compiled with -DSLMGS_TRACE): switch the per-phase clock stamps on / off, read them
extern "C" __attribute__((visibility("default"))) int slmgs_trace_enable(int on) {
    return (int)cudaMemcpyToSymbol(slmgs::slmgs_trace_on, &on, sizeof on);
}
extern "C" __attribute__((visibility("default"))) int slmgs_trace_read(long long* out) {
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, slmgs::slmgs_trace_buf, sizeof(long long) * 8 * 2 * 64);
}
#endif
