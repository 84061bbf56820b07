// slmgs_dispatch.h -- per-size launchers (one translation unit per N, see slmgs_inst.cu).
#pragma once
#include "slmgs_teams.h"
#include "slmgs_loop.h"

namespace slmgs {

#ifdef SLMGS_EMULATE
typedef void* rt_stream;
#else
typedef cudaStream_t rt_stream;
#endif

struct LaunchInfo {
    int E;       // points per thread
    int tpl;     // threads per line
    int maxt;    // max threads per block
    int padn;    // padded line length (complex elements)
    int ns;      // number of radix stages
    int r0, r1, r2, r3;  // stage radices of the plan (twiddle table layout); r3 = 1 for three-stage plans
    int p_npre;      // persistent TMA column kernel (ColKernelP): staged boxes per tile, 0 = not built for this size
    int p_ct;        // ... its tile width (columns) at maxt threads
    int p_box_rows;  // ... rows per box
    int teams;       // team kernels (slmgs_teams.h: ColKernelT / RowKernelT) are built for this size
};

#define SLMGS_DECL(N_)                                                                                        \
    int launch_row_##N_(int mode, int gx, int gy, int nthreads, rt_stream s, const RowArgs& a);               \
    int launch_col_##N_(int mode, int var, int gx, int gy, int nthreads, rt_stream s, const ColArgs& a);               \
    int launch_colp_##N_(int var, int dense, int gx, int gy, int nthreads, rt_stream s, const ColArgs& a);             \
    int launch_colt_##N_(int var, int dense, int gx, int gy, rt_stream s, const ColArgs& a, const void* tmap);         \
    int launch_rowt_##N_(int store, int dense, int gx, int gy, rt_stream s, const RowArgs& a);                        \
    int launch_loop_##N_(int li, int gx, int gy, int nthreads, rt_stream s, const LoopArgs& a, int query_blocks_per_sm); \
    LaunchInfo launch_info_##N_();
// 8192-point columns on the team design (ColKernelT8, slmgs_teams.h): mode COL_FUSED / COL_FWD; -1 = not available
int launch_colt8(int mode, int var, int gx, int gy, rt_stream s, const ColArgs& a, const void* tmap);
SLMGS_DECL(16)
SLMGS_DECL(32)
SLMGS_DECL(64)
SLMGS_DECL(128)
SLMGS_DECL(256)
SLMGS_DECL(512)
SLMGS_DECL(1024)
SLMGS_DECL(2048)
SLMGS_DECL(4096)
SLMGS_DECL(8192)
#undef SLMGS_DECL

}  // namespace slmgs
