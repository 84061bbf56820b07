// slmgs_api.cu -- C ABI (include/slmgs.h): context, state upload/download, fused and stepped
// GS / WGS loop.  Compiled by nvcc into libslmgs.so; compiled by g++ with -DSLMGS_EMULATE into
// the host-emulation library used by the CPU test-suite (tests/_emu, test infrastructure).
#include "../../include/slmgs.h"
#include "slmgs_dispatch.h"
#include "slmgs_pointwise.h"

#ifndef SLMGS_EMULATE
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint, libcuda is not linked)
#endif
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>

using namespace slmgs;

// ------------------------------------------------------------------------------------------
// runtime shim
// ------------------------------------------------------------------------------------------
#ifdef SLMGS_EMULATE
static int rt_set_device(int) { return 0; }
static int rt_malloc(void** p, size_t n) { *p = calloc(1, n ? n : 1); return *p ? 0 : 2; }
static int rt_free(void* p) { free(p); return 0; }
static int rt_h2d(void* d, const void* h, size_t n, rt_stream) { memcpy(d, h, n); return 0; }
static int rt_d2h(void* h, const void* d, size_t n, rt_stream) { memcpy(h, d, n); return 0; }
static int rt_memset(void* d, int v, size_t n, rt_stream) { memset(d, v, n); return 0; }
static int rt_sync(rt_stream) { return 0; }
static int rt_stream_create(rt_stream* s) { *s = nullptr; return 0; }
static int rt_stream_destroy(rt_stream) { return 0; }
static const char* rt_errstr(int e) { return e == 2 ? "out of memory" : "emulation error"; }
static bool rt_is_oom(int e) { return e == 2; }
static int rt_sm_count() { return 4; }
#else
static int rt_set_device(int d) { return (int)cudaSetDevice(d); }
static int rt_malloc(void** p, size_t n) { return (int)cudaMalloc(p, n ? n : 1); }
static int rt_free(void* p) { return (int)cudaFree(p); }
static int rt_h2d(void* d, const void* h, size_t n, rt_stream s) {
    int e = (int)cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s);
    if (e) return e;
    return (int)cudaStreamSynchronize(s);  // host buffer is only borrowed for the call
}
static int rt_d2h(void* h, const void* d, size_t n, rt_stream s) {
    int e = (int)cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s);
    if (e) return e;
    return (int)cudaStreamSynchronize(s);
}
static int rt_memset(void* d, int v, size_t n, rt_stream s) { return (int)cudaMemsetAsync(d, v, n, s); }
static int rt_sync(rt_stream s) { return (int)cudaStreamSynchronize(s); }
static int rt_stream_create(rt_stream* s) { return (int)cudaStreamCreateWithFlags(s, cudaStreamNonBlocking); }
static int rt_stream_destroy(rt_stream s) { return (int)cudaStreamDestroy(s); }
static const char* rt_errstr(int e) { return cudaGetErrorString((cudaError_t)e); }
static bool rt_is_oom(int e) { return e == (int)cudaErrorMemoryAllocation; }
static int rt_sm_count() {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
}
#endif

// ------------------------------------------------------------------------------------------
// per-size dispatch
// ------------------------------------------------------------------------------------------
#define SLMGS_FOR_SIZES(X) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096) X(8192)

static bool size_supported(int n) {
#define X(N_) if (n == N_) return true;
    SLMGS_FOR_SIZES(X)
#undef X
    return false;
}
static LaunchInfo size_info(int n) {
#define X(N_) if (n == N_) return launch_info_##N_();
    SLMGS_FOR_SIZES(X)
#undef X
    LaunchInfo z;
    memset(&z, 0, sizeof z);
    return z;
}
static int launch_row(int n, int mode, int gx, int gy, int nt, rt_stream s, const RowArgs& a) {
#define X(N_) if (n == N_) return launch_row_##N_(mode, gx, gy, nt, s, a);
    SLMGS_FOR_SIZES(X)
#undef X
    return -1;
}
static int launch_colt(int n, int var, int dense, int gx, int gy, rt_stream s, const ColArgs& a, const void* tmap) {
#define X(N_) if (n == N_) return launch_colt_##N_(var, dense, gx, gy, s, a, tmap);
    SLMGS_FOR_SIZES(X)
#undef X
    return -1;
}
static int launch_rowt(int n, int store, int dense, int gx, int gy, rt_stream s, const RowArgs& a) {
#define X(N_) if (n == N_) return launch_rowt_##N_(store, dense, gx, gy, s, a);
    SLMGS_FOR_SIZES(X)
#undef X
    return -1;
}
static int launch_loop(int n, int li, int gx, int gy, int nt, rt_stream s, const LoopArgs& a, int query) {
#define X(N_) if (n == N_) return launch_loop_##N_(li, gx, gy, nt, s, a, query);
    SLMGS_FOR_SIZES(X)
#undef X
    return query ? 0 : -1;
}
static int launch_colp(int n, int var, int dense, int gx, int gy, int nt, rt_stream s, const ColArgs& a) {
#define X(N_) if (n == N_) return launch_colp_##N_(var, dense, gx, gy, nt, s, a);
    SLMGS_FOR_SIZES(X)
#undef X
    return -1;
}
static int launch_col(int n, int mode, int var, int gx, int gy, int nt, rt_stream s, const ColArgs& a) {
#define X(N_) if (n == N_) return launch_col_##N_(mode, var, gx, gy, nt, s, a);
    SLMGS_FOR_SIZES(X)
#undef X
    return -1;
}

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
enum { ACC_W0 = 0, ACC_W1 = 1, ACC_FNORM = 2, ACC_MEAN = 3, ACC_S0 = 4, ACC_S1 = 5, ACC_S2 = 6, ACC_TMP = 7, ACC_N = 8 };

// a stream may be shared by the child contexts of a MultiplaneHologram: destroyed with its last user
struct StreamRef {
    rt_stream s;
    int refs;
};

struct slmgs_ctx {
    int device, B, H, W, h, w, i0, i2;
    rt_stream stream;
    StreamRef* sref;
    std::string err;
    long long launches;
    int sms;
    // launch geometry
    LaunchInfo irow, icol;
    int row_threads, row_gx, col_threads, col_gx;
    // device state
    cf* fld;
    cf* farfield;     // lazily allocated
    cf* stage_c;      // staging buffer for complex downloads (lazily allocated, B*H*W)
    float* stage_f;   // staging for rolled uploads/downloads (B*H*W floats)
    float *phase, *amp, *prop, *target, *weights, *phase_ff, *amp_ff;
    cf *twA_row, *twB_row, *twA_col, *twB_col;
    cf *twA_half, *twB_half;  // tables of H / 2 points (ColKernelT8), or nullptr
    double* acc;      // [B][ACC_N]
    float* winf;      // [B] 1/sqrt(sum w^2) of the pending normalisation, written by the row kernels of the fused loop
    double* partial;  // stats partials
    int* spot_x;
    int* spot_y;
    float* spot_amp;
    double* spot_pw;
    float* spot_wn;            // [B][N] updated spot weights before the scatter
    unsigned char* spot_keep;  // [N] last spot of every pixel
    int n_spots;
    // host-side state
    float amp_scalar;
    int amp_per_hologram;
    size_t amp_count;  // elements allocated for amp
    int target_shared;
    int w_pending;     // accumulator slot of a not-yet-applied weight normalisation, or -1
    bool ff_valid;     // farfield / amp_ff hold the transform of the current phase (stepped mode)
    double fnorm;      // ||nearfield||_2 == ||farfield||_2 (Parseval, ortho), from the amplitude
    float* phase_saved;
    cf* mp_sum;        // MultiplaneHologram accumulator, lazily allocated
    cf* zero_w;        // MRAF zero-region accumulator image, lazily allocated
    // sparse far field (fused loop): column tiles whose constrained far field can be non-zero
    int sparse_mode;               // 0 = off, 1 = automatic
    bool tiles_dirty;              // weights / target changed since the occupancy flags were computed
    int* tile_flags;               // device [ntiles]: bit 0 weights != 0, bit 1 target is NaN
    int* tile_list;                // device [ntiles]: active tiles in order
    unsigned char* tile_byte;      // device [ntiles]: 1 = active (row kernel filter)
    std::vector<int> tile_flags_h; // host copy of tile_flags
    std::vector<int> spot_x_h;     // host copy of the spot x coordinates (window tiles)
    int tile_key;                  // (mraf, spot width) the device list was built for, -1 = none
    int* tile_count;               // device [B]: active tiles per hologram
    int n_active;                  // most active tiles of any hologram (grid size of the sparse column kernels)
    long long n_active_total;      // active tiles summed over the batch
    bool sparse_now;               // the launches being issued use the tile list
    // grow-only device scratch for small downloads (gray levels, camera images): cudaMalloc / cudaFree per call
    // cost milliseconds once the process holds gigabytes of allocations
    void* scratch;
    size_t scratch_bytes;
    // camera sampling grid (slmgs_set_sample_grid)
    double* samp_y;
    double* samp_x;
    long long n_samp;
    bool last_sparse;              // the last slmgs_run used it
    // timing
    bool profiling;
    bool use_pdl;
    bool prefetch;
    bool pairs;                // fld is stored row-pair interleaved (slmgs_kernels.h, RowArgs)
    // the whole GS loop of a small square field in one cooperative kernel (slmgs_loop.h)
    unsigned* gbar;            // grid barrier counter (device, never reset)
    unsigned gbar_epoch;       // its value once every launch issued so far has finished
    int loop_bps;              // resident blocks per SM of the loop kernel at this geometry (-1 = not asked yet)
    bool populate_shortcut;    // slmgs_run: skip the row pass of _populate_results after a dense run (see run_sequence)
    bool weights_pristine;     // weights == nan_to_num(target) since the last slmgs_reset_weights: the next reset is a no-op
    // CUDA graphs of whole slmgs_run launch sequences (small fields are launch bound): see slmgs_run
    int launch_mode;           // 0 = launch, 1 = dry run: fold every kernel's arguments into `hash` instead
    unsigned long long hash;
    bool use_graphs;
#ifndef SLMGS_EMULATE
    struct GraphEntry {
        unsigned long long key;
        cudaGraphExec_t exec;
        int w_pending_out;
        long long launches;
        unsigned long long last_use;
    };
    std::vector<GraphEntry> graphs;
    std::vector<unsigned long long> graph_seen;   // keys run once eagerly (lazy allocations done): captured next time
    std::vector<unsigned long long> graph_bad;    // keys whose capture failed
    unsigned long long graph_clock;
#endif
    // persistent fused column kernel with TMA-staged tiles (ColKernelP)
    // team kernels (slmgs_teams.h): TMA-staged tiles, two compute teams per persistent block
    bool teams_col, teams_row;  // used for the COL_FUSED / ROW_FUSED launches of the dense-far-field loop
    bool teams_col8;            // dense 8192-point columns: COL_FUSED / COL_FWD as ColKernelT8 (two interleaved 4096-point lines)
    unsigned char tmap_col1[128] __attribute__((aligned(64)));  // ... its CUtensorMap: box {1 column x 2 row parities, 256 row pairs}
    int tb_pairs, tb_n, tb_lo, tb_hi0;  // TMA boxes of a column tile (ColArgs)
    unsigned char tmap_pairs[128] __attribute__((aligned(64)));  // host copy of the CUtensorMap over the row-pair interleaved fld
    bool colp;                 // used for COL_FUSED launches of this context
    bool colp_dense;           // slm rows == padded rows
    void* tmap_dev;            // device copy of the CUtensorMap over fld
    int n_boxes;               // boxes (groups of p_box_rows rows) holding SLM rows
    signed char box_slot[32];  // their staging slots, -1 = box without SLM rows
#ifndef SLMGS_EMULATE
    cudaEvent_t t0, t1;
    std::vector<cudaEvent_t> ev_pool;   // pairs
    std::vector<int> ev_class;
    size_t ev_used;
#endif
};

static std::string g_create_error;

static int fail(slmgs_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    else g_create_error = msg;
    return code;
}
static int rt_check(slmgs_ctx* c, int e, const char* what) {
    if (e == 0) return 0;
    return fail(c, rt_is_oom(e) ? SLMGS_ERR_OOM : SLMGS_ERR_CUDA, std::string(what) + ": " + rt_errstr(e));
}
#define RT(c, call)                                   \
    do {                                              \
        int e__ = rt_check((c), (call), #call);       \
        if (e__) return e__;                          \
    } while (0)
#define CHECK_CTX(c) \
    if (!(c)) return SLMGS_ERR_INVALID; \
    RT(c, rt_set_device((c)->device))

template <class T> static int dev_alloc(slmgs_ctx* c, T** p, size_t count) {
    void* q = nullptr;
    int e = rt_malloc(&q, count * sizeof(T));
    if (e) return rt_check(c, e, "device allocation");
    *p = (T*)q;
    return 0;
}

static int scratch_reserve(slmgs_ctx* c, size_t bytes, void** out) {
    if (bytes > c->scratch_bytes) {
        if (c->scratch) {
            RT(c, rt_sync(c->stream));
            rt_free(c->scratch);
            c->scratch = nullptr;
            c->scratch_bytes = 0;
        }
        int e = rt_check(c, rt_malloc(&c->scratch, bytes), "device allocation");
        if (e) return e;
        c->scratch_bytes = bytes;
    }
    *out = c->scratch;
    return 0;
}

static void make_twiddles(int n, std::vector<cf>& a, std::vector<cf>& b) {
    // twA[k0*M1 + j] = exp(-2 pi i j k0 / N), twB[k1*M2 + r] = exp(-2 pi i r k1 / M1) for the plan of this size
    // (M1 = R1 R2 R3, M2 = R2 R3), and for four-stage plans, behind twB: twC[k2*R3 + n3] = exp(-2 pi i n3 k2 / M2)
    const LaunchInfo li = size_info(n);
    const int r0 = li.r0, r1 = li.r1, r2 = li.r2, r3 = li.r3;
    const int m1 = r1 * r2 * r3, m2 = r2 * r3;
    a.resize(n);
    b.resize(m1 + (r3 > 1 ? m2 : 0));
    const double tau = 6.283185307179586476925286766559;
    for (int k0 = 0; k0 < r0; ++k0)
        for (int j = 0; j < m1; ++j) {
            const double ang = -tau * (double)((long long)j * k0 % n) / (double)n;
            a[k0 * m1 + j] = make_float2((float)cos(ang), (float)sin(ang));
        }
    for (int k1 = 0; k1 < r1; ++k1)
        for (int r = 0; r < m2; ++r) {
            const double ang = -tau * (double)((r * k1) % m1) / (double)m1;
            b[k1 * m2 + r] = make_float2((float)cos(ang), (float)sin(ang));
        }
    if (r3 > 1)
        for (int k2 = 0; k2 < r2; ++k2)
            for (int n3 = 0; n3 < r3; ++n3) {
                const double ang = -tau * (double)((n3 * k2) % m2) / (double)m2;
                b[m1 + k2 * r3 + n3] = make_float2((float)cos(ang), (float)sin(ang));
            }
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

// Persistent TMA column kernel: built for this column length, and every box of rows that holds SLM rows fits its
// staging buffer (dense fields keep the boxes that do not fit as direct loads).  Fills box_slot / n_boxes.
static bool colp_possible(slmgs_ctx* c) {
    const LaunchInfo& li = c->icol;
    if (li.p_npre <= 0 || env_int("SLMGS_TMA", 0) == 0 || c->pairs) return false;
    if (li.p_ct > c->W) return false;
    const int nbox = c->H / li.p_box_rows;
    if (nbox > 32) return false;
    c->colp_dense = c->h == c->H;
    c->n_boxes = 0;
    for (int m = 0; m < nbox; ++m) {
        bool any = false;
        for (int r = 0; r < li.p_box_rows && !any; ++r) {
            const int n = m * li.p_box_rows + r;
            const int sr = ((n + c->H / 2) & (c->H - 1)) - c->i0;
            any = sr >= 0 && sr < c->h;
        }
        c->box_slot[m] = any ? (signed char)c->n_boxes++ : (signed char)-1;
    }
    for (int m = nbox; m < 32; ++m) c->box_slot[m] = -1;
    return c->colp_dense || c->n_boxes <= li.p_npre;
}

static void choose_geometry(slmgs_ctx* c) {
    const int want = 2 * c->sms;  // at least two waves of blocks when the problem allows
    // rows: lines per block L = threads / tpl
    {
        const LaunchInfo& li = c->irow;
        int lo = li.tpl < 32 ? 32 : li.tpl;
        // two resident blocks per SM (<= 512 threads, <= ~70 KB shared memory each) overlap one block's
        // global-memory phases with the other's butterflies: measured 108 vs 124 us at 4096^2 on B200
        int nt = li.maxt > 512 && li.tpl <= 512 ? 512 : li.maxt;
        // (8192-point rows: a line is 64 KB, two of them fill an SM -- kept, one block per SM, because the row-pair
        // interleaved layout needs two lines per block and pays on the column side: configs[4] 18.9 -> 16.6 ms per 10 it)
        const bool pair8192 = c->W == 8192 && c->h >= 2 && env_int("SLMGS_PAIRS8192", 1) != 0;
        while (nt > lo && (size_t)(nt / li.tpl) * li.padn * sizeof(cf) > 113 * 1024 && !(pair8192 && nt / li.tpl == 2)) nt >>= 1;
        while (nt > lo) {
            const int lines = nt / li.tpl;
            const long long blocks = (long long)((c->h + lines - 1) / lines) * c->B;
            if (blocks >= want) break;
            nt >>= 1;
        }
        if (pair8192 && li.maxt / li.tpl >= 2) nt = 2 * li.tpl;  // (also with the four-stage plan: 512 threads per line)
        int o = env_int("SLMGS_ROW_THREADS", 0);
        if (o >= lo && o <= li.maxt && (o & (o - 1)) == 0) nt = o;
        c->row_threads = nt;
        const int lines = nt / li.tpl;
        c->row_gx = (c->h + lines - 1) / lines;
        // row-pair interleaved field layout: two lines per warp in the row kernels, so an even number of lines per block
        // (H >= 32: the column kernel steps through rows b + (H/R0) m and relies on an even H/R0)
        c->pairs = env_int("SLMGS_PAIRS", 1) != 0 && lines % 2 == 0 && c->H >= 32;
    }
    // columns: C = threads / tpl columns per block; keep >= 4 columns (one 32-byte sector per row)
    {
        const LaunchInfo& li = c->icol;
        int lo = li.tpl * 4;
        if (lo < 32) lo = 32;
        if (lo > li.maxt) lo = li.maxt;
        int nt = li.maxt;
        while (nt / li.tpl > c->W) nt >>= 1;
        while (nt > lo) {
            const long long blocks = (long long)(c->W / (nt / li.tpl)) * c->B;
            if (blocks >= want) break;
            nt >>= 1;
        }
        // zero-padded problems (SLM rows <= half the padded rows) move little field data per tile: two
        // half-width blocks per SM overlap better than one full-width block (+7.6 % on the bench workload),
        // while dense problems need the full 32-byte row segments (-8 % with half-width tiles)
        // With the row-pair interleaved field layout a half-width tile still reads whole sectors (2 rows x 2 columns), so
        // the two-blocks-per-SM geometry also wins on dense fields (+4 % on dense 4096^2 GS, B200).
        if (nt == li.maxt && (2 * c->h <= c->H || c->pairs) && li.maxt / li.tpl >= 4 && li.maxt >= 1024 && !colp_possible(c))
            nt = li.maxt / 2;
        int o = env_int("SLMGS_COL_THREADS", 0);
        if (o >= li.tpl && o >= 32 && o <= li.maxt && (o & (o - 1)) == 0 && o / li.tpl <= c->W) nt = o;
        c->col_threads = nt;
        c->col_gx = c->W / (nt / li.tpl);
        c->colp = nt == li.maxt && colp_possible(c);
    }
}

#ifndef SLMGS_EMULATE
typedef CUresult (*slmgs_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
#endif
// tensor map over fld for the persistent column kernel: dims {W, H, B} of 8-byte elements, box {C, box rows, 1}
static int make_tensor_map(slmgs_ctx* c) {
#ifndef SLMGS_EMULATE
    slmgs_encode_fn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres) != cudaSuccess || !encode) {
        cudaGetLastError();
        return 1;
    }
    CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)c->W, (cuuint64_t)c->H, (cuuint64_t)c->B};
    cuuint64_t strides[2] = {(cuuint64_t)c->W * sizeof(cf), (cuuint64_t)c->W * c->H * sizeof(cf)};
    cuuint32_t box[3] = {(cuuint32_t)c->icol.p_ct, (cuuint32_t)c->icol.p_box_rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    if (encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, c->fld, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return 1;
    if (cudaMalloc(&c->tmap_dev, sizeof tm) != cudaSuccess) {
        cudaGetLastError();
        c->tmap_dev = nullptr;
        return 1;
    }
    if (cudaMemcpy(c->tmap_dev, &tm, sizeof tm, cudaMemcpyHostToDevice) != cudaSuccess) return 1;
#else
    (void)c;
#endif
    return 0;
}

// Team column kernel: tensor map over the row-pair interleaved fld seen as {2 W elements, H / 2 row pairs, B} of 8-byte
// elements, box {2 columns x 2 row parities, tb_pairs row pairs, 1}; and the boxes of a column tile that hold SLM rows.
// The rolled rows that hold the SLM are [0, lo) and [hi, H) (lo = i0 + h - H/2, hi = i0 + H/2).
static bool setup_teams_col(slmgs_ctx* c) {
    const int H = c->H, h = c->h, i0 = c->i0;
    if (h == H) {
        c->tb_pairs = 256;
        if (c->tb_pairs > H / 2) c->tb_pairs = H / 2;
        c->tb_n = c->tb_lo = (H / 2) / c->tb_pairs;
        c->tb_hi0 = 0;
    } else {
        const int lo = i0 + h - H / 2, hi = i0 + H / 2;  // rows
        if (lo < 0 || hi > H || (lo & 1) || (hi & 1) || lo > hi) return false;
        // largest box (power of two, <= 256 row pairs) with at most 64 boxes per tile and little over-fetch
        int bp = 256;
        while (bp > 8) {
            const int nlo = (lo / 2 + bp - 1) / bp, nhi = ((H - hi) / 2 + bp - 1) / bp;
            const int fetched = (nlo + nhi) * bp * 2;
            if (fetched * 8 <= h * 9) break;  // <= 12.5 % extra rows
            bp >>= 1;
        }
        const int nlo = (lo / 2 + bp - 1) / bp, nhi = ((H - hi) / 2 + bp - 1) / bp;
        if (nlo + nhi > 64 || nlo + nhi < 1) return false;
        c->tb_pairs = bp;
        c->tb_lo = nlo;
        c->tb_n = nlo + nhi;
        c->tb_hi0 = H / 2 - nhi * bp;                      // the high boxes end at the last row pair
        if (c->tb_hi0 < nlo * bp) return false;            // the two ranges would overlap: not worth a special case
    }
#ifndef SLMGS_EMULATE
    slmgs_encode_fn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres) != cudaSuccess || !encode) {
        cudaGetLastError();
        return false;
    }
    static_assert(sizeof(CUtensorMap) <= sizeof(((slmgs_ctx*)0)->tmap_pairs), "tensor map copy");
    cuuint64_t dims[3] = {(cuuint64_t)c->W * 2, (cuuint64_t)c->H / 2, (cuuint64_t)c->B};
    cuuint64_t strides[2] = {(cuuint64_t)c->W * 2 * sizeof(cf), (cuuint64_t)c->W * c->H * sizeof(cf)};
    cuuint32_t box[3] = {4, (cuuint32_t)c->tb_pairs, 1};
    cuuint32_t es[3] = {1, 1, 1};
    if (encode(reinterpret_cast<CUtensorMap*>(c->tmap_pairs), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, c->fld, dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
#endif
    return true;
}

// ColKernelT8 (dense 8192-point columns as two interleaved 4096-point lines): twiddle tables of the half length and a
// tensor map over the row-pair interleaved fld with box {1 column x 2 row parities (16 bytes), 256 row pairs, 1}
static bool setup_teams_col8(slmgs_ctx* c) {
#ifndef SLMGS_EMULATE
    std::vector<cf> a, b;
    make_twiddles(c->H / 2, a, b);
    if (dev_alloc(c, &c->twA_half, a.size()) || dev_alloc(c, &c->twB_half, b.size())) return false;
    if (rt_h2d(c->twA_half, a.data(), a.size() * sizeof(cf), c->stream) || rt_h2d(c->twB_half, b.data(), b.size() * sizeof(cf), c->stream))
        return false;
    if (rt_sync(c->stream)) return false;  // (a, b are about to go out of scope)
    slmgs_encode_fn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres) != cudaSuccess || !encode) {
        cudaGetLastError();
        return false;
    }
    cuuint64_t dims[3] = {(cuuint64_t)c->W * 2, (cuuint64_t)c->H / 2, (cuuint64_t)c->B};
    cuuint64_t strides[2] = {(cuuint64_t)c->W * 2 * sizeof(cf), (cuuint64_t)c->W * c->H * sizeof(cf)};
    cuuint32_t box[3] = {2, 256, 1};
    cuuint32_t es[3] = {1, 1, 1};
    return encode(reinterpret_cast<CUtensorMap*>(c->tmap_col1), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, c->fld, dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
#else
    (void)c;
    return false;
#endif
}

extern "C" int slmgs_version(void) { return 100; }


extern "C" const char* slmgs_last_error(const slmgs_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

extern "C" int slmgs_create(slmgs_ctx** out, int device, int batch, int H, int W, int h, int w) {
    if (!out) return fail(nullptr, SLMGS_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (batch < 1) return fail(nullptr, SLMGS_ERR_INVALID, "batch must be >= 1");
    if (!size_supported(H) || !size_supported(W))
        return fail(nullptr, SLMGS_ERR_INVALID, "shape must be powers of two in [16, 8192] per dimension");
    if (h < 1 || w < 1 || h > H || w > W) return fail(nullptr, SLMGS_ERR_INVALID, "slm_shape must fit inside shape");
    int e = rt_set_device(device);
    if (e) return fail(nullptr, SLMGS_ERR_CUDA, std::string("cudaSetDevice: ") + rt_errstr(e));
    slmgs_ctx* c = new slmgs_ctx();
    c->device = device;
    c->B = batch; c->H = H; c->W = W; c->h = h; c->w = w;
    // toolbox.unpad, toolbox/__init__.py:1701-1709: floor of half the difference
    c->i0 = (H - h) / 2;
    c->i2 = (W - w) / 2;
    c->launches = 0;
    c->sms = rt_sm_count();
    c->irow = size_info(W);
    c->icol = size_info(H);
    c->fld = nullptr; c->farfield = nullptr; c->stage_c = nullptr; c->stage_f = nullptr;
    c->phase = c->amp = c->prop = c->target = c->weights = c->phase_ff = c->amp_ff = nullptr;
    c->twA_row = c->twB_row = c->twA_col = c->twB_col = nullptr;
    c->twA_half = c->twB_half = nullptr;
    c->acc = nullptr; c->partial = nullptr; c->winf = nullptr;
    c->spot_x = c->spot_y = nullptr; c->spot_amp = nullptr; c->spot_pw = nullptr; c->n_spots = 0;
    c->spot_wn = nullptr; c->spot_keep = nullptr;
    c->amp_scalar = (float)(1.0 / sqrt((double)h * (double)w));
    c->amp_per_hologram = 0;
    c->amp_count = 0;
    c->target_shared = 0;
    c->w_pending = -1;
    c->ff_valid = false;
    c->fnorm = 1.0;
    c->stream = nullptr;
    c->phase_saved = nullptr;
    c->mp_sum = nullptr;
    c->zero_w = nullptr;
    c->sparse_mode = env_int("SLMGS_SPARSE", 1) != 0 ? 1 : 0;
    c->tiles_dirty = true;
    c->tile_flags = nullptr; c->tile_list = nullptr; c->tile_byte = nullptr; c->tile_count = nullptr;
    c->n_active_total = 0;
    c->tile_key = -1;
    c->n_active = 0;
    c->sparse_now = false;
    c->samp_y = c->samp_x = nullptr;
    c->n_samp = 0;
    c->scratch = nullptr;
    c->scratch_bytes = 0;
    c->last_sparse = false;
    c->sref = nullptr;
    c->profiling = false;
    c->launch_mode = 0;
    c->hash = 0;
    c->use_graphs = env_int("SLMGS_GRAPHS", 0) != 0;  // opt-in: measured SLOWER than the PDL stream launches on B200
#ifndef SLMGS_EMULATE
    c->graph_clock = 0;
#endif
    c->teams_col = c->teams_row = false;
    c->teams_col8 = false;
    c->weights_pristine = false;
    c->gbar = nullptr;
    c->gbar_epoch = 0;
    c->loop_bps = -1;
    c->populate_shortcut = env_int("SLMGS_POPULATE_REBUILD", 0) == 0;
    c->tb_pairs = c->tb_n = c->tb_lo = c->tb_hi0 = 0;
    c->colp = false;
    c->colp_dense = false;
    c->tmap_dev = nullptr;
    c->n_boxes = 0;
    c->use_pdl = env_int("SLMGS_PDL", 1) != 0;
    // L2 prefetch of the next tile's field rows: +2.4 % on dense 4096^2, -0.4 % on zero-padded problems
    c->prefetch = env_int("SLMGS_PREFETCH", (h == H) ? 1 : 0) != 0;
#ifndef SLMGS_EMULATE
    c->t0 = c->t1 = nullptr;
    c->ev_used = 0;
#endif
    choose_geometry(c);
#define CR(call)                                                   \
    do {                                                           \
        int e__ = (call);                                          \
        if (e__) {                                                 \
            g_create_error = c->err;                               \
            slmgs_destroy(c);                                      \
            return e__;                                            \
        }                                                          \
    } while (0)
    CR(rt_check(c, rt_stream_create(&c->stream), "stream create"));
    c->sref = new StreamRef{c->stream, 1};
    const size_t P = (size_t)H * W, S = (size_t)h * w, Bz = (size_t)batch;
    CR(dev_alloc(c, &c->fld, Bz * P));
    CR(dev_alloc(c, &c->phase, Bz * S));
    CR(dev_alloc(c, &c->target, Bz * P));
    CR(dev_alloc(c, &c->weights, Bz * P));
    CR(dev_alloc(c, &c->phase_ff, Bz * P));
    CR(dev_alloc(c, &c->amp_ff, Bz * P));
    CR(dev_alloc(c, &c->stage_f, Bz * P));
    CR(dev_alloc(c, &c->acc, Bz * ACC_N));
    CR(dev_alloc(c, &c->winf, Bz));
    CR(dev_alloc(c, &c->partial, Bz * 256 * 8));
    CR(rt_check(c, rt_memset(c->fld, 0, Bz * P * sizeof(cf), c->stream), "memset"));
    CR(rt_check(c, rt_memset(c->phase, 0, Bz * S * sizeof(float), c->stream), "memset"));
    CR(rt_check(c, rt_memset(c->target, 0, Bz * P * sizeof(float), c->stream), "memset"));
    CR(rt_check(c, rt_memset(c->weights, 0, Bz * P * sizeof(float), c->stream), "memset"));
    CR(rt_check(c, rt_memset(c->phase_ff, 0, Bz * P * sizeof(float), c->stream), "memset"));
    CR(rt_check(c, rt_memset(c->amp_ff, 0, Bz * P * sizeof(float), c->stream), "memset"));
    CR(rt_check(c, rt_memset(c->acc, 0, Bz * ACC_N * sizeof(double), c->stream), "memset"));
    {
        std::vector<cf> a, b;
        make_twiddles(W, a, b);
        CR(dev_alloc(c, &c->twA_row, a.size()));
        CR(dev_alloc(c, &c->twB_row, b.size()));
        CR(rt_check(c, rt_h2d(c->twA_row, a.data(), a.size() * sizeof(cf), c->stream), "twiddle upload"));
        CR(rt_check(c, rt_h2d(c->twB_row, b.data(), b.size() * sizeof(cf), c->stream), "twiddle upload"));
        make_twiddles(H, a, b);
        CR(dev_alloc(c, &c->twA_col, a.size()));
        CR(dev_alloc(c, &c->twB_col, b.size()));
        CR(rt_check(c, rt_h2d(c->twA_col, a.data(), a.size() * sizeof(cf), c->stream), "twiddle upload"));
        CR(rt_check(c, rt_h2d(c->twB_col, b.data(), b.size() * sizeof(cf), c->stream), "twiddle upload"));
    }
    if (c->colp && make_tensor_map(c)) c->colp = false;  // no TMA descriptor: the plain column kernel takes over
    // team kernels: row-pair interleaved field, 2-column tiles (the image layout of the context), whole row pairs
    {
        const bool on = env_int("SLMGS_TEAMS", 1) != 0 && c->pairs && !c->colp;
        c->teams_col = on && c->icol.teams && c->col_threads == 2 * c->icol.tpl && (c->h % 2) == 0 && (c->i0 % 2) == 0 &&
                       setup_teams_col(c);
        c->teams_row = on && c->irow.teams && (c->h % 2) == 0 && (c->i0 % 2) == 0 && c->H >= 32;
        // a persistent block pays for its prologue (twiddle rows into shared memory, the first tile's latency): only worth it
        // with a few work items per team (configs[1] has 576 row pairs for 296 teams: the plain row kernel stays, 31 vs 33 us)
        const long long teams_gpu = 2LL * c->sms;
        if ((long long)(c->h / 2) * c->B < 4 * teams_gpu) c->teams_row = false;
        if ((long long)(c->W / 2) * c->B < 4 * teams_gpu) c->teams_col = false;
        // (opt-in: measured SLOWER than the plain 8192 kernels on B200 -- 16-byte TMA rows and half-used sectors)
        c->teams_col8 = on && env_int("SLMGS_TEAMS8", 0) != 0 && c->H == 8192 && c->h == c->H && c->col_threads == 2 * c->icol.tpl &&
                        (long long)c->W * c->B >= 4 * teams_gpu && setup_teams_col8(c);
    }
    CR(rt_check(c, rt_sync(c->stream), "sync"));
#undef CR
    *out = c;
    return SLMGS_OK;
}

extern "C" int slmgs_destroy(slmgs_ctx* c) {
    if (!c) return SLMGS_OK;
    rt_set_device(c->device);
    if (c->stream) rt_sync(c->stream);
    void* ptrs[] = {c->fld, c->farfield, c->stage_c, c->stage_f, c->phase, c->amp, c->prop, c->target,
                    c->weights, c->phase_ff, c->amp_ff, c->twA_row, c->twB_row, c->twA_col, c->twB_col, c->twA_half, c->twB_half, c->acc,
                    c->partial, c->spot_x, c->spot_y, c->spot_amp, c->spot_pw, c->spot_wn, c->spot_keep, c->phase_saved, c->mp_sum, c->zero_w,
                    c->tile_flags, c->tile_list, c->tile_byte, c->tile_count, c->samp_y, c->samp_x, c->scratch, c->winf,
                    c->tmap_dev, c->gbar};
    for (void* p : ptrs)
        if (p) rt_free(p);
#ifndef SLMGS_EMULATE
    for (auto& g : c->graphs) cudaGraphExecDestroy(g.exec);
    c->graphs.clear();
    if (c->t0) cudaEventDestroy(c->t0);
    if (c->t1) cudaEventDestroy(c->t1);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
#endif
    if (c->sref) {
        if (--c->sref->refs == 0) {
            rt_stream_destroy(c->sref->s);
            delete c->sref;
        }
    } else if (c->stream) {
        rt_stream_destroy(c->stream);
    }
    delete c;
    return SLMGS_OK;
}

extern "C" int slmgs_device_pci_bus_id(int device, char* out, int len) {
    if (!out || len < 13) return SLMGS_ERR_INVALID;
    out[0] = 0;
#ifndef SLMGS_EMULATE
    if (cudaDeviceGetPCIBusId(out, len, device) != cudaSuccess) {
        cudaGetLastError();
        out[0] = 0;
        return SLMGS_ERR_CUDA;
    }
#endif
    return SLMGS_OK;
}

extern "C" int slmgs_sync(slmgs_ctx* c) {
    CHECK_CTX(c);
    RT(c, rt_sync(c->stream));
    return SLMGS_OK;
}
extern "C" void* slmgs_phase_device_ptr(slmgs_ctx* c) { return c ? (void*)c->phase : nullptr; }
extern "C" void* slmgs_stream(slmgs_ctx* c) { return c ? (void*)c->stream : nullptr; }
extern "C" long long slmgs_launch_count(const slmgs_ctx* c) { return c ? c->launches : 0; }
extern "C" int slmgs_launch_geometry(const slmgs_ctx* c, int* out4) {
    if (!c || !out4) return SLMGS_ERR_INVALID;
    out4[0] = c->row_threads; out4[1] = c->row_gx; out4[2] = c->col_threads; out4[3] = c->col_gx;
    return SLMGS_OK;
}

// ------------------------------------------------------------------------------------------
// element-wise helpers
// ------------------------------------------------------------------------------------------
static ElemArgs elem_args(slmgs_ctx* c, const void* src, void* dst, long long n) {
    ElemArgs a;
    memset(&a, 0, sizeof a);
    a.src = src; a.dst = dst; a.n = n;
    a.src_bs = n; a.dst_bs = n; a.target_bs = c->target_shared ? 0 : n;
    a.target = c->target;
    a.acc = c->acc; a.acc_bs = ACC_N;
    a.H = c->H; a.W = c->W;
    a.C = c->col_threads / c->icol.tpl;
    a.unroll = 0;
    a.fnorm_slot = -1; a.mean_slot = -1;
    return a;
}
template <int OP> static int launch_elem(slmgs_ctx* c, const ElemArgs& a, int batch) {
    long long blocks = (a.n + 255) / 256;
    const long long cap = (long long)c->sms * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    c->launches++;
    if (a.dst == (void*)c->weights) c->weights_pristine = false;
    return rt_check(c, launch_kernel<ElemKernel<OP>>((int)blocks, batch, 256, 0, c->stream, a), "element-wise launch");
}
static int zero_slot(slmgs_ctx* c, int slot, int count = 1) {
    // slots of all holograms: strided -> one small memset per hologram would be B launches; clear via 2-D memset
#ifdef SLMGS_EMULATE
    for (int b = 0; b < c->B; ++b) memset(c->acc + (size_t)b * ACC_N + slot, 0, sizeof(double) * count);
    return 0;
#else
    return rt_check(c, (int)cudaMemset2DAsync(c->acc + slot, ACC_N * sizeof(double), 0, sizeof(double) * count, c->B,
                                               c->stream), "accumulator clear");
#endif
}

// upload a centred image -> rolled device image
static int upload_rolled(slmgs_ctx* c, const float* host, float* dst, int batch) {
    const long long P = (long long)c->H * c->W;
    RT(c, rt_h2d(c->stage_f, host, (size_t)batch * P * sizeof(float), c->stream));
    ElemArgs a = elem_args(c, c->stage_f, dst, P);
    return launch_elem<EW_ROLL_F32>(c, a, batch);
}
static int download_rolled(slmgs_ctx* c, const float* src, float* host, int batch) {
    const long long P = (long long)c->H * c->W;
    ElemArgs a = elem_args(c, src, c->stage_f, P);
    a.unroll = 1;
    int e = launch_elem<EW_ROLL_F32>(c, a, batch);  // roll by N/2 is an involution for even N
    if (e) return e;
    RT(c, rt_d2h(host, c->stage_f, (size_t)batch * P * sizeof(float), c->stream));
    return SLMGS_OK;
}

// resolve a pending weight normalisation (fused WGS leaves weights un-normalised by one scalar)
static int resolve_weights(slmgs_ctx* c) {
    if (c->w_pending < 0) return SLMGS_OK;
    ElemArgs a = elem_args(c, nullptr, c->weights, (long long)c->H * c->W);
    a.slot0 = c->w_pending;
    int e = launch_elem<EW_SCALE>(c, a, c->B);
    c->w_pending = -1;
    return e;
}

// ------------------------------------------------------------------------------------------
// state upload / download
// ------------------------------------------------------------------------------------------
extern "C" int slmgs_set_phase(slmgs_ctx* c, const float* phase) {
    CHECK_CTX(c);
    if (!phase) return fail(c, SLMGS_ERR_INVALID, "phase is NULL");
    RT(c, rt_h2d(c->phase, phase, (size_t)c->B * c->h * c->w * sizeof(float), c->stream));
    c->ff_valid = false;
    return SLMGS_OK;
}
extern "C" int slmgs_get_phase(slmgs_ctx* c, float* phase) {
    CHECK_CTX(c);
    if (!phase) return fail(c, SLMGS_ERR_INVALID, "phase is NULL");
    RT(c, rt_d2h(phase, c->phase, (size_t)c->B * c->h * c->w * sizeof(float), c->stream));
    return SLMGS_OK;
}
extern "C" int slmgs_set_amp_scalar(slmgs_ctx* c, float amp) {
    CHECK_CTX(c);
    if (c->amp) {
        RT(c, rt_sync(c->stream));
        rt_free(c->amp);
        c->amp = nullptr;
    }
    c->amp_scalar = amp;
    c->fnorm = (double)amp * sqrt((double)c->h * (double)c->w);
    c->ff_valid = false;
    return SLMGS_OK;
}
extern "C" int slmgs_set_amp_array(slmgs_ctx* c, const float* amp, int per_hologram) {
    CHECK_CTX(c);
    if (!amp) return fail(c, SLMGS_ERR_INVALID, "amp is NULL");
    const size_t S = (size_t)c->h * c->w, n = per_hologram ? S * c->B : S;
    if (c->amp && c->amp_count != n) {
        RT(c, rt_sync(c->stream));
        rt_free(c->amp);
        c->amp = nullptr;
    }
    if (!c->amp) {
        int e = dev_alloc(c, &c->amp, n);
        if (e) return e;
        c->amp_count = n;
    }
    RT(c, rt_h2d(c->amp, amp, n * sizeof(float), c->stream));  // stream-ordered after the kernels that read the old one
    c->amp_per_hologram = per_hologram ? 1 : 0;
    // Parseval: ||farfield|| = ||nearfield|| = ||amp|| (first hologram's amp is representative: the
    // constructor L2-normalises every amp, _hologram.py:404-405)
    double s = 0.0;
    for (size_t i = 0; i < S; ++i) s += (double)amp[i] * (double)amp[i];
    c->fnorm = sqrt(s);
    c->ff_valid = false;
    return SLMGS_OK;
}
extern "C" int slmgs_set_propagation(slmgs_ctx* c, const float* kernel) {
    CHECK_CTX(c);
    if (c->prop) {
        RT(c, rt_sync(c->stream));
        rt_free(c->prop);
        c->prop = nullptr;
    }
    if (kernel) {
        const size_t S = (size_t)c->h * c->w;
        int e = dev_alloc(c, &c->prop, S);
        if (e) return e;
        RT(c, rt_h2d(c->prop, kernel, S * sizeof(float), c->stream));
    }
    c->ff_valid = false;
    return SLMGS_OK;
}
extern "C" int slmgs_set_target(slmgs_ctx* c, const float* target, int shared) {
    CHECK_CTX(c);
    if (!target) return fail(c, SLMGS_ERR_INVALID, "target is NULL");
    c->target_shared = shared ? 1 : 0;
    c->tiles_dirty = true;
    c->weights_pristine = false;
    return upload_rolled(c, target, c->target, shared ? 1 : c->B);
}
extern "C" int slmgs_get_target(slmgs_ctx* c, float* target) {
    CHECK_CTX(c);
    if (!target) return fail(c, SLMGS_ERR_INVALID, "target is NULL");
    return download_rolled(c, c->target, target, c->target_shared ? 1 : c->B);
}
extern "C" int slmgs_reset_weights(slmgs_ctx* c) {
    CHECK_CTX(c);
    const long long P = (long long)c->H * c->W;
    // weights untouched since the last reset (GS never updates them): nothing to copy, the tile flags are current
    if (c->weights_pristine && c->w_pending < 0 && !c->zero_w && !c->tiles_dirty) return SLMGS_OK;
    ElemArgs a = elem_args(c, c->target, c->weights, P);
    a.src_bs = c->target_shared ? 0 : P;
    c->w_pending = -1;
    if (!c->tiles_dirty && !c->tile_flags_h.empty()) {
        // weights := nan_to_num(target): their occupancy is the target's (flag bit 2), already known -- no device pass
        // (the device lists are only rebuilt when that changes a flag: a reset between two runs on the same target
        // keeps them, so the next slmgs_run issues no copy and never synchronises)
        bool changed = false;
        for (int& f : c->tile_flags_h) {
            const int g = (f & ~1) | ((f & 4) ? 1 : 0);
            changed = changed || g != f;
            f = g;
        }
        if (changed) c->tile_key = -1;
    } else {
        c->tiles_dirty = true;
    }
    if (c->zero_w) RT(c, rt_memset(c->zero_w, 0, (size_t)c->B * P * sizeof(cf), c->stream));  // zero_weights *= 0, :609-610
    const int e = launch_elem<EW_FILL_NAN0>(c, a, c->B);
    c->weights_pristine = e == SLMGS_OK;
    return e;
}
extern "C" int slmgs_set_weights(slmgs_ctx* c, const float* weights) {
    CHECK_CTX(c);
    if (!weights) return fail(c, SLMGS_ERR_INVALID, "weights is NULL");
    c->w_pending = -1;
    c->tiles_dirty = true;
    c->weights_pristine = false;
    return upload_rolled(c, weights, c->weights, c->B);
}
extern "C" int slmgs_get_weights(slmgs_ctx* c, float* weights) {
    CHECK_CTX(c);
    if (!weights) return fail(c, SLMGS_ERR_INVALID, "weights is NULL");
    int e = resolve_weights(c);
    if (e) return e;
    return download_rolled(c, c->weights, weights, c->B);
}
extern "C" int slmgs_set_phase_ff(slmgs_ctx* c, const float* p) {
    CHECK_CTX(c);
    if (!p) return fail(c, SLMGS_ERR_INVALID, "phase_ff is NULL");
    return upload_rolled(c, p, c->phase_ff, c->B);
}
extern "C" int slmgs_get_phase_ff(slmgs_ctx* c, float* p) {
    CHECK_CTX(c);
    if (!p) return fail(c, SLMGS_ERR_INVALID, "phase_ff is NULL");
    return download_rolled(c, c->phase_ff, p, c->B);
}
extern "C" int slmgs_get_amp_ff(slmgs_ctx* c, float* p) {
    CHECK_CTX(c);
    if (!p) return fail(c, SLMGS_ERR_INVALID, "amp_ff is NULL");
    return download_rolled(c, c->amp_ff, p, c->B);
}

extern "C" int slmgs_get_phase_gray(slmgs_ctx* c, int bitdepth, const double* correction, void* out) {
    CHECK_CTX(c);
    if (!out) return fail(c, SLMGS_ERR_INVALID, "out is NULL");
    if (bitdepth < 1 || bitdepth > 16) return fail(c, SLMGS_ERR_INVALID, "bitdepth must be in [1, 16]");
    const long long S = (long long)c->h * c->w;
    const int out16 = bitdepth > 8;
    const size_t out_bytes = ((size_t)c->B * S * (out16 ? 2 : 1) + 7) & ~(size_t)7;
    int e;
    void* buf = nullptr;
    if ((e = scratch_reserve(c, out_bytes + (correction ? (size_t)S * sizeof(double) : 0), &buf))) return e;
    void* dout = buf;
    double* dcorr = correction ? reinterpret_cast<double*>(reinterpret_cast<char*>(buf) + out_bytes) : nullptr;
    if (correction) RT(c, rt_h2d(dcorr, correction, (size_t)S * sizeof(double), c->stream));
    ElemArgs a = elem_args(c, c->phase, dout, S);
    a.corr = dcorr;
    a.bitres = 1 << bitdepth;
    a.factor = -((double)a.bitres / 2.0 / 3.14159265358979323846);
    a.out16 = out16;
    if ((e = launch_elem<EW_PHASE2GRAY>(c, a, c->B))) return e;
    RT(c, rt_d2h(out, dout, (size_t)c->B * S * (out16 ? 2 : 1), c->stream));
    RT(c, rt_sync(c->stream));
    return SLMGS_OK;
}

// ------------------------------------------------------------------------------------------
// kernel argument builders
// ------------------------------------------------------------------------------------------
static RowArgs row_args(slmgs_ctx* c) {
    RowArgs a;
    memset(&a, 0, sizeof a);
    a.fld = c->fld;
    a.fld_bs = (long long)c->H * c->W;
    a.phase = c->phase;
    a.phase_bs = (long long)c->h * c->w;
    a.amp = c->amp;
    a.amp_bs = c->amp_per_hologram ? (long long)c->h * c->w : 0;
    a.prop = c->prop;
    a.twA = c->twA_row;
    a.twB = c->twB_row;
    a.amp_scalar = c->amp_scalar;
    a.scale = (float)(1.0 / sqrt((double)c->H * (double)c->W));
    a.H = c->H; a.W = c->W; a.h = c->h; a.w = c->w; a.i0 = c->i0; a.i2 = c->i2;
    a.store_phase = 0;
    a.mp_sum = nullptr;
    a.mp_weight = 0.f;
    a.mp_first = 0;
    a.zero_acc = nullptr;
    a.zero_bs = ACC_N;
    a.pdl = (c->use_pdl && !c->profiling) ? 1 : 0;
    a.pf_dist = c->prefetch ? c->sms * (c->row_threads <= 512 && c->irow.E <= 16 ? 2 : 1) : 0;  // (= blocks resident on the GPU)
    a.pairs = c->pairs ? 1 : 0;
    if (c->sparse_now) {
        a.colflag = c->tile_byte;
        a.colflag_bs = c->W / (c->col_threads / c->icol.tpl);
        a.ctile_shift = 0;
        while ((1 << a.ctile_shift) < c->col_threads / c->icol.tpl) ++a.ctile_shift;
        a.pf_dist = 0;
    }
    return a;
}
static ColArgs col_args(slmgs_ctx* c) {
    ColArgs a;
    memset(&a, 0, sizeof a);
    a.fld = c->fld;
    a.fld_bs = (long long)c->H * c->W;
    a.twA = c->twA_col;
    a.twB = c->twB_col;
    a.tw2A = c->twA_half;
    a.tw2B = c->twB_half;
    a.weights = c->weights;
    a.target = c->target;
    a.phase_ff = c->phase_ff;
    a.amp_ff = c->amp_ff;
    a.farfield = c->farfield;
    a.img_bs = (long long)c->H * c->W;
    a.target_bs = c->target_shared ? 0 : a.img_bs;
    a.acc = c->acc;
    a.acc_bs = ACC_N;
    a.w_in_slot = -1;
    a.w_out_slot = -1;
    a.ratio_slot = -1;
    a.wsq_slot = -1;
    a.ratio_extra = 0.0;
    a.inv_npix = 1.0 / ((double)c->H * (double)c->W);
    a.H = c->H; a.W = c->W; a.h = c->h; a.i0 = c->i0;
    a.scale = (float)(1.0 / sqrt((double)c->H * (double)c->W));
    a.wgs.method = METHOD_GS;
    a.wgs.p = 0.f; a.wgs.f = 0.f;
    a.wgs.inv_fnorm = (float)(1.0 / c->fnorm);
    a.wgs.neg_inv_mean = -1.0f;
    a.zero_w = c->zero_w;
    a.zero_factor = 1.0f;
    a.pdl = (c->use_pdl && !c->profiling) ? 1 : 0;
    a.pf_dist = c->prefetch ? c->sms * (c->col_threads <= 512 && c->icol.E <= 16 ? 2 : 1) : 0;  // (= blocks resident on the GPU)
    if (c->sparse_now) {
        a.tiles = c->tile_list;
        a.tile_count = c->tile_count;
        a.tiles_bs = c->W / (c->col_threads / c->icol.tpl);
    }
    a.pairs = c->pairs ? 1 : 0;
    a.tb_pairs = c->tb_pairs; a.tb_n = c->tb_n; a.tb_lo = c->tb_lo; a.tb_hi0 = c->tb_hi0;
    a.tmap = c->tmap_dev;
    a.n_boxes = c->n_boxes;
    memcpy(a.box_slot, c->box_slot, sizeof a.box_slot);
    return a;
}
// profiling: bracket a launch with events from a pool (class k: 0..2 row modes, 3..5 column modes)
static void prof_mark(slmgs_ctx* c, int klass, bool begin) {
#ifndef SLMGS_EMULATE
    if (!c->profiling) return;
    if (c->ev_used == c->ev_pool.size()) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        c->ev_pool.push_back(e);
    }
    if (begin) c->ev_class.push_back(klass);
    cudaEventRecord(c->ev_pool[c->ev_used++], c->stream);
#else
    (void)c; (void)klass; (void)begin;
#endif
}
static void hash_mix(slmgs_ctx* c, const void* data, size_t n) {  // FNV-1a
    const unsigned char* p = (const unsigned char*)data;
    unsigned long long h = c->hash;
    for (size_t i = 0; i < n; ++i) h = (h ^ p[i]) * 1099511628211ull;
    c->hash = h;
}
// persistent blocks per hologram of a team kernel: the resident blocks of the GPU (a block = two teams of 2 tpl
// threads) shared by the batch, at most one block per two work items
static int teams_blocks(slmgs_ctx* c, const LaunchInfo& li, int items) {
    const int per_sm = 1024 / (4 * li.tpl) > 0 ? 1024 / (4 * li.tpl) : 1;
    int gx = c->sms * per_sm / c->B;
    if (gx < 1) gx = 1;
    if (gx > (items + 1) / 2) gx = (items + 1) / 2;
    return gx;
}
static int run_row(slmgs_ctx* c, int mode, const RowArgs& a) {
    c->launches++;
    if (c->launch_mode == 1) {
        const int g[4] = {mode, c->row_gx, c->B, c->row_threads};
        hash_mix(c, g, sizeof g);
        hash_mix(c, &a, sizeof a);
        return 0;
    }
    prof_mark(c, mode, true);
    int e;
    if (mode == ROW_FUSED && c->teams_row && !a.colflag) {
        // persistent: one block (two teams) per SM over the whole batch, each team walking over row pairs
        int gx = teams_blocks(c, c->irow, (c->h + 1) / 2);
        const int dense = (a.h == a.H && a.w == a.W && !a.amp) ? 1 : 0;
        e = rt_check(c, launch_rowt(c->W, a.store_phase, dense, gx, c->B, c->stream, a), "team row kernel launch");
    } else {
        e = rt_check(c, launch_row(c->W, mode, c->row_gx, c->B, c->row_threads, c->stream, a), "row kernel launch");
    }
    prof_mark(c, mode, false);
    return e;
}
static int run_col(slmgs_ctx* c, int mode, const ColArgs& a) {
    c->launches++;
    if (a.wgs_update || a.zero_w) c->weights_pristine = false;
    if (c->launch_mode == 1) {
        const int g[6] = {16 + mode, a.tiles ? c->n_active : c->col_gx, c->B, c->col_threads, c->colp ? 1 : 0, c->sms};
        hash_mix(c, g, sizeof g);
        hash_mix(c, &a, sizeof a);
        return 0;
    }
    prof_mark(c, 3 + mode, true);
    // compile-time specialisation of the fused constraint (slmgs_kernels.h, VAR_*)
    int var = VAR_GENERAL;
    if (mode == COL_FUSED && !a.mraf && a.wsq_slot < 0) {  // (immediate normalisation only exists in the general variant)
        const bool pow_like = a.wgs.method == METHOD_LEONARDO || a.wgs.method == METHOD_KIM;
        if (!a.wgs_update && a.phase_mode == PHASE_COMPUTE) var = VAR_GS;
        else if (a.wgs_update && pow_like && a.phase_mode == PHASE_COMPUTE) var = VAR_POW;
        else if (a.wgs_update && pow_like && a.phase_mode == PHASE_STORED) var = VAR_POW_STORED;
    }
    const int gx = a.tiles ? c->n_active : c->col_gx;
    int e;
    if (mode == COL_FUSED && c->teams_col && !a.tiles) {
        const int pgx = teams_blocks(c, c->icol, c->W / 2);
        e = rt_check(c, launch_colt(c->H, var, c->h == c->H ? 1 : 0, pgx, c->B, c->stream, a, c->tmap_pairs),
                     "team column kernel launch");
    } else if ((mode == COL_FUSED || mode == COL_FWD) && c->teams_col8 && !a.tiles && !a.store_farfield && !a.store_phaseff) {
        int pgx = c->sms / c->B;
        if (pgx < 1) pgx = 1;
        e = rt_check(c, launch_colt8(mode, var, pgx, c->B, c->stream, a, c->tmap_col1), "team column kernel launch (8192)");
    } else if (mode == COL_FUSED && c->colp) {
        // persistent: about one block per SM over the whole batch, each walking over its share of the tiles
        int pgx = c->sms / c->B;
        if (pgx < 1) pgx = 1;
        if (pgx > gx) pgx = gx;
        e = rt_check(c, launch_colp(c->H, var, c->colp_dense ? 1 : 0, pgx, c->B, c->col_threads, c->stream, a),
                     "persistent column kernel launch");
    } else {
        e = rt_check(c, launch_col(c->H, mode, var, gx, c->B, c->col_threads, c->stream, a), "column kernel launch");
    }
    prof_mark(c, 3 + mode, false);
    return e;
}
static int ensure_farfield(slmgs_ctx* c) {
    if (c->farfield) return 0;
    return dev_alloc(c, &c->farfield, (size_t)c->B * c->H * c->W);
}
static void apply_params(ColArgs& a, const slmgs_params* p) {
    a.zero_factor = p->zero_factor;
    if (!(p->mraf && p->zero_weights)) a.zero_w = nullptr;
    a.wgs.method = p->method;
    a.wgs.p = p->feedback_exponent;
    a.wgs.f = p->feedback_factor;
    a.phase_mode = p->phase_mode;
    a.mraf = p->mraf;
    a.mraf_has_factor = p->mraf_has_factor;
    a.mraf_factor = p->mraf_factor;
}
static int ensure_zero_weights(slmgs_ctx* c) {
    if (c->zero_w) return 0;
    const size_t n = (size_t)c->B * c->H * c->W;
    int e = dev_alloc(c, &c->zero_w, n);
    if (e) return e;
    return rt_check(c, rt_memset(c->zero_w, 0, n * sizeof(cf), c->stream), "memset");
}

static int check_params(slmgs_ctx* c, const slmgs_params* p) {
    if (!p) return fail(c, SLMGS_ERR_INVALID, "params is NULL");
    if (p->mraf && p->zero_weights) {
        int e = ensure_zero_weights(c);
        if (e) return e;
    }
    if (p->method < SLMGS_GS || p->method > SLMGS_WGS_TANH) return fail(c, SLMGS_ERR_INVALID, "unknown method");
    if (p->phase_mode < 0 || p->phase_mode > 2) return fail(c, SLMGS_ERR_INVALID, "unknown phase_mode");
    return 0;
}

// ------------------------------------------------------------------------------------------
// fused loop
// ------------------------------------------------------------------------------------------
static int update_weights_pixel_impl(slmgs_ctx* c, const slmgs_params* p);
static int update_weights_spot_impl(slmgs_ctx* c, const slmgs_params* p, int width);

// ------------------------------------------------------------------------------------------
// Sparse far field.  farfield = weights * exp(i phase_ff) is zero wherever the weights are zero, and a zero
// weight stays zero under every update, so a column tile with all-zero weights (and no MRAF noise pixel, and no
// spot-feedback window) contributes nothing: the column kernels launch only the other tiles and the row kernels
// neither store nor load the skipped columns.  Results are identical to the dense loop; spot targets (the
// dominant use of the reference, cf. its CompressedSpotHologram rationale, _spots.py:222-241) run several times
// faster.  Returns with c->sparse_now set when the run should use the tile list.
// ------------------------------------------------------------------------------------------
static int prepare_sparse(slmgs_ctx* c, const slmgs_params* params, int n_iter) {
    c->sparse_now = false;
    if (!c->sparse_mode || n_iter < 1) return 0;
    int mraf = 0, spot_width = 0, nogrette = 0;
    for (int i = 0; i < n_iter; ++i) {
        const slmgs_params* p = params + i;
        if (p->mraf && p->zero_weights) return 0;   // farfield[zero] = zero_weights: dense by construction
        if (p->update_weights && p->feedback == 0 && p->mraf && p->method == SLMGS_WGS_NOGRETTE)
            return 0;                               // MRAF + Nogrette takes the element-wise route over every pixel
        // WGS-Nogrette's mean runs over the whole far field, but the ratio is exactly 1 wherever the target is zero:
        // tiles with a non-zero target stay active and the others are counted analytically
        if (p->update_weights && p->feedback == 0 && p->method == SLMGS_WGS_NOGRETTE) nogrette = 1;
        if (p->mraf) mraf = 1;
        if (p->update_weights && p->feedback == 1) {
            if (spot_width && spot_width != p->spot_width) return 0;
            spot_width = p->spot_width;
        }
    }
    const int C = c->col_threads / c->icol.tpl;
    const int ntiles = c->W / C;
    const size_t B = (size_t)c->B;
    if (ntiles < 4 || C * 16 > c->W) return 0;  // (the row kernel's flag layout needs whole tiles per W/16 columns)
    int e;
    if (!c->tile_flags) {
        if ((e = dev_alloc(c, &c->tile_flags, B * ntiles))) return e;
        if ((e = dev_alloc(c, &c->tile_list, B * ntiles))) return e;
        if ((e = dev_alloc(c, &c->tile_byte, B * ntiles))) return e;
        if ((e = dev_alloc(c, &c->tile_count, B))) return e;
        c->tiles_dirty = true;
    }
    if (c->tiles_dirty) {
        RT(c, rt_memset(c->tile_flags, 0, B * ntiles * sizeof(int), c->stream));
        TileArgs t;
        memset(&t, 0, sizeof t);
        t.weights = c->weights; t.target = c->target;
        t.img_bs = (long long)c->H * c->W; t.target_bs = c->target_shared ? 0 : t.img_bs;
        t.tile_elems = (long long)c->H * C;
        t.flags = c->tile_flags;
        t.flags_bs = ntiles;
        c->launches++;
        RT(c, launch_kernel<TileFlagKernel>(ntiles, c->B, 256, 0, c->stream, t));
        c->tile_flags_h.resize(B * ntiles);
        RT(c, rt_d2h(c->tile_flags_h.data(), c->tile_flags, B * ntiles * sizeof(int), c->stream));
        RT(c, rt_sync(c->stream));
        c->tiles_dirty = false;
        c->tile_key = -1;
    }
    const int key = mraf | (nogrette << 1) | (spot_width << 2);
    const int mask = 1 | (mraf ? 2 : 0) | (nogrette ? 4 : 0);
    if (key != c->tile_key) {
        // tiles under the analysis.take windows (SpotGatherKernel): x_n + floor(k - (w-1)/2), negative indices wrap;
        // the spot list is shared by the batch
        std::vector<unsigned char> win(ntiles, 0);
        if (spot_width > 0) {
            const int base = (spot_width & 1) ? -((spot_width - 1) / 2) : -(spot_width / 2);
            for (int x0 : c->spot_x_h)
                for (int d = 0; d < spot_width; ++d) {
                    int x = x0 + base + d;
                    if (x < 0) x += c->W;
                    if (x >= c->W) continue;  // out of range: the gather would fault in the reference too
                    win[((x + (c->W >> 1)) % c->W) / C] = 1;
                }
        }
        std::vector<unsigned char> on(B * ntiles, 0);
        std::vector<int> list(B * ntiles, 0), count(B, 0);
        long long total = 0;
        int most = 0;
        for (size_t b = 0; b < B; ++b) {
            int n = 0;
            const int ts = ntiles / 16;  // tiles per W/16 columns
            for (int t = 0; t < ntiles; ++t) {
                const bool a = (c->tile_flags_h[b * ntiles + t] & mask) || win[t];
                // row-kernel order (RowArgs::colflag): tile q + ts * m at byte 16 q + m
                on[b * ntiles + (size_t)(t % ts) * 16 + t / ts] = a ? 1 : 0;
                if (a) list[b * ntiles + n++] = t;
            }
            count[b] = n;
            total += n;
            if (n > most) most = n;
        }
        c->n_active = most;
        c->n_active_total = total;
        c->tile_key = key;
        if (most > 0) {
            RT(c, rt_h2d(c->tile_list, list.data(), list.size() * sizeof(int), c->stream));
            RT(c, rt_h2d(c->tile_byte, on.data(), on.size(), c->stream));
            RT(c, rt_h2d(c->tile_count, count.data(), count.size() * sizeof(int), c->stream));
        }
    }
    // worthwhile below half occupancy (the filtered row kernel costs a little more per element than the dense one);
    // every hologram needs at least one active tile (an empty one would leave stale columns in its field)
    c->sparse_now = c->n_active > 0 && 2 * c->n_active_total <= (long long)B * ntiles;
    if (c->sparse_now) {
        for (size_t b = 0; b < B && c->sparse_now; ++b) {
            bool any = false;
            for (int t = 0; t < ntiles && !any; ++t) any = (c->tile_flags_h[b * ntiles + t] & mask) != 0;
            if (!any && spot_width == 0) c->sparse_now = false;
        }
    }
    return 0;
}

extern "C" int slmgs_set_sparse(slmgs_ctx* c, int mode) {
    CHECK_CTX(c);
    c->sparse_mode = mode ? 1 : 0;
    return SLMGS_OK;
}
extern "C" int slmgs_sparse_info(const slmgs_ctx* c, int* out3) {
    if (!c || !out3) return SLMGS_ERR_INVALID;
    out3[0] = c->last_sparse ? 1 : 0;
    out3[1] = c->n_active;
    out3[2] = c->W / (c->col_threads / c->icol.tpl);
    return SLMGS_OK;
}


static int run_prepare(slmgs_ctx* c, const slmgs_params* params, int n_iter);
static int run_body(slmgs_ctx* c, const slmgs_params* params, int n_iter);
static int populate_body(slmgs_ctx* c, bool fld_ready = false);

// The launch sequence of a whole run (first row pass, n_iter x (column kernel, row kernel) [+ pre-passes],
// _populate_results) only depends on host state.  Small fields are launch bound (512^2: ~7 us of host time per
// launch against kernels of a few us), so a sequence that is run a second time with bit-identical kernel arguments is
// captured into a CUDA graph -- programmatic-dependent-launch edges included -- and replayed from then on.  The key is
// a hash over every kernel's argument block from a dry run of the same code that issues the launches.
#ifndef SLMGS_EMULATE
static bool graph_eligible(slmgs_ctx* c, const slmgs_params* params, int n_iter) {
    if (!c->use_graphs || c->profiling || n_iter < 1) return false;
    for (int i = 0; i < n_iter; ++i) {
        const slmgs_params* p = params + i;
        if (!p->update_weights) continue;
        if (p->feedback != 0) return false;                                   // spot feedback: element-wise kernels + scratch
        if (p->mraf && (p->zero_weights || p->method == SLMGS_WGS_NOGRETTE)) return false;  // element-wise route
    }
    return true;
}
static bool key_in(const std::vector<unsigned long long>& v, unsigned long long k) {
    for (unsigned long long x : v) if (x == k) return true;
    return false;
}
#endif

// first row pass + iterations [+ _populate_results, which sees the whole far field]; leaves sparse_now as it found it
static int run_sequence(slmgs_ctx* c, const slmgs_params* params, int n_iter, int populate) {
    const bool sp = c->sparse_now;
    int e = run_body(c, params, n_iter);
    c->sparse_now = false;
    // A dense run ends with the fused row kernel, whose forward half has just written the row spectrum of the final near
    // field (amp z/|z|, z = the field whose angle was stored as the phase): _populate_results only needs the column pass.
    // (A sparse run stored the active column tiles only, and the stored-phase rebuild is bit-faithful to the reference's
    // exp(i phase): SLMGS_POPULATE_REBUILD=1 keeps it.)
    if (!e && populate) e = populate_body(c, n_iter > 0 && !sp && c->populate_shortcut);
    c->sparse_now = sp;
    return e;
}

extern "C" int slmgs_run(slmgs_ctx* c, const slmgs_params* params, int n_iter, int populate) {
    CHECK_CTX(c);
    int e = run_prepare(c, params, n_iter);
    bool done = false;
#ifndef SLMGS_EMULATE
    if (!e && graph_eligible(c, params, n_iter)) {
        // dry run: hash of every launch of the sequence, and the host state it leaves behind
        const int w_in = c->w_pending;
        const long long l_in = c->launches;
        c->launch_mode = 1;
        c->hash = 1469598103934665603ull ^ (populate ? 0x9e3779b97f4a7c15ull : 0ull);
        e = run_sequence(c, params, n_iter, populate);
        c->launch_mode = 0;
        const unsigned long long key = c->hash;
        const int w_out = c->w_pending;
        const long long n_launch = c->launches - l_in;
        c->w_pending = w_in;
        c->launches = l_in;
        slmgs_ctx::GraphEntry* hit = nullptr;
        if (!e) {
            for (auto& g : c->graphs)
                if (g.key == key) hit = &g;
        }
        if (!e && !hit && key_in(c->graph_seen, key) && !key_in(c->graph_bad, key)) {
            cudaGraph_t graph = nullptr;
            cudaGraphExec_t exec = nullptr;
            bool ok = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
            if (ok) {
                const int be = run_sequence(c, params, n_iter, populate);
                ok = cudaStreamEndCapture(c->stream, &graph) == cudaSuccess && !be && graph;
            }
            if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
            if (graph) cudaGraphDestroy(graph);
            c->w_pending = w_in;
            c->launches = l_in;
            if (ok) {
                if (c->graphs.size() >= 16) {  // evict the least recently used
                    size_t lru = 0;
                    for (size_t i = 1; i < c->graphs.size(); ++i)
                        if (c->graphs[i].last_use < c->graphs[lru].last_use) lru = i;
                    cudaGraphExecDestroy(c->graphs[lru].exec);
                    c->graphs.erase(c->graphs.begin() + lru);
                }
                c->graphs.push_back({key, exec, w_out, n_launch, 0});
                hit = &c->graphs.back();
            } else {
                cudaGetLastError();
                c->err.clear();
                c->graph_bad.push_back(key);
            }
        }
        if (hit) {
            hit->last_use = ++c->graph_clock;
            e = rt_check(c, (int)cudaGraphLaunch(hit->exec, c->stream), "graph launch");
            c->w_pending = hit->w_pending_out;
            c->launches += hit->launches;
            done = true;
        } else if (!e && !key_in(c->graph_seen, key)) {
            if (c->graph_seen.size() >= 64) c->graph_seen.erase(c->graph_seen.begin());
            c->graph_seen.push_back(key);
        }
    }
#endif
    if (!e && !done) e = run_sequence(c, params, n_iter, populate);
    c->last_sparse = c->sparse_now;
    c->sparse_now = false;  // every stepped entry point sees the whole far field
    c->ff_valid = (!e && populate) ? (c->farfield != nullptr) : false;
    return e;
}

static int run_prepare(slmgs_ctx* c, const slmgs_params* params, int n_iter) {
    if (n_iter < 0) return fail(c, SLMGS_ERR_INVALID, "n_iter < 0");
    if (n_iter > 0 && !params) return fail(c, SLMGS_ERR_INVALID, "params is NULL");
    for (int i = 0; i < n_iter; ++i) {
        int e = check_params(c, params + i);
        if (e) return e;
        if (params[i].update_weights && params[i].method == SLMGS_GS) return fail(c, SLMGS_ERR_INVALID, "GS has no weight update");
        if (params[i].update_weights && params[i].feedback == 1 && c->n_spots < 1)
            return fail(c, SLMGS_ERR_STATE, "spot feedback without slmgs_set_spots");
    }
    return prepare_sparse(c, params, n_iter);
}

// ---- small square fields, GS: the whole loop in one cooperative kernel (slmgs_loop.h) --------------------------------
// Geometry: every block runs one column tile (the context's tile width: it fixes the image layout) and one row group
// with the same number of threads; all blocks must be resident at once.
static bool loop_eligible(slmgs_ctx* c, const slmgs_params* params, int n_iter) {
#ifdef SLMGS_EMULATE
    (void)c; (void)params; (void)n_iter;
    return false;
#else
    if (n_iter < 2 || c->H != c->W || c->H < 256 || c->H > 1024 || c->sparse_now || c->profiling || c->launch_mode != 0 ||
        c->w_pending >= 0 || c->prop || env_int("SLMGS_LOOP", 1) == 0)
        return false;
    for (int i = 0; i < n_iter; ++i) {
        const slmgs_params* p = params + i;
        if (p->update_weights || p->mraf || p->phase_mode != SLMGS_PHASE_COMPUTE) return false;
    }
    const int nt = c->col_threads;
    const int lines = nt / c->irow.tpl;
    if (lines < 1 || (c->pairs && (lines % 2))) return false;
    const int col_items = c->W / (nt / c->icol.tpl), row_items = (c->h + lines - 1) / lines;
    const int gx = col_items > row_items ? col_items : row_items;
    if (c->loop_bps < 0) {
        LoopArgs dummy;
        memset(&dummy, 0, sizeof dummy);
        c->loop_bps = launch_loop(c->H, c->pairs ? 2 : 1, gx, c->B, nt, c->stream, dummy, 1);
    }
    if ((long long)gx * c->B > (long long)c->loop_bps * c->sms) return false;
    if (!c->gbar) {
        if (dev_alloc(c, &c->gbar, 1)) return false;
        if (rt_memset(c->gbar, 0, sizeof(unsigned), c->stream)) return false;
        c->gbar_epoch = 0;
    }
    return true;
#endif
}
static int run_loop(slmgs_ctx* c, const slmgs_params* params, int n_iter) {
#ifdef SLMGS_EMULATE
    (void)c; (void)params; (void)n_iter;
    return SLMGS_ERR_STATE;
#else
    const int nt = c->col_threads;
    const int lines = nt / c->irow.tpl;
    LoopArgs la;
    memset(&la, 0, sizeof la);
    la.c = col_args(c);
    apply_params(la.c, params);
    la.c.wgs_update = 0;
    la.c.w_in_slot = -1;
    la.c.pf_dist = 0;
    la.c.pdl = 0;
    la.r = row_args(c);
    la.r.pf_dist = 0;
    la.r.pdl = 0;
    la.r_last = la.r;
    la.r_last.store_phase = 1;
    la.n_iter = n_iter;
    la.col_items = c->W / (nt / c->icol.tpl);
    la.row_items = (c->h + lines - 1) / lines;
    la.gbar = c->gbar;
    la.epoch0 = c->gbar_epoch;
    const int gx = la.col_items > la.row_items ? la.col_items : la.row_items;
    const int e = launch_loop(c->H, c->pairs ? 2 : 1, gx, c->B, nt, c->stream, la, 0);
    if (e) {
        cudaGetLastError();
        return SLMGS_ERR_CUDA;
    }
    c->gbar_epoch += (unsigned)(2 * n_iter - 1) * (unsigned)(gx * c->B);
    c->launches++;
    return SLMGS_OK;
#endif
}

static int run_body(slmgs_ctx* c, const slmgs_params* params, int n_iter) {
    int e;
    if (n_iter > 0) {
        // a weight update is done inside the fused kernel when it has no global dependency within the
        // iteration (the L2 renormalisation is deferred by one kernel, see DESIGN.md "Lazy normalisation")
        auto in_kernel_update = [&](const slmgs_params* p) {
            return p->update_weights && p->feedback == 0 && !p->mraf &&
                   (p->method == SLMGS_WGS_LEONARDO || p->method == SLMGS_WGS_KIM || p->method == SLMGS_WGS_WU ||
                    p->method == SLMGS_WGS_TANH || p->method == SLMGS_WGS_NOGRETTE);
        };
        // WGS-Nogrette needs mean(ratio) over the whole far field before any weight changes (:1851-1852): a forward
        // column pre-pass accumulates the sum of the ratio (no stores), the fused kernel then updates with that mean
        auto needs_ratio = [&](const slmgs_params* p) { return in_kernel_update(p) && p->method == SLMGS_WGS_NOGRETTE; };
        // MRAF + WGS with pixel feedback: the noise region passes the field through, so the weights must carry their
        // final normalisation inside the same iteration; a forward column pre-pass accumulates sum(w_new^2)
        auto mraf_fused = [&](const slmgs_params* p) {
            return p->update_weights && p->feedback == 0 && p->mraf && !p->zero_weights &&
                   (p->method == SLMGS_WGS_LEONARDO || p->method == SLMGS_WGS_KIM || p->method == SLMGS_WGS_WU ||
                    p->method == SLMGS_WGS_TANH);
        };
        auto presum_slot = [&](const slmgs_params* p) -> double* {
            if (needs_ratio(p)) return c->acc + ACC_MEAN;
            if (mraf_fused(p)) return c->acc + ACC_TMP;
            return nullptr;
        };
        // the accumulator slot that receives sum(w^2) alternates; the row kernel that precedes a column kernel
        // clears that kernel's slot (no memset node between the kernels)
        auto out_slot_after = [&](int pending) { return pending == ACC_W0 ? ACC_W1 : ACC_W0; };
        RowArgs ra = row_args(c);
        if (in_kernel_update(params)) ra.zero_acc = c->acc + out_slot_after(c->w_pending);
        // the row kernel in front of a column kernel turns the pending sum(w^2) into the float factor that kernel uses
        auto set_win = [&](RowArgs& r) {
            r.win_src = c->w_pending >= 0 ? c->acc + c->w_pending : nullptr;
            r.win_dst = c->w_pending >= 0 ? c->winf : nullptr;
        };
        set_win(ra);
        ra.zero_acc2 = presum_slot(params);
        if ((e = run_row(c, ROW_FIRST, ra))) return e;
        if (loop_eligible(c, params, n_iter)) {
            if (run_loop(c, params, n_iter) == SLMGS_OK) return SLMGS_OK;
            c->loop_bps = 0;  // the cooperative launch was refused: never again on this context, plain sequence from here
            c->err.clear();
        }
        for (int i = 0; i < n_iter; ++i) {
            const slmgs_params* p = params + i;
            ColArgs ca = col_args(c);
            apply_params(ca, p);
            const bool fused_mraf = mraf_fused(p);
            const bool in_kernel = in_kernel_update(p) || fused_mraf;
            const bool need_amp = p->update_weights && !in_kernel;
            const bool need_phase = ca.phase_mode == PHASE_COMPUTE_STORE;
            const bool need_ratio = needs_ratio(p);
            if (need_ratio) ca.ratio_slot = ACC_MEAN;  // cleared by the row kernel in front of this iteration
            if (need_amp || need_phase || need_ratio || fused_mraf) {
                // one forward column pass for |farfield| (global-dependency updates: Nogrette mean, per-spot
                // windows, MRAF + WGS) and/or angle(farfield) (the WGS-Kim iteration that fixes the phase)
                ColArgs fa = col_args(c);
                fa.store_ampff = need_amp;
                fa.store_phaseff = need_phase;
                if (need_ratio || fused_mraf) {
                    if (need_ratio) fa.ratio_slot = ACC_MEAN;
                    if (fused_mraf) {
                        fa.wsq_slot = ACC_TMP;
                        fa.w_in_slot = c->w_pending;
                        fa.win_f = c->winf;
                    }
                    fa.wgs.method = p->method;
                    fa.wgs.p = p->feedback_exponent;
                    fa.wgs.f = p->feedback_factor;
                }
                if ((e = run_col(c, COL_FWD, fa))) return e;
                if (need_phase) ca.phase_mode = PHASE_STORED;
            }
            if (need_amp) {
                if (p->feedback == 1) e = update_weights_spot_impl(c, p, p->spot_width);
                else e = update_weights_pixel_impl(c, p);
                if (e) return e;
            }
            ca.wgs_update = in_kernel ? 1 : 0;
            ca.w_in_slot = c->w_pending;
            ca.win_f = c->winf;
            if (ca.wgs_update && !fused_mraf) ca.w_out_slot = out_slot_after(c->w_pending);
            if (fused_mraf) ca.wsq_slot = ACC_TMP;
            if ((e = run_col(c, COL_FUSED, ca))) return e;
            if (ca.wgs_update) c->w_pending = fused_mraf ? -1 : ca.w_out_slot;
            ra.store_phase = (i == n_iter - 1);
            ra.zero_acc = nullptr;
            if (i + 1 < n_iter && in_kernel_update(params + i + 1)) ra.zero_acc = c->acc + out_slot_after(c->w_pending);
            set_win(ra);
            ra.zero_acc2 = (i + 1 < n_iter) ? presum_slot(params + i + 1) : nullptr;
            if ((e = run_row(c, ROW_FUSED, ra))) return e;
        }
    }
    return SLMGS_OK;
}

// ------------------------------------------------------------------------------------------
// stepped loop
// ------------------------------------------------------------------------------------------
static int forward_impl(slmgs_ctx* c, int store_ff, int store_amp, int store_phase, bool first_row) {
    int e;
    if (store_ff && (e = ensure_farfield(c))) return e;
    if (first_row) {
        RowArgs ra = row_args(c);
        if ((e = run_row(c, ROW_FIRST, ra))) return e;
    }
    ColArgs ca = col_args(c);
    ca.store_farfield = store_ff;
    ca.store_ampff = store_amp;
    ca.store_phaseff = store_phase;
    return run_col(c, COL_FWD, ca);
}

extern "C" int slmgs_forward(slmgs_ctx* c) {
    CHECK_CTX(c);
    int e = forward_impl(c, 1, 1, 0, true);
    if (e) return e;
    c->ff_valid = true;
    return SLMGS_OK;
}

static int populate_body(slmgs_ctx* c, bool fld_ready) {
    // after a fused run `fld` already holds the row transform of the new near field only if the last
    // kernel was a dense ROW_FUSED; rebuilding from the stored phase is always valid and costs one row pass.
    return forward_impl(c, c->farfield ? 1 : 0, 1, 1, !fld_ready);
}
extern "C" int slmgs_populate(slmgs_ctx* c) {
    CHECK_CTX(c);
    int e = populate_body(c);
    if (e) return e;
    c->ff_valid = c->farfield != nullptr;
    return SLMGS_OK;
}

extern "C" int slmgs_get_farfield(slmgs_ctx* c, float* out) {
    CHECK_CTX(c);
    if (!out) return fail(c, SLMGS_ERR_INVALID, "farfield is NULL");
    int e;
    if (!c->ff_valid) {
        if ((e = forward_impl(c, 1, 1, 0, true))) return e;
        c->ff_valid = true;
    }
    const long long P = (long long)c->H * c->W;
    if (!c->stage_c && (e = dev_alloc(c, &c->stage_c, (size_t)c->B * P))) return e;
    ElemArgs a = elem_args(c, c->farfield, c->stage_c, P);
    a.unroll = 1;
    if ((e = launch_elem<EW_ROLL_C64>(c, a, c->B))) return e;
    RT(c, rt_d2h(out, c->stage_c, (size_t)c->B * P * sizeof(cf), c->stream));
    return SLMGS_OK;
}

// ------------------------------------------------------------------------------------------
// camera sampling of |farfield|^2 (SimulatedCamera._get_image_hw, hardware/cameras/simulated.py:344-402)
// ------------------------------------------------------------------------------------------
extern "C" int slmgs_set_sample_grid(slmgs_ctx* c, long long n, const double* ky, const double* kx) {
    CHECK_CTX(c);
    if (n < 1 || !ky || !kx) return fail(c, SLMGS_ERR_INVALID, "bad sampling grid");
    RT(c, rt_sync(c->stream));
    if (c->samp_y) { rt_free(c->samp_y); rt_free(c->samp_x); c->samp_y = c->samp_x = nullptr; c->n_samp = 0; }
    int e;
    if ((e = dev_alloc(c, &c->samp_y, (size_t)n))) return e;
    if ((e = dev_alloc(c, &c->samp_x, (size_t)n))) return e;
    RT(c, rt_h2d(c->samp_y, ky, (size_t)n * sizeof(double), c->stream));
    RT(c, rt_h2d(c->samp_x, kx, (size_t)n * sizeof(double), c->stream));
    c->n_samp = n;
    return SLMGS_OK;
}

extern "C" int slmgs_sample_intensity(slmgs_ctx* c, float scale, float clip_max, int out_kind, void* out) {
    CHECK_CTX(c);
    if (!out) return fail(c, SLMGS_ERR_INVALID, "out is NULL");
    if (out_kind < 0 || out_kind > 2) return fail(c, SLMGS_ERR_INVALID, "out_kind must be 0 (float32), 1 (uint8) or 2 (uint16)");
    if (c->n_samp < 1) return fail(c, SLMGS_ERR_STATE, "sample_intensity without slmgs_set_sample_grid");
    int e;
    // far field of the current phase (like get_farfield, this refreshes amp_ff; _hologram.py:922-929)
    if ((e = forward_impl(c, c->farfield ? 1 : 0, 1, 0, true))) return e;
    c->ff_valid = c->farfield != nullptr;
    const size_t esz = out_kind == 0 ? 4 : out_kind == 1 ? 1 : 2;
    const size_t bytes = (size_t)c->B * (size_t)c->n_samp * esz;
    void* dout = nullptr;
    if ((e = scratch_reserve(c, bytes, &dout))) return e;
    SampleArgs a;
    memset(&a, 0, sizeof a);
    a.amp_ff = c->amp_ff; a.img_bs = (long long)c->H * c->W;
    a.ky = c->samp_y; a.kx = c->samp_x; a.n = c->n_samp;
    a.out = dout; a.out_kind = out_kind; a.scale = scale; a.clip_max = clip_max;
    a.H = c->H; a.W = c->W; a.C = c->col_threads / c->icol.tpl;
    long long blocks = (c->n_samp + 255) / 256;
    if (blocks > (long long)c->sms * 16) blocks = (long long)c->sms * 16;
    c->launches++;
    e = rt_check(c, launch_kernel<SampleKernel>((int)blocks, c->B, 256, 0, c->stream, a), "sample launch");
    if (!e) e = rt_check(c, rt_d2h(out, dout, bytes, c->stream), "d2h");
    rt_sync(c->stream);
    return e;
}

// pixel feedback on amp_ff (must hold |farfield| of the current iteration): _hologram.py:1822-1879
static int update_weights_pixel_impl(slmgs_ctx* c, const slmgs_params* p) {
    int e;
    if ((e = resolve_weights(c))) return e;
    const long long P = (long long)c->H * c->W;
    // ||feedback||, _hologram.py:1830-1831
    if ((e = zero_slot(c, ACC_FNORM, 2))) return e;  // FNORM and MEAN
    ElemArgs a = elem_args(c, c->amp_ff, nullptr, P);
    a.slot0 = ACC_FNORM;
    if ((e = launch_elem<EW_SUMSQ>(c, a, c->B))) return e;
    WgsParams q;
    q.method = p->method; q.p = p->feedback_exponent; q.f = p->feedback_factor; q.inv_fnorm = 1.f; q.neg_inv_mean = -1.f;
    if (p->method == SLMGS_WGS_NOGRETTE) {
        ElemArgs m = elem_args(c, c->amp_ff, nullptr, P);
        m.wgs = q; m.fnorm_slot = ACC_FNORM; m.slot0 = ACC_MEAN;
        if ((e = launch_elem<EW_RATIO_SUM>(c, m, c->B))) return e;
    }
    if ((e = zero_slot(c, ACC_W0))) return e;
    ElemArgs u = elem_args(c, c->amp_ff, c->weights, P);
    u.wgs = q; u.fnorm_slot = ACC_FNORM; u.mean_slot = (p->method == SLMGS_WGS_NOGRETTE) ? ACC_MEAN : -1; u.slot1 = ACC_W0;
    if ((e = launch_elem<EW_WGS_UPDATE>(c, u, c->B))) return e;
    ElemArgs s = elem_args(c, nullptr, c->weights, P);
    s.slot0 = ACC_W0;
    return launch_elem<EW_SCALE>(c, s, c->B);
}

extern "C" int slmgs_update_weights(slmgs_ctx* c, const slmgs_params* p) {
    CHECK_CTX(c);
    int e = check_params(c, p);
    if (e) return e;
    if (p->method == SLMGS_GS) return fail(c, SLMGS_ERR_INVALID, "Weighting is only for WGS.");
    if (!c->ff_valid) return fail(c, SLMGS_ERR_STATE, "update_weights needs a preceding forward()");
    return update_weights_pixel_impl(c, p);
}

extern "C" int slmgs_set_spots(slmgs_ctx* c, int n, const int* x, const int* y, const float* spot_amp) {
    CHECK_CTX(c);
    if (n < 1 || !x || !y || !spot_amp) return fail(c, SLMGS_ERR_INVALID, "bad spot arguments");
    for (int i = 0; i < n; ++i)
        if (x[i] < 0 || x[i] >= c->W || y[i] < 0 || y[i] >= c->H) return fail(c, SLMGS_ERR_INVALID, "spot outside the computational space");
    RT(c, rt_sync(c->stream));
    auto drop_spots = [&]() {
        void** ps[] = {(void**)&c->spot_x, (void**)&c->spot_y, (void**)&c->spot_amp, (void**)&c->spot_pw, (void**)&c->spot_wn,
                       (void**)&c->spot_keep};
        for (void** q : ps) {
            if (*q) rt_free(*q);
            *q = nullptr;
        }
        c->n_spots = 0;
    };
    drop_spots();
    int e;
    if ((e = dev_alloc(c, &c->spot_x, (size_t)n)) || (e = dev_alloc(c, &c->spot_y, (size_t)n)) ||
        (e = dev_alloc(c, &c->spot_amp, (size_t)n)) || (e = dev_alloc(c, &c->spot_pw, (size_t)n * c->B)) ||
        (e = dev_alloc(c, &c->spot_wn, (size_t)n * c->B)) || (e = dev_alloc(c, &c->spot_keep, (size_t)n))) {
        drop_spots();
        return e;
    }
    // two spots may round to the same pixel: the reference scatters the N-vector with a fancy index (_spots.py:1622-1624),
    // where the LAST occurrence wins
    std::vector<unsigned char> keep((size_t)n, 1);
    {
        std::vector<long long> key((size_t)n);
        for (int i = 0; i < n; ++i) key[i] = (long long)y[i] * c->W + x[i];
        std::vector<int> order((size_t)n);
        for (int i = 0; i < n; ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key[a] < key[b]; });
        for (int i = 0; i + 1 < n; ++i)
            if (key[order[i]] == key[order[i + 1]]) keep[order[i]] = 0;  // a later spot owns the pixel
    }
    RT(c, rt_h2d(c->spot_x, x, (size_t)n * sizeof(int), c->stream));
    RT(c, rt_h2d(c->spot_y, y, (size_t)n * sizeof(int), c->stream));
    RT(c, rt_h2d(c->spot_amp, spot_amp, (size_t)n * sizeof(float), c->stream));
    RT(c, rt_h2d(c->spot_keep, keep.data(), (size_t)n, c->stream));
    c->n_spots = n;
    c->spot_x_h.assign(x, x + n);
    c->tile_key = -1;
    return SLMGS_OK;
}

static SpotArgs spot_args(slmgs_ctx* c, int width) {
    SpotArgs a;
    memset(&a, 0, sizeof a);
    a.img = c->amp_ff; a.weights = c->weights; a.sx = c->spot_x; a.sy = c->spot_y; a.spot_amp = c->spot_amp;
    a.pw = c->spot_pw; a.wn = c->spot_wn; a.keep = c->spot_keep; a.img_bs = (long long)c->H * c->W; a.H = c->H; a.W = c->W; a.N = c->n_spots; a.width = width;
    a.C = c->col_threads / c->icol.tpl;
    return a;
}

static int update_weights_spot_impl(slmgs_ctx* c, const slmgs_params* p, int width) {
    int e;
    if (c->n_spots < 1) return fail(c, SLMGS_ERR_STATE, "no spots set");
    if (width < 1) return fail(c, SLMGS_ERR_INVALID, "width < 1");
    if ((e = resolve_weights(c))) return e;
    SpotArgs a = spot_args(c, width);
    a.wgs.method = p->method; a.wgs.p = p->feedback_exponent; a.wgs.f = p->feedback_factor;
    a.wgs.inv_fnorm = 1.f; a.wgs.neg_inv_mean = -1.f;
    c->launches++;
    e = rt_check(c, launch_kernel<SpotGatherKernel>((c->n_spots + 255) / 256, c->B, 256, 0, c->stream, a), "spot gather launch");
    if (e) return e;
    c->launches++;
    c->weights_pristine = false;
    return rt_check(c, launch_kernel<SpotUpdateKernel>(1, c->B, 1024, (1024 + 8 + 32) * sizeof(double), c->stream, a), "spot update launch");
}

extern "C" int slmgs_update_weights_spot(slmgs_ctx* c, const slmgs_params* p, int width) {
    CHECK_CTX(c);
    int e = check_params(c, p);
    if (e) return e;
    if (p->method == SLMGS_GS) return fail(c, SLMGS_ERR_INVALID, "Weighting is only for WGS.");
    if (!c->ff_valid) return fail(c, SLMGS_ERR_STATE, "update_weights_spot needs a preceding forward()");
    return update_weights_spot_impl(c, p, width);
}

extern "C" int slmgs_constrain_inverse(slmgs_ctx* c, const slmgs_params* p) {
    CHECK_CTX(c);
    int e = check_params(c, p);
    if (e) return e;
    if (!c->ff_valid || !c->farfield) return fail(c, SLMGS_ERR_STATE, "constrain_inverse needs a preceding forward()");
    if ((e = resolve_weights(c))) return e;
    ColArgs ca = col_args(c);
    apply_params(ca, p);
    ca.wgs_update = 0;
    if ((e = run_col(c, COL_INV, ca))) return e;
    RowArgs ra = row_args(c);
    if ((e = run_row(c, ROW_LAST, ra))) return e;
    c->ff_valid = false;
    return SLMGS_OK;
}

// ------------------------------------------------------------------------------------------
// MultiplaneHologram support
// ------------------------------------------------------------------------------------------
extern "C" int slmgs_share_stream(slmgs_ctx* c, slmgs_ctx* leader) {
    CHECK_CTX(c);
    if (!leader || leader->device != c->device) return fail(c, SLMGS_ERR_INVALID, "stream leader must live on the same device");
    if (leader == c) return SLMGS_OK;
    if (c->sref == leader->sref) return SLMGS_OK;
    RT(c, rt_sync(c->stream));
    if (c->sref && --c->sref->refs == 0) {
        rt_stream_destroy(c->sref->s);
        delete c->sref;
    }
    c->sref = leader->sref;
    c->sref->refs++;
    c->stream = leader->stream;
    return SLMGS_OK;
}

extern "C" void* slmgs_nearfield_sum_ptr(slmgs_ctx* c) {
    if (!c) return nullptr;
    if (rt_set_device(c->device)) return nullptr;
    if (!c->mp_sum && dev_alloc(c, &c->mp_sum, (size_t)c->B * c->h * c->w)) return nullptr;
    return c->mp_sum;
}

extern "C" int slmgs_constrain_accumulate(slmgs_ctx* c, const slmgs_params* p, float weight, void* sum, int first) {
    CHECK_CTX(c);
    int e = check_params(c, p);
    if (e) return e;
    if (!sum) return fail(c, SLMGS_ERR_INVALID, "sum is NULL");
    if (!c->ff_valid || !c->farfield) return fail(c, SLMGS_ERR_STATE, "constrain_accumulate needs a preceding forward()");
    if ((e = resolve_weights(c))) return e;
    ColArgs ca = col_args(c);
    apply_params(ca, p);
    ca.wgs_update = 0;
    if ((e = run_col(c, COL_INV, ca))) return e;
    RowArgs ra = row_args(c);
    ra.mp_sum = (cf*)sum;
    ra.mp_weight = weight;
    ra.mp_first = first ? 1 : 0;
    if ((e = run_row(c, ROW_LAST, ra))) return e;
    c->ff_valid = false;
    return SLMGS_OK;
}

extern "C" int slmgs_extract_phase_from_sum(slmgs_ctx* c, const void* sum) {
    CHECK_CTX(c);
    if (!sum) return fail(c, SLMGS_ERR_INVALID, "sum is NULL");
    const long long S = (long long)c->h * c->w;
    ElemArgs a = elem_args(c, sum, c->phase, S);
    c->ff_valid = false;
    return launch_elem<EW_ARG_C64>(c, a, c->B);
}

// One fused iteration of a MultiplaneHologram child (no callback / statistics): row first, the fused column
// kernel (with the same forward-pass detour as slmgs_run for updates that need |farfield| first), then the row
// inverse whose epilogue accumulates into `sum`.
extern "C" int slmgs_run_accumulate(slmgs_ctx* c, const slmgs_params* p, float weight, void* sum, int first) {
    CHECK_CTX(c);
    int e = check_params(c, p);
    if (e) return e;
    if (!sum) return fail(c, SLMGS_ERR_INVALID, "sum is NULL");
    if (p->update_weights && p->method == SLMGS_GS) return fail(c, SLMGS_ERR_INVALID, "GS has no weight update");
    if (p->update_weights && p->feedback == 1 && c->n_spots < 1)
        return fail(c, SLMGS_ERR_STATE, "spot feedback without slmgs_set_spots");
    // The children's near fields are SUMMED, so the scale of each child's far field matters: the deferred (one
    // kernel late) weight normalisation of slmgs_run is not allowed here.  Power-law / Wu / tanh updates with pixel
    // feedback use the pre-pass scheme of MRAF + WGS instead (a forward column pass accumulates sum(w_new^2), the
    // fused kernel updates and normalises at once); the others go through the forward pass + update kernels.
    const bool fused_update = p->update_weights && p->feedback == 0 && !(p->mraf && p->zero_weights) &&
                              (p->method == SLMGS_WGS_LEONARDO || p->method == SLMGS_WGS_KIM ||
                               p->method == SLMGS_WGS_WU || p->method == SLMGS_WGS_TANH);
    const bool need_amp = p->update_weights != 0 && !fused_update;
    if (!need_amp) {
        if ((e = prepare_sparse(c, p, 1))) return e;  // (one iteration: the flags persist between calls)
    } else {
        c->sparse_now = false;
    }
    struct SparseOff {  // the stepped entry points and the getters see the whole far field
        slmgs_ctx* c;
        ~SparseOff() { c->last_sparse = c->sparse_now; c->sparse_now = false; }
    } sparse_off{c};
    if ((e = resolve_weights(c))) return e;
    RowArgs ra = row_args(c);
    if (fused_update) ra.zero_acc2 = c->acc + ACC_TMP;
    if ((e = run_row(c, ROW_FIRST, ra))) return e;
    ColArgs ca = col_args(c);
    apply_params(ca, p);
    const bool need_phase = ca.phase_mode == PHASE_COMPUTE_STORE;
    if (need_amp || need_phase || fused_update) {
        ColArgs fa = col_args(c);
        fa.store_ampff = need_amp;
        fa.store_phaseff = need_phase;
        if (fused_update) {
            fa.wsq_slot = ACC_TMP;
            fa.wgs.method = p->method;
            fa.wgs.p = p->feedback_exponent;
            fa.wgs.f = p->feedback_factor;
        }
        if ((e = run_col(c, COL_FWD, fa))) return e;
        if (need_phase) ca.phase_mode = PHASE_STORED;
    }
    if (need_amp) {
        if (p->feedback == 1) e = update_weights_spot_impl(c, p, p->spot_width);
        else e = update_weights_pixel_impl(c, p);
        if (e) return e;
    }
    ca.wgs_update = fused_update ? 1 : 0;
    ca.w_in_slot = -1;
    if (fused_update) ca.wsq_slot = ACC_TMP;
    if ((e = run_col(c, COL_FUSED, ca))) return e;
    RowArgs rl = row_args(c);
    rl.mp_sum = (cf*)sum;
    rl.mp_weight = weight;
    rl.mp_first = first ? 1 : 0;
    if ((e = run_row(c, ROW_LAST, rl))) return e;
    c->ff_valid = false;
    return SLMGS_OK;
}

// ------------------------------------------------------------------------------------------
// statistics
// ------------------------------------------------------------------------------------------
extern "C" int slmgs_stats_pixel(slmgs_ctx* c, double* out8, double* out2) {
    CHECK_CTX(c);
    if (!out8 || !out2) return fail(c, SLMGS_ERR_INVALID, "output is NULL");
    int e;
    const long long P = (long long)c->H * c->W;
    if ((e = zero_slot(c, ACC_S0, 3))) return e;
    ElemArgs a = elem_args(c, c->amp_ff, nullptr, P);
    a.slot0 = ACC_S0; a.slot1 = ACC_S1; a.slot2 = ACC_S2;
    if ((e = launch_elem<EW_STATS1>(c, a, c->B))) return e;
    Stats2Args s;
    memset(&s, 0, sizeof s);
    s.f = c->amp_ff; s.t = c->target; s.partial = c->partial; s.acc = c->acc; s.n = P; s.f_bs = P;
    s.t_bs = c->target_shared ? 0 : P; s.acc_bs = ACC_N; s.fsum_slot = ACC_S0; s.tsum_slot = ACC_S1;
    const int gx = 256;
    c->launches++;
    e = rt_check(c, launch_kernel<Stats2Kernel>(gx, c->B, 256, 256 * 8 * sizeof(double), c->stream, s), "stats launch");
    if (e) return e;
    std::vector<double> part((size_t)c->B * gx * 8), acc((size_t)c->B * ACC_N);
    RT(c, rt_d2h(part.data(), c->partial, part.size() * sizeof(double), c->stream));
    RT(c, rt_d2h(acc.data(), c->acc, acc.size() * sizeof(double), c->stream));
    for (int b = 0; b < c->B; ++b) {
        double r[7] = {INFINITY, -INFINITY, INFINITY, -INFINITY, 0, 0, 0};
        for (int g = 0; g < gx; ++g) {
            const double* o = &part[((size_t)b * gx + g) * 8];
            if (o[0] < r[0]) r[0] = o[0];
            if (o[1] > r[1]) r[1] = o[1];
            if (o[2] < r[2]) r[2] = o[2];
            if (o[3] > r[3]) r[3] = o[3];
            r[4] += o[4]; r[5] += o[5]; r[6] += o[6];
        }
        double* o8 = out8 + (size_t)b * 8;
        o8[0] = acc[(size_t)b * ACC_N + ACC_S0];
        o8[1] = acc[(size_t)b * ACC_N + ACC_S1];
        o8[2] = acc[(size_t)b * ACC_N + ACC_S2];
        o8[3] = r[0]; o8[4] = r[1]; o8[5] = r[2]; o8[6] = r[3]; o8[7] = r[4];
        out2[(size_t)b * 2 + 0] = r[5];
        out2[(size_t)b * 2 + 1] = r[6];
    }
    return SLMGS_OK;
}

extern "C" int slmgs_window_power(slmgs_ctx* c, int n, const int* x, const int* y, int width, double* out, double* total) {
    CHECK_CTX(c);
    if (n < 1 || !x || !y || !out || width < 1) return fail(c, SLMGS_ERR_INVALID, "bad window_power arguments");
    for (int i = 0; i < n; ++i) {
        const int lo = (width & 1) ? -((width - 1) / 2) : -(width / 2), hi = lo + width - 1;
        // NumPy fancy indexing: negative indices wrap once, indices past the end raise IndexError
        if (x[i] + lo < -c->W || x[i] + hi >= c->W || y[i] + lo < -c->H || y[i] + hi >= c->H)
            return fail(c, SLMGS_ERR_INVALID, "index out of bounds in window integration");
    }
    int e;
    // grow-only scratch (a cudaMalloc / cudaFree pair per call costs milliseconds once the process holds gigabytes, and
    // this runs every iteration with stat_groups=["computational_spot"])
    const size_t ib = (((size_t)n * sizeof(int)) + 15) & ~(size_t)15;
    void* buf = nullptr;
    if ((e = scratch_reserve(c, 2 * ib + (size_t)n * c->B * sizeof(double), &buf))) return e;
    int* dx = reinterpret_cast<int*>(buf);
    int* dy = reinterpret_cast<int*>(reinterpret_cast<char*>(buf) + ib);
    double* dpw = reinterpret_cast<double*>(reinterpret_cast<char*>(buf) + 2 * ib);
    e = rt_check(c, rt_h2d(dx, x, (size_t)n * sizeof(int), c->stream), "h2d");
    if (!e) e = rt_check(c, rt_h2d(dy, y, (size_t)n * sizeof(int), c->stream), "h2d");
    if (!e) {
        SpotArgs a = spot_args(c, width);
        a.sx = dx; a.sy = dy; a.pw = dpw; a.N = n;
        c->launches++;
        e = rt_check(c, launch_kernel<SpotGatherKernel>((n + 255) / 256, c->B, 256, 0, c->stream, a), "spot gather launch");
    }
    if (!e) e = rt_check(c, rt_d2h(out, dpw, (size_t)n * c->B * sizeof(double), c->stream), "d2h");
    if (!e && total) {
        e = zero_slot(c, ACC_TMP);
        if (!e) {
            ElemArgs s = elem_args(c, c->amp_ff, nullptr, (long long)c->H * c->W);
            s.slot0 = ACC_TMP;
            e = launch_elem<EW_SUMSQ>(c, s, c->B);
        }
        std::vector<double> acc((size_t)c->B * ACC_N);
        if (!e) e = rt_check(c, rt_d2h(acc.data(), c->acc, acc.size() * sizeof(double), c->stream), "d2h");
        if (!e)
            for (int b = 0; b < c->B; ++b) total[b] = acc[(size_t)b * ACC_N + ACC_TMP];
    }
    rt_sync(c->stream);
    return e;
}

// ------------------------------------------------------------------------------------------
// kernel timing (bench / roofline): CUDA events on the launching stream
// ------------------------------------------------------------------------------------------
extern "C" int slmgs_time_kernel(slmgs_ctx* c, int which, int n, float* ms_out) {
    CHECK_CTX(c);
    if (n < 1 || !ms_out) return fail(c, SLMGS_ERR_INVALID, "bad time_kernel arguments");
#ifdef SLMGS_EMULATE
    *ms_out = 0.f;
    return fail(c, SLMGS_ERR_STATE, "kernel timing needs a GPU");
#else
    cudaEvent_t e0, e1;
    RT(c, (int)cudaEventCreate(&e0));
    RT(c, (int)cudaEventCreate(&e1));
    RowArgs ra = row_args(c);
    ColArgs ca = col_args(c);
    ca.store_ampff = 1; ca.store_phaseff = 1;
    int e = 0;
    // one untimed launch first (instruction cache, attribute set-up)
    for (int i = -1; i < n && !e; ++i) {
        if (i == 0) e = rt_check(c, (int)cudaEventRecord(e0, c->stream), "event record");
        if (e) break;
        switch (which) {
            case 0: e = run_row(c, ROW_FUSED, ra); break;
            case 1: e = run_col(c, COL_FUSED, ca); break;
            case 2: e = run_row(c, ROW_FIRST, ra); break;
            case 3: e = run_col(c, COL_FWD, ca); break;
            default: e = fail(c, SLMGS_ERR_INVALID, "unknown kernel selector");
        }
    }
    if (!e) e = rt_check(c, (int)cudaEventRecord(e1, c->stream), "event record");
    if (!e) e = rt_check(c, (int)cudaEventSynchronize(e1), "event sync");
    float ms = 0.f;
    if (!e) e = rt_check(c, (int)cudaEventElapsedTime(&ms, e0, e1), "event elapsed");
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_out = ms / (float)n;
    c->ff_valid = false;
    return e;
#endif
}

// ------------------------------------------------------------------------------------------
// phase snapshot, timers, per-kernel profile
// ------------------------------------------------------------------------------------------
extern "C" int slmgs_save_phase(slmgs_ctx* c) {
    CHECK_CTX(c);
    const size_t n = (size_t)c->B * c->h * c->w;
    int e;
    if (!c->phase_saved && (e = dev_alloc(c, &c->phase_saved, n))) return e;
#ifdef SLMGS_EMULATE
    memcpy(c->phase_saved, c->phase, n * sizeof(float));
#else
    RT(c, (int)cudaMemcpyAsync(c->phase_saved, c->phase, n * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
#endif
    return SLMGS_OK;
}
extern "C" int slmgs_restore_phase(slmgs_ctx* c) {
    CHECK_CTX(c);
    if (!c->phase_saved) return fail(c, SLMGS_ERR_STATE, "restore_phase without save_phase");
    const size_t n = (size_t)c->B * c->h * c->w;
#ifdef SLMGS_EMULATE
    memcpy(c->phase, c->phase_saved, n * sizeof(float));
#else
    RT(c, (int)cudaMemcpyAsync(c->phase, c->phase_saved, n * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
#endif
    c->ff_valid = false;
    return SLMGS_OK;
}
extern "C" int slmgs_timer_start(slmgs_ctx* c) {
    CHECK_CTX(c);
#ifndef SLMGS_EMULATE
    if (!c->t0) {
        RT(c, (int)cudaEventCreate(&c->t0));
        RT(c, (int)cudaEventCreate(&c->t1));
    }
    RT(c, (int)cudaEventRecord(c->t0, c->stream));
#endif
    return SLMGS_OK;
}
extern "C" int slmgs_timer_stop(slmgs_ctx* c, float* ms) {
    CHECK_CTX(c);
    if (!ms) return fail(c, SLMGS_ERR_INVALID, "ms is NULL");
    *ms = 0.f;
#ifndef SLMGS_EMULATE
    if (!c->t0) return fail(c, SLMGS_ERR_STATE, "timer_stop without timer_start");
    RT(c, (int)cudaEventRecord(c->t1, c->stream));
    RT(c, (int)cudaEventSynchronize(c->t1));
    RT(c, (int)cudaEventElapsedTime(ms, c->t0, c->t1));
#endif
    return SLMGS_OK;
}
extern "C" int slmgs_profile_enable(slmgs_ctx* c, int on) {
    CHECK_CTX(c);
    c->profiling = on != 0;
#ifndef SLMGS_EMULATE
    c->ev_used = 0;
    c->ev_class.clear();
#endif
    return SLMGS_OK;
}
extern "C" int slmgs_profile_read(slmgs_ctx* c, float* ms6, int* count6) {
    CHECK_CTX(c);
    if (!ms6 || !count6) return fail(c, SLMGS_ERR_INVALID, "output is NULL");
    for (int k = 0; k < 6; ++k) { ms6[k] = 0.f; count6[k] = 0; }
#ifndef SLMGS_EMULATE
    RT(c, rt_sync(c->stream));
    for (size_t i = 0; i < c->ev_class.size() && 2 * i + 1 < c->ev_used; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->ev_pool[2 * i], c->ev_pool[2 * i + 1]) == cudaSuccess) {
            ms6[c->ev_class[i]] += ms;
            count6[c->ev_class[i]]++;
        }
    }
    c->ev_used = 0;
    c->ev_class.clear();
#endif
    return SLMGS_OK;
}

// ==========================================================================================
// CompressedSpotHologram ("next" row 4): include/slmgs.h "compressed spot hologram"
// ==========================================================================================
#include "slmgs_compressed.h"

struct slmgs_comp {
    int device, h, w, N, M, MT;
    long long S;
    rt_stream stream;
    std::string err;
    long long launches;
    int sms;
    double *mono, *cw, *facc;
    float *phase, *amp, *target, *weights, *amp_ff, *phase_ff;
    cf *nf, *far, *far_norm;
    float amp_scalar;
#ifndef SLMGS_EMULATE
    cudaEvent_t t0, t1;
#endif
};

static std::string g_comp_error;
static int cfail(slmgs_comp* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    else g_comp_error = msg;
    return code;
}
static int crt(slmgs_comp* c, int e, const char* what) {
    if (e == 0) return 0;
    return cfail(c, rt_is_oom(e) ? SLMGS_ERR_OOM : SLMGS_ERR_CUDA, std::string(what) + ": " + rt_errstr(e));
}
#define CRT(c, call)                                \
    do {                                            \
        int e__ = crt((c), (call), #call);          \
        if (e__) return e__;                        \
    } while (0)
#define CHECK_COMP(c) \
    if (!(c)) return SLMGS_ERR_INVALID; \
    CRT(c, rt_set_device((c)->device))

extern "C" const char* slmgs_comp_last_error(const slmgs_comp* c) { return c ? c->err.c_str() : g_comp_error.c_str(); }

extern "C" int slmgs_comp_destroy(slmgs_comp* c) {
    if (!c) return SLMGS_OK;
    rt_set_device(c->device);
    if (c->stream) rt_sync(c->stream);
    void* ptrs[] = {c->mono, c->cw, c->facc, c->phase, c->amp, c->target, c->weights, c->amp_ff, c->phase_ff,
                    c->nf, c->far, c->far_norm};
    for (void* p : ptrs)
        if (p) rt_free(p);
#ifndef SLMGS_EMULATE
    if (c->t0) cudaEventDestroy(c->t0);
    if (c->t1) cudaEventDestroy(c->t1);
#endif
    if (c->stream) rt_stream_destroy(c->stream);
    delete c;
    return SLMGS_OK;
}

extern "C" int slmgs_comp_create(slmgs_comp** out, int device, int h, int w, int n_spots, int n_monomials) {
    if (!out) return cfail(nullptr, SLMGS_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (h < 1 || w < 1 || n_spots < 1) return cfail(nullptr, SLMGS_ERR_INVALID, "slm_shape and the number of spots must be positive");
    if (n_monomials < 1 || n_monomials > 10)
        return cfail(nullptr, SLMGS_ERR_INVALID, "the basis must have 1..10 functions");
    int e = rt_set_device(device);
    if (e) return cfail(nullptr, SLMGS_ERR_CUDA, std::string("cudaSetDevice: ") + rt_errstr(e));
    slmgs_comp* c = new slmgs_comp();
    c->device = device; c->h = h; c->w = w; c->N = n_spots; c->M = n_monomials;
    c->MT = n_monomials <= 2 ? 2 : n_monomials <= 3 ? 3 : n_monomials <= 6 ? 6 : 10;
    c->S = (long long)h * w;
    c->launches = 0;
    c->sms = rt_sm_count();
    c->stream = nullptr;
    c->amp_scalar = (float)(1.0 / sqrt((double)c->S));
    c->mono = c->cw = c->facc = nullptr;
    c->phase = c->amp = c->target = c->weights = c->amp_ff = c->phase_ff = nullptr;
    c->nf = c->far = c->far_norm = nullptr;
#ifndef SLMGS_EMULATE
    c->t0 = c->t1 = nullptr;
#endif
    const size_t S = (size_t)c->S, N = (size_t)n_spots;
    int err = 0;
    void* p;
#define CA(field, type, count)                                                              \
    if (!err) {                                                                             \
        err = rt_malloc(&p, (count) * sizeof(type));                                        \
        if (!err) { c->field = (type*)p; err = rt_memset(p, 0, (count) * sizeof(type), nullptr); } \
    }
    err = rt_stream_create(&c->stream);
    CA(mono, double, (size_t)c->MT * S)
    CA(cw, double, (size_t)c->MT * N)
    CA(facc, double, 2 * N)
    CA(phase, float, S)
    CA(target, float, N)
    CA(weights, float, N)
    CA(amp_ff, float, N)
    CA(phase_ff, float, N)
    CA(nf, cf, S)
    CA(far, cf, N)
    CA(far_norm, cf, N)
#undef CA
    if (!err) err = rt_sync(nullptr);
    if (err) {
        g_comp_error = std::string("compressed context allocation: ") + rt_errstr(err);
        const int code = rt_is_oom(err) ? SLMGS_ERR_OOM : SLMGS_ERR_CUDA;
        slmgs_comp_destroy(c);
        return code;
    }
    *out = c;
    return SLMGS_OK;
}

extern "C" int slmgs_comp_sync(slmgs_comp* c) {
    CHECK_COMP(c);
    CRT(c, rt_sync(c->stream));
    return SLMGS_OK;
}
extern "C" long long slmgs_comp_launch_count(const slmgs_comp* c) { return c ? c->launches : 0; }

// mono: [M][h*w] float64 basis-function values; cw: [M][N] float64 per-spot weights of the basis functions in RADIANS
extern "C" int slmgs_comp_set_basis(slmgs_comp* c, const double* mono, const double* cw) {
    CHECK_COMP(c);
    if (!mono || !cw) return cfail(c, SLMGS_ERR_INVALID, "mono / cw is NULL");
    CRT(c, rt_h2d(c->mono, mono, (size_t)c->M * c->S * sizeof(double), c->stream));
    std::vector<double> turns((size_t)c->M * c->N);
    const double inv = 1.0 / 6.283185307179586476925286766559;
    for (size_t i = 0; i < turns.size(); ++i) turns[i] = cw[i] * inv;
    CRT(c, rt_h2d(c->cw, turns.data(), turns.size() * sizeof(double), c->stream));
    return SLMGS_OK;
}
extern "C" int slmgs_comp_set_phase(slmgs_comp* c, const float* phase) {
    CHECK_COMP(c);
    if (!phase) return cfail(c, SLMGS_ERR_INVALID, "phase is NULL");
    CRT(c, rt_h2d(c->phase, phase, (size_t)c->S * sizeof(float), c->stream));
    return SLMGS_OK;
}
extern "C" int slmgs_comp_get_phase(slmgs_comp* c, float* phase) {
    CHECK_COMP(c);
    if (!phase) return cfail(c, SLMGS_ERR_INVALID, "phase is NULL");
    CRT(c, rt_d2h(phase, c->phase, (size_t)c->S * sizeof(float), c->stream));
    return SLMGS_OK;
}
extern "C" int slmgs_comp_set_amp_scalar(slmgs_comp* c, float amp) {
    CHECK_COMP(c);
    if (c->amp) { CRT(c, rt_sync(c->stream)); rt_free(c->amp); c->amp = nullptr; }
    c->amp_scalar = amp;
    return SLMGS_OK;
}
extern "C" int slmgs_comp_set_amp_array(slmgs_comp* c, const float* amp) {
    CHECK_COMP(c);
    if (!amp) return cfail(c, SLMGS_ERR_INVALID, "amp is NULL");
    if (!c->amp) {
        void* p = nullptr;
        CRT(c, rt_malloc(&p, (size_t)c->S * sizeof(float)));
        c->amp = (float*)p;
    }
    CRT(c, rt_h2d(c->amp, amp, (size_t)c->S * sizeof(float), c->stream));
    return SLMGS_OK;
}
#define COMP_VEC_SETTER(name, field)                                                          \
    extern "C" int slmgs_comp_set_##name(slmgs_comp* c, const float* v) {                     \
        CHECK_COMP(c);                                                                        \
        if (!v) return cfail(c, SLMGS_ERR_INVALID, #name " is NULL");                         \
        CRT(c, rt_h2d(c->field, v, (size_t)c->N * sizeof(float), c->stream));                 \
        return SLMGS_OK;                                                                      \
    }
#define COMP_VEC_GETTER(name, field, type)                                                    \
    extern "C" int slmgs_comp_get_##name(slmgs_comp* c, float* v) {                           \
        CHECK_COMP(c);                                                                        \
        if (!v) return cfail(c, SLMGS_ERR_INVALID, #name " is NULL");                         \
        CRT(c, rt_d2h(v, c->field, (size_t)c->N * sizeof(type), c->stream));                  \
        return SLMGS_OK;                                                                      \
    }
COMP_VEC_SETTER(target, target)
COMP_VEC_SETTER(weights, weights)
COMP_VEC_SETTER(phase_ff, phase_ff)
COMP_VEC_GETTER(weights, weights, float)
COMP_VEC_GETTER(amp_ff, amp_ff, float)
COMP_VEC_GETTER(phase_ff, phase_ff, float)
COMP_VEC_GETTER(farfield, far_norm, cf)

static CompArgs comp_args(slmgs_comp* c) {
    CompArgs a;
    memset(&a, 0, sizeof a);
    a.S = c->S; a.N = c->N; a.M = c->M;
    a.mono = c->mono; a.cw = c->cw; a.phase = c->phase; a.amp = c->amp; a.amp_scalar = c->amp_scalar;
    a.nf = c->nf; a.facc = c->facc; a.far = c->far; a.phase_out = c->phase;
    return a;
}
template <template <int> class K> static int comp_launch_mt(slmgs_comp* c, int gx, int gy, const CompArgs& a) {
    c->launches++;
    if (c->MT == 2) return launch_kernel<K<2>>(gx, gy, 256, 0, c->stream, a);
    if (c->MT == 3) return launch_kernel<K<3>>(gx, gy, 256, 0, c->stream, a);
    if (c->MT == 6) return launch_kernel<K<6>>(gx, gy, 256, 0, c->stream, a);
    return launch_kernel<K<10>>(gx, gy, 256, 0, c->stream, a);
}
// nearfield -> farfield accumulators (un-normalised sums in facc)
static int comp_near2far(slmgs_comp* c) {
    CompArgs a = comp_args(c);
    long long blocks = (c->S + 255) / 256;
    if (blocks > (long long)c->sms * 8) blocks = (long long)c->sms * 8;
    c->launches++;
    CRT(c, launch_kernel<CompBuildKernel>((int)blocks, 1, 256, 0, c->stream, a));
    const int gy = (c->N + COMP_SPOTS - 1) / COMP_SPOTS;
    const long long span = 256LL * COMP_PPT;
    long long gx = (c->S + span - 1) / span;
    // enough blocks to fill the GPU a few times over, few enough that every spot sees O(100) atomic adds per block row
    long long want = ((long long)c->sms * 4 + gy - 1) / gy;
    if (want < 1) want = 1;
    if (gx > want) gx = want;
    CRT(c, comp_launch_mt<CompNear2FarKernel>(c, (int)gx, gy, a));
    return SLMGS_OK;
}
static int comp_vec(slmgs_comp* c, const slmgs_params* p, int finalize) {
    CompVecArgs v;
    memset(&v, 0, sizeof v);
    v.facc = c->facc; v.far_norm = c->far_norm; v.far = c->far; v.amp_ff = c->amp_ff; v.phase_ff = c->phase_ff;
    v.weights = c->weights; v.target = c->target; v.N = c->N; v.finalize = finalize;
    v.wgs.method = METHOD_GS; v.wgs.inv_fnorm = 1.f; v.wgs.neg_inv_mean = -1.f;
    if (p) {
        v.update = p->update_weights; v.phase_mode = p->phase_mode; v.mraf = p->mraf;
        v.mraf_has_factor = p->mraf_has_factor; v.mraf_factor = p->mraf_factor;
        v.wgs.method = p->method; v.wgs.p = p->feedback_exponent; v.wgs.f = p->feedback_factor;
    }
    c->launches++;
    CRT(c, launch_kernel<CompVecKernel>(1, 1, 1024, (1024 + 8) * sizeof(double), c->stream, v));
    return SLMGS_OK;
}

// _nearfield2farfield + _midloop_cleaning (+ phase_ff = angle(farfield), _populate_results :934-949)
extern "C" int slmgs_comp_forward(slmgs_comp* c, int populate) {
    CHECK_COMP(c);
    int e;
    if ((e = comp_near2far(c))) return e;
    return comp_vec(c, nullptr, populate ? 2 : 1);
}

// ---- the same loop in pieces, for a hologram whose PIXELS are sharded over several GPUs ---------------------------
// Each rank holds a slab of SLM rows (its own context with S_local pixels, all N spots).  near -> far is a sum over
// pixels, so the per-rank accumulators are partial sums: the caller all-reduces the [N][2] float64 buffer
// (slmgs_comp_facc_ptr, 16 N bytes) between slmgs_comp_near2far and slmgs_comp_constrain_far2near; every rank then
// runs the identical N-vector stage and projects its own slab.  One tiny collective per iteration, nothing else.
extern "C" void* slmgs_comp_facc_ptr(slmgs_comp* c) { return c ? (void*)c->facc : nullptr; }
extern "C" void* slmgs_comp_stream(slmgs_comp* c) { return c ? (void*)c->stream : nullptr; }
extern "C" int slmgs_comp_near2far(slmgs_comp* c) {
    CHECK_COMP(c);
    return comp_near2far(c);
}
extern "C" int slmgs_comp_finalize(slmgs_comp* c, int populate) {
    CHECK_COMP(c);
    return comp_vec(c, nullptr, populate ? 2 : 1);
}
static int comp_check_params(slmgs_comp* c, const slmgs_params* p) {
    if (!p) return cfail(c, SLMGS_ERR_INVALID, "params is NULL");
    if (p->method < SLMGS_GS || p->method > SLMGS_WGS_TANH) return cfail(c, SLMGS_ERR_INVALID, "unknown method");
    if (p->phase_mode < 0 || p->phase_mode > 2) return cfail(c, SLMGS_ERR_INVALID, "unknown phase_mode");
    if (p->update_weights && p->method == SLMGS_GS) return cfail(c, SLMGS_ERR_INVALID, "GS has no weight update");
    if (p->mraf && p->zero_weights) return cfail(c, SLMGS_ERR_INVALID, "the MRAF zero_factor accumulator is not supported for compressed holograms");
    return 0;
}
extern "C" int slmgs_comp_constrain_far2near(slmgs_comp* c, const slmgs_params* p) {
    CHECK_COMP(c);
    int e;
    if ((e = comp_check_params(c, p))) return e;
    if ((e = comp_vec(c, p, 0))) return e;
    CompArgs a = comp_args(c);
    const long long span = 256LL * COMP_PPT;
    CRT(c, comp_launch_mt<CompFar2NearKernel>(c, (int)((c->S + span - 1) / span), 1, a));
    return SLMGS_OK;
}

// optimize_gs for the compressed maps: n_iter iterations (+ _populate_results)
extern "C" int slmgs_comp_run(slmgs_comp* c, const slmgs_params* params, int n_iter, int populate) {
    CHECK_COMP(c);
    if (n_iter < 0) return cfail(c, SLMGS_ERR_INVALID, "n_iter < 0");
    if (n_iter > 0 && !params) return cfail(c, SLMGS_ERR_INVALID, "params is NULL");
    for (int i = 0; i < n_iter; ++i) {
        const slmgs_params* p = params + i;
        if (p->method < SLMGS_GS || p->method > SLMGS_WGS_TANH) return cfail(c, SLMGS_ERR_INVALID, "unknown method");
        if (p->phase_mode < 0 || p->phase_mode > 2) return cfail(c, SLMGS_ERR_INVALID, "unknown phase_mode");
        if (p->update_weights && p->method == SLMGS_GS) return cfail(c, SLMGS_ERR_INVALID, "GS has no weight update");
        if (p->mraf && p->zero_weights) return cfail(c, SLMGS_ERR_INVALID, "the MRAF zero_factor accumulator is not supported for compressed holograms");
    }
    int e;
    for (int i = 0; i < n_iter; ++i) {
        if ((e = comp_near2far(c))) return e;
        if ((e = comp_vec(c, params + i, 0))) return e;
        CompArgs a = comp_args(c);
        const long long span = 256LL * COMP_PPT;
        CRT(c, comp_launch_mt<CompFar2NearKernel>(c, (int)((c->S + span - 1) / span), 1, a));
    }
    if (populate) return slmgs_comp_forward(c, 1);
    return SLMGS_OK;
}

extern "C" int slmgs_comp_timer(slmgs_comp* c, int start, float* ms) {
    CHECK_COMP(c);
#ifndef SLMGS_EMULATE
    if (!c->t0) {
        CRT(c, (int)cudaEventCreate(&c->t0));
        CRT(c, (int)cudaEventCreate(&c->t1));
    }
    if (start) {
        CRT(c, (int)cudaEventRecord(c->t0, c->stream));
        return SLMGS_OK;
    }
    CRT(c, (int)cudaEventRecord(c->t1, c->stream));
    CRT(c, (int)cudaEventSynchronize(c->t1));
    if (ms) CRT(c, (int)cudaEventElapsedTime(ms, c->t0, c->t1));
#else
    if (ms) *ms = 0.f;
    (void)start;
#endif
    return SLMGS_OK;
}
