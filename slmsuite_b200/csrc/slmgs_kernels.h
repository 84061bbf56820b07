// slmgs_kernels.h -- the two fused kernels of the GS / WGS loop.
//
// Index space.  The reference computes farfield = fftshift(fft2(fftshift(nearfield)))
// (slmsuite/holography/algorithms/_hologram.py:1048) and the inverse with ifftshift (:1070).
// For even sizes both shifts are the same roll by N/2, so every device array lives in
// ROLLED index space (stored index = (centred index + N/2) mod N) and the loop is a plain
// unshifted 2-D DFT; only uploads/downloads roll.  The SLM-sized arrays (phase, amp,
// propagation kernel) stay in natural (h, w) order; the centred crop of
// toolbox.unpad (toolbox/__init__.py:1665-1712) is index arithmetic inside the row kernel.
//
// One GS/WGS iteration = two kernels over the padded field `fld` (H x W complex64):
//   ColKernel<H, COL_FUSED>: columns:  forward FFT along y  ->  far-field constraint (+ weight
//        update, _hologram.py:1550-1653, :1822-1879)  ->  inverse FFT along y
//   RowKernel<W, ROW_FUSED>: rows:     inverse FFT along x  ->  near-field phase-only projection
//        (_hologram.py:1026-1036 + :1000-1011)  ->  forward FFT along x
// Only the h rows that hold the SLM are ever non-zero after the row pass / needed before
// it, so both kernels touch h*W complex values of `fld`, not H*W.
#pragma once

#include "slmgs_fft.h"
#include "slmgs_launch.h"

#include <float.h>

namespace slmgs {

enum { METHOD_GS = 0, METHOD_LEONARDO = 1, METHOD_KIM = 2, METHOD_NOGRETTE = 3, METHOD_WU = 4, METHOD_TANH = 5 };
enum { ROW_FIRST = 0, ROW_FUSED = 1, ROW_LAST = 2 };
enum { COL_FWD = 0, COL_FUSED = 1, COL_INV = 2 };
enum { PHASE_COMPUTE = 0, PHASE_COMPUTE_STORE = 1, PHASE_STORED = 2 };
enum { VAR_GENERAL = 0, VAR_GS = 1, VAR_POW = 2, VAR_POW_STORED = 3 };

// ------------------------------------------------------------------------------------------
// WGS weight multiplier: _hologram.py:1822-1867 (`_update_weights_generic_cupy`), element-wise
// part.  `famp` is the feedback amplitude, `t` the target amplitude.
// ------------------------------------------------------------------------------------------
struct WgsParams {
    int method;
    float p;             // feedback_exponent
    float f;             // feedback_factor
    float inv_fnorm;     // 1 / ||feedback||_2          (:1830-1831)
    float neg_inv_mean;  // Nogrette: -(1 / nanmean(fc)) (:1852)
};

// fc after the normalisation / division / fix-ups, before the method's nonlinearity (:1830-1843)
SLMGS_HD float wgs_ratio(float famp, float t, const WgsParams& q) {
    float fc = famp * q.inv_fnorm;
    if (q.method == METHOD_WU || q.method == METHOD_TANH) {
        fc = fc * (-q.p);
        fc = fc + t;
    } else {
        fc = fc / t;
        if (fc == INFINITY) fc = 1.0f;
        if (t == 0.0f) fc = 1.0f;
        if (fc != fc) fc = 1.0f;
    }
    return fc;
}

// Fast device math for the fused kernel (MUFU based, ~3e-7 relative): keeps the fully unrolled
// point-wise code small enough for 64 registers / the instruction cache.  The stepped path uses the
// accurate library functions.
#if defined(__CUDACC__) && !defined(SLMGS_EMULATE)
SLMGS_DEVICE float fast_pow(float x, float y) {
    float l, r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y * l));
    return r;
}
SLMGS_DEVICE float fast_lg2(float x) {
    float l;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
    return l;
}
SLMGS_DEVICE float fast_ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// flush-to-zero reciprocal square root: one MUFU, without the denormal rescaling sequence rsqrtf() carries
SLMGS_DEVICE float fast_rsqrt(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
SLMGS_DEVICE float fast_exp(float x) { return __expf(x); }
SLMGS_DEVICE float fast_tanh(float x) {
    const float c = fminf(fmaxf(x, -15.0f), 15.0f);  // (drops a NaN: restored below)
    const float e = __expf(2.0f * c);
    const float r = __fdividef(e - 1.0f, e + 1.0f);
    // a NaN must survive: with an MRAF target the additive methods see T = NaN in the noise region and the reference
    // ends up with nan_to_num(0 * NaN) = 1e-4 there (_hologram.py:1834-1835, :1870-1873)
    return (x != x) ? x : r;
}
SLMGS_DEVICE void fast_sincos(float x, float* s, float* c) { __sincosf(x, s, c); }
#else
inline float __fdividef(float a, float b) { return a / b; }
inline float fast_pow(float x, float y) { return powf(x, y); }
inline float fast_lg2(float x) { return log2f(x); }
inline float fast_ex2(float x) { return exp2f(x); }
inline float fast_rsqrt(float x) { return 1.0f / sqrtf(x); }
inline float fast_exp(float x) { return expf(x); }
inline float fast_tanh(float x) { return tanhf(x); }
inline void fast_sincos(float x, float* s, float* c) { *s = sinf(x); *c = cosf(x); }
#endif

SLMGS_HD float wgs_multiplier(float famp, float t, const WgsParams& q) {
    float fc = wgs_ratio(famp, t, q);
    switch (q.method) {
        case METHOD_LEONARDO:
        case METHOD_KIM: fc = powf(fc, -q.p); break;  // :1848
        case METHOD_NOGRETTE:                         // :1851-1855
            fc = fc * q.neg_inv_mean;
            fc = fc + 1.0f;
            fc = fc * (-q.f);
            fc = fc + 1.0f;
            fc = 1.0f / fc;
            break;
        case METHOD_WU: fc = expf(q.p * fc); break;  // :1857
        case METHOD_TANH:                            // :1859-1860
            fc = q.f * tanhf(q.p * fc);
            fc = fc + 1.0f;
            break;
        default: break;
    }
    if (fc == INFINITY) fc = 1.0f;  // :1867
    return fc;
}

// Leonardo / Kim: (F/||F|| / T)^-p with the reference's fix-ups (inf, T == 0, nan -> 1 before the power,
// inf -> 1 after) folded into one select.
SLMGS_DEVICE float wgs_multiplier_pow_fast(float famp, float t, const WgsParams& q) {
    float fc = __fdividef(famp * q.inv_fnorm, t);
    fc = fast_pow(fc, -q.p);
    return (t == 0.0f || !(fc < INFINITY)) ? 1.0f : fc;
}

// The same from |F|^2: ((|F| s / T)^-p with s = ortho scale / ||F||) = 2^(-p (lg2(|F|^2)/2 + lg2 s - lg2 T)): two
// logarithms and one exponential instead of a reciprocal square root, a division and a power (lg2s = lg2 s).
SLMGS_DEVICE float wgs_multiplier_pow_log(float m2, float t, float lg2s, float p) {
    const float x = fmaf(0.5f, fast_lg2(m2), lg2s) - fast_lg2(t);
    const float fc = fast_ex2(-p * x);
    return (t == 0.0f || !(fc < INFINITY)) ? 1.0f : fc;
}

SLMGS_DEVICE float wgs_multiplier_fast(float famp, float t, const WgsParams& q) {
    if (q.method == METHOD_LEONARDO || q.method == METHOD_KIM) return wgs_multiplier_pow_fast(famp, t, q);
    float fc = wgs_ratio(famp, t, q);
    switch (q.method) {
        case METHOD_LEONARDO:
        case METHOD_KIM: fc = fast_pow(fc, -q.p); break;
        case METHOD_NOGRETTE:
            fc = fc * q.neg_inv_mean;
            fc = fc + 1.0f;
            fc = fc * (-q.f);
            fc = fc + 1.0f;
            fc = 1.0f / fc;
            break;
        case METHOD_WU: fc = fast_exp(q.p * fc); break;
        case METHOD_TANH:
            fc = q.f * fast_tanh(q.p * fc);
            fc = fc + 1.0f;
            break;
        default: break;
    }
    if (fc == INFINITY) fc = 1.0f;
    return fc;
}

// weights *= fc; nan_to_num(nan=1e-4)  (:1870-1873; +-inf -> +-FLT_MAX as numpy does)
SLMGS_HD float wgs_apply(float w, float fc) {
    w = w * fc;
    // branch-free: clamp +-inf to +-FLT_MAX (fminf / fmaxf return the non-NaN operand), then replace NaN
    const float c = fminf(fmaxf(w, -FLT_MAX), FLT_MAX);
    return (w != w) ? 1.0e-4f : c;
}

// ==========================================================================================
// Row kernel
// ==========================================================================================
// Layout of the working field `fld` (private to the row / column kernels):
//   pairs == 0: row-major [H][W];
//   pairs == 1: ROW-PAIR INTERLEAVED [H/2][W][2]: element (r, c) at ((r >> 1) W + c) 2 + (r & 1).  A 32-byte sector
//               then holds 2 rows x 2 columns and a 64-byte segment 2 rows x 4 columns, so the column kernel's
//               strided tile access touches half as many lines per request (or serves half-width tiles, two blocks
//               per SM, with fully used sectors), while a row block that interleaves two lines in every warp
//               (RowKernel<.., LI = 2>) still reads and writes contiguous 256-byte runs.
struct RowArgs {
    cf* fld;             // [B][H][W], rolled index space (layout: see above)
    long long fld_bs;    // batch stride (elements)
    float* phase;        // [B][h][w] near-field phase (natural SLM order)
    long long phase_bs;
    const float* amp;    // [h][w] (or [B][h][w]) source amplitude, or nullptr -> amp_scalar
    long long amp_bs;
    const float* prop;   // [h][w] propagation kernel or nullptr (shared by the batch)
    const cf* twA;       // twiddle tables for N = W
    const cf* twB;
    float amp_scalar;
    float scale;         // ortho scale 1/sqrt(H W) of the inverse transform (MultiplaneHologram sum)
    int H, W, h, w, i0, i2;
    int store_phase;     // ROW_FUSED: also write the phase this iteration
    cf* mp_sum;          // ROW_LAST, MultiplaneHologram: accumulate weight * nearfield * exp(-i kernel) here instead of
                         // extracting the phase (_multiplane.py:261-279); [B][h][w], or nullptr
    float mp_weight;
    int mp_first;        // first child of the sum: overwrite instead of add
    double* zero_acc;    // accumulator slot to clear for the next column kernel ([B] slots, stride zero_bs), or nullptr
    double* zero_acc2;   // a second one (WGS-Nogrette's ratio sum), or nullptr
    int zero_bs;
    const double* win_src;  // accumulator slot holding sum(w^2) of a pending weight normalisation ([B], stride zero_bs): this
    float* win_dst;         // kernel converts it to 1/sqrt(.) in float32 for the column kernel that follows ([B]), or nullptr
    int pdl;             // launch with programmatic dependent launch
    int pf_dist;         // L2 prefetch distance in tiles (blocks resident on the GPU), 0 = off
    long long colflag_bs;          // per-hologram stride of colflag
    // The flags are stored in the order the row kernel reads them: a thread's last-stage butterfly b touches columns
    // b + (W/16) m, m = 0..15, i.e. column tiles q + TS m with q = b / C and TS = (W/16) / C; tile q + TS m sits at byte
    // 16 q + m, so the thread fetches its sixteen flags with ONE 16-byte load.
    const unsigned char* colflag;  // sparse far field: one byte per column tile of the column kernel (1 = the tile is
                                   // processed by the column kernel); columns of other tiles are identically zero after
                                   // the far-field constraint, so they are neither stored nor loaded.  nullptr = dense
    int ctile_shift;               // log2(columns per column tile)
    int pairs;                     // fld layout (must match LI of the kernel: LI == 2 <=> pairs)
};

// STORE (ROW_FUSED only): this launch also writes the phase (last iteration of a fused run)
// SPARSE (ROW_FIRST / ROW_FUSED): spectrum columns are filtered through a.colflag
// LI: lines interleaved in a warp (thread -> line tid % LI): 1 = row-major fld, 2 = row-pair interleaved fld
// DENSE (ROW_FUSED): the SLM covers the whole padded field and the source amplitude is a scalar -- no row / column
//        range tests, no amplitude loads: the projection is ~8 instructions per point
template <int N, int MODE, bool STORE = false, bool SPARSE = false, int LI = 1, bool DENSE = false> struct RowKernel {
    typedef Fft<N> F;
    static constexpr int TRACE_CLASS = 20 + MODE;  // diagnostic builds (-DSLMGS_TRACE)
    typedef RowArgs Args;
    static constexpr int E = F::E, NS = F::NS;
    static constexpr int MAXT = 16384 / E;
    static constexpr int NPHASE = (MODE == ROW_FUSED) ? 2 * NS - 1 : NS;
    struct State {
        cf v[E];
    };

    static size_t smem_bytes(int nthreads) { return NS > 1 ? (size_t)(nthreads / F::TPL) * F::PADN * sizeof(cf) : 0; }
    static constexpr int TEAM = LI * F::TPL;  // threads that exchange data: the LI interleaved lines of a warp set

#ifndef SLMGS_EMULATE
    // Lines are independent: only the threads of one team exchange data, so they synchronise on their own
    // named barrier (one per team) instead of the whole block when a team owns whole warps.
    static SLMGS_DEVICE void barrier(const ThreadId& id) {
        if constexpr (TEAM >= 128) {
            asm volatile("bar.sync %0, %1;" ::"r"(1 + id.tid / TEAM), "n"(TEAM) : "memory");
        } else {
            __syncthreads();
        }
    }
#endif

    struct Loc {
        int lt, sr, fr;  // thread-in-line, SLM row (may be >= h: idle), field row
        cf* s;           // shared-memory line base
        long long fbase, pbase;
        long long cbase;  // this hologram's offset into colflag
        bool active;
    };
    static SLMGS_DEVICE Loc locate(const Args& a, cf* smem, const ThreadId& id) {
        Loc L;
        const int team = id.tid / TEAM;
        const int line = team * LI + id.tid % LI;
        const int lines = id.nthreads / F::TPL;
        L.lt = (id.tid % TEAM) / LI;
        L.sr = id.bx * lines + line;
        L.active = DENSE || L.sr < a.h;
        L.fr = (L.sr + a.i0 + (a.H >> 1)) & (a.H - 1);
        L.s = smem + (size_t)team * (F::PADN * LI) + id.tid % LI;
        // element k of the row sits at fbase + k * LI
        if (LI == 2) L.fbase = (long long)id.by * a.fld_bs + (long long)(L.fr >> 1) * a.W * 2 + (L.fr & 1);
        else L.fbase = (long long)id.by * a.fld_bs + (long long)L.fr * a.W;
        L.pbase = (long long)id.by * a.phase_bs + (long long)L.sr * a.w;
        L.cbase = (long long)id.by * a.colflag_bs;
        return L;
    }
    // SLM column of rolled x index n, or -1 outside the SLM
    static SLMGS_DEVICE int slm_col(const Args& a, int n) {
        const unsigned sc = (unsigned)(((n + (a.W >> 1)) & (a.W - 1)) - a.i2);
        return sc < (unsigned)a.w ? (int)sc : -1;
    }
    static SLMGS_DEVICE float amp_at(const Args& a, const ThreadId& id, const Loc& L, int sc) {
        return a.amp ? __ldg(a.amp + (long long)id.by * a.amp_bs + (long long)L.sr * a.w + sc) : a.amp_scalar;
    }

    // the sixteen column-tile flags of last-stage butterfly u of this thread (see RowArgs::colflag)
    struct Flags16 {
        unsigned w[4];
    };
    static SLMGS_DEVICE Flags16 load_flags(const Args& a, const Loc& L, int u) {
        Flags16 f;
        const int q = (L.lt + F::TPL * u) >> a.ctile_shift;
        const unsigned char* p = a.colflag + L.cbase + (long long)q * 16;
#if defined(__CUDACC__) && !defined(SLMGS_EMULATE)
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
        f.w[0] = v.x; f.w[1] = v.y; f.w[2] = v.z; f.w[3] = v.w;
#else
        memcpy(f.w, p, 16);
#endif
        return f;
    }
    static SLMGS_DEVICE bool flag_byte(const Flags16& f, int m) { return ((f.w[m >> 2] >> ((m & 3) * 8)) & 0xffu) != 0; }

    // v <- spectrum of this row (inverse input), natural last-stage order
    static SLMGS_DEVICE void load_spectrum(State& st, const Args& a, const Loc& L) {
        constexpr int R = F::last_radix();
        SLMGS_UNROLL
        for (int u = 0; u < E / R; ++u) {
            Flags16 fl;
            if (SPARSE) fl = load_flags(a, L, u);
            SLMGS_UNROLL
            for (int m = 0; m < R; ++m) {
                const int k = F::last_index(L.lt + F::TPL * u, m);
                bool on = L.active;
                if (SPARSE) on = on && flag_byte(fl, m);
                st.v[u * R + m] = on ? ld_stream(a.fld + L.fbase + k * LI) : cmake(0.f, 0.f);
            }
        }
    }
    static SLMGS_DEVICE void store_spectrum(State& st, const Args& a, const Loc& L) {
        constexpr int R = F::last_radix();
        if (!L.active) return;
        SLMGS_UNROLL
        for (int u = 0; u < E / R; ++u) {
            Flags16 fl;
            if (SPARSE) fl = load_flags(a, L, u);
            SLMGS_UNROLL
            for (int m = 0; m < R; ++m) {
                const int k = F::last_index(L.lt + F::TPL * u, m);
                if (SPARSE && !flag_byte(fl, m)) continue;
                a.fld[L.fbase + k * LI] = st.v[u * R + m];
            }
        }
    }
    // v <- amp * exp(i (phase + prop)) zero-padded: _hologram.py:1000-1011
    static SLMGS_DEVICE void build_nearfield(State& st, const Args& a, const ThreadId& id, const Loc& L) {
        constexpr int R = F::R0;
        SLMGS_UNROLL
        for (int u = 0; u < E / R; ++u) {
            SLMGS_UNROLL
            for (int m = 0; m < R; ++m) {
                const int n = F::first_index(L.lt + F::TPL * u, m);
                const int sc = slm_col(a, n);
                cf val = cmake(0.f, 0.f);
                if (L.active && sc >= 0) {
                    float ph = a.phase[L.pbase + sc];
                    if (a.prop) ph += __ldg(a.prop + (long long)L.sr * a.w + sc);
                    float sn, cs;
                    sincosf(ph, &sn, &cs);
                    const float am = amp_at(a, id, L, sc);
                    val = cmake(am * cs, am * sn);
                }
                st.v[u * R + m] = val;
            }
        }
    }
    // phase-only projection: _hologram.py:1026-1036 followed by :1000-1011 of the next iteration.
    // amp * exp(i * arctan2(im, re)) == amp * z / |z|  (z == 0 -> phase 0 -> amp)
    // REBUILD: also build the next near field (ROW_FUSED); WRITE: store the phase (last iteration / ROW_LAST).
    // The hot variant (REBUILD, !WRITE) is branch-free: select instead of branch, no arctan2 in the code.
    template <bool REBUILD, bool WRITE>
    static SLMGS_DEVICE void project(State& st, const Args& a, const ThreadId& id, const Loc& L) {
        constexpr int R = F::R0;
        if constexpr (DENSE && REBUILD && !WRITE) {
            const float am = a.amp_scalar;
            SLMGS_UNROLL
            for (int i = 0; i < E; ++i) {
                const cf z = st.v[i];
                const float m2 = z.x * z.x + z.y * z.y;
                const bool nz = m2 > 1.0e-37f;  // flush-to-zero MUFU rsqrt; |z|^2 below 1e-37 counts as zero (phase 0)
                const float r = fast_rsqrt(m2) * am;
                st.v[i] = nz ? cscale(z, r) : cmake(am, 0.f);
            }
            return;
        }
        const bool full = a.w == a.W;
        SLMGS_UNROLL
        for (int u = 0; u < E / R; ++u) {
            SLMGS_UNROLL
            for (int m = 0; m < R; ++m) {
                const int n = F::first_index(L.lt + F::TPL * u, m);
                const unsigned scu = (unsigned)(((n + (a.W >> 1)) & (a.W - 1)) - a.i2);
                const bool inside = L.active && (full || scu < (unsigned)a.w);
                const int sc = inside ? (int)scu : 0;
                const cf z = st.v[u * R + m];
                if (WRITE && !REBUILD && a.mp_sum) {
                    if (inside) {  // child of a MultiplaneHologram: complex sum over the children, no extraction
                        cf val = cscale(z, a.scale * a.mp_weight);
                        if (a.prop) {
                            float sn, cs;
                            sincosf(__ldg(a.prop + (long long)L.sr * a.w + sc), &sn, &cs);
                            val = cmulc(val, cmake(cs, sn));
                        }
                        cf* dst = a.mp_sum + (long long)id.by * a.phase_bs + (long long)L.sr * a.w + sc;
                        *dst = a.mp_first ? val : cadd(*dst, val);
                    }
                } else if (WRITE) {
                    if (inside) {
                        float ph = atan2f(z.y, z.x);
                        if (a.prop) ph -= __ldg(a.prop + (long long)L.sr * a.w + sc);
                        a.phase[L.pbase + sc] = ph;
                    }
                }
                cf val = cmake(0.f, 0.f);
                if (REBUILD) {
                    const float m2 = z.x * z.x + z.y * z.y;
                    const float am = a.amp ? (L.active ? __ldg(a.amp + (long long)id.by * a.amp_bs + (long long)L.sr * a.w + sc) : 0.f)
                                           : a.amp_scalar;
                    // flush-to-zero MUFU rsqrt (no denormal rescaling sequence); |z|^2 below 1e-37 counts as zero
                    const bool nz = m2 > 1.0e-37f;
                    const float r = nz ? fast_rsqrt(m2) * am : 0.f;
                    val = nz ? cmake(z.x * r, z.y * r) : cmake(am, 0.f);
                    if (!inside) val = cmake(0.f, 0.f);
                }
                st.v[u * R + m] = val;
            }
        }
    }

    template <int P> static SLMGS_DEVICE void phase(State& st, const Args& a, cf* smem, const ThreadId& id) {
        const Loc L = locate(a, smem, id);
        if constexpr (P == 0) {
            if (a.zero_acc && id.bx == 0 && id.tid == 0) a.zero_acc[(long long)id.by * a.zero_bs] = 0.0;
            if (a.zero_acc2 && id.bx == 0 && id.tid == 0) a.zero_acc2[(long long)id.by * a.zero_bs] = 0.0;
            if (a.win_dst && id.bx == 0 && id.tid == 0)
                a.win_dst[id.by] = (float)(1.0 / sqrt(a.win_src[(long long)id.by * a.zero_bs]));
            if (MODE != ROW_FIRST && LI == 2 && N >= 8192 && a.pf_dist > 0 && !a.colflag) {
                // 8192-point rows, row-pair interleaved field: the block's two lines are one contiguous 2 W block; pull the
                // row pair this SM runs next into L2 (one 512-thread block per SM: nothing else hides the head of a tile;
                // fused row kernel of configs[4] 426 -> 406 us)
                const int nsr = (id.bx + a.pf_dist) * 2;
                if (nsr < a.h) {
                    const int nfr = (nsr + a.i0 + (a.H >> 1)) & (a.H - 1);
                    const char* base = reinterpret_cast<const char*>(a.fld + (long long)id.by * a.fld_bs + (long long)(nfr >> 1) * a.W * 2);
                    for (int i = id.tid; i < (int)(2 * N * sizeof(cf) / 128); i += id.nthreads) prefetch_l2(base + (size_t)i * 128);
                }
            }
            if (MODE != ROW_FIRST && LI == 1 && a.pf_dist > 0) {
                // pull the rows of the tile this SM will run next into L2 while this tile computes
                const int lines = id.nthreads / F::TPL;
                const int nsr = (id.bx + a.pf_dist) * lines + id.tid / F::TPL;
                if (nsr < a.h) {
                    const int nfr = (nsr + a.i0 + (a.H >> 1)) & (a.H - 1);
                    const char* base = reinterpret_cast<const char*>(a.fld + (long long)id.by * a.fld_bs + (long long)nfr * a.W);
                    for (int i = id.tid % F::TPL; i < (int)(N * sizeof(cf) / 128); i += F::TPL) prefetch_l2(base + (size_t)i * 128);
                }
            }
        }
        if constexpr (MODE == ROW_FIRST) {
            if constexpr (P == 0) build_nearfield(st, a, id, L);
            F::template fwd_stage<P>(st.v, L.lt, a.twA, a.twB, L.s, LI);
            if constexpr (P == NS - 1) store_spectrum(st, a, L);
        } else if constexpr (MODE == ROW_LAST) {
            if constexpr (P == 0) load_spectrum(st, a, L);
            F::template inv_stage<NS - 1 - P>(st.v, L.lt, a.twA, a.twB, L.s, LI);
            if constexpr (P == NS - 1) project<false, true>(st, a, id, L);
        } else {
            if constexpr (P == 0) load_spectrum(st, a, L);
            if constexpr (P < NS - 1) {
                F::template inv_stage<NS - 1 - P>(st.v, L.lt, a.twA, a.twB, L.s, LI);
            } else if constexpr (P == NS - 1) {
                F::template inv_stage<0>(st.v, L.lt, a.twA, a.twB, L.s, LI);
                project<true, STORE>(st, a, id, L);
                F::template fwd_stage<0>(st.v, L.lt, a.twA, a.twB, L.s, LI);
            } else {
                F::template fwd_stage<P - (NS - 1)>(st.v, L.lt, a.twA, a.twB, L.s, LI);
            }
            if constexpr (P == NPHASE - 1) store_spectrum(st, a, L);
        }
    }

};

// ==========================================================================================
// Column kernel
// ==========================================================================================
// Layout of the far-field-shaped images (target, weights, phase_ff, amp_ff, farfield): the column
// kernel of a context always works on tiles of C adjacent columns, so these private device buffers
// are stored tile-major, [W/C tiles][H rows][C columns] in rolled coordinates: a tile is one
// contiguous block and a warp's access (8 rows x 4 columns, or 2 x 16, ...) is one fully used
// 128-byte line instead of 8 half-used sectors.  Only uploads/downloads and the spot gather need
// the mapping; element-wise kernels are layout agnostic.
SLMGS_HD long long image_index(int ry, int rx, int H, int C) {
    return (long long)(rx / C) * H * C + (long long)ry * C + (rx % C);
}

struct ColArgs {
    cf* fld;  // [B][H][W] rolled; only SLM rows are read / written
    long long fld_bs;
    const cf* twA;  // twiddle tables for N = H
    const cf* twB;
    const cf* tw2A;  // twiddle tables for N = H / 2 (ColKernelT8: 8192-point columns as two interleaved 4096-point lines), or nullptr
    const cf* tw2B;
    float* weights;       // [B][H][W] rolled
    const float* target;  // [B][H][W] rolled (or shared: target_bs == 0)
    float* phase_ff;      // [B][H][W] rolled
    float* amp_ff;        // [B][H][W] rolled
    cf* farfield;         // [B][H][W] rolled, ortho-scaled
    long long img_bs;     // batch stride of weights / phase_ff / amp_ff / farfield
    long long target_bs;
    double* acc;          // [B][acc_bs] accumulators
    int acc_bs;
    int w_in_slot;        // accumulator slot holding sum(w^2) of a pending normalisation, or -1
    const float* win_f;   // [B] 1/sqrt of that slot, prepared by the preceding row kernel (fused loop), or nullptr
    int w_out_slot;       // accumulator slot receiving sum(w_new^2), or -1
    int H, W, h, i0;
    float scale;          // 1/sqrt(H W): ortho normalisation of the forward transform
    WgsParams wgs;
    int wgs_update;       // apply the WGS weight update this iteration (_hologram.py:1552)
    int phase_mode;       // PHASE_COMPUTE / PHASE_COMPUTE_STORE / PHASE_STORED
    int mraf;             // target carries NaN noise region (_hologram.py:1495-1548)
    int mraf_has_factor;
    float mraf_factor;
    cf* zero_w;           // MRAF zero-region accumulator image ([B][H][W], tile-major), or nullptr (_hologram.py:1613-1616)
    float zero_factor;
    int store_ampff, store_phaseff, store_farfield;  // COL_FWD outputs
    int ratio_slot;       // COL_FWD: accumulate sum(wgs_ratio(|F| / ||F||, target)) here (WGS-Nogrette's mean, :1851-1852), or -1
                          // COL_FUSED: Nogrette mean = (acc[ratio_slot] + ratio_extra) * inv_npix, or -1
    int wsq_slot;         // MRAF + WGS in the fused loop (the noise region fixes the scale of the weights, so their
                          // normalisation cannot be deferred, :1877 before :1643-1653).  COL_FWD: accumulate
                          // sum(w_new^2) of the updated (not stored) weights here; COL_FUSED: scale the updated
                          // weights by 1 / sqrt(acc[wsq_slot]) at once.  -1 = off
    double ratio_extra;   // added to the sum (pixels that were not visited and whose ratio is known to be 1), normally 0
    double inv_npix;      // 1 / (H W)
    int pdl;              // launch with programmatic dependent launch
    int pf_dist;          // L2 prefetch distance in tiles (blocks resident on the GPU), 0 = off
    const int* tiles;     // sparse far field: blockIdx.x -> column tile (only tiles whose constrained far field can be
                          // non-zero are launched), or nullptr = every tile, in order.  [B][tiles_bs]
    const int* tile_count;  // [B] active tiles of each hologram: blocks with blockIdx.x >= count exit at once
    int tiles_bs;
    // persistent column kernel (ColKernelP): TMA staging of the field tiles
    int pairs;            // fld layout: 0 row-major, 1 row-pair interleaved (see RowArgs)
    const void* tmap;     // device copy of the CUtensorMap over fld: dims {W, H, B}, box {C, N/R0 rows, 1}
    int n_boxes;          // zero-padded field: boxes (groups of N/R0 rows) that hold SLM rows ...
    signed char box_slot[32];  // ... and the staging slot of each box (-1 = no SLM row inside: reads as zero)
    // team column kernel (slmgs_teams.h): TMA boxes of tb_pairs row pairs; tb_n boxes per tile, the first tb_lo of them
    // from row pair 0 upwards, the others from row pair tb_hi0 upwards
    int tb_pairs, tb_n, tb_lo, tb_hi0;
};

// CT: columns per tile known at compile time (block of MAXT threads), 0 = derived from blockDim at run time
// (small problems).  With CT fixed, every shared-memory access is base register + immediate offset.
// DENSE: the SLM rows cover the whole column (h == H) -- no row-range tests in the loads / stores
template <int N, int MODE, int VAR = 0, int CT = 0, bool DENSE = false> struct ColKernel {
    typedef Fft<N> F;
    static constexpr int TRACE_CLASS = 10 + MODE;  // diagnostic builds (-DSLMGS_TRACE)
    typedef ColArgs Args;
    static constexpr int E = F::E, NS = F::NS;
    static constexpr int MAXT = 16384 / E;
    static constexpr int NPHASE = (MODE == COL_FUSED) ? 2 * NS - 1 : NS;
    struct State {
        cf v[E];
    };

    static size_t smem_bytes(int nthreads) { return NS > 1 ? (size_t)(nthreads / F::TPL) * F::PADN * sizeof(cf) : 0; }

#ifndef SLMGS_EMULATE
    static SLMGS_DEVICE void barrier(const ThreadId&) { __syncthreads(); }
#endif
    // sparse far field, batch: the grid is sized for the hologram with the most active tiles
    static SLMGS_DEVICE bool skip(const Args& a, const ThreadId& id) {
        return a.tiles != nullptr && id.bx >= __ldg(a.tile_count + id.by);
    }

    struct Loc {
        int lt, col, C;  // thread-in-line, column inside the tile, columns per tile
        int gc;          // global column
        cf* s;
        long long fbase, ibase, tbase;
    };
    static SLMGS_DEVICE Loc locate(const Args& a, cf* smem, const ThreadId& id) { return locate_q(a, smem, id, id.bx); }
    // q: position of the block's tile in the launch order (the block index, or further tiles of a persistent block)
    static SLMGS_DEVICE Loc locate_q(const Args& a, cf* smem, const ThreadId& id, int q) {
        Loc L;
        L.C = CT > 0 ? CT : id.nthreads / F::TPL;  // a power of two
        L.col = id.tid & (L.C - 1);
        L.lt = id.tid >> ilog2(L.C);
        const int tile = a.tiles ? __ldg(a.tiles + (long long)id.by * a.tiles_bs + q) : q;
        L.gc = tile * L.C + L.col;
        L.s = smem + L.col;
        L.fbase = (long long)id.by * a.fld_bs + L.gc;
        // far-field-shaped images are tile-major: [W/C tiles][H rows][C columns] (image_index below)
        L.ibase = (long long)id.by * a.img_bs + (long long)tile * a.H * L.C + L.col;
        L.tbase = (long long)id.by * a.target_bs + (long long)tile * a.H * L.C + L.col;
        return L;
    }
    // Offset of element (row b, this thread's column) of `fld`; rows b + (N/R0) m follow at a constant step of
    // (N/R0) W elements in both layouts (N/R0 is even).
    static SLMGS_DEVICE long long row_base(const Args& a, const Loc& L, int b) {
        const long long hb = L.fbase - L.gc;  // hologram base
        if (a.pairs) return hb + ((long long)(b >> 1) * a.W + L.gc) * 2 + (b & 1);
        return hb + (long long)b * a.W + L.gc;
    }
    // Rows of `fld` that hold the SLM (rolled index n): ((n + H/2) mod H) - i0 in [0, h).  Row n of butterfly
    // b is b + (N/R0) m, so the row pointer advances by a constant and the test is one unsigned compare.
    static SLMGS_DEVICE void load_rows(State& st, const Args& a, const Loc& L) {
        constexpr int R = F::R0;
        const bool full = DENSE || a.h == a.H;
        const long long step = (long long)(N / R) * a.W;
        SLMGS_UNROLL
        for (int u = 0; u < E / R; ++u) {
            const int b = L.lt + F::TPL * u;
            const cf* p = a.fld + row_base(a, L, b);
            const unsigned r0 = (unsigned)(b + (a.H >> 1) - a.i0);
            SLMGS_UNROLL
            for (int m = 0; m < R; ++m) {
                const unsigned sr = ((r0 + (unsigned)((N / R) * m) + (unsigned)a.i0) & (unsigned)(a.H - 1)) - (unsigned)a.i0;
                st.v[u * R + m] = (full || sr < (unsigned)a.h) ? ld_stream(p) : cmake(0.f, 0.f);
                p += step;
            }
        }
    }
    static SLMGS_DEVICE void store_rows(State& st, const Args& a, const Loc& L) {
        constexpr int R = F::R0;
        const bool full = DENSE || a.h == a.H;
        const long long step = (long long)(N / R) * a.W;
        SLMGS_UNROLL
        for (int u = 0; u < E / R; ++u) {
            const int b = L.lt + F::TPL * u;
            cf* p = a.fld + row_base(a, L, b);
            const unsigned r0 = (unsigned)(b + (a.H >> 1) - a.i0);
            SLMGS_UNROLL
            for (int m = 0; m < R; ++m) {
                const unsigned sr = ((r0 + (unsigned)((N / R) * m) + (unsigned)a.i0) & (unsigned)(a.H - 1)) - (unsigned)a.i0;
                if (full || sr < (unsigned)a.h) *p = st.v[u * R + m];
                p += step;
            }
        }
    }
    static SLMGS_DEVICE void load_farfield(State& st, const Args& a, const Loc& L) {
        constexpr int R = F::last_radix();
        SLMGS_UNROLL
        for (int u = 0; u < E / R; ++u) {
            SLMGS_UNROLL
            for (int m = 0; m < R; ++m) {
                const int k = F::last_index(L.lt + F::TPL * u, m);
                st.v[u * R + m] = ld_stream(a.farfield + L.ibase + (long long)k * L.C);
            }
        }
    }
    // COL_FWD epilogue: _hologram.py:951-953 (amp_ff), :934-949 (phase_ff), farfield (ortho-scaled)
    static SLMGS_DEVICE void store_farfield(State& st, const Args& a, const ThreadId& id, const Loc& L) {
        constexpr int R = F::last_radix();
        SLMGS_UNROLL
        for (int u = 0; u < E / R; ++u) {
            SLMGS_UNROLL
            for (int m = 0; m < R; ++m) {
                const long long off = (long long)F::last_index(L.lt + F::TPL * u, m) * L.C;
                const cf z = cscale(st.v[u * R + m], a.scale);
                if (a.store_farfield) a.farfield[L.ibase + off] = z;
                if (a.store_ampff) a.amp_ff[L.ibase + off] = sqrtf(z.x * z.x + z.y * z.y);
                if (a.store_phaseff) a.phase_ff[L.ibase + off] = atan2f(z.y, z.x);
            }
        }
        if (a.ratio_slot >= 0) {  // every thread of the block gets here (accum_add needs full warps)
            WgsParams q = a.wgs;
            double rs = 0.0;
            SLMGS_UNROLL
            for (int u = 0; u < E / R; ++u) {
                SLMGS_UNROLL
                for (int m = 0; m < R; ++m) {
                    const long long off = (long long)F::last_index(L.lt + F::TPL * u, m) * L.C;
                    const cf z = cscale(st.v[u * R + m], a.scale);
                    rs += (double)wgs_ratio(sqrtf(z.x * z.x + z.y * z.y), ld_stream(a.target + L.tbase + off), q);
                }
            }
            accum_add(a.acc + (long long)id.by * a.acc_bs + a.ratio_slot, rs);
        }
        if (a.wsq_slot >= 0) {  // pre-pass of MRAF + WGS: the update exactly as the fused kernel will apply it
            double* acc = a.acc + (long long)id.by * a.acc_bs;
            float win = 1.0f;
            if (a.w_in_slot >= 0) win = a.win_f ? __ldg(a.win_f + id.by) : (float)(1.0 / sqrt(acc[a.w_in_slot]));
            double ws = 0.0;
            SLMGS_UNROLL
            for (int u = 0; u < E / R; ++u) {
                SLMGS_UNROLL
                for (int m = 0; m < R; ++m) {
                    const long long off = (long long)F::last_index(L.lt + F::TPL * u, m) * L.C;
                    const cf z = st.v[u * R + m];
                    const float m2 = z.x * z.x + z.y * z.y;
                    const float famp = (m2 > 1.0e-37f ? m2 * fast_rsqrt(m2) : 0.f) * a.scale;
                    const float w = wgs_apply(ld_stream(a.weights + L.ibase + off) * win,
                                              wgs_multiplier_fast(famp, ld_stream(a.target + L.tbase + off), a.wgs));
                    ws += (double)(w * w);
                }
            }
            accum_add(acc + a.wsq_slot, ws);
        }
    }

    // Far-field constraint (+ fused weight update): _hologram.py:1550-1653.
    // SCALED: v already carries the ortho scale (COL_INV, accurate math) or not (COL_FUSED, fast math).
    // VAR specialises the fused kernel at compile time (fewer branches, fewer live registers):
    //   VAR_GENERAL everything decided at run time; VAR_GS no update, phase from the field, no MRAF;
    //   VAR_POW Leonardo/Kim update, phase from the field; VAR_POW_STORED the same with the stored phase.
    // The loads of weights / target / phase_ff are software pipelined two elements ahead of their use:
    // the weights store of element i may alias later loads as far as the compiler knows, so without
    // the explicit prefetch every element would pay a full memory round trip.
#ifndef SLMGS_IMG_PF
#define SLMGS_IMG_PF 5
#endif
    // L1 prefetch distance of the fused constraint's image loads, in elements.  Measured on B200:
    // 5 gives +3 % on dense 4096^2 WGS-Kim and +1.5 % on the bench workload; small problems (launch bound) lose a
    // few percent to the extra instructions, so it is only compiled in for long columns.
    static constexpr int IMG_PF = (N >= 2048) ? SLMGS_IMG_PF : 0;
#ifndef SLMGS_STAGE
#define SLMGS_STAGE 0
#endif
    // EXPERIMENT, compiled out (measured slower on B200): staging of weights / target through the
    // thread's PRIVATE exchange slots.  The 16 shared-memory slots a thread
    // reads in the last forward stage are the ones it writes in the first inverse stage, and nobody else touches
    // them in between: that window is exactly the fused constraint.  So right after its reads the thread fires
    // 4-byte cp.async copies of its weights (low word of slot i) and target (high word) values; they land while the
    // radix-16 butterflies run, hold no registers and need no extra shared memory or barrier.  (The register pipeline
    // that remains for phase_ff covers two elements of latency; the ncu source view put 20 % of this kernel's
    // warp-time in waits on these loads.)  Result: -7 % on the bench workload and on dense WGS-Kim: the 32 extra
    // LDGSTS + 16 LDS per thread land on the shared-memory / L1 data pipe, which is the busier resource.
    static constexpr bool STAGE = SLMGS_STAGE != 0 && MODE == COL_FUSED && NS > 1 && N >= 2048;
    static SLMGS_DEVICE void stage_images(const Args& a, const Loc& L) {
        if constexpr (STAGE) {
            constexpr int R = F::last_radix();
            const bool update = VAR == VAR_GENERAL ? (a.wgs_update != 0) : (VAR == VAR_POW || VAR == VAR_POW_STORED);
            const bool need_t = update || (VAR == VAR_GENERAL && a.mraf != 0);
            SLMGS_UNROLL
            for (int i = 0; i < E; ++i) {
                const int off = F::last_index(L.lt + F::TPL * (i / R), i % R) * L.C;
                float* slot = reinterpret_cast<float*>(L.s + F::last_slot(L.lt, i) * L.C);
                cp_async_f32(slot, a.weights + L.ibase + off);
                if (need_t) cp_async_f32(slot + 1, a.target + L.tbase + off);
            }
        }
    }

    // COL_FUSED: pull the image values of the first SLMGS_IMG_PF elements into L1 before the last forward stage,
    // whose butterflies then hide the DRAM latency (the constraint continues the prefetch at the same distance)
    static SLMGS_DEVICE void prefetch_images_head(const Args& a, const Loc& L) {
        constexpr int R = F::last_radix();
        constexpr int PF = IMG_PF;
        if constexpr (PF > 0) {
            const bool update = VAR == VAR_GENERAL ? (a.wgs_update != 0) : (VAR == VAR_POW || VAR == VAR_POW_STORED);
            const bool stored = VAR == VAR_GENERAL ? (a.phase_mode == PHASE_STORED) : (VAR == VAR_POW_STORED);
            const bool need_t = update || (VAR == VAR_GENERAL && a.mraf != 0);
            SLMGS_UNROLL
            for (int e = 0; e < PF && e < E; ++e) {
                const int off = F::last_index(L.lt + F::TPL * (e / R), e % R) * L.C;
                if (!STAGE) prefetch_l1(a.weights + L.ibase + off);
                if (!STAGE && need_t) prefetch_l1(a.target + L.tbase + off);
                if (stored) prefetch_l1(a.phase_ff + L.ibase + off);
            }
        }
    }

    template <bool SCALED> static SLMGS_DEVICE void constrain(State& st, const Args& a, const ThreadId& id, const Loc& L) {
        constexpr int R = F::last_radix();
        constexpr bool GEN = SCALED || VAR == VAR_GENERAL;
        const bool update = GEN ? (a.wgs_update != 0) : (VAR == VAR_POW || VAR == VAR_POW_STORED);
        const bool stored = GEN ? (a.phase_mode == PHASE_STORED) : (VAR == VAR_POW_STORED);
        const bool mraf = GEN ? (a.mraf != 0) : false;
        const bool need_t = update || mraf;
        double* acc = a.acc + (long long)id.by * a.acc_bs;
        float win = 1.0f;
        if (a.w_in_slot >= 0) win = (!SCALED && a.win_f) ? __ldg(a.win_f + id.by) : (float)(1.0 / sqrt(acc[a.w_in_slot]));
        const float fscale = SCALED ? 1.0f : a.scale;
        const float lg2s = fast_lg2(fscale * a.wgs.inv_fnorm);
        WgsParams wq_params = a.wgs;
        float wnorm = 1.0f;
        if (GEN && !SCALED && a.wsq_slot >= 0) wnorm = (float)(1.0 / sqrt(acc[a.wsq_slot]));
        if (GEN && !SCALED && a.ratio_slot >= 0) {  // WGS-Nogrette in the fused loop: mean of the ratio from the pre-pass
            // sparse far field: the tiles that were not launched hold target == 0 only, where the ratio is exactly 1 (:1841)
            const double skipped = a.tiles ? (double)a.H * (double)(a.W - __ldg(a.tile_count + id.by) * L.C) : 0.0;
            wq_params.neg_inv_mean = -(1.0f / (float)((acc[a.ratio_slot] + a.ratio_extra + skipped) * a.inv_npix));
        }
        float wsum = 0.0f;
        const float* SLMGS_RESTRICT wp = a.weights + L.ibase;
        const float* SLMGS_RESTRICT tp = a.target + L.tbase;
        const float* SLMGS_RESTRICT pp = a.phase_ff + L.ibase;
        float wq[E], tq[E], pq[E];
        constexpr int AHEAD = 2;
        if (STAGE && !SCALED) cp_async_wait_all();
        // L1 prefetch distance (elements) of the fused kernel's image loads: the register pipeline above only
        // covers AHEAD elements, a fraction of the DRAM latency; prefetch.global.L1 needs no registers
        constexpr int PF = SCALED ? 0 : IMG_PF;
        SLMGS_UNROLL
        for (int e = 0; e < E + AHEAD; ++e) {
            if (PF > 0 && e + PF - AHEAD < E && e >= AHEAD) {
                const int off = F::last_index(L.lt + F::TPL * ((e + PF - AHEAD) / R), (e + PF - AHEAD) % R) * L.C;
                if (!(STAGE && !SCALED)) prefetch_l1(wp + off);
                if (!(STAGE && !SCALED) && need_t) prefetch_l1(tp + off);
                if (stored) prefetch_l1(pp + off);
            }
            if (e < E) {  // issue the loads of element e
                const int off = F::last_index(L.lt + F::TPL * (e / R), e % R) * L.C;
                if (STAGE && !SCALED) {  // weights / target were staged in this thread's private exchange slots
                    const cf wt = L.s[F::last_slot(L.lt, e) * L.C];
                    wq[e] = wt.x;
                    tq[e] = need_t ? wt.y : 1.0f;
                    pq[e] = stored ? (PF > 0 ? ld_cached(pp + off) : ld_stream(pp + off)) : 0.0f;
                } else if (PF > 0) {
                    wq[e] = ld_cached(wp + off);
                    tq[e] = need_t ? ld_cached(tp + off) : 1.0f;
                    pq[e] = stored ? ld_cached(pp + off) : 0.0f;
                } else {
                    wq[e] = ld_stream(wp + off);
                    tq[e] = need_t ? ld_stream(tp + off) : 1.0f;
                    pq[e] = stored ? ld_stream(pp + off) : 0.0f;
                }
            }
            if (e >= AHEAD) {  // consume element i
                const int i = e - AHEAD;
                const int off = F::last_index(L.lt + F::TPL * (i / R), i % R) * L.C;
                const cf z = st.v[i];
                const float m2 = z.x * z.x + z.y * z.y;
                constexpr bool POWV = !SCALED && (VAR == VAR_POW || VAR == VAR_POW_STORED);
                // the stored-phase power-law variant needs |F| only inside the logarithm: no reciprocal square root
                float rinv = 0.f;
                // (fused path: flush-to-zero MUFU; |F|^2 below 1e-37 is treated like an exact zero, phase 0)
                const bool nz = SCALED ? (m2 > 0.f) : (m2 > 1.0e-37f);
                if (!(POWV && VAR == VAR_POW_STORED)) rinv = nz ? (SCALED ? rsqrtf(m2) : fast_rsqrt(m2)) : 0.f;
                float w = wq[i] * win;
                const float t = tq[i];
                if (update) {
                    float fc;
                    if (POWV) {
                        fc = wgs_multiplier_pow_log(m2, t, lg2s, a.wgs.p);
                    } else {
                        const float famp = m2 * rinv * fscale;  // |F| (ortho-scaled)
                        if (SCALED) fc = wgs_multiplier(famp, t, a.wgs);
                        else fc = wgs_multiplier_fast(famp, t, wq_params);
                    }
                    w = wgs_apply(w, fc);
                    if (GEN && !SCALED) w *= wnorm;
                    a.weights[L.ibase + off] = w;
                    wsum += w * w;
                }
                const bool zero_region = mraf && t == 0.0f;
                cf unit;
                if (stored) {
                    float sn, cs;
                    if (SCALED) sincosf(pq[i], &sn, &cs);
                    else fast_sincos(pq[i], &sn, &cs);
                    unit = cmake(cs, sn);
                } else {
                    unit = nz ? cmake(z.x * rinv, z.y * rinv) : cmake(1.0f, 0.f);
                    if (zero_region) unit = cmake(1.0f, 0.f);  // angle taken after farfield[zero] = 0 (:1613-1622)
                    // the fused kernel never stores: the host runs a COL_FWD pass for that iteration instead
                    if (SCALED && a.phase_mode == PHASE_COMPUTE_STORE) a.phase_ff[L.ibase + off] = atan2f(unit.y, unit.x);
                }
                cf g = cscale(unit, w);
                if (mraf) {
                    if (zero_region) {
                        g = cmake(0.f, 0.f);
                        if (a.zero_w) {  // zero_weights -= zero_factor * |F| * F; farfield[zero] = zero_weights
                            const cf fz = cscale(z, fscale);
                            const float k = a.zero_factor * sqrtf(fz.x * fz.x + fz.y * fz.y);
                            cf zw = a.zero_w[L.ibase + off];
                            zw = cmake(zw.x - k * fz.x, zw.y - k * fz.y);
                            a.zero_w[L.ibase + off] = zw;
                            g = zw;
                        }
                    } else if (t != t) {  // noise region keeps the (scaled) field (:1643-1653)
                        const float q = a.mraf_has_factor ? fscale * a.mraf_factor : fscale;
                        g = cscale(z, q);
                    }
                }
                st.v[i] = g;
            }
        }
        if (update && a.w_out_slot >= 0) accum_add(acc + a.w_out_slot, (double)wsum);
    }

    template <int P> static SLMGS_DEVICE void phase(State& st, const Args& a, cf* smem, const ThreadId& id) {
        const Loc L = locate(a, smem, id);
        if constexpr (MODE == COL_FWD) {
            if constexpr (P == 0) load_rows(st, a, L);
            F::template fwd_stage<P>(st.v, L.lt, a.twA, a.twB, L.s, L.C);
            if constexpr (P == NS - 1) store_farfield(st, a, id, L);
        } else if constexpr (MODE == COL_INV) {
            if constexpr (P == 0) {
                load_farfield(st, a, L);
                constrain<true>(st, a, id, L);
            }
            F::template inv_stage<NS - 1 - P>(st.v, L.lt, a.twA, a.twB, L.s, L.C);
            if constexpr (P == NS - 1) store_rows(st, a, L);
        } else {
            if constexpr (P == 0) {
                load_rows(st, a, L);
                // (8192-point columns, row-pair interleaved field: one prefetch.global.L2 per 32-byte sector of the next tile
                // was measured SLOWER -- fused column kernel 531 -> 562 us -- the head of a tile is not what that kernel waits for)
                if (a.pf_dist > 0 && a.h == a.H && !a.tiles && !a.pairs) {
                    // dense field: pull the rows of the tile group this SM will run next into L2 (a 128-byte line
                    // holds 16 columns = several tiles, so one tile of each line-sharing group issues the prefetch)
                    const int per_line = 16 / L.C > 0 ? 16 / L.C : 1;
                    const int nb = id.bx + a.pf_dist;
                    if ((id.bx % per_line) == 0 && nb < id.gx && L.col == 0) {
                        const cf* base = a.fld + (long long)id.by * a.fld_bs + (long long)nb * L.C;
                        SLMGS_UNROLL
                        for (int m = 0; m < E; ++m) prefetch_l2(base + (long long)(L.lt + F::TPL * m) * a.W);
                    }
                }
            }
            if constexpr (P < NS - 1) {
                F::template fwd_stage<P>(st.v, L.lt, a.twA, a.twB, L.s, L.C);
            } else if constexpr (P == NS - 1) {
                prefetch_images_head(a, L);
                if constexpr (STAGE) {
                    F::fwd_last_load(st.v, L.lt, L.s, L.C);
                    stage_images(a, L);
                    F::fwd_last_compute(st.v);
                } else {
                    F::template fwd_stage<NS - 1>(st.v, L.lt, a.twA, a.twB, L.s, L.C);
                }
                constrain<false>(st, a, id, L);
                F::template inv_stage<NS - 1>(st.v, L.lt, a.twA, a.twB, L.s, L.C);
            } else {
                F::template inv_stage<2 * NS - 2 - P>(st.v, L.lt, a.twA, a.twB, L.s, L.C);
            }
            if constexpr (P == NPHASE - 1) store_rows(st, a, L);
        }
    }

};

// ==========================================================================================
// Persistent fused column kernel with TMA-staged tiles
// ==========================================================================================
// Same arithmetic as ColKernel<N, COL_FUSED, VAR, CT>, different data movement.  One block per SM walks over
// tiles q = blockIdx.x + it * gridDim.x.  The shared memory that the exchange buffer leaves free holds NPRE
// "boxes" of the NEXT tile -- box m = rows [m N/R0, (m+1) N/R0) x CT columns, exactly the rows that form element m
// of the first-stage butterflies -- fetched by TMA (cp.async.bulk.tensor, completion on an mbarrier) while the
// current tile runs its six radix stages.  A thread then reads its first-stage elements from the staging buffer
// (conflict-free LDS at immediate offsets) instead of issuing strided global loads and waiting for them:
//   * no exposed DRAM latency at the head of a tile for the staged boxes (dense 4096^2: 11 of 16; the last five
//     do not fit next to the 137 KB exchange buffer and stay direct loads, issued before the mbarrier wait);
//   * an LSU request of the strided pattern costs one L1 wavefront per row (8 per warp) -- the staged LDS costs 2;
//   * no row-range predicates or 64-bit address arithmetic for the staged elements.
// Zero-padded fields (DENSE = false): rows outside the SLM are never written and stay zero in `fld`, so whole
// boxes are fetched; boxes without any SLM row are not fetched at all and read as zero.  The host only selects
// this kernel when all boxes that hold SLM rows fit the staging buffer.
template <int N, int VAR, int CT, bool DENSE> struct ColKernelP : ColKernel<N, COL_FUSED, VAR, CT> {
    typedef ColKernel<N, COL_FUSED, VAR, CT> Base;
    typedef typename Base::F F;
    static constexpr int TRACE_CLASS = 30;
    typedef ColArgs Args;
    typedef typename Base::State State;
    typedef typename Base::Loc Loc;
    static constexpr int E = F::E, NS = F::NS, R0 = F::R0;
    static constexpr int MAXT = Base::MAXT;
    static constexpr int NPHASE = Base::NPHASE;
    static_assert(CT > 0 && NS == 3, "persistent column kernel: compile-time tile width, three radix stages");
    static constexpr int BOX_ROWS = N / R0;
    static constexpr int BOX_ELEMS = BOX_ROWS * CT;
    static constexpr size_t BOX_BYTES = (size_t)BOX_ELEMS * sizeof(cf);
    static constexpr size_t EXCH_BYTES = (((size_t)F::PADN * CT * sizeof(cf)) + 127) / 128 * 128;
    static constexpr size_t SMEM_MAX = 227 * 1024;
    static constexpr int NPRE_FIT = (int)((SMEM_MAX - 16 - EXCH_BYTES) / BOX_BYTES);
    static constexpr int NPRE = NPRE_FIT < R0 ? NPRE_FIT : R0;
    static_assert(NPRE >= 1, "no room for a staging box");
    static_assert(R0 <= 32, "box_slot table");

    static size_t smem_bytes(int) { return EXCH_BYTES + (size_t)NPRE * BOX_BYTES + 16; }
    static SLMGS_DEVICE cf* stage_of(cf* smem) { return reinterpret_cast<cf*>(reinterpret_cast<char*>(smem) + EXCH_BYTES); }
    static SLMGS_DEVICE unsigned long long* bar_of(cf* smem) {
        return reinterpret_cast<unsigned long long*>(stage_of(smem) + (size_t)NPRE * BOX_ELEMS);
    }

    static SLMGS_DEVICE int tile_count(const Args& a, const ThreadId& id) {
        return a.tiles ? __ldg(a.tile_count + id.by) : a.W / CT;
    }
    static SLMGS_DEVICE bool skip(const Args& a, const ThreadId& id) { return id.bx >= tile_count(a, id); }
    static SLMGS_DEVICE int iterations(const Args& a, const ThreadId& id) {
        return (tile_count(a, id) - id.bx + id.gx - 1) / id.gx;
    }
    static SLMGS_DEVICE void init(const Args&, cf* smem, const ThreadId& id) {
        if (id.tid == 0) mbar_init(bar_of(smem), 1);
    }

    // one thread: fetch the staged boxes of tile position q into the staging buffer
    static SLMGS_DEVICE void prefetch(const Args& a, cf* smem, const ThreadId& id, int q) {
        const int tile = a.tiles ? __ldg(a.tiles + (long long)id.by * a.tiles_bs + q) : q;
        cf* stage = stage_of(smem);
#if defined(__CUDACC__) && !defined(SLMGS_EMULATE)
        unsigned long long* bar = bar_of(smem);
        if (DENSE) {
            mbar_expect_tx(bar, (unsigned)(NPRE * BOX_BYTES));
#pragma unroll 1
            for (int m = 0; m < NPRE; ++m) tma_load_box(stage + (size_t)m * BOX_ELEMS, a.tmap, tile * CT, m * BOX_ROWS, id.by, bar);
        } else {
            mbar_expect_tx(bar, (unsigned)(a.n_boxes * BOX_BYTES));
#pragma unroll 1
            for (int m = 0; m < R0; ++m) {
                const int sl = a.box_slot[m];
                if (sl >= 0) tma_load_box(stage + (size_t)sl * BOX_ELEMS, a.tmap, tile * CT, m * BOX_ROWS, id.by, bar);
            }
        }
#else
        for (int m = 0; m < R0; ++m) {
            const int sl = DENSE ? (m < NPRE ? m : -1) : a.box_slot[m];
            if (sl < 0) continue;
            for (int r = 0; r < BOX_ROWS; ++r)
                for (int cc = 0; cc < CT; ++cc)
                    stage[(size_t)sl * BOX_ELEMS + r * CT + cc] =
                        a.fld[(long long)id.by * a.fld_bs + (long long)(m * BOX_ROWS + r) * a.W + tile * CT + cc];
        }
#endif
    }

    // first-stage elements: staged boxes from shared memory, the rest (dense fields only) straight from global memory
    static SLMGS_DEVICE void load_direct(State& st, const Args& a, const Loc& L) {
        if constexpr (DENSE && NPRE < R0) {
            const long long step = (long long)BOX_ROWS * a.W;
            SLMGS_UNROLL
            for (int u = 0; u < E / R0; ++u) {
                const cf* p = a.fld + L.fbase + (long long)(L.lt + F::TPL * u) * a.W + (long long)NPRE * step;
                SLMGS_UNROLL
                for (int m = NPRE; m < R0; ++m) {
                    st.v[u * R0 + m] = ld_stream(p);
                    p += step;
                }
            }
        }
    }
    static SLMGS_DEVICE void load_staged(State& st, const Args& a, cf* smem, const ThreadId& id, const Loc& L) {
        const cf* sp = stage_of(smem) + id.tid;  // (lt * CT + col) == tid
        SLMGS_UNROLL
        for (int u = 0; u < E / R0; ++u) {
            SLMGS_UNROLL
            for (int m = 0; m < R0; ++m) {
                if (DENSE) {
                    if (m < NPRE) st.v[u * R0 + m] = sp[(size_t)m * BOX_ELEMS + (size_t)u * F::TPL * CT];
                } else {
                    const int sl = a.box_slot[m];
                    st.v[u * R0 + m] = sl >= 0 ? sp[(size_t)sl * BOX_ELEMS + (size_t)u * F::TPL * CT] : cmake(0.f, 0.f);
                }
            }
        }
    }
    static SLMGS_DEVICE void store_rows_p(State& st, const Args& a, const Loc& L) {
        if constexpr (DENSE) {
            const long long step = (long long)BOX_ROWS * a.W;
            SLMGS_UNROLL
            for (int u = 0; u < E / R0; ++u) {
                cf* p = a.fld + L.fbase + (long long)(L.lt + F::TPL * u) * a.W;
                SLMGS_UNROLL
                for (int m = 0; m < R0; ++m) {
                    *p = st.v[u * R0 + m];
                    p += step;
                }
            }
        } else {
            Base::store_rows(st, a, L);
        }
    }

    template <int P> static SLMGS_DEVICE void phase(State& st, const Args& a, cf* smem, const ThreadId& id) {
        const int q = id.bx + id.it * id.gx;
        const Loc L = Base::locate_q(a, smem, id, q);
        if constexpr (P == 0) {
            if (id.it == 0 && id.tid == 0) prefetch(a, smem, id, q);  // the block's first tile: nobody fetched it yet
            load_direct(st, a, L);
            mbar_wait(bar_of(smem), (unsigned)(id.it & 1));
            load_staged(st, a, smem, id, L);
            F::template fwd_stage<0>(st.v, L.lt, a.twA, a.twB, L.s, L.C);
        } else if constexpr (P == 1) {
            // every thread has read the staging buffer (barrier behind phase 0): fetch the next tile into it
            if (id.tid == 0 && id.it + 1 < iterations(a, id)) prefetch(a, smem, id, q + id.gx);
            F::template fwd_stage<1>(st.v, L.lt, a.twA, a.twB, L.s, L.C);
        } else if constexpr (P == NS - 1) {
            Base::prefetch_images_head(a, L);
            F::template fwd_stage<NS - 1>(st.v, L.lt, a.twA, a.twB, L.s, L.C);
            Base::template constrain<false>(st, a, id, L);
            F::template inv_stage<NS - 1>(st.v, L.lt, a.twA, a.twB, L.s, L.C);
        } else {
            F::template inv_stage<2 * NS - 2 - P>(st.v, L.lt, a.twA, a.twB, L.s, L.C);
            if constexpr (P == NPHASE - 1) store_rows_p(st, a, L);
        }
    }
};

}  // namespace slmgs
