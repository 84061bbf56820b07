// slmgs_compressed.h -- kernels of the kernel-based ("compressed") spot hologram.
//
// "Next" row 4 of SURVEY.md 8f: CompressedSpotHologram (slmsuite/holography/algorithms/_spots.py:178-1019).  Every
// spot n owns a phase kernel phi_n(pix) = sum_d a[d, n] Z_d(x_pix, y_pix) (Zernike polynomials of the basis on the
// aperture-scaled SLM grid, _spots.py:595-636) and the two maps of the GS loop are direct sums, O(S N):
//     farfield[n] = sum_pix nearfield[pix] exp(-i phi_n(pix)) / sqrt(S),  then  farfield /= ||farfield||   (:767-824)
//     nearfield[pix] = sum_n farfield[n] exp(+i phi_n(pix)) / sqrt(S)                                      (:887-915)
// (the reference's own CUDA pair for this is toolbox/cuda.cu:95-288: one thread per pixel, per-spot 1024-thread
// shared-memory tree reduction).  Here:
//   * the host evaluates the basis functions once (mono[m][pix] = Z_m(x_pix, y_pix), float64, from their monomial
//     expansion) and passes the spot coefficients in turns, cw[m][n] = a[m, n] / (2 pi), so phi/(2 pi) = sum_m cw[m][n]
//     mono[m][pix] is M = D double FMAs per (pixel, spot); the turn count is reduced in double before the MUFU sine /
//     cosine, so the phase is exact to float32 rounding however many radians it spans (the reference accumulates it
//     in float32);
//   * near -> far: a thread keeps PPT pixels (monomials + near field) in registers and 16 spot accumulators, loads the
//     spot weights once per PPT pixels (L1 broadcast), and a block contributes one warp-reduced atomicAdd(double) per
//     warp and spot; far -> near: a thread keeps PPT pixels and loops over all spots; the phase-only projection
//     (arctan2) is fused into its epilogue, the near-field amplitude is never stored;
//   * everything on the N-vectors (normalisation, WGS update, WGS-Kim phase, MRAF) is one single-block kernel.
// SFU / FP64-issue bound, not HBM bound: S N sincos evaluations per map.
#pragma once

#include "slmgs_kernels.h"

namespace slmgs {

enum { COMP_SPOTS = 16, COMP_PPT = 4 };

struct CompArgs {
    long long S;          // SLM pixels
    int N;                // spots
    int M;                // basis functions in use (<= MT of the instantiation)
    const double* mono;   // [MT][S]   basis-function values per pixel (rows >= M are zero)
    const double* cw;     // [MT][N]   per-spot weights of the basis functions in turns (rows >= M are zero)
    const float* phase;   // [S]
    const float* amp;     // [S] or nullptr
    float amp_scalar;
    cf* nf;               // [S] near field amp * exp(i phase)
    double* facc;         // [N][2] far-field accumulators
    const cf* far;        // [N] constrained far field
    float* phase_out;     // [S]
};

#if defined(__CUDACC__) && !defined(SLMGS_EMULATE)
SLMGS_DEVICE void sincos_turns(double turns, float* s, float* c) {
    // round to nearest integer with two FP64 adds (|turns| < 2^51) instead of FRND.F64, which shares the 16-lane XU
    // pipe with the two MUFU evaluations and the F2F below -- the unit that bounds these kernels
    const double r = (turns + 6755399441055744.0) - 6755399441055744.0;
    const double t = turns - r;  // [-0.5, 0.5]
    // MUFU sine / cosine on an argument already reduced to [-pi, pi]: absolute error 2^-21.4 (CUDA programming guide),
    // the size of one float32 rounding of the result; the reduction above is where the accuracy comes from
    __sincosf(6.283185307179586f * (float)t, s, c);
}
#else
inline void sincos_turns(double turns, float* s, float* c) {
    const double t = turns - rint(turns);
    const double a = 6.283185307179586476925286766559 * t;
    *s = (float)sin(a);
    *c = (float)cos(a);
}
#endif

// nf = amp * exp(i phase): _build_nearfield with shape == slm_shape, _hologram.py:1000-1011
struct CompBuildKernel {
#ifndef SLMGS_EMULATE
    static SLMGS_DEVICE void barrier(const ThreadId&) { __syncthreads(); }
#endif
    typedef CompArgs Args;
    static constexpr int MAXT = 256;
    static constexpr int NPHASE = 1;
    struct State {};
    template <int P> static SLMGS_DEVICE void phase(State&, const Args& a, cf*, const ThreadId& id) {
        const long long stride = (long long)id.gx * id.nthreads;
        for (long long i = (long long)id.bx * id.nthreads + id.tid; i < a.S; i += stride) {
            float sn, cs;
            sincosf(a.phase[i], &sn, &cs);
            const float am = a.amp ? a.amp[i] : a.amp_scalar;
            a.nf[i] = cmake(am * cs, am * sn);
        }
    }
};

// near -> far.  grid (gx, ceil(N / COMP_SPOTS)); block y handles COMP_SPOTS spots, the blocks of a row stride over pixels.
template <int MT> struct CompNear2FarKernel {
#ifndef SLMGS_EMULATE
    static SLMGS_DEVICE void barrier(const ThreadId&) { __syncthreads(); }
#endif
    typedef CompArgs Args;
    static constexpr int MAXT = 256;
    static constexpr int MINB = MT <= 6 ? 2 : 1;  // <= 128 registers: the unrolled spot loop must not hoist all 16 weight sets
    static constexpr int NPHASE = 1;
    struct State {};
    template <int P> static SLMGS_DEVICE void phase(State&, const Args& a, cf*, const ThreadId& id) {
        float ax[COMP_SPOTS], ay[COMP_SPOTS];
        SLMGS_UNROLL
        for (int s = 0; s < COMP_SPOTS; ++s) ax[s] = ay[s] = 0.0f;
        const int n0 = id.by * COMP_SPOTS;
        const long long span = (long long)id.nthreads * COMP_PPT;
        for (long long base = (long long)id.bx * span; base < a.S; base += (long long)id.gx * span) {
            double mono[MT][COMP_PPT];
            cf nf[COMP_PPT];
            SLMGS_UNROLL
            for (int k = 0; k < COMP_PPT; ++k) {
                const long long i = base + (long long)k * id.nthreads + id.tid;
                const bool in = i < a.S;
                nf[k] = in ? a.nf[i] : cmake(0.f, 0.f);
                SLMGS_UNROLL
                for (int m = 0; m < MT; ++m) mono[m][k] = in ? a.mono[(long long)m * a.S + i] : 0.0;
            }
            SLMGS_UNROLL
            for (int s = 0; s < COMP_SPOTS; ++s) {
                const int n = n0 + s < a.N ? n0 + s : a.N - 1;  // clamped: out-of-range spots are dropped below
                double w[MT];
                SLMGS_UNROLL
                for (int m = 0; m < MT; ++m) w[m] = __ldg(a.cw + (long long)m * a.N + n);
                SLMGS_UNROLL
                for (int k = 0; k < COMP_PPT; ++k) {
                    double t = 0.0;
                    SLMGS_UNROLL
                    for (int m = 0; m < MT; ++m) t = fma(w[m], mono[m][k], t);
                    float sn, cs;
                    sincos_turns(t, &sn, &cs);
                    // nf * exp(-i phi)
                    ax[s] = fmaf(nf[k].x, cs, fmaf(nf[k].y, sn, ax[s]));
                    ay[s] = fmaf(nf[k].y, cs, fmaf(-nf[k].x, sn, ay[s]));
                }
            }
        }
        SLMGS_UNROLL
        for (int s = 0; s < COMP_SPOTS; ++s) {
            const int n = n0 + s;
            // every thread of the block calls accum_add (full warps); out-of-range spots add to a dummy of zero weight
            const bool ok = n < a.N;
            double* slot = a.facc + 2LL * (ok ? n : 0);
            accum_add(slot, ok ? (double)ax[s] : 0.0);
            accum_add(slot + 1, ok ? (double)ay[s] : 0.0);
        }
    }
};

// far -> near with the phase-only projection fused: phase = arctan2(nearfield), _hologram.py:1026-1036
template <int MT> struct CompFar2NearKernel {
#ifndef SLMGS_EMULATE
    static SLMGS_DEVICE void barrier(const ThreadId&) { __syncthreads(); }
#endif
    typedef CompArgs Args;
    static constexpr int MAXT = 256;
    static constexpr int NPHASE = 1;
    struct State {};
    template <int P> static SLMGS_DEVICE void phase(State&, const Args& a, cf*, const ThreadId& id) {
        const long long base = (long long)id.bx * id.nthreads * COMP_PPT;
        double mono[MT][COMP_PPT];
        float ax[COMP_PPT], ay[COMP_PPT];
        SLMGS_UNROLL
        for (int k = 0; k < COMP_PPT; ++k) {
            const long long i = base + (long long)k * id.nthreads + id.tid;
            ax[k] = ay[k] = 0.0f;
            SLMGS_UNROLL
            for (int m = 0; m < MT; ++m) mono[m][k] = i < a.S ? a.mono[(long long)m * a.S + i] : 0.0;
        }
        for (int n = 0; n < a.N; ++n) {
            const cf f = __ldg(a.far + n);
            double w[MT];
            SLMGS_UNROLL
            for (int m = 0; m < MT; ++m) w[m] = __ldg(a.cw + (long long)m * a.N + n);
            SLMGS_UNROLL
            for (int k = 0; k < COMP_PPT; ++k) {
                double t = 0.0;
                SLMGS_UNROLL
                for (int m = 0; m < MT; ++m) t = fma(w[m], mono[m][k], t);
                float sn, cs;
                sincos_turns(t, &sn, &cs);
                // f * exp(+i phi)
                ax[k] = fmaf(f.x, cs, fmaf(-f.y, sn, ax[k]));
                ay[k] = fmaf(f.y, cs, fmaf(f.x, sn, ay[k]));
            }
        }
        SLMGS_UNROLL
        for (int k = 0; k < COMP_PPT; ++k) {
            const long long i = base + (long long)k * id.nthreads + id.tid;
            if (i < a.S) a.phase_out[i] = atan2f(ay[k], ax[k]);
        }
    }
};

// ------------------------------------------------------------------------------------------
// N-vector stage between the two maps (single block): normalisation (_spots.py:822), amp_ff (_hologram.py:951-953),
// WGS update on the spot vector (_spots.py:950-989 -> _hologram.py:1822-1879), WGS-Kim phase (:1556-1585),
// constraint incl. MRAF (:1587-1653), and clearing of the accumulators for the next near -> far pass.
// ------------------------------------------------------------------------------------------
struct CompVecArgs {
    double* facc;        // [N][2] in: raw sums; cleared on exit
    cf* far_norm;        // [N] out: normalised far field (what Hologram.farfield holds after _nearfield2farfield)
    cf* far;             // [N] out: constrained far field
    float* amp_ff;       // [N]
    float* phase_ff;     // [N]
    float* weights;      // [N]
    const float* target; // [N] (NaN = MRAF noise point, 0 = null point)
    int N;
    int finalize;        // 1: forward only (normalise, amp_ff); 2: _populate_results (also phase_ff = angle(farfield))
    int update;          // apply the weight update this iteration
    int phase_mode;      // PHASE_COMPUTE / PHASE_COMPUTE_STORE / PHASE_STORED
    int mraf, mraf_has_factor;
    float mraf_factor;
    WgsParams wgs;
};

struct CompVecKernel {
#ifndef SLMGS_EMULATE
    static SLMGS_DEVICE void barrier(const ThreadId&) { __syncthreads(); }
#endif
    typedef CompVecArgs Args;
    static constexpr int MAXT = 1024;
    static constexpr int NPHASE = 9;
    struct State {};
    // smem doubles: [0..nthreads) scratch, then scal[0..3]: ||F||^2, sum amp_ff^2, Nogrette ratio sum, sum w^2
    template <int P> static SLMGS_DEVICE void phase(State&, const Args& a, cf* smem, const ThreadId& id) {
        double* sm = reinterpret_cast<double*>(smem);
        double* scal = sm + id.nthreads;
        WgsParams q = a.wgs;
        if (P == 0) {
            double s = 0.0;
            for (int n = id.tid; n < a.N; n += id.nthreads) {
                const float x = (float)a.facc[2 * n], y = (float)a.facc[2 * n + 1];
                s += (double)(x * x) + (double)(y * y);
            }
            sm[id.tid] = s;
        } else if (P == 1 || P == 3 || P == 5 || P == 7) {
            if (id.tid == 0) {
                double s = 0.0;
                for (int t = 0; t < id.nthreads; ++t) s += sm[t];
                scal[P >> 1] = s;
            }
        } else if (P == 2) {
            const float inv = (float)(1.0 / sqrt(scal[0]));
            double s = 0.0;
            for (int n = id.tid; n < a.N; n += id.nthreads) {
                const cf f = cmake((float)a.facc[2 * n] * inv, (float)a.facc[2 * n + 1] * inv);
                a.far_norm[n] = f;
                const float am = sqrtf(f.x * f.x + f.y * f.y);
                a.amp_ff[n] = am;
                if (am == am) s += (double)am * (double)am;
                if (a.finalize == 2) a.phase_ff[n] = atan2f(f.y, f.x);
                a.facc[2 * n] = 0.0;
                a.facc[2 * n + 1] = 0.0;
            }
            sm[id.tid] = s;
        } else if (P == 4) {
            double s = 0.0;
            if (!a.finalize && a.update && q.method == METHOD_NOGRETTE) {
                q.inv_fnorm = (float)(1.0 / sqrt(scal[1]));
                for (int n = id.tid; n < a.N; n += id.nthreads) s += (double)wgs_ratio(a.amp_ff[n], a.target[n], q);
            }
            sm[id.tid] = s;
        } else if (P == 6) {
            double s = 0.0;
            if (!a.finalize && a.update) {
                q.inv_fnorm = (float)(1.0 / sqrt(scal[1]));
                q.neg_inv_mean = -(1.0f / (float)(scal[2] / (double)a.N));
                for (int n = id.tid; n < a.N; n += id.nthreads) {
                    const float w = wgs_apply(a.weights[n], wgs_multiplier(a.amp_ff[n], a.target[n], q));
                    a.weights[n] = w;
                    s += (double)w * (double)w;
                }
            }
            sm[id.tid] = s;
        } else if (P == 8) {
            if (a.finalize) return;
            const float sc = a.update ? (float)(1.0 / sqrt(scal[3])) : 1.0f;
            for (int n = id.tid; n < a.N; n += id.nthreads) {
                float w = a.weights[n];
                if (a.update) {
                    w *= sc;
                    a.weights[n] = w;
                }
                const cf f = a.far_norm[n];
                const float t = a.target[n];
                const bool noise = a.mraf && t != t;
                const bool zero = a.mraf && t == 0.0f;
                float ph;
                if (a.phase_mode == PHASE_STORED) {
                    ph = a.phase_ff[n];
                } else {
                    ph = zero ? 0.0f : atan2f(f.y, f.x);  // MRAF: angle taken after farfield[zero] = 0 (:1613-1622)
                    a.phase_ff[n] = ph;
                }
                cf g;
                if (zero) {
                    g = cmake(0.f, 0.f);
                } else if (noise) {
                    g = a.mraf_has_factor ? cscale(f, a.mraf_factor) : f;
                } else {
                    float sn, cs;
                    sincosf(ph, &sn, &cs);
                    g = cmake(w * cs, w * sn);
                }
                a.far[n] = g;
            }
        }
    }
};

}  // namespace slmgs
