// slmgs_fft.h -- register/shared-memory 1-D FFT of a line of N = R0*R1*R2 points.
//
// Data model.  A tile is LINES lines of N points.  Each thread owns E points (E = 16, or 32
// for N = 8192) in registers for the whole kernel; shared memory is only the exchange medium
// between radix stages.  A line is served by TPL = N/E threads; "lt" is the thread's index
// inside its line.  A stage of radix R has N/R butterflies per line; thread lt executes the
// Q = E/R butterflies b = lt + TPL*u, u = 0..Q-1, keeping butterfly u in v[u*R .. u*R+R-1].
//
// Index algebra (n = spatial index, k = frequency index):
//     n = n0*(R1 R2) + n1*R2 + n2            k = k0 + R0*k1 + R0*R1*k2
//   stage A (radix R0): butterfly b = j = n1*R2+n2, element m=n0/k0 at line index  m*M1 + j
//   stage B (radix R1): butterfly b = k0*R2 + n2,   element m=n1/k1 at line index  k0*M1 + m*R2 + n2
//   stage C (radix R2): butterfly b = k0 + R0*k1,   element m=n2/k2 at line index  k0*M1 + k1*R2 + m
// with M1 = R1*R2.  Every stage reads and writes the SAME R line positions (in place), so one
// barrier per stage suffices.  First-stage elements sit at spatial index n = b + (N/R0)*m and
// last-stage elements at frequency k = b + (N/Rlast)*m, i.e. consecutive threads touch
// consecutive addresses at both ends: global loads and stores are coalesced straight from
// registers with no staging pass.
//
//   forward:  A, *W_N^{j k0}, B, *W_M1^{n2 k1}, C            (DIF)
//   inverse:  C^-1, *conj W_M1, B^-1, *conj W_N, A^-1        (exact mirror, DIT)
// so the inverse consumes exactly the register layout the forward produces and vice versa:
// forward -> pointwise -> inverse (column kernel) and inverse -> pointwise -> forward (row
// kernel) chain without any data movement in between.
//
// Four stages (N = 8192 = 2*16*16*16): one more digit, n = ((n0 R1 + n1) R2 + n2) R3 + n3, k = k0 + R0 (k1 + R1 (k2 + R2 k3)):
//   stage C becomes a middle stage (butterfly b = (k0 + R0*k1)*R3 + n3, elements along n2/k2) and stage D
//   (radix R3, butterfly b = k0 + R0*k1 + R0*R1*k2) the last one; with R3 = 1 every formula below reduces to the
//   three-stage one.  A third table follows twB in the same allocation: twC[k2*R3 + n3] = exp(-2 pi i n3 k2 / (R2 R3)).
//
// Twiddles come from two small host-computed tables (double precision, rounded once):
//   twA[k0*M1 + j] = exp(-2 pi i j k0 / N)   (N entries, indexed by the line index itself)
//   twB[k1*R2 + n2] = exp(-2 pi i n2 k1 / M1) (M1 entries)
// Consecutive threads read consecutive entries; the tables stay L1 resident.
#pragma once

#include "slmgs_common.h"

namespace slmgs {

// L1-burst hooks of the *_sy stage functions (see below): no-ops outside the ping-pong kernels
struct NoSync {
    SLMGS_DEVICE void acquire() {}
    SLMGS_DEVICE void release() {}
};

// twiddle tables in shared memory (compact layout, see Fft::twiddle)
struct SmemTw {
    const cf* p;
};

template <int N> struct Plan;
// E: points per thread; R0,R1,R2: stage radices (smallest first so the two big stages keep 16 consecutive
// lanes on consecutive addresses); P2: pad of the k0 stride (see "Shared memory layout" below).
#define SLMGS_PLAN4(N_, E_, A_, B_, C_, D_, P2_)                                         \
    template <> struct Plan<N_> {                                                        \
        static constexpr int E = E_, R0 = A_, R1 = B_, R2 = C_, R3 = D_, P2 = P2_;       \
    };
#define SLMGS_PLAN(N_, E_, A_, B_, C_, P2_) SLMGS_PLAN4(N_, E_, A_, B_, C_, 1, P2_)
SLMGS_PLAN(16, 16, 16, 1, 1, 0)
SLMGS_PLAN(32, 16, 2, 16, 1, 1)
SLMGS_PLAN(64, 16, 4, 16, 1, 1)
SLMGS_PLAN(128, 16, 8, 16, 1, 1)
SLMGS_PLAN(256, 16, 16, 16, 1, 1)
SLMGS_PLAN(512, 16, 2, 16, 16, 8)
SLMGS_PLAN(1024, 16, 4, 16, 16, 4)
SLMGS_PLAN(2048, 16, 8, 16, 16, 2)
#ifndef SLMGS_E4096
#define SLMGS_E4096 16
#endif
SLMGS_PLAN(4096, SLMGS_E4096, 16, 16, 16, 1)
// 8192 points: 32 points per thread, three stages (128 registers; the hot kernels spill a few registers).  The
// four-stage plan 2*16*16*16 (-DSLMGS_8192_4STAGE: 16 points per thread, 64 registers, no spills in the hot kernels) was
// measured SLOWER on B200 (configs[4], 10 iterations: 17.9 vs 16.6 ms): its extra exchange and twiddle layer cost more than
// the spills and the halved occupancy of this one.
#ifdef SLMGS_8192_4STAGE
SLMGS_PLAN4(8192, 16, 2, 16, 16, 16, 8)
#else
SLMGS_PLAN(8192, 32, 32, 16, 16, 1)
#endif
#undef SLMGS_PLAN4
#undef SLMGS_PLAN

template <int N> struct Fft {
    typedef Plan<N> P;
    static constexpr int E = P::E, R0 = P::R0, R1 = P::R1, R2 = P::R2, R3 = P::R3;
    static constexpr int NS = 1 + (R1 > 1) + (R2 > 1) + (R3 > 1);
    static constexpr int M1 = R1 * R2 * R3;  // points behind the first digit
    static constexpr int M2 = R2 * R3;       // ... behind the second
    static constexpr int TPL = N / E;
    static_assert(R0 * R1 * R2 * R3 == N, "bad plan");
    static_assert(R3 == 1 || (R1 > 1 && R2 > 1), "a fourth stage needs the other three");
    // Shared memory layout.  A line index i = ((d0 R1 + d1) R2 + d2) R3 + d3 is stored at
    //     d0*T2 + d1*T1 + d2*T0 + d3,   T0 = R3 + (R3 > 1),   T1 = R2*T0 + (R2 > 1),   T2 = R1*T1 + P2
    // i.e. plain digit strides with small pads (three stages: R3 = 1, T0 = 1).  Element m of a butterfly is then always at
    // base(b) + m*const, so the 16 shared-memory accesses of a stage use one address register with
    // compile-time offsets.  Pads are chosen so that within every half-warp (16 lanes, 8-byte accesses)
    // the access patterns hit 16 distinct bank pairs:
    //   lanes along the last digit (all stages but the last): consecutive;    last stage: lanes along d0 (stride T2 = P2
    //   mod 16) and d1 (stride T1 = 1 mod 16): P2*d0 + d1 distinct for the R0 x 16/R0 lanes of a half-warp.
    static constexpr int T0 = R3 + (R3 > 1 ? 1 : 0);
    static constexpr int T1 = R2 * T0 + (R2 > 1 ? 1 : 0);
    static constexpr int T2 = R1 * T1 + P::P2;
    static constexpr int PADN = R0 * T2 + 1;

    template <int S> static constexpr int radix() { return S == 0 ? R0 : S == 1 ? R1 : S == 2 ? R2 : R3; }
    static constexpr int last_radix() { return radix<NS - 1>(); }

    // shared-memory position of element 0 of butterfly b at stage S, and the stride between elements
    //   S = 0: b = (d1, d2, d3);  S = 1: b = (d0, d2, d3);  S = 2: b = ((d0 + R0 d1), d3);  S = 3: b = d0 + R0 (d1 + R1 d2)
    template <int S> static SLMGS_HD int sbase(int b) {
        if (S == 0) return (b / M2) * T1 + ((b / R3) % R2) * T0 + (b % R3);
        if (S == 1) return (b / M2) * T2 + ((b / R3) % R2) * T0 + (b % R3);
        if (S == 2) return ((b / R3) % R0) * T2 + ((b / R3) / R0) * T1 + (b % R3);
        return (b % R0) * T2 + ((b / R0) % R1) * T1 + (b / (R0 * R1)) * T0;
    }
    template <int S> static constexpr int sstride() { return S == 0 ? T2 : S == 1 ? T1 : S == 2 ? T0 : 1; }
    // twiddle applied after forward stage S (S < NS-1) to output m of butterfly b
    template <int S> static SLMGS_DEVICE cf twiddle(const cf* SLMGS_RESTRICT twA, const cf* SLMGS_RESTRICT twB, int b,
                                                   int m) {
        if (S == 0) return __ldg(twA + m * M1 + b);
        if (S == 1) return __ldg(twB + m * M2 + (b % M2));
        return __ldg(twB + M1 + m * R3 + (b % R3));  // third table, stored behind the second
    }
    // The same from a compact copy of the tables in shared memory (persistent kernels, slmgs_teams.h): with
    // SLMGS_TW_PRODUCTS only the rows m = 1, 2, 4, 8 of either table are ever read; row m is stored at log2(m).
    static constexpr int TWS_ROWS = 4;
    static constexpr int TWS_A = TWS_ROWS * M1, TWS_B = TWS_ROWS * R2;  // entries of the compact tables
    static SLMGS_HD int tws_row(int m) { return m == 1 ? 0 : m == 2 ? 1 : m == 4 ? 2 : 3; }
    template <int S> static SLMGS_DEVICE cf twiddle(SmemTw twA, SmemTw twB, int b, int m) {
        static_assert(R3 == 1, "compact shared-memory tables: three-stage plans");
        if (S == 0) return twA.p[tws_row(m) * M1 + b];
        return twB.p[tws_row(m) * R2 + (b % R2)];
    }
    // spatial index of element m of first-stage butterfly b / frequency of element m of last-stage butterfly b
    static SLMGS_HD int first_index(int b, int m) { return b + (N / R0) * m; }
    static SLMGS_HD int last_index(int b, int m) { return b + (N / last_radix()) * m; }

    // ---- twiddle generation ---------------------------------------------------------------
    // The R-1 twiddles w^k of a butterfly (w = table entry of exponent 1) are visited as fn(IC<k>, w^k).
    // SLMGS_TW_PRODUCTS: only the powers of two are loaded (w, w^2, w^4, w^8, ...), the others are products of
    // at most three loaded values (<= 2.5 ulp): 4 instead of 15 table loads per radix-16 stage, which takes
    // load off the L1 data pipe that the shared-memory exchange already keeps half busy.
    template <int K> struct IC {
        static constexpr int value = K;
    };
    static constexpr int high_pow2(int g) { return g >= 16 ? 16 : g >= 8 ? 8 : g >= 4 ? 4 : g >= 2 ? 2 : 1; }
    template <int S, int G, int NG, class Fn, class TP>
    static SLMGS_DEVICE void tw_groups(cf* W, cf w1, cf w2, cf w3, int b, TP twA, TP twB, Fn& fn) {
        if constexpr (G < NG) {
            constexpr int hp = high_pow2(G);
            if constexpr (hp == G) W[G] = twiddle<S>(twA, twB, b, 4 * G);
            else W[G] = cmul(W[hp], W[G - hp]);
            fn(IC<4 * G>(), W[G]);
            fn(IC<4 * G + 1>(), cmul(W[G], w1));
            fn(IC<4 * G + 2>(), cmul(W[G], w2));
            fn(IC<4 * G + 3>(), cmul(W[G], w3));
            tw_groups<S, G + 1, NG>(W, w1, w2, w3, b, twA, twB, fn);
        }
    }
    template <int S, int K, class Fn, class TP> static SLMGS_DEVICE void tw_table(int b, TP twA, TP twB, Fn& fn) {
        if constexpr (K < radix<S>()) {
            fn(IC<K>(), twiddle<S>(twA, twB, b, K));
            tw_table<S, K + 1>(b, twA, twB, fn);
        }
    }
    template <int S, class Fn, class TP> static SLMGS_DEVICE void for_each_twiddle(int b, TP twA, TP twB, Fn fn) {
        constexpr int R = radix<S>();
#ifdef SLMGS_TW_PRODUCTS
        if constexpr (R >= 8) {
            const cf w1 = twiddle<S>(twA, twB, b, 1);
            const cf w2 = twiddle<S>(twA, twB, b, 2);
            const cf w3 = cmul(w1, w2);
            fn(IC<1>(), w1);
            fn(IC<2>(), w2);
            fn(IC<3>(), w3);
            cf W[R / 4];
            tw_groups<S, 1, R / 4>(W, w1, w2, w3, b, twA, twB, fn);
        } else {
            tw_table<S, 1>(b, twA, twB, fn);
        }
#else
        tw_table<S, 1>(b, twA, twB, fn);
#endif
    }

    // ---- compile-time loops --------------------------------------------------------------
    template <int S, int U> static SLMGS_DEVICE void fwd_store_all(cf* v, int lt, const cf* twA, const cf* twB, cf* s, int si) {
        constexpr int R = radix<S>();
        const int b = lt + TPL * U;
        cf* sp = s + sbase<S>(b) * si;
        sp[0] = v[U * R + RegFFT<R>::pos(0)];
        for_each_twiddle<S>(b, twA, twB, [&](auto k, cf w) {
            constexpr int K = decltype(k)::value;
            sp[K * sstride<S>() * si] = cmul(v[U * R + RegFFT<R>::pos(K)], w);
        });
    }
    template <int S, int U, int K> static SLMGS_DEVICE void load_elems(cf* v, int lt, const cf* s, int si) {
        constexpr int R = radix<S>();
        if constexpr (K < R) {
            const int b = lt + TPL * U;
            v[U * R + K] = s[(sbase<S>(b) + K * sstride<S>()) * si];
            load_elems<S, U, K + 1>(v, lt, s, si);
        }
    }
    template <int S, int U, class TP> static SLMGS_DEVICE void inv_twiddle_all(cf* v, int lt, TP twA, TP twB) {
        constexpr int R = radix<S>();
        for_each_twiddle<S>(lt + TPL * U, twA, twB, [&](auto k, cf w) {
            constexpr int K = decltype(k)::value;
            v[U * R + K] = cmulc(v[U * R + K], w);
        });
    }
    template <int S, int U, int K> static SLMGS_DEVICE void inv_store(cf* v, int lt, cf* s, int si) {
        constexpr int R = radix<S>();
        if constexpr (K < R) {
            const int b = lt + TPL * U;
            s[(sbase<S>(b) + K * sstride<S>()) * si] = v[U * R + RegFFT<R>::pos(K)];
            inv_store<S, U, K + 1>(v, lt, s, si);
        }
    }
    template <int R, int U, int K> static SLMGS_DEVICE void unscramble(const cf* v, cf* o) {
        if constexpr (K < R) {
            o[U * R + K] = v[U * R + RegFFT<R>::pos(K)];
            unscramble<R, U, K + 1>(v, o);
        }
    }

    // ---- forward stage S ------------------------------------------------------------------
    // S == 0     : v holds the spatial samples (v[u*R0+m] = x[first_index(b_u, m)])
    // S  > 0     : elements are read from shared memory first
    // S  < NS-1  : results are twiddled and written back to shared memory (caller barriers)
    // S == NS-1  : results stay in v, natural order: v[u*R+m] = X[last_index(b_u, m)]
    template <int S, int U> static SLMGS_DEVICE void fwd_stage_u(cf* v, int lt, const cf* twA, const cf* twB, cf* s,
                                                                 int si) {
        constexpr int R = radix<S>();
        constexpr int Q = E / R;
        if constexpr (U < Q) {
            if constexpr (S > 0) load_elems<S, U, 0>(v, lt, s, si);
            RegFFT<R>::template run<1, 1>(v + U * R);
            if constexpr (S < NS - 1) fwd_store_all<S, U>(v, lt, twA, twB, s, si);
            fwd_stage_u<S, U + 1>(v, lt, twA, twB, s, si);
        }
    }
    // The last forward stage split in two, so that a caller can issue work between the shared-memory reads and the
    // butterflies: fwd_last_load reads the elements, fwd_last_compute runs the butterflies on them.
    template <int S, int U> static SLMGS_DEVICE void fwd_load_u(cf* v, int lt, const cf* s, int si) {
        if constexpr (U < E / radix<S>()) {
            load_elems<S, U, 0>(v, lt, s, si);
            fwd_load_u<S, U + 1>(v, lt, s, si);
        }
    }
    template <int S, int U> static SLMGS_DEVICE void fwd_compute_u(cf* v) {
        if constexpr (U < E / radix<S>()) {
            RegFFT<radix<S>()>::template run<1, 1>(v + U * radix<S>());
            fwd_compute_u<S, U + 1>(v);
        }
    }
    static SLMGS_DEVICE void fwd_last_load(cf* v, int lt, const cf* s, int si) {
        static_assert(NS > 1, "needs an exchange stage");
        fwd_load_u<NS - 1, 0>(v, lt, s, si);
    }
    static SLMGS_DEVICE void fwd_last_compute(cf* v) {
        fwd_compute_u<NS - 1, 0>(v);
        cf o[E];
        unscramble_all<radix<NS - 1>(), 0>(v, o);
        SLMGS_UNROLL
        for (int i = 0; i < E; ++i) v[i] = o[i];
    }
    // Shared-memory slot (in cf units, before the column interleave) that the last forward stage read its element i
    // from and that the first inverse stage will write its element i to: PRIVATE to the thread in between.
    static SLMGS_HD int last_slot(int lt, int i) {
        constexpr int R = radix<NS - 1>();
        return sbase<NS - 1>(lt + TPL * (i / R)) + (i % R) * sstride<NS - 1>();
    }

    template <int S> static SLMGS_DEVICE void fwd_stage(cf* v, int lt, const cf* twA, const cf* twB, cf* s, int si) {
        fwd_stage_u<S, 0>(v, lt, twA, twB, s, si);
        if constexpr (S == NS - 1) {
            cf o[E];
            unscramble_all<radix<S>(), 0>(v, o);
            SLMGS_UNROLL
            for (int i = 0; i < E; ++i) v[i] = o[i];
        }
    }
    template <int R, int U> static SLMGS_DEVICE void unscramble_all(const cf* v, cf* o) {
        if constexpr (U < E / R) {
            unscramble<R, U, 0>(v, o);
            unscramble_all<R, U + 1>(v, o);
        }
    }

    // ---- stages split at their shared-memory bursts (team kernels, slmgs_teams.h) -------------------
    // Same arithmetic as fwd_stage / inv_stage, ordered as  [shared-memory reads] release | butterflies + twiddles |
    // acquire [shared-memory writes].  `sy` are hooks around the bursts on the L1 / shared-memory data pipe: no-ops
    // (NoSync) in the product kernels; the token / mutex experiments of DESIGN.md 4.10 (tools/micro/pp_token.h) plugged
    // their hand-offs in here.  Stage S == 0 of the forward (S == NS-1 of the inverse) has no read burst.
    template <int S, int U, class TP> static SLMGS_DEVICE void fwd_twiddle_u(cf* v, int lt, TP twA, TP twB) {
        constexpr int R = radix<S>();
        if constexpr (U < E / R) {
            for_each_twiddle<S>(lt + TPL * U, twA, twB, [&](auto k, cf w) {
                constexpr int K = decltype(k)::value;
                v[U * R + RegFFT<R>::pos(K)] = cmul(v[U * R + RegFFT<R>::pos(K)], w);
            });
            fwd_twiddle_u<S, U + 1>(v, lt, twA, twB);
        }
    }
    template <int S, int U> static SLMGS_DEVICE void store_scrambled_u(cf* v, int lt, cf* s, int si) {
        if constexpr (U < E / radix<S>()) {
            inv_store<S, U, 0>(v, lt, s, si);  // s[element K] = v[pos(K)]
            store_scrambled_u<S, U + 1>(v, lt, s, si);
        }
    }
    template <int S, int U, class TP> static SLMGS_DEVICE void inv_twiddle_u(cf* v, int lt, TP twA, TP twB) {
        if constexpr (U < E / radix<S>()) {
            inv_twiddle_all<S, U>(v, lt, twA, twB);
            inv_twiddle_u<S, U + 1>(v, lt, twA, twB);
        }
    }
    template <int S, int U> static SLMGS_DEVICE void inv_compute_u(cf* v) {
        if constexpr (U < E / radix<S>()) {
            RegFFT<radix<S>()>::template run<-1, 1>(v + U * radix<S>());
            inv_compute_u<S, U + 1>(v);
        }
    }
    template <int S, class Sy, class TP>
    static SLMGS_DEVICE void fwd_stage_sy(cf* v, int lt, TP twA, TP twB, cf* s, int si, Sy& sy) {
        if constexpr (S > 0) {
            fwd_load_u<S, 0>(v, lt, s, si);
            sy.release();
        }
        fwd_compute_u<S, 0>(v);
        if constexpr (S < NS - 1) {
            fwd_twiddle_u<S, 0>(v, lt, twA, twB);
            sy.acquire();
            store_scrambled_u<S, 0>(v, lt, s, si);
        } else {
            cf o[E];
            unscramble_all<radix<S>(), 0>(v, o);
            SLMGS_UNROLL
            for (int i = 0; i < E; ++i) v[i] = o[i];
        }
    }
    // (S == 0: the caller acquires before it stores the results to global memory)
    template <int S, class Sy, class TP>
    static SLMGS_DEVICE void inv_stage_sy(cf* v, int lt, TP twA, TP twB, cf* s, int si, Sy& sy) {
        if constexpr (S < NS - 1) {
            fwd_load_u<S, 0>(v, lt, s, si);
            sy.release();
            inv_twiddle_u<S, 0>(v, lt, twA, twB);
        }
        inv_compute_u<S, 0>(v);
        if constexpr (S > 0) {
            sy.acquire();
            store_scrambled_u<S, 0>(v, lt, s, si);
        } else {
            cf o[E];
            unscramble_all<R0, 0>(v, o);
            SLMGS_UNROLL
            for (int i = 0; i < E; ++i) v[i] = o[i];
        }
    }

    // ---- inverse (mirror) stage S -----------------------------------------------------------
    // S == NS-1 : v holds the spectrum in natural order (v[u*R+m] = X[last_index(b_u, m)])
    // S  < NS-1 : elements are read from shared memory and multiplied by conj(twiddle_S)
    // S  > 0    : results are written back to shared memory (caller barriers)
    // S == 0    : results stay in v, natural order: v[u*R0+m] = x[first_index(b_u, m)] (unnormalised)
    template <int S, int U> static SLMGS_DEVICE void inv_stage_u(cf* v, int lt, const cf* twA, const cf* twB, cf* s,
                                                                 int si) {
        constexpr int R = radix<S>();
        constexpr int Q = E / R;
        if constexpr (U < Q) {
            if constexpr (S < NS - 1) {
                load_elems<S, U, 0>(v, lt, s, si);
                inv_twiddle_all<S, U>(v, lt, twA, twB);
            }
            RegFFT<R>::template run<-1, 1>(v + U * R);
            if constexpr (S > 0) inv_store<S, U, 0>(v, lt, s, si);
            inv_stage_u<S, U + 1>(v, lt, twA, twB, s, si);
        }
    }
    template <int S> static SLMGS_DEVICE void inv_stage(cf* v, int lt, const cf* twA, const cf* twB, cf* s, int si) {
        inv_stage_u<S, 0>(v, lt, twA, twB, s, si);
        if constexpr (S == 0) {
            cf o[E];
            unscramble_all<R0, 0>(v, o);
            SLMGS_UNROLL
            for (int i = 0; i < E; ++i) v[i] = o[i];
        }
    }
};

}  // namespace slmgs
