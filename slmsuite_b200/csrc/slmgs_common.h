// slmgs_common.h -- shared device / host-emulation primitives for the GS/WGS hot path.
//
// Every kernel in this library is written as a sequence of barrier-delimited "phases"
// (see slmgs_launch.h).  Under nvcc the phases are chained with __syncthreads() inside one
// sm_100a kernel.  Under a plain C++ compiler with -DSLMGS_EMULATE the very same phase
// functions are executed by a host loop over (block, phase, thread): that build is TEST
// INFRASTRUCTURE (tests/_emu) used to validate index math and arithmetic without a GPU;
// the product library (libslmgs.so) never contains it.
#pragma once

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__) && !defined(SLMGS_EMULATE)
#define SLMGS_DEVICE __device__ __forceinline__
#define SLMGS_HD __host__ __device__ __forceinline__
#define SLMGS_RESTRICT __restrict__
#define SLMGS_UNROLL _Pragma("unroll")
#else
#ifndef SLMGS_EMULATE
#define SLMGS_EMULATE 1
#endif
#define SLMGS_DEVICE inline
#define SLMGS_HD inline
#define SLMGS_RESTRICT
#define SLMGS_UNROLL
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline void sincosf_emu(float a, float* s, float* c) { *s = sinf(a); *c = cosf(a); }
#define sincosf(a, s, c) sincosf_emu(a, s, c)
template <class T> static inline T __ldg(const T* p) { return *p; }
#endif

namespace slmgs {

typedef float2 cf;  // complex float, (re, im)

// Streaming global loads: the field and the far-field images are touched once per kernel, so they
// must not evict the twiddle tables from L1 (ld.global.L1::no_allocate).
#if defined(__CUDACC__) && !defined(SLMGS_EMULATE)
SLMGS_DEVICE float ld_stream(const float* p) {
    float v;
    asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
SLMGS_DEVICE float2 ld_stream(const float2* p) {
    float2 v;
    asm volatile("ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
SLMGS_DEVICE void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
SLMGS_DEVICE void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// asynchronous 4-byte copy global -> shared (LDGSTS): no register is held while the load is in flight
SLMGS_DEVICE void cp_async_f32(float* smem_dst, const float* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
SLMGS_DEVICE void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// cached (L1-allocating) load for data that was prefetched into L1
SLMGS_DEVICE float ld_cached(const float* p) {
    float v;
    asm volatile("ld.global.ca.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
#else
inline float ld_stream(const float* p) { return *p; }
inline float2 ld_stream(const float2* p) { return *p; }
inline void prefetch_l2(const void*) {}
inline void prefetch_l1(const void*) {}
inline void cp_async_f32(float* smem_dst, const float* gsrc) { *smem_dst = *gsrc; }
inline void cp_async_wait_all() {}
inline float ld_cached(const float* p) { return *p; }
#endif

// ---------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) + mbarrier: asynchronous staging of field tiles in shared memory.
// Under emulation the copy is done synchronously by the issuing thread (phases run one after the
// other, so the data is there when the consumers' phase starts) and the barrier is a no-op.
// ---------------------------------------------------------------------------------------
#if defined(__CUDACC__) && !defined(SLMGS_EMULATE)
SLMGS_DEVICE unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
SLMGS_DEVICE void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
SLMGS_DEVICE void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
SLMGS_DEVICE void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
// box {C columns, 256 rows, 1 hologram} of the row-major field at (x, y, z) -> dense [256][C] block in shared memory
SLMGS_DEVICE void tma_load_box(void* smem_dst, const void* tmap, int x, int y, int z, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}
SLMGS_DEVICE void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
#else
inline void mbar_init(unsigned long long*, int) {}
inline void mbar_expect_tx(unsigned long long*, unsigned) {}
inline void mbar_wait(unsigned long long*, unsigned) {}
inline void fence_proxy_async() {}
#endif

// log2 of a power of two
#if defined(__CUDACC__) && !defined(SLMGS_EMULATE)
SLMGS_DEVICE int ilog2(int x) { return __ffs(x) - 1; }
#else
inline int ilog2(int x) { return __builtin_ctz((unsigned)x); }
#endif

SLMGS_HD cf cmake(float re, float im) { return make_float2(re, im); }
#if defined(__CUDA_ARCH__) && !defined(SLMGS_EMULATE) && defined(SLMGS_PACKED_F32X2)
// Blackwell packed FP32x2 (FADD2 / FMUL2 / FFMA2): one issue slot per complex add / subtract, two per complex
// multiply.  ptxas folds the operand shapes used below into the instruction's own modifiers -- a scalar broadcast
// (`R.F32`) and a swapped pair with one half negated (`R.F32x2.LO_HI.NP`) -- so no MOV is spent on them:
//     a * w       = (a.x, a.y) * w.x + (-a.y,  a.x) * w.y        FMUL2 + FFMA2
//     a * conj(w) = (a.x, a.y) * w.x + ( a.y, -a.x) * w.y        FMUL2 + FFMA2
// Rounding is that of the scalar FMUL + FFMA sequence the compiler contracts the plain expression to.
SLMGS_HD cf cadd(cf a, cf b) { return __fadd2_rn(a, b); }
SLMGS_HD cf csub(cf a, cf b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }  // FADD2 R, R, -R (the negation is an operand modifier)
SLMGS_HD cf cmul(cf a, cf b) {
    return __ffma2_rn(make_float2(-a.y, a.x), make_float2(b.y, b.y), __fmul2_rn(a, make_float2(b.x, b.x)));
}
// a * conj(b)
SLMGS_HD cf cmulc(cf a, cf b) {
    return __ffma2_rn(make_float2(a.y, -a.x), make_float2(b.y, b.y), __fmul2_rn(a, make_float2(b.x, b.x)));
}
SLMGS_HD cf cscale(cf a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
// a + s * b, a - i s b ... helpers for the constant twiddles of the in-register butterflies
SLMGS_HD cf caxpy(float s, cf b, cf a) { return __ffma2_rn(b, make_float2(s, s), a); }
#else
SLMGS_HD cf cadd(cf a, cf b) { return make_float2(a.x + b.x, a.y + b.y); }
SLMGS_HD cf csub(cf a, cf b) { return make_float2(a.x - b.x, a.y - b.y); }
SLMGS_HD cf cmul(cf a, cf b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// a * conj(b)
SLMGS_HD cf cmulc(cf a, cf b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
SLMGS_HD cf cscale(cf a, float s) { return make_float2(a.x * s, a.y * s); }
SLMGS_HD cf caxpy(float s, cf b, cf a) { return make_float2(a.x + s * b.x, a.y + s * b.y); }
#endif
// a * w (DIR=+1) or a * conj(w) (DIR=-1)
template <int DIR> SLMGS_HD cf cmul_dir(cf a, cf w) { return DIR > 0 ? cmul(a, w) : cmulc(a, w); }
// multiply by -i (DIR=+1, forward) or +i (DIR=-1, inverse)
template <int DIR> SLMGS_HD cf cmul_mi(cf a) {
    return DIR > 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}

// ---------------------------------------------------------------------------------------
// 32nd roots of unity, cos(2 pi k / 32), k = 0..8, correctly rounded to fp32.  In-register
// butterflies only need constants up to radix 32.
// ---------------------------------------------------------------------------------------
constexpr float root32_tab(int j) {
    return j == 0 ? 1.0f : j == 1 ? 0.98078528040323043f : j == 2 ? 0.92387953251128674f
         : j == 3 ? 0.83146961230254524f : j == 4 ? 0.70710678118654757f : j == 5 ? 0.55557023301960218f
         : j == 6 ? 0.38268343236508978f : j == 7 ? 0.19509032201612825f : 0.0f;
}
constexpr float root32_cos(int k) {
    return k <= 8 ? root32_tab(k) : k <= 16 ? -root32_tab(16 - k) : k <= 24 ? -root32_tab(k - 16) : root32_tab(32 - k);
}
constexpr float root32_sin(int k) {
    return k <= 8 ? root32_tab(8 - k) : k <= 16 ? root32_tab(k - 8) : k <= 24 ? -root32_tab(24 - k) : -root32_tab(k - 24);
}

// v * exp(-i * DIR * 2 pi E / R)   (compile-time constant twiddle; trivial cases folded)
template <int DIR, int R, int E> SLMGS_HD cf ctwiddle_const(cf v) {
    constexpr int k = ((E * (32 / R)) % 32 + 32) % 32;
    if (k == 0) return v;
    if (k == 8) return cmul_mi<DIR>(v);
    if (k == 16) return make_float2(-v.x, -v.y);
    if (k == 24) return cmul_mi<-DIR>(v);
    constexpr float c = root32_cos(k);
    constexpr float s = DIR > 0 ? -root32_sin(k) : root32_sin(k);
    if (k == 4 || k == 12 || k == 20 || k == 28) {
        // |c| == |s| == sqrt(1/2):  (x + i y)(sc + i ss) q = ((x, y) + (ss/sc) (-y, x)) * (sc q): one packed add with a
        // swapped operand + one packed multiply by a scalar
        constexpr float q = 0.70710678118654757f;
        constexpr float sc = c > 0 ? 1.0f : -1.0f, ss = s > 0 ? 1.0f : -1.0f;
        return cscale(caxpy(ss * sc, make_float2(-v.y, v.x), v), sc * q);
    }
    // (x + i y)(c + i s) = (x, y) c + (-y, x) s
    return caxpy(s, make_float2(-v.y, v.x), cscale(v, c));
}

// ---------------------------------------------------------------------------------------
// In-register DFT of R points (R = 1, 2, 4, 8, 16, 32), decimation in frequency.  Input in
// natural order in v[0], v[S], v[2S], ...  Output: frequency k lives in v[S * pos(k)].
// DIR=+1: forward kernel exp(-i...), DIR=-1: inverse kernel exp(+i...).  Unnormalised.
// ---------------------------------------------------------------------------------------
template <int R> struct RegFFT;

template <> struct RegFFT<1> {
    static constexpr int pos(int k) { return k; }
    template <int DIR, int S> static SLMGS_HD void run(cf*) {}
};

template <> struct RegFFT<2> {
    static constexpr int pos(int k) { return k; }
    template <int DIR, int S> static SLMGS_HD void run(cf* v) {
        cf a = v[0], b = v[S];
        v[0] = cadd(a, b);
        v[S] = csub(a, b);
    }
};

template <> struct RegFFT<4> {
    static constexpr int pos(int k) { return k; }
    template <int DIR, int S> static SLMGS_HD void run(cf* v) {
        cf a = v[0], b = v[S], c = v[2 * S], d = v[3 * S];
        cf t0 = cadd(a, c), t1 = csub(a, c), t2 = cadd(b, d), t3 = cmul_mi<DIR>(csub(b, d));
        v[0] = cadd(t0, t2);
        v[S] = cadd(t1, t3);
        v[2 * S] = csub(t0, t2);
        v[3 * S] = csub(t1, t3);
    }
};

// R = 4 * M: radix-4 over elements (j, j+M, j+2M, j+3M), twiddle, then M-point DFTs on
// each contiguous block of M.  X[k0 + 4 k'] sits at k0*M + RegFFT<M>::pos(k').
template <int R> struct RegFFT {
    static constexpr int M = R / 4;
    static constexpr int pos(int k) { return (k % 4) * M + RegFFT<M>::pos(k / 4); }

    template <int DIR, int S, int J> static SLMGS_HD void first(cf* v) {
        if constexpr (J < M) {
            RegFFT<4>::template run<DIR, S * M>(v + J * S);
            v[(J + M) * S] = ctwiddle_const<DIR, R, J>(v[(J + M) * S]);
            v[(J + 2 * M) * S] = ctwiddle_const<DIR, R, 2 * J>(v[(J + 2 * M) * S]);
            v[(J + 3 * M) * S] = ctwiddle_const<DIR, R, 3 * J>(v[(J + 3 * M) * S]);
            first<DIR, S, J + 1>(v);
        }
    }
    template <int DIR, int S> static SLMGS_HD void run(cf* v) {
        first<DIR, S, 0>(v);
        RegFFT<M>::template run<DIR, S>(v);
        RegFFT<M>::template run<DIR, S>(v + M * S);
        RegFFT<M>::template run<DIR, S>(v + 2 * M * S);
        RegFFT<M>::template run<DIR, S>(v + 3 * M * S);
    }
};

template <> struct RegFFT<8> {
    // 8 = 2 * 4: radix-2 over (j, j+4), twiddle w8^j on the odd half, then two 4-point DFTs.
    // X[k0 + 2 k'] at k0*4 + k'
    static constexpr int pos(int k) { return (k % 2) * 4 + (k / 2); }
    template <int DIR, int S> static SLMGS_HD void run(cf* v) {
        cf a0 = v[0], a1 = v[S], a2 = v[2 * S], a3 = v[3 * S];
        cf b0 = v[4 * S], b1 = v[5 * S], b2 = v[6 * S], b3 = v[7 * S];
        v[0] = cadd(a0, b0);
        v[S] = cadd(a1, b1);
        v[2 * S] = cadd(a2, b2);
        v[3 * S] = cadd(a3, b3);
        v[4 * S] = csub(a0, b0);
        v[5 * S] = ctwiddle_const<DIR, 8, 1>(csub(a1, b1));
        v[6 * S] = ctwiddle_const<DIR, 8, 2>(csub(a2, b2));
        v[7 * S] = ctwiddle_const<DIR, 8, 3>(csub(a3, b3));
        RegFFT<4>::template run<DIR, S>(v);
        RegFFT<4>::template run<DIR, S>(v + 4 * S);
    }
};

}  // namespace slmgs
