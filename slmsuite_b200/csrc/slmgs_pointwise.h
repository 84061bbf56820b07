// slmgs_pointwise.h -- element-wise / reduction kernels used by the stepped (callback, stats,
// Nogrette, MRAF+WGS, spot-feedback) path and by uploads/downloads.  Same phase structure as
// the FFT kernels so the host emulation covers them too.
#pragma once

#include "slmgs_kernels.h"

namespace slmgs {

enum {
    EW_ROLL_F32 = 0,    // upload: dst[image_index(roll(y,x))] = src[y,x]   (centred row-major -> rolled tile-major)
    EW_ROLL_C64 = 1,    // download (a.unroll): dst[y,x] = src[image_index(roll(y,x))]
    EW_SUMSQ = 2,       // acc[slot0] += nansum(src^2)
    EW_RATIO_SUM = 3,   // acc[slot0] += sum(wgs_ratio(src=amp_ff, target))           (Nogrette mean)
    EW_WGS_UPDATE = 4,  // dst=weights <- wgs_apply(...); acc[slot1] += sum(w^2)
    EW_SCALE = 5,       // dst *= 1/sqrt(acc[slot0])
    EW_STATS1 = 6,      // acc[slot0..2] += sum f^2, nansum t^2, nansum t f           (_stats.py:51-76)
    EW_FILL_NAN0 = 7,   // dst = nan_to_num(src, nan=0)                                (reset_weights, :608-614)
    EW_ABS_C64 = 8,     // dst(f32) = |src(c64)|
    EW_PHASE2GRAY = 9,  // dst(u8/u16) = SLM gray level of get_phase() = src + pi (hardware/slms/slm.py:695-743)
    EW_ARG_C64 = 10,    // dst(f32) = arctan2(src.imag, src.real)   (MultiplaneHologram._nearfield_extract)
};

struct ElemArgs {
    const void* src;
    void* dst;
    const float* target;
    double* acc;  // [B][acc_bs]
    long long n;  // elements per hologram
    long long src_bs, dst_bs, target_bs;
    int acc_bs;
    int H, W;  // EW_ROLL_*
    int C;     // EW_ROLL_*: column-tile width of the device image layout
    int unroll;  // EW_ROLL_*: 0 = upload (host layout -> device layout), 1 = download
    int slot0, slot1, slot2;
    WgsParams wgs;
    int fnorm_slot;  // EW_WGS_UPDATE / EW_RATIO_SUM: inv_fnorm = 1/sqrt(acc[fnorm_slot]) if >= 0
    int mean_slot;   // EW_WGS_UPDATE: Nogrette mean = acc[mean_slot] / n
    // EW_PHASE2GRAY
    const double* corr;  // optional wavefront correction [n] (source["phase"], float64), shared by the batch
    double factor;       // -(bitresolution / 2 pi)
    int bitres;          // 2^bitdepth
    int out16;           // output uint16 (bitdepth > 8) or uint8
};

template <int OP> struct ElemKernel {
    typedef ElemArgs Args;
    static constexpr int MAXT = 256;
    static constexpr int NPHASE = 1;
    struct State {};
#ifndef SLMGS_EMULATE
    static SLMGS_DEVICE void barrier(const ThreadId&) { __syncthreads(); }
#endif

    template <int P> static SLMGS_DEVICE void phase(State&, const Args& a, cf*, const ThreadId& id) {
        double* acc = a.acc ? a.acc + (long long)id.by * a.acc_bs : nullptr;
        const float* srcf = reinterpret_cast<const float*>(a.src) + (long long)id.by * a.src_bs;
        const cf* srcc = reinterpret_cast<const cf*>(a.src) + (long long)id.by * a.src_bs;
        float* dstf = reinterpret_cast<float*>(a.dst) + (long long)id.by * a.dst_bs;
        cf* dstc = reinterpret_cast<cf*>(a.dst) + (long long)id.by * a.dst_bs;
        const float* tgt = a.target ? a.target + (long long)id.by * a.target_bs : nullptr;
        WgsParams q = a.wgs;
        if (OP == EW_WGS_UPDATE || OP == EW_RATIO_SUM) {
            if (a.fnorm_slot >= 0) q.inv_fnorm = (float)(1.0 / sqrt(acc[a.fnorm_slot]));
            if (OP == EW_WGS_UPDATE && a.mean_slot >= 0) q.neg_inv_mean = -(1.0f / (float)(acc[a.mean_slot] / (double)a.n));
        }
        float sc = 1.0f;
        if (OP == EW_SCALE) sc = (float)(1.0 / sqrt(acc[a.slot0]));
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        const long long stride = (long long)id.gx * id.nthreads;
        for (long long i = (long long)id.bx * id.nthreads + id.tid; i < a.n; i += stride) {
            if (OP == EW_ROLL_F32 || OP == EW_ROLL_C64) {
                // H, W and C are powers of two: shifts and masks instead of 64-bit divisions (which made this kernel
                // instruction bound: 124 instructions per element, 113 us for a 4096^2 image)
                const int lw = ilog2(a.W), lc = ilog2(a.C);
                const int y = (int)(i >> lw), x = (int)(i & (a.W - 1));
                const int ry = (y + (a.H >> 1)) & (a.H - 1), rx = (x + (a.W >> 1)) & (a.W - 1);
                const long long j = (((long long)(rx >> lc) * a.H + ry) << lc) + (rx & (a.C - 1));  // == image_index(ry, rx, H, C)
                if (OP == EW_ROLL_F32) {
                    if (a.unroll) dstf[i] = srcf[j];
                    else dstf[j] = srcf[i];
                } else {
                    if (a.unroll) dstc[i] = srcc[j];
                    else dstc[j] = srcc[i];
                }
            } else if (OP == EW_SUMSQ) {
                const float v = srcf[i];
                if (v == v) s0 += (double)v * (double)v;
            } else if (OP == EW_RATIO_SUM) {
                s0 += (double)wgs_ratio(srcf[i], tgt[i], q);
            } else if (OP == EW_WGS_UPDATE) {
                const float w = wgs_apply(dstf[i], wgs_multiplier(srcf[i], tgt[i], q));
                dstf[i] = w;
                s1 += (double)w * (double)w;
            } else if (OP == EW_SCALE) {
                dstf[i] *= sc;
            } else if (OP == EW_STATS1) {
                const float f = srcf[i], t = tgt[i];
                s0 += (double)f * (double)f;
                if (t == t) {
                    s1 += (double)t * (double)t;
                    s2 += (double)t * (double)f;
                }
            } else if (OP == EW_FILL_NAN0) {
                const float v = srcf[i];
                dstf[i] = (v == v) ? v : 0.0f;
            } else if (OP == EW_ABS_C64) {
                const cf z = srcc[i];
                dstf[i] = sqrtf(z.x * z.x + z.y * z.y);
            } else if (OP == EW_ARG_C64) {
                const cf z = srcc[i];
                dstf[i] = atan2f(z.y, z.x);
            } else if (OP == EW_PHASE2GRAY) {
                // get_phase() is float32 (phase + pi, _hologram.py:807-811); SLM.set_phase copies it into a float64
                // cache, adds the float64 correction, scales, rounds half-to-even, casts, subtracts one and masks.
                // The reference's shift by a multiple of 2*bitresolution (slm.py:729-733) is a no-op modulo
                // bitresolution and commutes with rint, so it is not reproduced.
                const float g = srcf[i] + 3.14159274101257324f;
                double p = (double)g;
                if (a.corr) p += a.corr[i];
                p = rint(p * a.factor);
                const long long v = (long long)p - 1;
                const unsigned u = (unsigned)(v & (long long)(a.bitres - 1));
                if (a.out16) reinterpret_cast<unsigned short*>(a.dst)[(long long)id.by * a.dst_bs + i] = (unsigned short)u;
                else reinterpret_cast<unsigned char*>(a.dst)[(long long)id.by * a.dst_bs + i] = (unsigned char)u;
            }
        }
        if (OP == EW_SUMSQ || OP == EW_RATIO_SUM) accum_add(acc + a.slot0, s0);
        if (OP == EW_WGS_UPDATE) accum_add(acc + a.slot1, s1);
        if (OP == EW_STATS1) {
            accum_add(acc + a.slot0, s0);
            accum_add(acc + a.slot1, s1);
            accum_add(acc + a.slot2, s2);
        }
    }
};

// ------------------------------------------------------------------------------------------
// Second statistics pass (_stats.py:78-101): over pixels with target power != 0 and not NaN,
//   ratio = (f^2/fsum)/(t^2/tsum) -> min, max;  err = t^2/tsum - f^2/fsum -> min, max, sum, sum^2, count
// One partial record of 8 doubles per block; the host finishes the reduction.
// ------------------------------------------------------------------------------------------
struct Stats2Args {
    const float* f;
    const float* t;
    double* partial;  // [B][gx][8]
    const double* acc;
    long long n, f_bs, t_bs;
    int acc_bs, fsum_slot, tsum_slot;
};

struct Stats2Kernel {
#ifndef SLMGS_EMULATE
    static SLMGS_DEVICE void barrier(const ThreadId&) { __syncthreads(); }
#endif
    typedef Stats2Args Args;
    static constexpr int MAXT = 256;
    static constexpr int NPHASE = 2;
    struct State {};

    template <int P> static SLMGS_DEVICE void phase(State&, const Args& a, cf* smem, const ThreadId& id) {
        double* sm = reinterpret_cast<double*>(smem);  // [nthreads][8]
        if (P == 0) {
            const double* acc = a.acc + (long long)id.by * a.acc_bs;
            const double fi = 1.0 / acc[a.fsum_slot], ti = 1.0 / acc[a.tsum_slot];
            const float* f = a.f + (long long)id.by * a.f_bs;
            const float* t = a.t + (long long)id.by * a.t_bs;
            double rmin = INFINITY, rmax = -INFINITY, emin = INFINITY, emax = -INFINITY, es = 0, es2 = 0, cnt = 0;
            const long long stride = (long long)id.gx * id.nthreads;
            for (long long i = (long long)id.bx * id.nthreads + id.tid; i < a.n; i += stride) {
                const float tv = t[i];
                const double tp = (double)tv * (double)tv * ti;
                if (tv == tv && tp != 0.0) {
                    const double fp = (double)f[i] * (double)f[i] * fi;
                    const double r = fp / tp, e = tp - fp;
                    rmin = r < rmin ? r : rmin;
                    rmax = r > rmax ? r : rmax;
                    emin = e < emin ? e : emin;
                    emax = e > emax ? e : emax;
                    es += e;
                    es2 += e * e;
                    cnt += 1;
                }
            }
            double* o = sm + (size_t)id.tid * 8;
            o[0] = rmin; o[1] = rmax; o[2] = emin; o[3] = emax; o[4] = es; o[5] = es2; o[6] = cnt; o[7] = 0;
        } else if (id.tid == 0) {
            double r[8] = {INFINITY, -INFINITY, INFINITY, -INFINITY, 0, 0, 0, 0};
            for (int t = 0; t < id.nthreads; ++t) {
                const double* o = sm + (size_t)t * 8;
                r[0] = o[0] < r[0] ? o[0] : r[0];
                r[1] = o[1] > r[1] ? o[1] : r[1];
                r[2] = o[2] < r[2] ? o[2] : r[2];
                r[3] = o[3] > r[3] ? o[3] : r[3];
                r[4] += o[4]; r[5] += o[5]; r[6] += o[6];
            }
            double* out = a.partial + ((long long)id.by * id.gx + id.bx) * 8;
            for (int k = 0; k < 8; ++k) out[k] = r[k];
        }
    }
};

// ------------------------------------------------------------------------------------------
// Spot feedback (SpotHologram._update_weights, _spots.py:1573-1624 + analysis.take,
// analysis/__init__.py:61-204 with centered=True, integrate=True, clip=False).
//   gather : pw[n] = sum over the w x w window round (x_n, y_n) of img^2, float64 accumulation,
//            offsets floor(arange(w) - (w-1)/2), negative indices wrap (NumPy semantics).
//            Coordinates are in the reference's centred convention; the image is rolled.
//   update : N-vector WGS update against spot_amp, normalised over the N spots, scattered back.
// ------------------------------------------------------------------------------------------
struct SpotArgs {
    const float* img;     // amp_ff [B][H][W] rolled
    float* weights;       // [B][H][W] rolled
    const int* sx;        // [N] integer spot x (centred convention)
    const int* sy;
    const float* spot_amp;  // [N] target amplitudes (host float64 in the reference, rounded once here)
    double* pw;           // [B][N] window powers (output of gather, input of update)
    float* wn;            // [B][N] updated spot weights before the scatter (the reference's N-vector, _spots.py:1617-1624)
    const unsigned char* keep;  // [N] 1 = last spot that rounds to its pixel: numpy's fancy-index scatter keeps that one
    long long img_bs;
    int H, W, N, width;
    int C;  // column-tile width of the image layout
    WgsParams wgs;
};

struct SpotGatherKernel {
#ifndef SLMGS_EMULATE
    static SLMGS_DEVICE void barrier(const ThreadId&) { __syncthreads(); }
#endif
    typedef SpotArgs Args;
    static constexpr int MAXT = 256;
    static constexpr int NPHASE = 1;
    struct State {};
    template <int P> static SLMGS_DEVICE void phase(State&, const Args& a, cf*, const ThreadId& id) {
        const float* img = a.img + (long long)id.by * a.img_bs;
        for (int n = id.bx * id.nthreads + id.tid; n < a.N; n += id.gx * id.nthreads) {
            // floor(k - (w-1)/2): for odd w this is k - (w-1)/2, for even w it is k - w/2
            const int base = (a.width & 1) ? -((a.width - 1) / 2) : -(a.width / 2);
            double s = 0.0;
            for (int dy = 0; dy < a.width; ++dy) {
                int y = a.sy[n] + base + dy;
                if (y < 0) y += a.H;  // NumPy negative-index wrap
                const int ry = (y + (a.H >> 1)) % a.H;
                for (int dx = 0; dx < a.width; ++dx) {
                    int x = a.sx[n] + base + dx;
                    if (x < 0) x += a.W;
                    const int rx = (x + (a.W >> 1)) % a.W;
                    const float v = img[image_index(ry, rx, a.H, a.C)];
                    s += (double)(v * v);  // reference squares in float32, sums in float64
                }
            }
            a.pw[(long long)id.by * a.N + n] = s;
        }
    }
};

// single block per hologram
struct SpotUpdateKernel {
#ifndef SLMGS_EMULATE
    static SLMGS_DEVICE void barrier(const ThreadId&) { __syncthreads(); }
#endif
    typedef SpotArgs Args;
    static constexpr int MAXT = 1024;
    static constexpr int NPHASE = 10;
    struct State {};

    static SLMGS_DEVICE long long pix(const Args& a, int n) {
        const int ry = (a.sy[n] + (a.H >> 1)) % a.H, rx = (a.sx[n] + (a.W >> 1)) % a.W;
        return image_index(ry, rx, a.H, a.C);
    }
    // smem doubles: [0..nthreads) scratch, then [nthreads + k] block scalars (k < 8), then [nthreads + 8 + j] partial sums.
    // A block sum takes two phases: 32 threads add 1/32 of the scratch each (fixed order), then thread 0 adds the 32
    // partial sums -- two dependent chains of 32 double additions instead of one of 1024 (which was 5 us per sum on B200).
    template <int P> static SLMGS_DEVICE void phase(State&, const Args& a, cf* smem, const ThreadId& id) {
        double* sm = reinterpret_cast<double*>(smem);
        double* scal = sm + id.nthreads;  // 0: sum f^2, 1: sum ratio, 2: sum w^2
        double* part = scal + 8;
        const double* pw = a.pw + (long long)id.by * a.N;
        float* wts = a.weights + (long long)id.by * a.img_bs;
        float* wn = a.wn + (long long)id.by * a.N;
        WgsParams q = a.wgs;
        if (P == 0) {  // partial sum of feedback^2 (feedback = float32(sqrt(pw)))
            double s = 0.0;
            for (int n = id.tid; n < a.N; n += id.nthreads) {
                const float f = (float)sqrt(pw[n]);
                if (f == f) s += (double)f * (double)f;
            }
            sm[id.tid] = s;
        } else if (P == 1 || P == 4 || P == 7) {
            if (id.tid < 32) {
                const int chunk = (id.nthreads + 31) / 32;
                double s = 0.0;
                for (int t = id.tid * chunk; t < (id.tid + 1) * chunk && t < id.nthreads; ++t) s += sm[t];
                part[id.tid] = s;
            }
        } else if (P == 2 || P == 5 || P == 8) {
            if (id.tid == 0) {
                double s = 0.0;
                for (int t = 0; t < 32; ++t) s += part[t];
                scal[P / 3] = s;
            }
        } else if (P == 3) {  // Nogrette mean of the ratio
            q.inv_fnorm = (float)(1.0 / sqrt(scal[0]));
            double s = 0.0;
            if (q.method == METHOD_NOGRETTE)
                for (int n = id.tid; n < a.N; n += id.nthreads) s += (double)wgs_ratio((float)sqrt(pw[n]), a.spot_amp[n], q);
            sm[id.tid] = s;
        } else if (P == 6) {  // gather + update on the N-vector (no pixel is written yet: two spots may round to the
                              // same pixel, and both must start from its old weight), partial sum of w^2
            q.inv_fnorm = (float)(1.0 / sqrt(scal[0]));
            q.neg_inv_mean = -(1.0f / (float)(scal[1] / (double)a.N));
            double s = 0.0;
            for (int n = id.tid; n < a.N; n += id.nthreads) {
                const float w = wgs_apply(wts[pix(a, n)], wgs_multiplier((float)sqrt(pw[n]), a.spot_amp[n], q));
                wn[n] = w;
                s += (double)w * (double)w;
            }
            sm[id.tid] = s;
        } else if (P == 9) {  // normalise over the N spots (:1877 on the N-vector) and scatter: the last duplicate wins
            const float sc = (float)(1.0 / sqrt(scal[2]));
            for (int n = id.tid; n < a.N; n += id.nthreads)
                if (a.keep[n]) wts[pix(a, n)] = wn[n] * sc;
        }
    }
};

// ------------------------------------------------------------------------------------------
// Column-tile occupancy for the sparse fused loop.  The constrained far field farfield = weights *
// exp(i phase_ff) (_hologram.py:1601-1605) is identically zero wherever weights == 0 -- and a weight that
// is zero stays zero under every WGS update (:1870, W *= fc with fc finite) -- except in an MRAF noise
// region (target is NaN, :1643-1653), which passes the field through.  One block per (tile, hologram):
//   flags[tile] |= 1 if any weight of the tile is non-zero (or NaN), |= 2 if any target value is NaN,
//                |= 4 if any target value is non-zero (WGS-Nogrette's mean runs over the ratio |F| / T, which is 1 where T == 0).
// The images are tile-major, so a tile is one contiguous block of H*C floats.
// ------------------------------------------------------------------------------------------
struct TileArgs {
    const float* weights;
    const float* target;
    long long img_bs, target_bs;
    long long tile_elems;  // H * C
    int* flags;            // [B][flags_bs], zeroed by the caller
    int flags_bs;          // W / C
};

#ifndef SLMGS_EMULATE
SLMGS_DEVICE void atomic_or_int(int* p, int v) { atomicOr(p, v); }
#else
inline void atomic_or_int(int* p, int v) { *p |= v; }
#endif

struct TileFlagKernel {
#ifndef SLMGS_EMULATE
    static SLMGS_DEVICE void barrier(const ThreadId&) { __syncthreads(); }
#endif
    typedef TileArgs Args;
    static constexpr int MAXT = 256;
    static constexpr int NPHASE = 1;
    struct State {};
    template <int P> static SLMGS_DEVICE void phase(State&, const Args& a, cf*, const ThreadId& id) {
        const float* w = a.weights + (long long)id.by * a.img_bs + (long long)id.bx * a.tile_elems;
        const float* t = a.target + (long long)id.by * a.target_bs + (long long)id.bx * a.tile_elems;
        int f = 0;
        for (long long i = id.tid; i < a.tile_elems; i += id.nthreads) {
            const float wv = ld_stream(w + i), tv = ld_stream(t + i);
            if (!(wv == 0.0f)) f |= 1;
            if (tv != tv) f |= 2;
            else if (tv != 0.0f) f |= 4;
        }
        if (f) atomic_or_int(a.flags + (long long)id.by * a.flags_bs + id.bx, f);
    }
};

// ------------------------------------------------------------------------------------------
// Camera sampling of the far-field intensity ("next" row):
// SimulatedCamera._get_image_hw, hardware/cameras/simulated.py:344-402:
//     img = map_coordinates(|farfield|^2, knm_cam, order=0);  img *= exposure * gain;
//     img[img > bitresolution - 1] = bitresolution - 1;  img.astype(dtype)
// scipy.ndimage.map_coordinates(order=0, mode="constant", cval=0): a coordinate outside [0, len-1] gives 0,
// otherwise the sample at floor(c + 0.5).  Coordinates are float64 (y, x) in the reference's centred far-field
// convention; amp_ff is rolled and tile-major on the device.
// ------------------------------------------------------------------------------------------
struct SampleArgs {
    const float* amp_ff;  // [B][H][W] rolled, tile-major
    long long img_bs;
    const double* ky;     // [n]
    const double* kx;     // [n]
    long long n;
    void* out;            // [B][n] float32 / uint8 / uint16
    int out_kind;         // 0 float32, 1 uint8, 2 uint16
    float scale;          // exposure * gain (applied in float32 like the reference's in-place multiply)
    float clip_max;       // bitresolution - 1, or < 0 for no clipping
    int H, W, C;
};

struct SampleKernel {
#ifndef SLMGS_EMULATE
    static SLMGS_DEVICE void barrier(const ThreadId&) { __syncthreads(); }
#endif
    typedef SampleArgs Args;
    static constexpr int MAXT = 256;
    static constexpr int NPHASE = 1;
    struct State {};
    template <int P> static SLMGS_DEVICE void phase(State&, const Args& a, cf*, const ThreadId& id) {
        const float* img = a.amp_ff + (long long)id.by * a.img_bs;
        const long long stride = (long long)id.gx * id.nthreads;
        for (long long i = (long long)id.bx * id.nthreads + id.tid; i < a.n; i += stride) {
            const double cy = a.ky[i], cx = a.kx[i];
            float v = 0.0f;
            if (cy >= 0.0 && cy <= (double)(a.H - 1) && cx >= 0.0 && cx <= (double)(a.W - 1)) {
                const int y = (int)floor(cy + 0.5), x = (int)floor(cx + 0.5);
                const float f = img[image_index((y + (a.H >> 1)) % a.H, (x + (a.W >> 1)) % a.W, a.H, a.C)];
                v = f * f;
            }
            v = v * a.scale;
            if (a.clip_max >= 0.0f && v > a.clip_max) v = a.clip_max;
            const long long o = (long long)id.by * a.n + i;
            if (a.out_kind == 0) reinterpret_cast<float*>(a.out)[o] = v;
            else if (a.out_kind == 1) reinterpret_cast<unsigned char*>(a.out)[o] = (unsigned char)v;
            else reinterpret_cast<unsigned short*>(a.out)[o] = (unsigned short)v;
        }
    }
};

}  // namespace slmgs
