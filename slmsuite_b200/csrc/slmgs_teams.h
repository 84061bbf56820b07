// slmgs_teams.h -- the fused column kernel of the hot loop with TMA-staged tiles and two compute teams per SM.
//
// Why.  The plain fused column kernel (ColKernel<N, COL_FUSED, .., 2>, two 512-thread blocks per SM) spends about
// half of a tile's time in its global-memory phases: with the row-pair interleaved field a warp's load or store touches
// eight 128-byte lines (one 32-byte sector each), the LSU needs eight wavefronts per request and the two blocks of an SM,
// which start together and do equal work, stay in lock step -- both wait for memory, then both compete for the FMA
// pipe (clock-stamp traces: 7-8 k of a tile's 33 k cycles are the store + load burst, 5 k more the wait for the data;
// ncu: L1 data pipe ~60 % + FMA pipe ~40 % busy, adding up to the whole time instead of overlapping).
//
// What.  One persistent block per SM, two teams of 512 threads.  Each team owns an exchange buffer (the padded line
// layout of slmgs_fft.h) and walks over column tiles (2 columns x N rows).  The field never goes through the LSU:
//   * a tile comes in as eight TMA boxes {2 rows x 2 columns (one 32-byte sector) x 256 row pairs} of the row-pair
//     interleaved field (cp.async.bulk.tensor + mbarrier) into ONE staging buffer that the two teams use in turn:
//     the team that has read tile i out of it (conflict-free LDS) issues the load of tile i + 1 for its partner, so a
//     tile is always requested most of a tile-time before it is needed and the teams run half a period apart --
//     one team's butterflies run under the other team's shared-memory exchange;
//   * the result leaves through the team's own exchange buffer (free between the last inverse exchange and the next
//     tile's first one): dense store with STS, fence.proxy.async, eight TMA box stores.
// Arithmetic, register layout and the fused constraint are those of ColKernel (same functions): results are bit
// identical.  Under SLMGS_EMULATE the launcher runs the plain kernel.
#pragma once

#include "slmgs_kernels.h"

#ifndef SLMGS_EMULATE
#include <cuda.h>
#endif

namespace slmgs {

#ifndef SLMGS_EMULATE
SLMGS_DEVICE void tma_load_box2(void* smem_dst, const void* tmap, int x, int y, int z, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}
SLMGS_DEVICE void tma_store_box2(const void* tmap, int x, int y, int z, const void* smem_src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(smem_src))
                 : "memory");
}
SLMGS_DEVICE void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
SLMGS_DEVICE void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
SLMGS_DEVICE void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
SLMGS_DEVICE void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
SLMGS_DEVICE void bulk_prefetch_l2(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// The two teams of a block should run about half a tile-time apart (in phase they queue their exchanges together and
// then compete for the FMA pipe).  The offset is set once, by when team 0 requests team 1's FIRST tile: after barrier
// number SLMGS_TEAMS_STAGGER of its own first tile (0 = as early as possible); it then persists, both teams doing
// equal work per tile.
#ifndef SLMGS_TEAMS_STAGGER
#define SLMGS_TEAMS_STAGGER 0
#endif
template <int N, int VAR, bool DENSE> struct ColKernelT {
    typedef ColKernel<N, COL_FUSED, VAR, 2, DENSE> Base;
    typedef typename Base::F F;
    typedef ColArgs Args;
    static constexpr int C = 2, E = F::E, NS = F::NS;
    static constexpr int T = C * F::TPL;                   // threads of a team
    static constexpr int TILE_BYTES = N * C * (int)sizeof(cf);
    static constexpr int NWARP = T / 32;
    static constexpr int EXCH_BYTES = ((F::PADN * C * (int)sizeof(cf)) + 127) / 128 * 128;
    static_assert(NS == 3 && E == 16 && N >= 1024, "team column kernel: three radix stages, 16 points per thread");
    static_assert(EXCH_BYTES >= TILE_BYTES, "the output tile is staged in the exchange buffer");
    static constexpr int TW_BYTES = (F::TWS_A + F::TWS_B) * (int)sizeof(cf);  // compact twiddle tables (slmgs_fft.h)
    static size_t smem_bytes() { return (size_t)TILE_BYTES + 2 * (size_t)EXCH_BYTES + TW_BYTES + 64; }

    // dense [row pair][column][row parity] layout of a staged tile: index of (row n, column col)
    static SLMGS_DEVICE int sidx(int n, int col) { return (n >> 1) * 4 + col * 2 + (n & 1); }

    static SLMGS_DEVICE void team_bar(int team) {
        SLMGS_PP_STAMP(team, 2);
        asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "n"(T) : "memory");
        SLMGS_PP_STAMP(team, 3);
    }

    // The boxes of a tile.  A box is {2 rows x 2 columns} x a.tb_pairs row pairs (the tensor map's box); the rolled rows that
    // hold the SLM are [0, lo) and [hi, N), so boxes 0 .. tb_lo-1 start at row pair j tb_pairs and the others at
    // tb_hi0 + (j - tb_lo) tb_pairs.  Dense fields: 8 boxes of 256 row pairs.  A box is issued by lane 0 of warp
    // j mod 16 of the team (a TMA issue holds the issuing thread for a few hundred cycles).
    static SLMGS_DEVICE int box_pair0(const Args& a, int j) { return j < a.tb_lo ? j * a.tb_pairs : a.tb_hi0 + (j - a.tb_lo) * a.tb_pairs; }
    static SLMGS_DEVICE void issue_load(const Args& a, const void* tmap, cf* stage, unsigned long long* bar, int q, int by, int tid) {
        if ((tid & 31) != 0) return;
        if (tid == 0) mbar_expect_tx(bar, (unsigned)(a.tb_n * a.tb_pairs * 4 * (int)sizeof(cf)));
        for (int j = tid >> 5; j < a.tb_n; j += NWARP) {
            const int p0 = box_pair0(a, j);
            tma_load_box2(stage + (size_t)p0 * 4, tmap, q * 4, p0, by, bar);
        }
    }
    static SLMGS_DEVICE void issue_store(const Args& a, const void* tmap, const cf* exch, int q, int by, int tid) {
        if ((tid & 31) != 0) return;
        bool any = false;
        for (int j = tid >> 5; j < a.tb_n; j += NWARP) {
            const int p0 = box_pair0(a, j);
            tma_store_box2(tmap, q * 4, p0, by, exch + (size_t)p0 * 4);
            any = true;
        }
        if (any) bulk_commit();
    }

    // the far-field image blocks of a tile (tile-major: contiguous) into L2, one tile ahead of their use
    static SLMGS_DEVICE void prefetch_images(const Args& a, int q, int by) {
        constexpr unsigned BYTES = (unsigned)(N * C * sizeof(float));
        const long long off = (long long)q * N * C;
        bulk_prefetch_l2(a.weights + (long long)by * a.img_bs + off, BYTES);
        if (VAR == VAR_POW || VAR == VAR_POW_STORED || VAR == VAR_GENERAL)
            bulk_prefetch_l2(a.target + (long long)by * a.target_bs + off, BYTES);
        if (VAR == VAR_POW_STORED) bulk_prefetch_l2(a.phase_ff + (long long)by * a.img_bs + off, BYTES);
    }

    static SLMGS_DEVICE void run(const Args& a, const void* tmap, unsigned char* smem_raw) {
        cf* stage = reinterpret_cast<cf*>(smem_raw);
        const int team = threadIdx.x / T;
        SLMGS_PP_STAMP(team, 6);
        cf* exch = reinterpret_cast<cf*>(smem_raw + TILE_BYTES + (size_t)team * EXCH_BYTES);
        cf* tws = reinterpret_cast<cf*>(smem_raw + TILE_BYTES + 2 * (size_t)EXCH_BYTES);
        unsigned long long* full = reinterpret_cast<unsigned long long*>(smem_raw + TILE_BYTES + 2 * (size_t)EXCH_BYTES + TW_BYTES);
        // the rows of the twiddle tables the radix stages read (m = 1, 2, 4, 8), once per block
        for (int e = threadIdx.x; e < F::TWS_A + F::TWS_B; e += 2 * T) {
            if (e < F::TWS_A) tws[e] = __ldg(a.twA + (1 << (e / F::M1)) * F::M1 + e % F::M1);
            else tws[e] = __ldg(a.twB + (1 << ((e - F::TWS_A) / F::R2)) * F::R2 + (e - F::TWS_A) % F::R2);
        }
        SmemTw twA, twB;
        twA.p = tws;
        twB.p = tws + F::TWS_A;
        ThreadId id;
        id.tid = threadIdx.x % T;
        id.nthreads = T;
        id.by = blockIdx.y;
        id.gx = a.W / C;
        id.it = 0;
        const int ntiles = a.W / C;
        const int n_my = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // tiles of this block
        const bool elected = (id.tid & 31) == 0;
        if (threadIdx.x == 0) {
            mbar_init(full + 0, 1);
            mbar_init(full + 1, 1);
        }
        // Programmatic dependent launch: everything above (twiddle rows, barriers, descriptor prefetch) overlaps the tail
        // of the previous kernel; no access to the field or the images before it has completed and flushed.
        asm volatile("griddepcontrol.wait;" ::: "memory");
        __syncthreads();
        if (team == 0 && n_my > 0) issue_load(a, tmap, stage, full + 0, blockIdx.x, id.by, id.tid);
        typename Base::State st;
        const cf* sp = stage + sidx(id.tid >> 1, id.tid & 1);  // (lt, col) = (tid / 2, tid % 2)
        for (int i = team, k = 0; i < n_my; i += 2, ++k) {
            const int q = (int)blockIdx.x + i * (int)gridDim.x;
            id.bx = q;
            const typename Base::Loc L = Base::locate_q(a, exch, id, q);
            // ---- tile in: staging buffer -> registers (first-stage elements: rows lt + (N/16) m) ----
            SLMGS_PP_STAMP(team, 0);
            mbar_wait(full + team, (unsigned)(k & 1));
            SLMGS_PP_STAMP(team, 1);
            SLMGS_UNROLL
            for (int u = 0; u < E / F::R0; ++u) {
                // SLM rows (rolled index n): ((n + N/2) mod N) - i0 in [0, h), as in ColKernel::load_rows
                const unsigned r0 = (unsigned)(L.lt + F::TPL * u + (N >> 1) - a.i0);
                SLMGS_UNROLL
                for (int m = 0; m < F::R0; ++m) {
                    const unsigned sr = ((r0 + (unsigned)((N / F::R0) * m) + (unsigned)a.i0) & (unsigned)(N - 1)) - (unsigned)a.i0;
                    const cf x = sp[(u * F::TPL + (N / F::R0) * m) * 2];
                    st.v[u * F::R0 + m] = (DENSE || sr < (unsigned)a.h) ? x : cmake(0.f, 0.f);
                }
            }
            F::template fwd_compute_u<0, 0>(st.v);
            F::template fwd_twiddle_u<0, 0>(st.v, L.lt, twA, twB);
            if (elected) bulk_wait_read0();  // this team's previous tile has left the exchange buffer
            team_bar(team);                  // ... and every thread of the team has read the staging buffer
            auto request_next = [&](int point) {
                if (i + 1 < n_my && point == (i == 0 ? SLMGS_TEAMS_STAGGER : 0)) {
                    issue_load(a, tmap, stage, full + (team ^ 1), q + (int)gridDim.x, id.by, id.tid);
                    if (id.tid == 32 * (NWARP - 1)) prefetch_images(a, q + (int)gridDim.x, id.by);
                }
            };
            request_next(0);
            F::template store_scrambled_u<0, 0>(st.v, L.lt, L.s, C);
            team_bar(team);
            request_next(1);
            NoSync sy;
            F::template fwd_stage_sy<1>(st.v, L.lt, twA, twB, L.s, C, sy);
            team_bar(team);
            request_next(2);
            Base::prefetch_images_head(a, L);
            F::template fwd_stage_sy<2>(st.v, L.lt, twA, twB, L.s, C, sy);
            Base::template constrain<false>(st, a, id, L);
            F::template inv_stage_sy<2>(st.v, L.lt, twA, twB, L.s, C, sy);
            team_bar(team);
            request_next(3);
            F::template inv_stage_sy<1>(st.v, L.lt, twA, twB, L.s, C, sy);
            team_bar(team);
            request_next(4);
            F::template inv_stage_sy<0>(st.v, L.lt, twA, twB, L.s, C, sy);
            team_bar(team);  // the exchange buffer has been read: it now takes the output tile
            cf* op = exch + sidx(id.tid >> 1, id.tid & 1);
            SLMGS_UNROLL
            for (int u = 0; u < E / F::R0; ++u) {
                SLMGS_UNROLL
                for (int m = 0; m < F::R0; ++m) op[(u * F::TPL + (N / F::R0) * m) * 2] = st.v[u * F::R0 + m];
            }
            fence_async_smem();
            team_bar(team);
            issue_store(a, tmap, exch, q, id.by, id.tid);
        }
        if (elected) bulk_wait0();
    }
};

// ------------------------------------------------------------------------------------------------------------------
// 8192-point columns on the team design (dense fields): decimation in time over the row parity.
//
// A 2-column tile of 8192 rows is 128 KB -- no room for a staging buffer and two exchange buffers in one SM, which is why
// the plain 8192 kernels (32 points per thread, one 512-thread block per SM, load -> compute -> store in series) run at a
// quarter of the HBM rate.  But in the row-pair interleaved field the even and the odd rows of ONE column sit side by side
// (16 bytes per row pair), so a 1-column tile is 64 KB and is exactly the shape the 4096 team kernel works on: two lines
// of 4096 points (line p = rows 2 j + p), interleaved element by element.  The team runs Fft<4096> on both lines -- same
// exchange buffer, same bank-conflict-free layout, same 16 points per thread in 64 registers -- and one radix-2 step
// between lane pairs (thread (lt, 0) and (lt, 1) are neighbours in a warp: SHFL.BFLY 1) joins them,
//     X[k]        = E[k] + W_8192^k O[k]          E'[k] =  X[k] + X[k + 4096]
//     X[k + 4096] = E[k] - W_8192^k O[k]          O'[k] = (X[k] - X[k + 4096]) conj(W_8192^k)
// with k = lt + 256 m, so W_8192^k = W_8192^lt (one table value per thread for the whole kernel) x W_32^m (a compile-time
// constant).  Thread (lt, p) then holds the far field at rows k + 4096 p of its column: the constraint and the far-field
// stores are ColKernel's own functions with the image offsets of those rows.  Tiles move as 16 TMA boxes of
// {1 column x 2 row parities (16 bytes), 256 row pairs}; the other half of every 32-byte sector belongs to the neighbouring
// column, which the neighbouring block works on at the same time (L2 serves the second request).
// MODE: COL_FUSED (forward, constraint, inverse, tile out through the exchange buffer) or COL_FWD (forward + far-field
// outputs: the pre-pass of the global-dependency weight updates, e.g. per-spot feedback).
//
// MEASURED AND LEFT OPT-IN (SLMGS_TEAMS8=1; tools/ab_teams8.py, B200, dense 8192^2): results agree with the plain kernels
// (1e-6 .. 1e-7 on well-conditioned targets) but it is slower -- GS 1229 -> 1306 us per iteration, WGS-Kim 1365 -> 2090,
// spot feedback 1574 -> 1511: the 16-byte TMA rows double the number of row requests per tile and every field / image
// access uses half a sector (the plain kernels with 1-column tiles lose in the same way: fused column kernel 548 -> 783 us).
// ------------------------------------------------------------------------------------------------------------------
template <int MODE, int VAR> struct ColKernelT8 {
    static constexpr int N2 = 8192, N = 4096;
    typedef ColKernel<N, MODE, VAR, 2, true> Base;
    typedef typename Base::F F;
    typedef ColArgs Args;
    static constexpr int C = 2, E = F::E, NS = F::NS;
    static constexpr int T = C * F::TPL;
    static constexpr int TILE_BYTES = N2 * (int)sizeof(cf);
    static constexpr int NWARP = T / 32;
    static constexpr int BOX_PAIRS = 256, NBOX = N / BOX_PAIRS;
    static constexpr int EXCH_BYTES = ((F::PADN * C * (int)sizeof(cf)) + 127) / 128 * 128;
    static_assert(NS == 3 && E == 16 && F::R0 == 16 && F::last_radix() == 16 && NBOX == NWARP, "Fft<4096>: 16 * 16 * 16");
    static_assert(EXCH_BYTES >= TILE_BYTES, "the output tile is staged in the exchange buffer");
    static constexpr int TW_BYTES = (F::TWS_A + F::TWS_B) * (int)sizeof(cf);
    static size_t smem_bytes() { return (size_t)TILE_BYTES + 2 * (size_t)EXCH_BYTES + TW_BYTES + 64; }

    static SLMGS_DEVICE void team_bar(int team) { asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "n"(T) : "memory"); }
    static SLMGS_DEVICE cf shfl1(cf x) {
        return make_float2(__shfl_xor_sync(0xffffffffu, x.x, 1), __shfl_xor_sync(0xffffffffu, x.y, 1));
    }
    // box j = row pairs [256 j, 256 j + 256) of column q: 512 consecutive staging elements; issued by lane 0 of warp j
    static SLMGS_DEVICE void issue_load(const void* tmap, cf* stage, unsigned long long* bar, int q, int by, int tid) {
        if ((tid & 31) != 0) return;
        if (tid == 0) mbar_expect_tx(bar, (unsigned)TILE_BYTES);
        const int j = tid >> 5;
        tma_load_box2(stage + (size_t)j * BOX_PAIRS * 2, tmap, q * 2, j * BOX_PAIRS, by, bar);
    }
    static SLMGS_DEVICE void issue_store(const void* tmap, const cf* exch, int q, int by, int tid) {
        if ((tid & 31) != 0) return;
        const int j = tid >> 5;
        tma_store_box2(tmap, q * 2, j * BOX_PAIRS, by, exch + (size_t)j * BOX_PAIRS * 2);
        bulk_commit();
    }
    // the far-field image block of a column PAIR (tile-major images, 2 columns per tile) into L2
    static SLMGS_DEVICE void prefetch_images(const Args& a, int q, int by) {
        if (q & 1) return;  // (the even neighbour fetches the block for both)
        constexpr unsigned BYTES = (unsigned)(N2 * 2 * sizeof(float));
        const long long off = (long long)(q >> 1) * N2 * 2;
        bulk_prefetch_l2(a.weights + (long long)by * a.img_bs + off, BYTES);
        if (MODE == COL_FWD || VAR == VAR_POW || VAR == VAR_POW_STORED || VAR == VAR_GENERAL)
            bulk_prefetch_l2(a.target + (long long)by * a.target_bs + off, BYTES);
        if (VAR == VAR_POW_STORED) bulk_prefetch_l2(a.phase_ff + (long long)by * a.img_bs + off, BYTES);
    }
    template <int M> static SLMGS_DEVICE void join(cf* v, cf wb, bool odd) {
        if constexpr (M < 16) {
            const cf t = ctwiddle_const<1, 32, M>(cmul(v[M], wb));  // W^k O[k] (odd lanes)
            const cf r = shfl1(odd ? t : v[M]);
            v[M] = odd ? csub(r, t) : cadd(v[M], r);
            join<M + 1>(v, wb, odd);
        }
    }
    template <int M> static SLMGS_DEVICE void split(cf* v, cf wb, bool odd) {
        if constexpr (M < 16) {
            const cf r = shfl1(v[M]);
            const cf d = odd ? csub(r, v[M]) : cadd(v[M], r);
            v[M] = odd ? ctwiddle_const<-1, 32, M>(cmulc(d, wb)) : d;
            split<M + 1>(v, wb, odd);
        }
    }

    static SLMGS_DEVICE void run(const Args& a, const void* tmap, unsigned char* smem_raw) {
        cf* stage = reinterpret_cast<cf*>(smem_raw);
        const int team = threadIdx.x / T;
        cf* exch = reinterpret_cast<cf*>(smem_raw + TILE_BYTES + (size_t)team * EXCH_BYTES);
        cf* tws = reinterpret_cast<cf*>(smem_raw + TILE_BYTES + 2 * (size_t)EXCH_BYTES);
        unsigned long long* full = reinterpret_cast<unsigned long long*>(smem_raw + TILE_BYTES + 2 * (size_t)EXCH_BYTES + TW_BYTES);
        // twiddle rows of the 4096-point lines (a.tw2A / a.tw2B: the tables of N / 2)
        for (int e = threadIdx.x; e < F::TWS_A + F::TWS_B; e += 2 * T) {
            if (e < F::TWS_A) tws[e] = __ldg(a.tw2A + (1 << (e / F::M1)) * F::M1 + e % F::M1);
            else tws[e] = __ldg(a.tw2B + (1 << ((e - F::TWS_A) / F::R2)) * F::R2 + (e - F::TWS_A) % F::R2);
        }
        SmemTw twA, twB;
        twA.p = tws;
        twB.p = tws + F::TWS_A;
        ThreadId id;
        id.tid = threadIdx.x % T;
        id.nthreads = T;
        id.by = blockIdx.y;
        id.gx = a.W;
        id.it = 0;
        const int ntiles = a.W;
        const int n_my = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const bool elected = (id.tid & 31) == 0;
        const bool odd = (id.tid & 1) != 0;
        const int lt = id.tid >> 1;
        // W_8192^lt: row k0 = 1 of the first twiddle table of the 8192-point plan (twA[k0 M1 + j] = W_N^(j k0), lt < M1)
        const cf wb = __ldg(a.twA + Fft<N2>::M1 + lt);
        if (threadIdx.x == 0) {
            mbar_init(full + 0, 1);
            mbar_init(full + 1, 1);
        }
        asm volatile("griddepcontrol.wait;" ::: "memory");  // (see ColKernelT::run)
        __syncthreads();
        if (team == 0 && n_my > 0) issue_load(tmap, stage, full + 0, blockIdx.x, id.by, id.tid);
        typename Base::State st;
        const cf* sp = stage + id.tid;  // staged element (row pair j, parity p) at 2 j + p; this thread: j = lt + 256 m
        for (int i = team, k = 0; i < n_my; i += 2, ++k) {
            const int q = (int)blockIdx.x + i * (int)gridDim.x;
            id.bx = q;
            typename Base::Loc L;
            L.C = C;
            L.col = id.tid & 1;
            L.lt = lt;
            L.gc = q;
            L.s = exch + L.col;
            L.fbase = 0;
            {   // image_index(k + 4096 p, q, 8192, 2): rows k + 4096 p of column q in the tile-major images
                const long long ioff = (long long)(q >> 1) * N2 * 2 + (long long)L.col * N * 2 + (q & 1);
                L.ibase = (long long)id.by * a.img_bs + ioff;
                L.tbase = (long long)id.by * a.target_bs + ioff;
            }
            mbar_wait(full + team, (unsigned)(k & 1));
            SLMGS_UNROLL
            for (int m = 0; m < 16; ++m) st.v[m] = sp[2 * F::TPL * m];
            F::template fwd_compute_u<0, 0>(st.v);
            F::template fwd_twiddle_u<0, 0>(st.v, lt, twA, twB);
            if (MODE == COL_FUSED && elected) bulk_wait_read0();  // this team's previous tile has left the exchange buffer
            team_bar(team);                                       // ... and every thread of the team has read the staging buffer
            if (i + 1 < n_my) {
                issue_load(tmap, stage, full + (team ^ 1), q + (int)gridDim.x, id.by, id.tid);
                if (id.tid == 32 * (NWARP - 1) + 1) prefetch_images(a, q + (int)gridDim.x, id.by);
            }
            F::template store_scrambled_u<0, 0>(st.v, lt, L.s, C);
            team_bar(team);
            NoSync sy;
            F::template fwd_stage_sy<1>(st.v, lt, twA, twB, L.s, C, sy);
            team_bar(team);
            if constexpr (MODE == COL_FUSED) Base::prefetch_images_head(a, L);
            F::template fwd_stage_sy<2>(st.v, lt, twA, twB, L.s, C, sy);
            join<0>(st.v, wb, odd);
            if constexpr (MODE == COL_FWD) {
                Base::store_farfield(st, a, id, L);
            } else {
                Base::template constrain<false>(st, a, id, L);
                split<0>(st.v, wb, odd);
                F::template inv_stage_sy<2>(st.v, lt, twA, twB, L.s, C, sy);
                team_bar(team);
                F::template inv_stage_sy<1>(st.v, lt, twA, twB, L.s, C, sy);
                team_bar(team);
                F::template inv_stage_sy<0>(st.v, lt, twA, twB, L.s, C, sy);
                team_bar(team);  // the exchange buffer has been read: it now takes the output tile
                cf* op = exch + id.tid;
                SLMGS_UNROLL
                for (int m = 0; m < 16; ++m) op[2 * F::TPL * m] = st.v[m];
                fence_async_smem();
                team_bar(team);
                issue_store(tmap, exch, q, id.by, id.tid);
            }
        }
        if (MODE == COL_FUSED && elected) bulk_wait0();
    }
};

// ------------------------------------------------------------------------------------------------------------------
// The fused row kernel with the same structure.  A team owns one ROW PAIR: in the row-pair interleaved field that is
// one contiguous block of 2 W complex values, so tile in / tile out are single 1-D bulk copies (cp.async.bulk), and a
// thread's element k of line l sits at staging index 2 k + l = tid + 2 TPL m: lane-linear, conflict free.
// ------------------------------------------------------------------------------------------------------------------
SLMGS_DEVICE void bulk_load_1d(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
SLMGS_DEVICE void bulk_store_1d(void* gdst, const void* smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}

template <int N, bool STORE, bool DENSE> struct RowKernelT {
    typedef RowKernel<N, ROW_FUSED, STORE, false, 2, DENSE> Base;
    typedef typename Base::F F;
    typedef RowArgs Args;
    static constexpr int LI = 2, E = F::E, NS = F::NS;
    static constexpr int T = LI * F::TPL;
    static constexpr int NCHUNK = 8;                       // bulk copies per tile (one issuing lane each)
    static constexpr int TILE_BYTES = N * LI * (int)sizeof(cf);
    static constexpr int CHUNK_BYTES = TILE_BYTES / NCHUNK;
    static constexpr int EXCH_BYTES = ((F::PADN * LI * (int)sizeof(cf)) + 127) / 128 * 128;
    static constexpr int TW_BYTES = (F::TWS_A + F::TWS_B) * (int)sizeof(cf);
    static_assert(NS == 3 && E == 16 && N >= 2048, "team row kernel: three radix stages, 16 points per thread");
    static size_t smem_bytes() { return (size_t)TILE_BYTES + 2 * (size_t)EXCH_BYTES + TW_BYTES + 64; }

    static SLMGS_DEVICE void team_bar(int team) {
        SLMGS_PP_STAMP(team, 2);
        asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "n"(T) : "memory");
        SLMGS_PP_STAMP(team, 3);
    }
    static SLMGS_DEVICE bool issuer(int tid) { return (tid & 31) == 0 && (tid >> 5) < NCHUNK; }
    // the row pair of work item q (SLM rows 2 q, 2 q + 1) in the field
    static SLMGS_DEVICE cf* pair_ptr(const Args& a, int q, int by) {
        const int fr = (2 * q + a.i0 + (a.H >> 1)) & (a.H - 1);
        return a.fld + (long long)by * a.fld_bs + (long long)(fr >> 1) * a.W * 2;
    }
    static SLMGS_DEVICE void issue_load(const Args& a, cf* stage, unsigned long long* bar, int q, int by, int tid) {
        const int b = tid >> 5;
        if (b == 0) mbar_expect_tx(bar, (unsigned)TILE_BYTES);
        bulk_load_1d(reinterpret_cast<char*>(stage) + (size_t)b * CHUNK_BYTES,
                     reinterpret_cast<const char*>(pair_ptr(a, q, by)) + (size_t)b * CHUNK_BYTES, (unsigned)CHUNK_BYTES, bar);
    }

    static SLMGS_DEVICE void run(const Args& a, const void*, unsigned char* smem_raw) {
        cf* stage = reinterpret_cast<cf*>(smem_raw);
        const int team = threadIdx.x / T;
        SLMGS_PP_STAMP(team, 6);
        cf* exch = reinterpret_cast<cf*>(smem_raw + TILE_BYTES + (size_t)team * EXCH_BYTES);
        cf* tws = reinterpret_cast<cf*>(smem_raw + TILE_BYTES + 2 * (size_t)EXCH_BYTES);
        unsigned long long* full = reinterpret_cast<unsigned long long*>(smem_raw + TILE_BYTES + 2 * (size_t)EXCH_BYTES + TW_BYTES);
        for (int e = threadIdx.x; e < F::TWS_A + F::TWS_B; e += 2 * T) {
            if (e < F::TWS_A) tws[e] = __ldg(a.twA + (1 << (e / F::M1)) * F::M1 + e % F::M1);
            else tws[e] = __ldg(a.twB + (1 << ((e - F::TWS_A) / F::R2)) * F::R2 + (e - F::TWS_A) % F::R2);
        }
        SmemTw twA, twB;
        twA.p = tws;
        twB.p = tws + F::TWS_A;
        ThreadId id;
        id.tid = threadIdx.x % T;
        id.nthreads = T;
        id.by = blockIdx.y;
        const int items = (a.h + LI - 1) / LI;
        id.gx = items;
        id.it = 0;
        const int n_my = (items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const bool elected = issuer(id.tid);
        if (threadIdx.x == 0) {
            mbar_init(full + 0, 1);
            mbar_init(full + 1, 1);
        }
        asm volatile("griddepcontrol.wait;" ::: "memory");  // (see ColKernelT::run)
        if (threadIdx.x == 0 && blockIdx.x == 0) {  // (RowKernel::phase<0>: accumulator slots of the column kernel that follows)
            if (a.zero_acc) a.zero_acc[(long long)id.by * a.zero_bs] = 0.0;
            if (a.zero_acc2) a.zero_acc2[(long long)id.by * a.zero_bs] = 0.0;
            if (a.win_dst) a.win_dst[id.by] = (float)(1.0 / sqrt(a.win_src[(long long)id.by * a.zero_bs]));
        }
        __syncthreads();
        if (team == 0 && elected && n_my > 0) issue_load(a, stage, full + 0, blockIdx.x, id.by, id.tid);
        typename Base::State st;
        constexpr int RL = F::last_radix();
        for (int i = team, k = 0; i < n_my; i += 2, ++k) {
            const int q = (int)blockIdx.x + i * (int)gridDim.x;
            id.bx = q;
            const typename Base::Loc L = Base::locate(a, exch, id);
            SLMGS_PP_STAMP(team, 0);
            mbar_wait(full + team, (unsigned)(k & 1));
            SLMGS_PP_STAMP(team, 1);
            SLMGS_UNROLL
            for (int u = 0; u < E / RL; ++u) {
                SLMGS_UNROLL
                for (int m = 0; m < RL; ++m) {
                    const cf x = stage[id.tid + (u * F::TPL + (N / RL) * m) * LI];
                    st.v[u * RL + m] = (DENSE || L.active) ? x : cmake(0.f, 0.f);
                }
            }
            F::template inv_compute_u<NS - 1, 0>(st.v);
            if (elected) bulk_wait_read0();
            team_bar(team);
            auto request_next = [&](int point) {
                if (elected && i + 1 < n_my && point == (i == 0 ? SLMGS_TEAMS_STAGGER : 0))
                    issue_load(a, stage, full + (team ^ 1), q + (int)gridDim.x, id.by, id.tid);
            };
            request_next(0);
            SLMGS_PP_STAMP(team, 4);
            F::template store_scrambled_u<NS - 1, 0>(st.v, L.lt, L.s, LI);
            team_bar(team);
            request_next(1);
            NoSync sy;
            F::template inv_stage_sy<1>(st.v, L.lt, twA, twB, L.s, LI, sy);
            team_bar(team);
            request_next(2);
            F::template inv_stage_sy<0>(st.v, L.lt, twA, twB, L.s, LI, sy);
            Base::template project<true, STORE>(st, a, id, L);
            F::template fwd_stage_sy<0>(st.v, L.lt, twA, twB, L.s, LI, sy);
            team_bar(team);
            request_next(3);
            F::template fwd_stage_sy<1>(st.v, L.lt, twA, twB, L.s, LI, sy);
            team_bar(team);
            request_next(4);
            F::template fwd_stage_sy<2>(st.v, L.lt, twA, twB, L.s, LI, sy);
            team_bar(team);
            SLMGS_UNROLL
            for (int u = 0; u < E / RL; ++u) {
                SLMGS_UNROLL
                for (int m = 0; m < RL; ++m) exch[id.tid + (u * F::TPL + (N / RL) * m) * LI] = st.v[u * RL + m];
            }
            fence_async_smem();
            team_bar(team);
            if (elected) {
                const int b = id.tid >> 5;
                bulk_store_1d(reinterpret_cast<char*>(pair_ptr(a, q, id.by)) + (size_t)b * CHUNK_BYTES,
                              reinterpret_cast<const char*>(exch) + (size_t)b * CHUNK_BYTES, (unsigned)CHUNK_BYTES);
                bulk_commit();
            }
            SLMGS_PP_STAMP(team, 5);
        }
        if (elected) bulk_wait0();
    }
};

template <class K>
__global__ void __launch_bounds__(2 * K::T, 1) slmgs_kernel_teams_col(const typename K::Args a, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) unsigned char slmgs_smem_teams[];
    if (threadIdx.x == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    asm volatile("griddepcontrol.launch_dependents;");
    K::run(a, &tmap, slmgs_smem_teams);
}
template <class K> __global__ void __launch_bounds__(2 * K::T, 1) slmgs_kernel_teams_row(const typename K::Args a) {
    extern __shared__ __align__(128) unsigned char slmgs_smem_teams[];
    asm volatile("griddepcontrol.launch_dependents;");
    K::run(a, nullptr, slmgs_smem_teams);
}

template <class KernelFn, class... Params>
int launch_teams_impl(KernelFn kernel, bool* attr_set, int threads, size_t smem, int gx, int gy, cudaStream_t stream, bool pdl,
                      Params... params) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (e != cudaSuccess) return (int)e;
        attr_set[dev & 63] = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(gx, gy, 1);
    cfg.blockDim = dim3(threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return (int)cudaLaunchKernelEx(&cfg, kernel, params...);
}
// gx persistent blocks per hologram (gx * gy <= resident blocks of the GPU)
template <class K>
int launch_kernel_teams(int gx, int gy, cudaStream_t stream, const typename K::Args& a, const CUtensorMap& tmap, bool pdl = false) {
    static bool attr_set[64] = {false};
    return launch_teams_impl(slmgs_kernel_teams_col<K>, attr_set, 2 * K::T, K::smem_bytes(), gx, gy, stream, pdl, a, tmap);
}
template <class K> int launch_kernel_teams(int gx, int gy, cudaStream_t stream, const typename K::Args& a, bool pdl = false) {
    static bool attr_set[64] = {false};
    return launch_teams_impl(slmgs_kernel_teams_row<K>, attr_set, 2 * K::T, K::smem_bytes(), gx, gy, stream, pdl, a);
}
#endif  // !SLMGS_EMULATE

}  // namespace slmgs
