// pp_bench.cu -- development microbenchmark (build tool, not product): the two hot kernels of the dense GS loop at
// 4096^2 (ColKernel<4096, COL_FUSED, VAR_GS, 2, dense>, RowKernel<4096, ROW_FUSED, LI = 2, dense>) launched the
// way libslmgs.so launches them (two 512-thread blocks per SM) against the ping-pong-team launch (slmgs_kernel_pp:
// one persistent 1024-thread block per SM, two teams handing a token for the L1 data pipe back and forth).
// Checks that both give bit-identical fields, then times them with CUDA events.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DSLMGS_PACKED_F32X2
//             -DSLMGS_TW_PRODUCTS -I slmsuite_b200/csrc -o tools/micro/pp_bench tools/micro/pp_bench.cu
#include "slmgs_teams.h"
#include "pp_token.h"

#include <cuda.h>

#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

using namespace slmgs;

#define CK(x) do { cudaError_t e_ = (cudaError_t)(x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

#ifndef NN
#define NN 4096
#endif
typedef Fft<NN> F;

static void make_twiddles(std::vector<cf>& a, std::vector<cf>& b) {
    const int n = NN, r0 = F::R0, r1 = F::R1, r2 = F::R2, m1 = r1 * r2;
    a.resize(n);
    b.resize(m1);
    const double tau = 6.283185307179586476925286766559;
    for (int k0 = 0; k0 < r0; ++k0)
        for (int j = 0; j < m1; ++j) {
            const double ang = -tau * (double)((long long)j * k0 % n) / (double)n;
            a[k0 * m1 + j] = make_float2((float)cos(ang), (float)sin(ang));
        }
    for (int k1 = 0; k1 < r1; ++k1)
        for (int n2 = 0; n2 < r2; ++n2) {
            const double ang = -tau * (double)((n2 * k1) % m1) / (double)m1;
            b[k1 * r2 + n2] = make_float2((float)cos(ang), (float)sin(ang));
        }
}

typedef ColKernel<NN, COL_FUSED, VAR_GS, 2, true> KCol;
typedef RowKernel<NN, ROW_FUSED, false, false, 2, true> KRow;
typedef ColKernelT<NN, VAR_GS, true> KColT;
typedef RowKernelT<NN, false, true> KRowT;
typedef PPCol<KCol> KColPP;
typedef PPRow<KRow, 2> KRowPP;
typedef ColKernel<NN, COL_FUSED, VAR_POW, 2, true> KColPow;
typedef ColKernelT<NN, VAR_POW, true> KColPowT;
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// tensor map over the row-pair interleaved field: {W * 2 elements, H / 2 row pairs, 1}, box {4, 256, 1}
static CUtensorMap make_tmap(cf* fld, int H, int W) {
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)W * 2, (cuuint64_t)H / 2, 1};
    cuuint64_t strides[2] = {(cuuint64_t)W * 2 * sizeof(cf), (cuuint64_t)W * H * sizeof(cf)};
    cuuint32_t box[3] = {4, 256, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, fld, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
    return tm;
}

template <class Fn> static float time_ms(int reps, Fn fn) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    fn();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) fn();
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main(int argc, char** argv) {
    const int H = NN, W = NN;
    const size_t P = (size_t)H * W;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int reps = argc > 1 ? atoi(argv[1]) : 20;
    int pgx = argc > 2 ? atoi(argv[2]) : sms;

    std::vector<cf> ta, tb;
    make_twiddles(ta, tb);
    cf *twA, *twB, *fld0, *fld1, *fld2;
    float* weights;
    double* acc;
    CK(cudaMalloc(&twA, ta.size() * sizeof(cf)));
    CK(cudaMalloc(&twB, tb.size() * sizeof(cf)));
    CK(cudaMemcpy(twA, ta.data(), ta.size() * sizeof(cf), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(twB, tb.data(), tb.size() * sizeof(cf), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&fld0, P * sizeof(cf)));
    CK(cudaMalloc(&fld1, P * sizeof(cf)));
    CK(cudaMalloc(&fld2, P * sizeof(cf)));
    CK(cudaMalloc(&weights, P * sizeof(float)));
    CK(cudaMalloc(&acc, 64 * sizeof(double)));
    CK(cudaMemset(acc, 0, 64 * sizeof(double)));
    {
        std::vector<cf> h(P);
        std::vector<float> w(P);
        unsigned s = 12345u;
        auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)(s >> 8) * (1.0f / 16777216.0f); };
        for (size_t i = 0; i < P; ++i) { h[i] = make_float2(rnd() - 0.5f, rnd() - 0.5f); w[i] = rnd() + 0.01f; }
        CK(cudaMemcpy(fld0, h.data(), P * sizeof(cf), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(weights, w.data(), P * sizeof(float), cudaMemcpyHostToDevice));
    }

    ColArgs ca;
    memset(&ca, 0, sizeof ca);
    ca.fld_bs = (long long)P; ca.twA = twA; ca.twB = twB;
    ca.weights = weights; ca.target = weights; ca.phase_ff = weights; ca.amp_ff = weights;
    ca.img_bs = (long long)P; ca.target_bs = (long long)P; ca.acc = acc; ca.acc_bs = 8;
    ca.w_in_slot = -1; ca.w_out_slot = -1; ca.ratio_slot = -1; ca.wsq_slot = -1;
    ca.inv_npix = 1.0 / (double)P; ca.H = H; ca.W = W; ca.h = H; ca.i0 = 0;
    ca.scale = (float)(1.0 / sqrt((double)P));
    ca.wgs.method = METHOD_GS; ca.wgs.inv_fnorm = 1.0f; ca.wgs.neg_inv_mean = -1.0f; ca.zero_factor = 1.0f;
    ca.pairs = 1;
    ca.tb_pairs = 256; ca.tb_n = NN / 512; ca.tb_lo = NN / 512; ca.tb_hi0 = 0;
    RowArgs ra;
    memset(&ra, 0, sizeof ra);
    ra.fld_bs = (long long)P; ra.twA = twA; ra.twB = twB; ra.amp_scalar = 1.0f / 4096.0f;
    ra.scale = ca.scale; ra.H = H; ra.W = W; ra.h = H; ra.w = W; ra.zero_bs = 8; ra.pairs = 1;

    const size_t col_smem = KCol::smem_bytes(512), row_smem = KRow::smem_bytes(512);
    printf("smem per team: col %zu row %zu; SMs %d, persistent blocks %d\n", col_smem, row_smem, sms, pgx);

    // ---- correctness: plain vs ping-pong on the same input --------------------------------------------------------
    auto check = [&](const char* what) {
        std::vector<cf> a(P), b(P);
        CK(cudaMemcpy(a.data(), fld1, P * sizeof(cf), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(b.data(), fld2, P * sizeof(cf), cudaMemcpyDeviceToHost));
        size_t bad = 0;
        double nrm = 0;
        for (size_t i = 0; i < P; ++i) {
            if (memcmp(&a[i], &b[i], sizeof(cf)) != 0) ++bad;
            nrm += (double)a[i].x * a[i].x + (double)a[i].y * a[i].y;
        }
        printf("%s: %zu of %zu elements differ (norm %.6g)\n", what, bad, P, sqrt(nrm));
    };
    CK(cudaMemcpy(fld1, fld0, P * sizeof(cf), cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(fld2, fld0, P * sizeof(cf), cudaMemcpyDeviceToDevice));
    ca.fld = fld1;
    CK(launch_kernel<KCol>(W / 2, 1, 512, col_smem, 0, ca));
    ca.fld = fld2;
    CK(launch_kernel_pp<KColPP>(pgx, 1, col_smem, 0, ca));
    CK(cudaDeviceSynchronize());
    check("column kernel");
    ra.fld = fld1;
    CK(launch_kernel<KRow>(H / 2, 1, 512, row_smem, 0, ra));
    ra.fld = fld2;
    CK(launch_kernel_pp<KRowPP>(pgx, 1, row_smem, 0, ra));
    CK(cudaDeviceSynchronize());
    check("row kernel");
    CUtensorMap tmap1 = make_tmap(fld1, H, W);
    CUtensorMap tmap2 = make_tmap(fld2, H, W);
    CK(cudaMemcpy(fld1, fld0, P * sizeof(cf), cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(fld2, fld0, P * sizeof(cf), cudaMemcpyDeviceToDevice));
    ca.fld = fld1;
    CK(launch_kernel<KCol>(W / 2, 1, 512, col_smem, 0, ca));
    ca.fld = fld2;
    CK(launch_kernel_teams<KColT>(pgx, 1, 0, ca, tmap2));
    CK(cudaDeviceSynchronize());
    check("column kernel, TMA teams");
    ra.fld = fld1;
    CK(launch_kernel<KRow>(H / 2, 1, 512, row_smem, 0, ra));
    ra.fld = fld2;
    CK(launch_kernel_teams<KRowT>(pgx, 1, 0, ra));
    CK(cudaDeviceSynchronize());
    check("row kernel, TMA teams");
#ifdef SLMGS_PP_TRACE
    {   // one traced launch of the ping-pong column kernel, then of the row kernel
        for (int which = 0; which < 4; ++which) {
            int zero[2] = {0, 0};
            CK(cudaMemcpyToSymbol(slmgs_pp_trace_n, zero, sizeof zero));
            if (which == 3) { ra.fld = fld2; CK(launch_kernel_teams<KRowT>(pgx, 1, 0, ra)); }
            else if (which == 2) { ca.fld = fld2; CK(launch_kernel_teams<KColT>(pgx, 1, 0, ca, tmap2)); }
            else if (which == 0) { ca.fld = fld2; CK(launch_kernel_pp<KColPP>(pgx, 1, col_smem, 0, ca)); }
            else { ra.fld = fld2; CK(launch_kernel_pp<KRowPP>(pgx, 1, row_smem, 0, ra)); }
            CK(cudaDeviceSynchronize());
            static long long buf[2 * 512];
            int n[2];
            CK(cudaMemcpyFromSymbol(buf, slmgs_pp_trace, sizeof buf));
            CK(cudaMemcpyFromSymbol(n, slmgs_pp_trace_n, sizeof n));
            const long long t0 = buf[0] / 8;
            printf("== trace %s kernel (block 0; cycles since team 0's first stamp; a = wait for token, A = token acquired, R = released, b = team barrier passed)\n", which == 3 ? "row TMA teams" : which == 2 ? "column TMA teams (a = wait for tile, A = tile arrived, R = at barrier, b = barrier passed)" : which ? "row" : "column");
            for (int team = 0; team < 2; ++team) {
                printf("team %d:", team);
                for (int i = 0; i < n[team] && i < 80; ++i) {
                    const long long v = buf[team * 512 + i];
                    printf(" %c%lld", "aARbsScC"[v & 7], v / 8 - t0);
                }
                printf("\n");
            }
        }
    }
#endif
    ca.fld = fld1;

    // ---- timing ---------------------------------------------------------------------------------------------------
    ca.fld = fld1;
    ra.fld = fld1;
    const double colB = 20.0 * P, rowB = 16.0 * P;
    float t;
    t = time_ms(reps, [&]() { CK(launch_kernel<KCol>(W / 2, 1, 512, col_smem, 0, ca)); });
    printf("col plain     : %8.2f us  %7.1f GB/s\n", t * 1e3, colB / t * 1e-6);
    t = time_ms(reps, [&]() { CK(launch_kernel_pp<KColPP>(pgx, 1, col_smem, 0, ca)); });
    printf("col ping-pong : %8.2f us  %7.1f GB/s\n", t * 1e3, colB / t * 1e-6);
    t = time_ms(reps, [&]() { CK(launch_kernel_teams<KColT>(pgx, 1, 0, ca, tmap1)); });
    printf("col TMA teams : %8.2f us  %7.1f GB/s\n", t * 1e3, colB / t * 1e-6);
    t = time_ms(reps, [&]() { CK(launch_kernel<KRow>(H / 2, 1, 512, row_smem, 0, ra)); });
    printf("row plain     : %8.2f us  %7.1f GB/s\n", t * 1e3, rowB / t * 1e-6);
    t = time_ms(reps, [&]() { CK(launch_kernel_teams<KRowT>(pgx, 1, 0, ra)); });
    printf("row TMA teams : %8.2f us  %7.1f GB/s\n", t * 1e3, rowB / t * 1e-6);
    t = time_ms(reps, [&]() { CK(launch_kernel_pp<KRowPP>(pgx, 1, row_smem, 0, ra)); });
    printf("row ping-pong : %8.2f us  %7.1f GB/s\n", t * 1e3, rowB / t * 1e-6);
    {   // WGS (power-law update): weights read + written, target read -- timing only (the weights evolve)
        float* target;
        CK(cudaMalloc(&target, P * sizeof(float)));
        CK(cudaMemcpy(target, weights, P * sizeof(float), cudaMemcpyDeviceToDevice));
        ColArgs cw = ca;
        cw.target = target;
        cw.wgs_update = 1; cw.wgs.method = METHOD_KIM; cw.wgs.p = 0.8f; cw.w_out_slot = 0; cw.wgs.inv_fnorm = 1.0f;
        const double powB = 28.0 * P;
        t = time_ms(reps, [&]() { CK(launch_kernel<KColPow>(W / 2, 1, 512, col_smem, 0, cw)); });
        printf("col WGS plain     : %8.2f us  %7.1f GB/s\n", t * 1e3, powB / t * 1e-6);
        t = time_ms(reps, [&]() { CK(launch_kernel_teams<KColPowT>(pgx, 1, 0, cw, tmap1)); });
        printf("col WGS TMA teams : %8.2f us  %7.1f GB/s\n", t * 1e3, powB / t * 1e-6);
    }
    t = time_ms(reps, [&]() {
        CK(launch_kernel<KCol>(W / 2, 1, 512, col_smem, 0, ca, true));
        CK(launch_kernel<KRow>(H / 2, 1, 512, row_smem, 0, ra, true));
    });
    printf("iteration plain (PDL)     : %8.2f us -> %6.0f it/s\n", t * 1e3, 1e3 / t);
    t = time_ms(reps, [&]() {
        CK(launch_kernel_pp<KColPP>(pgx, 1, col_smem, 0, ca, true));
        CK(launch_kernel_pp<KRowPP>(pgx, 1, row_smem, 0, ra, true));
    });
    printf("iteration ping-pong (PDL) : %8.2f us -> %6.0f it/s\n", t * 1e3, 1e3 / t);
    t = time_ms(reps, [&]() {
        CK(launch_kernel_teams<KColT>(pgx, 1, 0, ca, tmap1, true));
        CK(launch_kernel<KRow>(H / 2, 1, 512, row_smem, 0, ra, true));
    });
    printf("iteration col TMA teams + row plain (PDL) : %8.2f us -> %6.0f it/s\n", t * 1e3, 1e3 / t);
    t = time_ms(reps, [&]() {
        CK(launch_kernel_teams<KColT>(pgx, 1, 0, ca, tmap1, true));
        CK(launch_kernel_teams<KRowT>(pgx, 1, 0, ra, true));
    });
    printf("iteration col + row TMA teams (PDL)       : %8.2f us -> %6.0f it/s\n", t * 1e3, 1e3 / t);
    return 0;
}
