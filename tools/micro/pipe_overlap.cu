// pipe_overlap.cu -- microbenchmark (build tool, not product): do packed-FP32 math (FFMA2) and shared-memory exchange
// traffic (LDS.64 / STS.64) overlap on one SM when different warps issue them?  One 1024-thread block per SM; the
// first half of the warps runs a math loop, the second half a shared-memory loop; each is also run alone.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/pipe_overlap tools/micro/pipe_overlap.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

template <int MATHMODE>  // 0 = FFMA2 (packed), 1 = scalar FFMA
__device__ __forceinline__ float2 math_loop(int iters, float2 seed) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(seed.x + i, seed.y - i);
    const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(1e-3f, -1e-3f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MATHMODE == 0) a[i] = __ffma2_rn(a[i], m, c);
                else if (MATHMODE == 2) a[i] = __fadd2_rn(a[i], a[(i + 3) & 7]);
                else if (MATHMODE == 3) a[i] = __ffma2_rn(a[i], a[(i + 1) & 7], a[(i + 3) & 7]);
                else if (MATHMODE == 4) a[i] = __fmul2_rn(a[i], a[(i + 3) & 7]);
                else if (MATHMODE == 6) {  // full complex multiply, the library's form: FMUL2 + FFMA2 sharing the operand a
                    const float2 x = a[i], w = a[(i + 3) & 7];
                    a[i] = __ffma2_rn(make_float2(-x.y, x.x), make_float2(w.y, w.y), __fmul2_rn(x, make_float2(w.x, w.x)));
                } else if (MATHMODE == 7) {  // the same with scalar instructions
                    const float2 x = a[i], w = a[(i + 3) & 7];
                    a[i] = make_float2(fmaf(-x.y, w.y, x.x * w.x), fmaf(x.y, w.x, x.x * w.y));
                }
                else if (MATHMODE == 5) a[i] = __ffma2_rn(make_float2(-a[(i + 1) & 7].y, a[(i + 1) & 7].x), make_float2(a[(i + 2) & 7].y, a[(i + 2) & 7].y), a[i]);
                else { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }
            }
        }
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s.x += a[i].x; s.y += a[i].y; }
    return s;
}
// radix-2 butterflies: (a, b) -> (a + b, a - b): both FADD2 of a pair read the same two register pairs (operand reuse cache)
__device__ __forceinline__ float2 bfly_loop(int iters, float2 seed) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(seed.x + i, seed.y - i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int j = i + 4;
                const float2 x = __fadd2_rn(a[i], a[j]);
                const float2 y = __fadd2_rn(a[i], make_float2(-a[j].x, -a[j].y));
                a[i] = x;
                a[j] = make_float2(y.x * 0.5f, y.y * 0.5f) ;
                a[i] = make_float2(a[i].x * 0.5f, a[i].y * 0.5f);
            }
        }
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s.x += a[i].x; s.y += a[i].y; }
    return s;
}
__global__ void __launch_bounds__(1024, 1) kb(int mi, float2* out) {
    const int w = threadIdx.x >> 5;
    float2 r = make_float2(0.f, 0.f);
    if (w < 16) r = bfly_loop(mi, make_float2((float)threadIdx.x, 1.0f));
    if (r.x == 123.456f) out[threadIdx.x] = r;
}

// 16 x (LDS.64 + STS.64) per iteration, conflict free (consecutive lanes, consecutive 8-byte words)
__device__ __forceinline__ float2 smem_loop(int iters, float2* s, int lane_base) {
    float2 v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = make_float2((float)i, (float)lane_base);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) s[lane_base + i * 512] = v[i];
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = s[lane_base + ((i + 1) & 15) * 512];
        __syncwarp();
    }
    float2 r = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 16; ++i) { r.x += v[i].x; r.y += v[i].y; }
    return r;
}

// mode bit 0: math warps active, bit 1: smem warps active
template <int MATHMODE> __global__ void __launch_bounds__(1024, 1) k(int mode, int mi, int si, float2* out) {
    extern __shared__ float2 sm[];
    const int w = threadIdx.x >> 5;
    float2 r = make_float2(0.f, 0.f);
    if (w < 16) {
        if (mode & 1) r = math_loop<MATHMODE>(mi, make_float2((float)threadIdx.x, 1.0f));
    } else {
        if (mode & 2) r = smem_loop(si, sm, (threadIdx.x - 512));
    }
    if (r.x == 123.456f) out[threadIdx.x] = r;
}

template <int MATHMODE> static float run(int mode, int mi, int si, float2* out) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    k<MATHMODE><<<148, 1024, 16 * 512 * 8>>>(mode, mi, si, out);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    k<MATHMODE><<<148, 1024, 16 * 512 * 8>>>(mode, mi, si, out);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms * 1e3f;
}

int main() {
    float2* out;
    CK(cudaMalloc(&out, 1024 * sizeof(float2)));
    CK(cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 512 * 8));
    CK(cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 512 * 8));
    CK(cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 512 * 8));
    CK(cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 512 * 8));
    CK(cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 512 * 8));
    CK(cudaFuncSetAttribute(k<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 512 * 8));
    CK(cudaFuncSetAttribute(k<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 512 * 8));
    CK(cudaFuncSetAttribute(k<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 512 * 8));
    const int mi = 2000, si = 1000;
    // math: 16 warps x mi x 32 instructions; smem: 16 warps x si x 32 instructions (2 wavefronts each)
    {
        const float tm = run<0>(1, mi, si, out), ts = run<0>(2, mi, si, out), tb = run<0>(3, mi, si, out);
        printf("FFMA2 : math alone %.1f us (%.2f clk per warp-instr per SMSP), smem alone %.1f us (%.2f clk per LDS/STS.64 per SM), both %.1f us (sum %.1f, max %.1f)\n",
               tm, tm * 1965.0 / (4.0 * mi * 32), ts, ts * 1965.0 / (16.0 * si * 32), tb, tm + ts, tm > ts ? tm : ts);
    }
    {
        const float t2 = run<2>(1, mi, si, out), t3 = run<3>(1, mi, si, out), t4 = run<4>(1, mi, si, out), t5 = run<5>(1, mi, si, out);
        const float b2 = run<2>(3, mi, si, out), b3 = run<3>(3, mi, si, out);
        const double f = 1965.0 / (4.0 * mi * 32);
        printf("distinct operands, clk per warp-instr per SMSP: FADD2 %.2f  FFMA2(3 regs) %.2f  FMUL2 %.2f  FFMA2(swizzled cmul form) %.2f\n", t2 * f, t3 * f, t4 * f, t5 * f);
        {
            const float t6 = run<6>(1, mi, si, out), t7 = run<7>(1, mi, si, out);
            printf("complex multiply: packed (FMUL2 + FFMA2) %.2f clk per multiply, scalar (2 FMUL + 2 FFMA) %.2f clk per multiply\n", t6 * f, t7 * f);
        }
        printf("with the shared-memory warps running too: FADD2 %.1f us (alone %.1f), FFMA2 %.1f us (alone %.1f)\n", b2, t2, b3, t3);
    }
    {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        kb<<<148, 1024>>>(mi, out);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        kb<<<148, 1024>>>(mi, out);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        // per iteration and warp: 16 butterflies = 32 FADD2 + 32 FMUL2 (scaling by 0.5 keeps the values finite)
        printf("butterfly pairs (a+b, a-b) + 2 FMUL2 by an immediate each: %.2f clk per packed instruction per SMSP\n",
               ms * 1e3 * 1965.0 / (4.0 * mi * 64));
    }
    {
        const float tm = run<1>(1, mi, si, out), ts = run<1>(2, mi, si, out), tb = run<1>(3, mi, si, out);
        printf("FFMA  : math alone %.1f us (%.2f clk per warp-instr per SMSP), smem alone %.1f us, both %.1f us (sum %.1f, max %.1f)\n",
               tm, tm * 1965.0 / (4.0 * mi * 64), ts, tb, tm + ts, tm > ts ? tm : ts);
    }
    return 0;
}
