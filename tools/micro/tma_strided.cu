// tma_strided.cu -- microbenchmark (build tool, not product): how fast does TMA move NARROW column tiles of a
// row-major complex64 field?  One block loads a tile of C columns x ROWS rows as boxes {C x 256 rows} into shared
// memory (cp.async.bulk.tensor.2d, mbarrier), then stores it back with TMA; persistent over tiles.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_strided tma_strided.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) copy_kernel(const __grid_constant__ CUtensorMap src, const __grid_constant__ CUtensorMap dst,
                                                   int ntiles, int rows, int cbytes, int nbuf) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar[4];
    const int tile_bytes = rows * cbytes;
    const int ccols = cbytes / 8;
    if (threadIdx.x == 0) {
        for (int i = 0; i < nbuf; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    // software pipeline over this block's tiles with nbuf buffers: load(i + nbuf - 1) is in flight while tile i is stored
    int n_my = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) ++n_my;
    unsigned phase_bits = 0;
    auto issue_load = [&](int k) {
        const int t = blockIdx.x + k * gridDim.x;
        const int b = k % nbuf;
        // the buffer was read by an earlier store: wait for those reads
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[b])), "r"(tile_bytes) : "memory");
        for (int r = 0; r < rows; r += 256) {
            asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(s32(smem + (size_t)b * tile_bytes + (size_t)r * cbytes)), "l"(&src), "r"(t * ccols), "r"(r), "r"(s32(&bar[b]))
                         : "memory");
        }
    };
    const int pre = nbuf - 1 > 0 ? nbuf - 1 : 1;
    for (int k = 0; k < pre && k < n_my; ++k) issue_load(k);
    for (int k = 0; k < n_my; ++k) {
        const int b = k % nbuf;
        if (nbuf > 1 && k + pre < n_my) issue_load(k + pre);
        // wait for tile k
        const unsigned ph = (phase_bits >> b) & 1u;
        unsigned done = 0;
        while (!done) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(s32(&bar[b])), "r"(ph) : "memory");
        }
        phase_bits ^= (1u << b);
        const int t = blockIdx.x + k * gridDim.x;
        for (int r = 0; r < rows; r += 256) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                         ::"l"(&dst), "r"(t * ccols), "r"(r), "r"(s32(smem + (size_t)b * tile_bytes + (size_t)r * cbytes)) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (nbuf == 1 && k + 1 < n_my) issue_load(k + 1);
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// plain LDG/STG copy of the same tiles for comparison: C columns x rows, lanes = (rows x C)
__global__ void __launch_bounds__(1024) ldg_kernel(const float2* src, float2* dst, int W, int H, int C) {
    const int tile = blockIdx.x;
    const int col = threadIdx.x % C, r0 = threadIdx.x / C, rstep = blockDim.x / C;
    for (int r = r0; r < H; r += rstep) {
        const size_t i = (size_t)r * W + tile * C + col;
        dst[i] = src[i];
    }
}

int main(int argc, char** argv) {
    const int H = 4096, W = 4096;
    float2 *a, *b;
    CK(cudaMalloc(&a, (size_t)H * W * 8));
    CK(cudaMalloc(&b, (size_t)H * W * 8));
    CK(cudaMemset(a, 1, (size_t)H * W * 8));
    CK(cudaMemset(b, 0, (size_t)H * W * 8));
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    CK(cudaFuncSetAttribute(copy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int cs[] = {2, 4, 8, 16};
    for (int ci = 0; ci < 4; ++ci) {
        const int C = cs[ci];
        for (int swz = 0; swz < 1; ++swz) {
            CUtensorMap ms, md;
            cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
            cuuint64_t strides[1] = {(cuuint64_t)W * 8};
            cuuint32_t box[2] = {(cuuint32_t)C, 256};
            cuuint32_t es[2] = {1, 1};
            CUtensorMapSwizzle sw = swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
            // element = 8 bytes: use FLOAT64 as an opaque 8-byte type
            CUresult r1 = encode(&ms, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, a, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                 CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            CUresult r2 = encode(&md, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, b, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                 CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r1 || r2) { printf("encode failed C=%d swz=%d: %d %d\n", C, swz, (int)r1, (int)r2); continue; }
            for (int rows = 1024; rows <= 4096; rows *= 2) {
                const int tile_bytes = rows * C * 8;
                for (int nbuf = 1; nbuf <= 3; ++nbuf) {
                    for (int bps = 1; bps <= 2; ++bps) {
                        const size_t smem = (size_t)tile_bytes * nbuf;
                        if (smem * bps > 220 * 1024) continue;
                        if (rows != 4096) continue;  // tiles are whole columns here (a 2-D tile loop would be needed otherwise)
                        const int ntiles = W / C;
                        const int grid = 148 * bps;
                        for (int it = 0; it < 3; ++it) copy_kernel<<<grid, 128, smem>>>(ms, md, ntiles, rows, C * 8, nbuf);
                        CK(cudaDeviceSynchronize());
                        CK(cudaEventRecord(e0));
                        const int reps = 10;
                        for (int it = 0; it < reps; ++it) copy_kernel<<<grid, 128, smem>>>(ms, md, ntiles, rows, C * 8, nbuf);
                        CK(cudaEventRecord(e1));
                        CK(cudaEventSynchronize(e1));
                        float ms_ = 0;
                        CK(cudaEventElapsedTime(&ms_, e0, e1));
                        const double us = ms_ * 1000.0 / reps;
                        printf("TMA  C=%2d (%3d B rows) swz=%d nbuf=%d blocks/SM=%d : %8.1f us  %7.1f GB/s (read+write)\n", C, C * 8, swz, nbuf,
                               bps, us, 2.0 * H * W * 8 / us * 1e-3);
                    }
                }
            }
        }
    }
    for (int ci = 0; ci < 4; ++ci) {
        const int C = cs[ci];
        for (int it = 0; it < 3; ++it) ldg_kernel<<<W / C, 1024>>>(a, b, W, H, C);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int it = 0; it < 10; ++it) ldg_kernel<<<W / C, 1024>>>(a, b, W, H, C);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms_ = 0;
        CK(cudaEventElapsedTime(&ms_, e0, e1));
        printf("LDG  C=%2d : %8.1f us  %7.1f GB/s\n", C, ms_ * 100.0, 2.0 * H * W * 8 / (ms_ * 100.0) * 1e-3);
    }
    // check one value made it
    float2 hv;
    CK(cudaMemcpy(&hv, b + 12345, 8, cudaMemcpyDeviceToHost));
    printf("check: %08x\n", *(unsigned*)&hv.x);
    return 0;
}
