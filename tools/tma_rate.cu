// TMA throughput for the U access pattern: per CTA, stream tiles [rows x 128 points] of a time-major matrix U[m][ld].
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
// mode 0: 2D tensor box {128, rows}; mode 1: 1D bulk copies of 512 B per row
__global__ void __launch_bounds__(128) tma_kernel(const __grid_constant__ CUtensorMap tm, const float* U, long long ld, int m, int rows, int stages, int mode,
                                                  long long ntiles, long long* out, int P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[128];
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages * P; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bars[i])), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < P) {
        const int pid = threadIdx.x >> 5;
        const uint32_t stage_bytes = rows * 512;
        const int chunks_per_tile = m / rows;
        long long t0 = clock64();
        long long issued = 0, waited = 0;
        const long long total = ((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x) * chunks_per_tile / P;
        // incremental bookkeeping (no divisions in the loop)
        long long gi = pid; long long tile = blockIdx.x; int chunk = pid;      // chunk index inside the tile
        while (chunk >= chunks_per_tile) { chunk -= chunks_per_tile; tile += gridDim.x; }
        int st_i = 0, st_w = 0; uint32_t par_w = 0;
        while (waited < total) {
            while (issued < total && issued - waited < stages) {
                const int row0 = chunk * rows;
                const uint32_t dst = smem_u32(smem) + (pid * stages + st_i) * stage_bytes, bar = smem_u32(&bars[pid * stages + st_i]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(stage_bytes) : "memory");
                if (mode == 0) {
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                 :: "r"(dst), "l"(&tm), "r"((int)(tile * 128)), "r"(row0), "r"(bar) : "memory");
                } else {
                    for (int r = 0; r < rows; ++r)
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     :: "r"(dst + r * 512), "l"(U + (long long)(row0 + r) * ld + tile * 128), "r"(512u), "r"(bar) : "memory");
                }
                ++issued; gi += P; chunk += P;
                while (chunk >= chunks_per_tile) { chunk -= chunks_per_tile; tile += gridDim.x; }
                if (++st_i == stages) st_i = 0;
            }
            mbar_wait(smem_u32(&bars[pid * stages + st_w]), par_w);
            ++waited;
            if (++st_w == stages) { st_w = 0; par_w ^= 1; }
        }
        if (blockIdx.x == 0 && pid == 0) out[0] = clock64() - t0;
    }
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    const long long ld = 1 << 20; const int m = 1024;   // 4 GB
    float* U; CK(cudaMalloc(&U, sizeof(float) * ld * m)); CK(cudaMemset(U, 0, sizeof(float) * ld * m));
    long long* out; CK(cudaMallocManaged(&out, 64));
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int cfgs[][4] = {{8, 12, 0, 1}, {8, 6, 0, 2}, {8, 3, 0, 4}, {8, 6, 0, 4}, {16, 3, 0, 2}, {16, 3, 0, 4}, {32, 3, 0, 1}, {32, 2, 0, 2}, {8, 3, 1, 4}};
    for (auto& c : cfgs) {
        const int rows = c[0], stages = c[1], mode = c[2], P = c[3];
        CUtensorMap tm;
        const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)m}; const cuuint64_t str[1] = {(cuuint64_t)ld * 4};
        const cuuint32_t box[2] = {128, (cuuint32_t)rows}; const cuuint32_t es[2] = {1, 1};
        if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, U, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 3; }
        const long long ntiles = ld / 128;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        tma_kernel<<<148, 128, P * stages * rows * 512 + 1024>>>(tm, U, ld, m, rows, stages, mode, ntiles, out, P);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        tma_kernel<<<148, 128, P * stages * rows * 512 + 1024>>>(tm, U, ld, m, rows, stages, mode, ntiles, out, P);
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%s rows=%2d stages/thread=%2d producers=%d (%3d KB in flight/SM): %.3f ms  %.0f GB/s\n", mode ? "bulk-1D" : "tensor2D", rows, stages, P, P * stages * rows / 2, ms,
               (double)ld * m * 4 / ms / 1e6);
    }
    return 0;
}
