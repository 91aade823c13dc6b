// Measures the cost of back-to-back tcgen05.mma kind::f16 (bf16) instructions for the operand layouts of the fused kernel.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t idesc_bf16(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\t@px mov.s32 %0, 1;\n\t}\n" : "+r"(pred) : "r"(0xffffffffu));
    return pred != 0;
}
// mode: a_mn, b_mn, N; nmma MMAs each K=16 cycling over 8 k-steps of a 128-deep operand; A tile [128 x 128] bf16 = 32 KB, B [N x 128]
__global__ void __launch_bounds__(128) rate_kernel(int a_mn, int b_mn, int N, int nmma, int same_addr, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_s;
    uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < (65536 + 32768) / 4; i += 128) ((uint32_t*)base)[i] = 0x3c003c00u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar)), "r"(1u) : "memory"); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_s;
    if (warp == 1 && elect_one()) {
        const uint32_t a0 = smem_u32(base), b0 = smem_u32(base + 65536);
        const uint32_t idesc = idesc_bf16(128, N, a_mn, b_mn);
        uint64_t ad[8], bd[8];
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            ad[ks] = a_mn ? make_desc(a0 + ks * 2048, 16384, 1024) : make_desc(a0 + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024);
            bd[ks] = b_mn ? make_desc(b0 + ks * 2048, 16384, 1024) : make_desc(b0 + (ks >> 2) * (N * 128) + (ks & 3) * 32, 16, 1024);
        }
        long long t0 = clock64();
        for (int i = 0; i < nmma; i += 8) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                uint32_t acc = (i + ks) > 0;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                             :: "r"(tmem + (uint32_t)(same_addr ? 0 : ((i >> 3) & 1) * 128)), "l"(ad[ks]), "l"(bd[ks]), "r"(idesc), "r"(acc) : "memory");
            }
        }
        long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
        uint32_t done = 0;
        while (!done) asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512u) : "memory");
}
int main() {
    long long* out; CK(cudaMallocManaged(&out, 64));
    CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110000));
    const int cfgs[][3] = {{0, 0, 32}, {0, 0, 64}, {0, 0, 96}, {0, 0, 128}, {1, 0, 32}, {1, 0, 64}, {1, 0, 128}, {1, 1, 128}, {1, 1, 32}, {0, 1, 128}};
    for (auto& c : cfgs)
        for (int grid : {1, 148})
            for (int same : {0, 1}) {
                const int nmma = 2048;
                rate_kernel<<<grid, 128, 110000>>>(c[0], c[1], c[2], nmma, same, out);
                CK(cudaDeviceSynchronize());
                printf("A %s B %s N=%3d grid=%3d sameacc=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (math floor %d)\n", c[0] ? "MN" : "K ", c[1] ? "MN" : "K ", c[2], grid, same,
                       (double)out[0] / nmma, (double)out[1] / nmma, c[2] / 2);
            }
    return 0;
}
