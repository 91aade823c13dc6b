// Fused residual + gradient pass on the 5th-gen tensor cores (DESMO_PATH_TC): tcgen05.mma kind::f16 with every fp32 operand
// split into three bf16 terms (x = x1 + x2 + x3 carries all 24 significand bits; the six products with i + j <= 4 are kept, so
// each GEMM is accurate to ~2^-24 relative, i.e. fp32-class, at the cost of 6 bf16 MMAs = 3 TF32-equivalents).
//
// One persistent CTA per SM; a CTA owns tiles of 128 mesh points and walks all snapshots in slabs of 128:
//   G1  Rec[p x t]    = G[p x lib] W[lib x t]           -> TMEM cols [0,128)                      (CYL:548,565-572)
//   epi r = Rec - U   (U read once from HBM, coalesced, straight into registers; R never leaves the SM) (CYL:722)
//       r -> three bf16 planes in smem  R_s[p rows][t contiguous]  (128B-swizzled, K-major for G3 AND MN-major for G4)
//   G3  D[p x lib]   += R[p x t] W^T[t x lib]           -> TMEM cols [128,160), accumulated over the slabs of a tile
//   G4  E^T[t x lib] += R^T[t x p] G[p x lib]           -> TMEM cols [256+32*slab, +32), accumulated over ALL tiles of the CTA
// Warp roles: warp 0 = TMA producer (W slab planes), warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4..11 = epilogue
// (thread <-> mesh point == TMEM lane; warps 4-7 take snapshots 0-63 of the slab, warps 8-11 snapshots 64-127).
// The MMA issuer runs G1 of slab s+1 ahead of G3/G4 of slab s, so the tensor pipe works while the epilogue forms R.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace desmo {

namespace tc {
constexpr int KP = 32;          // padded library size handled by this kernel
constexpr int BP = 128;         // points per tile  (MMA M of G1/G3, K of G4)
constexpr int BT = 128;         // snapshots per slab (MMA N of G1, K of G3, M of G4)
constexpr int MAXSLAB = 8;      // TMEM: 8 x 32 columns of E accumulators
constexpr int THREADS = 384;
constexpr uint32_t R_PLANE = 2 * BP * 128;            // one bf16 plane of R: 2 boxes [128 p rows x 128 B (64 t)]
constexpr uint32_t W_PLANE = 2 * KP * 128;            // one bf16 plane of a W slab: 2 boxes [32 lib rows x 128 B (64 t)]
constexpr uint32_t G_PLANE = 2 * KP * 128;            // one bf16 plane of G: 2 boxes [32 lib rows x 128 B (64 p)]
constexpr uint32_t R_OFF = 0;
constexpr uint32_t W_OFF = R_OFF + 3 * R_PLANE;        // 98304
constexpr uint32_t G_OFF = W_OFF + 2 * 3 * W_PLANE;    // 147456
constexpr uint32_t RED_OFF = G_OFF + 3 * G_PLANE;      // 172032  (4 warps x kScal doubles)
constexpr uint32_t SMEM_BYTES = RED_OFF + 4 * kScal * 8 + 1024;  // + alignment slack
constexpr uint32_t TMEM_REC = 0, TMEM_D = 128, TMEM_E = 256;
}  // namespace tc

struct TcArgs {
    const float* U;
    const float* P;
    const float* phi;
    const float* omega;
    float* dphi;
    float* Epart;
    double* Spart;
    long long n, ld;
    int m, mld, r, T, K, nslab, kp_out;
    float scale;
    MonoTable mt;
};

// ------------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t sw128(uint32_t row, uint32_t byte_in_row) {
    return row * 128u + ((((byte_in_row >> 4) ^ (row & 7u)) << 4) | (byte_in_row & 15u));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);  // version 1, SWIZZLE_128B
}
__device__ __forceinline__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// x = b1 + b2 + b3 (bf16 each, round-to-nearest): packs two consecutive elements (lo = first) per 32-bit word
__device__ __forceinline__ void split3_pair(float x0, float x1, uint32_t& w1, uint32_t& w2, uint32_t& w3) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w1) : "f"(x1), "f"(x0));
    const float e0 = x0 - __uint_as_float(w1 << 16), e1 = x1 - __uint_as_float(w1 & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w2) : "f"(e1), "f"(e0));
    const float f0 = e0 - __uint_as_float(w2 << 16), f1 = e1 - __uint_as_float(w2 & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w3) : "f"(f1), "f"(f0));
}

// six (a-plane, b-plane) products kept by the 3-way split, smallest contributions first
__device__ __constant__ int kPairA[6] = {2, 0, 1, 1, 0, 0};
__device__ __constant__ int kPairB[6] = {0, 2, 1, 0, 1, 0};

__global__ void __launch_bounds__(tc::THREADS, 1) fused_tc_kernel(const TcArgs a, const __grid_constant__ CUtensorMap tmW) {
    using namespace tc;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[16];
    __shared__ uint32_t tmem_base_s;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    double* red_s = reinterpret_cast<double*>(smem + RED_OFF);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    enum { W_FULL0 = 0, W_FULL1, W_EMPTY0, W_EMPTY1, REC_FULL, REC_EMPTY, R_FULL, R_EMPTY, G_FULL, G_EMPTY, D_FULL, D_EMPTY };
    auto bar = [&](int i) { return smem_u32(&bars[i]); };

    const int nslab = a.nslab;
    const long long ntiles = a.ld / BP;
    const int my_tiles = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const int total = my_tiles * nslab;

    if (tid == 32) {
        mbar_init(bar(W_FULL0), 1); mbar_init(bar(W_FULL1), 1); mbar_init(bar(W_EMPTY0), 1); mbar_init(bar(W_EMPTY1), 1);
        mbar_init(bar(REC_FULL), 1); mbar_init(bar(REC_EMPTY), 256); mbar_init(bar(R_FULL), 256); mbar_init(bar(R_EMPTY), 1);
        mbar_init(bar(G_FULL), 128); mbar_init(bar(G_EMPTY), 1); mbar_init(bar(D_FULL), 1); mbar_init(bar(D_EMPTY), 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < 4 * kScal; i += THREADS) red_s[i] = 0.0;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    if (warp == 0) {
        // ================================================ TMA producer: W slab planes ================================================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
            for (int it = 0; it < total; ++it) {
                const int buf = it & 1, slab = it % nslab;
                if (it >= 2) mbar_wait(bar(W_EMPTY0 + buf), ((it >> 1) & 1) ^ 1);
                mbar_expect_tx(bar(W_FULL0 + buf), 3 * W_PLANE);
                for (int s = 0; s < 3; ++s)
                    for (int h = 0; h < 2; ++h)
                        tma_load_2d(sbase + W_OFF + buf * 3 * W_PLANE + s * W_PLANE + h * (KP * 128), &tmW, slab * BT + h * 64, s * KP,
                                    bar(W_FULL0 + buf));
            }
        }
    } else if (warp == 1) {
        // ================================================ MMA issuer ================================================
        if (lane == 0) {
            constexpr uint32_t idesc_g1 = make_idesc_bf16(BP, BT, 1, 1);
            constexpr uint32_t idesc_g3 = make_idesc_bf16(BP, KP, 0, 0);
            constexpr uint32_t idesc_g4 = make_idesc_bf16(BT, KP, 1, 0);
            auto issue_g1 = [&](int it) {
                const int buf = it & 1, slab = it % nslab, tl = it / nslab;
                mbar_wait(bar(W_FULL0 + buf), (it >> 1) & 1);
                if (it > 0) mbar_wait(bar(REC_EMPTY), (it - 1) & 1);
                if (slab == 0) mbar_wait(bar(G_FULL), tl & 1);
                tc_fence_after();
                const uint32_t gb = sbase + G_OFF, wb = sbase + W_OFF + buf * 3 * W_PLANE;
                uint32_t acc = 0;
#pragma unroll 1
                for (int pr = 0; pr < 6; ++pr) {
                    const uint32_t ga = gb + kPairA[pr] * G_PLANE, wa = wb + kPairB[pr] * W_PLANE;
#pragma unroll
                    for (int ks = 0; ks < KP / 16; ++ks) {
                        mma_bf16(tmem + TMEM_REC, make_desc(ga + ks * 2048, KP * 128, 1024), make_desc(wa + ks * 2048, KP * 128, 1024),
                                 idesc_g1, acc);
                        acc = 1;
                    }
                }
                umma_commit(bar(REC_FULL));
            };
            auto issue_g34 = [&](int it) {
                const int buf = it & 1, slab = it % nslab, tl = it / nslab;
                mbar_wait(bar(R_FULL), it & 1);
                if (slab == 0 && tl > 0) mbar_wait(bar(D_EMPTY), (tl - 1) & 1);
                tc_fence_after();
                const uint32_t rb = sbase + R_OFF, wb = sbase + W_OFF + buf * 3 * W_PLANE, gb = sbase + G_OFF;
                uint32_t acc = slab > 0 ? 1u : 0u;
#pragma unroll 1
                for (int pr = 0; pr < 6; ++pr) {  // G3: D += R W^T   (A = R K-major, B = W K-major, K = snapshots)
                    const uint32_t ra = rb + kPairA[pr] * R_PLANE, wa = wb + kPairB[pr] * W_PLANE;
#pragma unroll
                    for (int ks = 0; ks < BT / 16; ++ks) {
                        mma_bf16(tmem + TMEM_D, make_desc(ra + (ks >> 2) * (BP * 128) + (ks & 3) * 32, 16, 1024),
                                 make_desc(wa + (ks >> 2) * (KP * 128) + (ks & 3) * 32, 16, 1024), idesc_g3, acc);
                        acc = 1;
                    }
                }
                acc = tl > 0 ? 1u : 0u;
#pragma unroll 1
                for (int pr = 0; pr < 6; ++pr) {  // G4: E^T += R^T G  (A = R MN-major, B = G K-major, K = points)
                    const uint32_t ra = rb + kPairA[pr] * R_PLANE, ga = gb + kPairB[pr] * G_PLANE;
#pragma unroll
                    for (int ks = 0; ks < BP / 16; ++ks) {
                        mma_bf16(tmem + TMEM_E + slab * KP, make_desc(ra + ks * 2048, BP * 128, 1024),
                                 make_desc(ga + (ks >> 2) * (KP * 128) + (ks & 3) * 32, 16, 1024), idesc_g4, acc);
                        acc = 1;
                    }
                }
                umma_commit(bar(R_EMPTY));
                umma_commit(bar(W_EMPTY0 + buf));
                if (slab == nslab - 1) {
                    umma_commit(bar(D_FULL));
                    umma_commit(bar(G_EMPTY));
                }
            };
            if (total > 0) issue_g1(0);
            for (int it = 0; it < total; ++it) {
                const bool next_same_tile = (it + 1 < total) && ((it + 1) % nslab != 0);
                if (next_same_tile) issue_g1(it + 1);  // run ahead: the tensor pipe works on G1(s+1) while the epilogue forms R(s)
                issue_g34(it);
                if (!next_same_tile && it + 1 < total) issue_g1(it + 1);  // new tile: its G needs G4 of the old tile finished
            }
        }
    } else if (warp >= 4) {
        // ================================================ epilogue warps ================================================
        const int q = warp & 3, h = (warp - 4) >> 2;
        const int p = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        double loss_acc = 0.0;
        float lat[kMaxR], dl[tc::KP], dph[kMaxR], dom[3 * kMaxR];

        auto chain_and_store = [&](int tile_local, long long tile) {
            // D of the finished tile -> d mse/d phi, d omega, Phi^T Phi   (warps 8..11)
            const long long x = tile * BP + p;
            mbar_wait(bar(D_FULL), tile_local & 1);
            tc_fence_after();
            uint32_t v[16];
            tmem_ld16(tmem + lane_addr + TMEM_D, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) dl[j] = __uint_as_float(v[j]) * a.scale;
            tmem_ld16(tmem + lane_addr + TMEM_D + 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) dl[16 + j] = __uint_as_float(v[j]) * a.scale;
            tc_fence_before();
            mbar_arrive(bar(D_EMPTY));
            for (int i = 0; i < a.r; ++i) lat[i] = a.phi[(long long)i * a.ld + x] * a.P[(long long)i * a.ld + x];
            chain_rule_point(a.mt, a.r, a.T, a.omega, lat, dl, dph, dom, 1);
            for (int i = 0; i < a.r; ++i) a.dphi[(long long)i * a.ld + x] = dph[i] * a.P[(long long)i * a.ld + x];
            const bool xin = x < a.n;
            for (int i = 0; i < 3 * a.r; ++i) {
                const float s = warp_sum(xin ? dom[i] : 0.0f);
                if (lane == 0) red_s[q * kScal + 1 + kMaxR * kMaxR + i] += (double)s;
            }
            for (int i = 0; i < a.r; ++i)
                for (int j = i; j < a.r; ++j) {
                    const float s = warp_sum(lat[i] * lat[j]);
                    if (lane == 0) red_s[q * kScal + 1 + i * kMaxR + j] += (double)s;
                }
        };

        int it = 0;
        for (int tl = 0; tl < my_tiles; ++tl) {
            const long long tile = blockIdx.x + (long long)tl * gridDim.x;
            const long long x = tile * BP + p;
            const bool xin = x < a.n;
            if (h == 0) {
                // ---- library row of this point (CYL:538-548,565-567), split into bf16 planes, G_s[lib rows][p contiguous] ----
                for (int i = 0; i < a.r; ++i) lat[i] = a.phi[(long long)i * a.ld + x] * a.P[(long long)i * a.ld + x];
                float g[KP];
#pragma unroll
                for (int j = 0; j < KP; ++j) {
                    float v = 0.0f;
                    if (j < a.T) {
                        v = monomial(a.mt, j, lat, 1);
                    } else if (j < a.K) {
                        const int b = (j - a.T) / a.r, i = (j - a.T) - b * a.r;
                        const float arg = a.omega[3 * i + b] * lat[i];
                        v = (b == 0) ? sinf(arg) : (b == 1) ? cosf(arg) : tanhf(arg);
                    }
                    g[j] = v;
                }
                if (tl > 0) mbar_wait(bar(G_EMPTY), (tl - 1) & 1);
                uint8_t* gs = smem + G_OFF + (p >> 6) * (KP * 128);
                const uint32_t pb = (p & 63) * 2;
#pragma unroll
                for (int j = 0; j < KP; ++j) {
                    const __nv_bfloat16 b1 = __float2bfloat16_rn(g[j]);
                    const float e1 = g[j] - __bfloat162float(b1);
                    const __nv_bfloat16 b2 = __float2bfloat16_rn(e1);
                    const __nv_bfloat16 b3 = __float2bfloat16_rn(e1 - __bfloat162float(b2));
                    const uint32_t off = sw128(j, pb);
                    *reinterpret_cast<__nv_bfloat16*>(gs + off) = b1;
                    *reinterpret_cast<__nv_bfloat16*>(gs + G_PLANE + off) = b2;
                    *reinterpret_cast<__nv_bfloat16*>(gs + 2 * G_PLANE + off) = b3;
                }
                fence_async_smem();
                mbar_arrive(bar(G_FULL));
            } else if (tl > 0) {
                chain_and_store(tl - 1, tile - gridDim.x);
            }

            for (int slab = 0; slab < nslab; ++slab, ++it) {
                const int t0 = slab * BT + h * 64;
                float u[64];
#pragma unroll
                for (int j = 0; j < 64; ++j) {
                    const int t = t0 + j;
                    u[j] = (xin && t < a.m) ? __ldg(a.U + (long long)t * a.ld + x) : 0.0f;
                }
                mbar_wait(bar(REC_FULL), it & 1);
                tc_fence_after();
                float lsum = 0.0f;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t v[16];
                    tmem_ld16(tmem + lane_addr + TMEM_REC + h * 64 + c * 16, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int t = t0 + c * 16 + j;
                        const float rr = (xin && t < a.m) ? __uint_as_float(v[j]) - u[c * 16 + j] : 0.0f;
                        u[c * 16 + j] = rr;
                        lsum = fmaf(rr, rr, lsum);
                    }
                }
                loss_acc += (double)lsum;
                tc_fence_before();
                mbar_arrive(bar(REC_EMPTY));
                if (it > 0) mbar_wait(bar(R_EMPTY), (it - 1) & 1);
                // ---- r -> three bf16 planes, own row p of box h (64 snapshots = 128 B = 8 chunks of 16 B) ----
                uint8_t* rs = smem + R_OFF + h * (BP * 128) + p * 128;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    uint32_t w1[4], w2[4], w3[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) split3_pair(u[c * 8 + 2 * e], u[c * 8 + 2 * e + 1], w1[e], w2[e], w3[e]);
                    const uint32_t off = ((uint32_t)(c ^ (p & 7))) << 4;
                    *reinterpret_cast<uint4*>(rs + off) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
                    *reinterpret_cast<uint4*>(rs + R_PLANE + off) = make_uint4(w2[0], w2[1], w2[2], w2[3]);
                    *reinterpret_cast<uint4*>(rs + 2 * R_PLANE + off) = make_uint4(w3[0], w3[1], w3[2], w3[3]);
                }
                fence_async_smem();
                mbar_arrive(bar(R_FULL));
            }
        }
        // last tile's chain rule, then the E accumulators of this CTA
        if (h == 1 && my_tiles > 0) chain_and_store(my_tiles - 1, blockIdx.x + (long long)(my_tiles - 1) * gridDim.x);
        if (total > 0) mbar_wait(bar(R_EMPTY), (total - 1) & 1);
        tc_fence_after();
        float* Eo = a.Epart + (long long)blockIdx.x * a.kp_out * a.mld;
        for (int slab = h; slab < nslab; slab += 2) {
            const int t = slab * BT + p;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t v[16];
                tmem_ld16(tmem + lane_addr + TMEM_E + slab * KP + c * 16, v);
                tmem_ld_wait();
                if (t < a.mld) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c * 16 + j < a.kp_out) Eo[(long long)(c * 16 + j) * a.mld + t] = __uint_as_float(v[j]);
                }
            }
        }
        loss_acc = warp_sum(loss_acc);
        if (lane == 0) atomicAdd(&red_s[q * kScal + 0], loss_acc);
        tc_fence_before();
    }
    __syncthreads();
    for (int i = tid; i < kScal; i += THREADS) {
        double s = 0.0;
        for (int w = 0; w < 4; ++w) s += red_s[w * kScal + i];
        a.Spart[(long long)blockIdx.x * kScal + i] = s;
    }
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int fused_tc_supported(const desmo_shape* s, int Kp) {
    return (Kp <= tc::KP && s->mld <= tc::MAXSLAB * tc::BT && s->ld % 128 == 0 && s->mld % 8 == 0) ? 1 : 0;
}

void reduce_partials_launch(const float* Epart, int nx, long long ecount, const double* Spart, int nslots, int r, float* red, cudaStream_t st);

int fused_tc(const desmo_shape* s, const MonoTable& mt, int T, int Kp, const float* U, const float* P, const float* phi,
             const float* omega, const float* W, float* dphi, float* red, const Workspace& ws, cudaStream_t st) {
    (void)W;
    if (!fused_tc_supported(s, Kp)) { set_error("tcgen05 path: unsupported shape"); return DESMO_ERR_UNSUPPORTED; }
    EncodeTiledFn enc = encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled not available"); return DESMO_ERR_CUDA; }
    int dev = 0, sms = 0;
    DESMO_CUDA(cudaGetDevice(&dev));
    DESMO_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // bf16 planes of W written by build_w: Wb[3][Kp_tc][mld]
    CUtensorMap tm;
    const cuuint64_t dims[2] = {(cuuint64_t)s->mld, (cuuint64_t)(3 * tc::KP)};
    const cuuint64_t strides[1] = {(cuuint64_t)s->mld * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)tc::KP};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)ws.tc, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)cr); return DESMO_ERR_CUDA; }
    TcArgs a{};
    a.U = U; a.P = P; a.phi = phi; a.omega = omega; a.dphi = dphi; a.Epart = ws.Epart; a.Spart = ws.Spart;
    a.n = s->n; a.ld = s->ld; a.m = s->m; a.mld = s->mld; a.r = s->r; a.T = T; a.K = T + 3 * s->r;
    a.nslab = (s->m + tc::BT - 1) / tc::BT;
    a.kp_out = Kp;
    a.scale = (float)(2.0 / ((double)s->n_global * (double)s->m));
    a.mt = mt;
    const long long ntiles = s->ld / tc::BP;
    const int grid = (int)(ntiles < sms ? ntiles : sms);
    DESMO_CUDA(cudaFuncSetAttribute(fused_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES));
    fused_tc_kernel<<<grid, tc::THREADS, tc::SMEM_BYTES, st>>>(a, tm);
    DESMO_CUDA(cudaGetLastError());
    reduce_partials_launch(ws.Epart, grid, (long long)Kp * s->mld, ws.Spart, grid, s->r, red, st);
    DESMO_CUDA(cudaGetLastError());
    return DESMO_OK;
}

int pod_gram_tc(const desmo_shape* s, const float* U, float* C, void* workspace, cudaStream_t st) {
    (void)s; (void)U; (void)C; (void)workspace; (void)st;
    return DESMO_ERR_UNSUPPORTED;
}

}  // namespace desmo
