// Evaluation-only kernels: materialised reconstruction (forward()'s first output, CYL:572-576) and the squared column
// norms of the spatial library G used by the post-hoc term norms (poly_norm / nonlinear_norm, CYL:624-692).
#include "common.cuh"

namespace desmo {

__device__ __forceinline__ void library_row_to_smem(const EvalArgs& a, long long x, int tid, float* Phi_s, float* G_s) {
    for (int i = 0; i < a.r; ++i)  // P == nullptr: the library of the raw phi_list, as the reference's post-hoc norms evaluate it (CYL:1192-1194)
        Phi_s[i * 256 + tid] = (x < a.ld) ? (a.P ? a.phi[(long long)i * a.ld + x] * a.P[(long long)i * a.ld + x] : a.phi[(long long)i * a.ld + x]) : 0.0f;
    for (int j = 0; j < a.K; ++j) {
        float v;
        if (j < a.T) {
            v = monomial(a.mt, j, Phi_s + tid, 256);
        } else {
            const int b = (j - a.T) / a.r, i = (j - a.T) - b * a.r;
            const float arg = a.omega[3 * i + b] * Phi_s[i * 256 + tid];
            v = (b == 0) ? sinf(arg) : (b == 1) ? cosf(arg) : tanhf(arg);
        }
        G_s[j * 256 + tid] = v;
    }
}

__global__ void __launch_bounds__(256) reconstruct_kernel(const EvalArgs a) {
    extern __shared__ float sm[];
    float* Phi_s = sm;
    float* G_s = sm + kMaxR * 256;
    const int tid = threadIdx.x;
    const long long x = (long long)blockIdx.x * 256 + tid;
    library_row_to_smem(a, x, tid, Phi_s, G_s);
    for (int t = blockIdx.y; t < a.m; t += gridDim.y) {
        float rec = 0.0f;
        for (int k = 0; k < a.K; ++k) rec = fmaf(G_s[k * 256 + tid], __ldg(a.W + (long long)k * a.mld + t), rec);
        if (x < a.ld) a.out[(long long)t * a.ld + x] = (x < a.n) ? rec : 0.0f;
    }
}

__global__ void __launch_bounds__(256) colnorm2_kernel(const EvalArgs a) {
    extern __shared__ float sm[];
    float* Phi_s = sm;
    float* G_s = sm + kMaxR * 256;
    const int tid = threadIdx.x;
    const long long ntiles = (a.ld + 255) / 256;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long x = tile * 256 + tid;
        library_row_to_smem(a, x, tid, Phi_s, G_s);
        for (int k = 0; k < a.K; ++k) {
            const float g = (x < a.n) ? G_s[k * 256 + tid] : 0.0f;
            const float s = warp_sum(g * g);
            if ((tid & 31) == 0) atomicAdd(a.out + k, s);
        }
    }
}

// Post-hoc term norms (poly_norm / nonlinear_norm, CYL:624-692; FCYL:644-720) in closed form:
//   norm_j = |gate_j| * ||G_j||_2 * ||z_j||_2       (= torch.norm(gate * (G_j z_j^T)), never materialising the n x m term)
// fourier_quirk: the Fourier scripts stack the polynomial series as (T, m) but slice `zs[:, i:i+1]` (FCYL:652,659), so polynomial
// term i < T is weighted by sqrt(sum_{j<T} z_j(t_i)^2) -- all T series at time index i -- instead of its own series' norm.
__global__ void __launch_bounds__(256) term_norms_kernel(const float* __restrict__ g2, const float* __restrict__ gates,
                                                         const float* __restrict__ rows, int T, int K, int m, int mld, int fourier_quirk,
                                                         double* __restrict__ out) {
    __shared__ double sh[8];
    const int k = blockIdx.x, tid = threadIdx.x;
    double acc = 0.0;
    if (fourier_quirk && k < T) {
        for (int j = tid; j < T; j += 256) { const double v = rows[(long long)j * mld + k]; acc += v * v; }
    } else {
        for (int t = tid; t < m; t += 256) { const double v = rows[(long long)k * mld + t]; acc += v * v; }
    }
    acc = warp_sum(acc);
    if ((tid & 31) == 0) sh[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += sh[w];
        out[k] = fabs((double)gates[k]) * sqrt((double)g2[k]) * sqrt(s);
    }
}

int launch_term_norms(const float* g2, const float* gates, const float* rows, int T, int K, int m, int mld, int fourier_quirk, double* out,
                      cudaStream_t st) {
    term_norms_kernel<<<K, 256, 0, st>>>(g2, gates, rows, T, K, m, mld, fourier_quirk, out);
    DESMO_CUDA(cudaGetLastError());
    return DESMO_OK;
}

int launch_reconstruct(const EvalArgs& a, cudaStream_t st) {
    const size_t smem = (size_t)(kMaxR + a.K) * 256 * sizeof(float);
    DESMO_CUDA(cudaFuncSetAttribute(reconstruct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned gx = (unsigned)((a.ld + 255) / 256);
    unsigned gy = 1;
    while (gx * gy < 296 && gy < (unsigned)a.m) gy *= 2;
    reconstruct_kernel<<<dim3(gx, gy), 256, smem, st>>>(a);
    DESMO_CUDA(cudaGetLastError());
    return DESMO_OK;
}

int launch_colnorm2(const EvalArgs& a, cudaStream_t st) {
    const size_t smem = (size_t)(kMaxR + a.K) * 256 * sizeof(float);
    DESMO_CUDA(cudaFuncSetAttribute(colnorm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DESMO_CUDA(cudaMemsetAsync(a.out, 0, sizeof(float) * a.K, st));
    long long g = (a.ld + 255) / 256;
    if (g > 296) g = 296;
    colnorm2_kernel<<<(unsigned)g, 256, smem, st>>>(a);
    DESMO_CUDA(cudaGetLastError());
    return DESMO_OK;
}

}  // namespace desmo
