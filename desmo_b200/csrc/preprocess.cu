// Snapshot pre-processing on the device: what the reference does in numpy between the VTK reader and
// the POD / training loop --
//   convert3Dto2D_data (CYL:88-106)  : drop the w component of a 2-D flow          -> d_use = 2 of d_in = 3
//   convertToMagnitude (CYL:109-133) : |u| per point and snapshot, in float64       -> DESMO_PRE_MAGNITUDE
//   subtract_mean      (CYL:136-149) : remove the temporal mean of every point      -> DESMO_PRE_SUBTRACT_MEAN
//   ... * 1/sqrt(m)    (ANEU:143)    : the aneurysm scripts' extra scaling          -> DESMO_PRE_SCALE_SQRT_M
//   X[:, 0::2]         (TURB:189)    : keep every t_stride-th snapshot AFTER the mean was taken over all of them
//   X.T -> FloatTensor (CYL:356,708) : time-major fp32 snapshot, here straight into the padded U[m][ld] layout
// The raw input is X.T as the reader produces it: V[m_in][n * d_in], the components of a point adjacent.  One thread owns one
// point; pass 1 accumulates the fp64 temporal sum, pass 2 recomputes the magnitude and rounds (|u| - mean) * scale to fp32 ONCE,
// exactly where the reference's float64 -> float32 cast sits.  HBM-bound: (2 * d_in * m_in * sizeof(in) + 4 * m) bytes / point.
#include "common.cuh"

namespace desmo {

template <typename TIn>
__device__ __forceinline__ double point_value(const TIn* __restrict__ p, int d_use, bool magnitude) {
    if (!magnitude) return (double)p[0];
    // np.sum(np.square(Ui), 1) then np.sqrt, float64, no contraction into FMAs
    double s = __dmul_rn((double)p[0], (double)p[0]);
    for (int c = 1; c < d_use; ++c) s = __dadd_rn(s, __dmul_rn((double)p[c], (double)p[c]));
    return sqrt(s);
}

template <typename TIn>
__global__ void __launch_bounds__(256) preprocess_kernel(const TIn* __restrict__ V, long long v_ld, long long n, long long ld, int m_in,
                                                         int m_out, int t_stride, int d_in, int d_use, int flags,
                                                         float* __restrict__ U, double* __restrict__ mean_out) {
    const long long x = (long long)blockIdx.x * 256 + threadIdx.x;
    if (x >= ld) return;
    if (x >= n) {
        for (int t = 0; t < m_out; ++t) U[(long long)t * ld + x] = 0.0f;
        return;
    }
    const bool magnitude = flags & DESMO_PRE_MAGNITUDE;
    const TIn* col = V + x * d_in;
    double mean = 0.0;
    if (flags & DESMO_PRE_SUBTRACT_MEAN) {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        int t = 0;
        for (; t + 4 <= m_in; t += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u] += point_value(col + (long long)(t + u) * v_ld, d_use, magnitude);
        }
        for (; t < m_in; ++t) acc[0] += point_value(col + (long long)t * v_ld, d_use, magnitude);
        mean = ((acc[0] + acc[1]) + (acc[2] + acc[3])) / (double)m_in;
    }
    if (mean_out) mean_out[x] = mean;
    const double scale = (flags & DESMO_PRE_SCALE_SQRT_M) ? 1.0 / sqrt((double)m_in) : 1.0;
#pragma unroll 4
    for (int t = 0; t < m_out; ++t) {
        const double v = point_value(col + (long long)t * t_stride * v_ld, d_use, magnitude);
        U[(long long)t * ld + x] = (float)(scale * (v - mean));  // (1/np.sqrt(m)) * (X - mean), one rounding to fp32
    }
}

int preprocess(const desmo_shape* s, const void* V, int v_dtype, long long v_ld, int m_in, int t_stride, int d_in, int d_use, int flags,
               float* U, double* mean, cudaStream_t st) {
    const unsigned grid = (unsigned)((s->ld + 255) / 256);
    if (v_dtype == DESMO_DTYPE_F64)
        preprocess_kernel<double><<<grid, 256, 0, st>>>((const double*)V, v_ld, s->n, s->ld, m_in, s->m, t_stride, d_in, d_use, flags, U, mean);
    else
        preprocess_kernel<float><<<grid, 256, 0, st>>>((const float*)V, v_ld, s->n, s->ld, m_in, s->m, t_stride, d_in, d_use, flags, U, mean);
    DESMO_CUDA(cudaGetLastError());
    return DESMO_OK;
}

}  // namespace desmo
