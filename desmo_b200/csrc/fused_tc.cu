// tcgen05 (3xTF32) fused residual + gradient pass -- see DESIGN.md "tensor-core path".
#include "common.cuh"

namespace desmo {

int fused_tc_supported(const desmo_shape* s, int Kp) {
    (void)s; (void)Kp;
    return 0;
}

int fused_tc(const desmo_shape* s, const MonoTable& mt, int T, int Kp, const float* U, const float* P, const float* phi,
             const float* omega, const float* W, float* dphi, float* red, const Workspace& ws, cudaStream_t st) {
    (void)s; (void)mt; (void)T; (void)Kp; (void)U; (void)P; (void)phi; (void)omega; (void)W; (void)dphi; (void)red; (void)ws; (void)st;
    set_error("tcgen05 path not available for this shape");
    return DESMO_ERR_UNSUPPORTED;
}

int pod_gram_tc(const desmo_shape* s, const float* U, float* C, void* workspace, cudaStream_t st) {
    (void)s; (void)U; (void)C; (void)workspace; (void)st;
    return DESMO_ERR_UNSUPPORTED;
}

}  // namespace desmo
