// Host-buffer entry points: a device-resident training session fed from HOST memory, mirroring what one process of the
// reference does per epoch (CYL:706-778): upload the (m, n) snapshot batch, one fused step, read the losses back.
#include <string.h>

#include <vector>

#include "common.cuh"

using namespace desmo;

struct desmo_session {
    desmo_shape s{};
    Dims d{};
    int width = 0;  // columns of rows_or_coefs on the host side
    cudaStream_t st = nullptr;
    float *U = nullptr, *P = nullptr, *phi = nullptr, *phi_m = nullptr, *phi_u = nullptr, *dphi = nullptr;
    float *gates = nullptr, *gates_m = nullptr, *gates_u = nullptr, *rows = nullptr, *rows_m = nullptr, *rows_u = nullptr;
    float *coefs = nullptr, *coefs_m = nullptr, *coefs_u = nullptr, *periods = nullptr, *periods_m = nullptr, *periods_u = nullptr;
    float *omega = nullptr, *omega_m = nullptr, *omega_u = nullptr, *W = nullptr, *red = nullptr, *hyper = nullptr, *losses = nullptr;
    int32_t* step = nullptr;
    void* ws = nullptr;
    float* pinned_losses = nullptr;
    desmo_allreduce_fn allreduce = nullptr;  // multi-rank sessions: sums `red` over the ranks on the session's stream
    void* allreduce_user = nullptr;
    std::vector<void*> allocs;
};

static int dalloc(desmo_session* ss, void** p, size_t bytes) {
    DESMO_CUDA(cudaMalloc(p, bytes ? bytes : 4));
    DESMO_CUDA(cudaMemsetAsync(*p, 0, bytes ? bytes : 4, ss->st));
    ss->allocs.push_back(*p);
    return DESMO_OK;
}
#define DALLOC(field, count) do { int _r = dalloc(ss, reinterpret_cast<void**>(&ss->field), sizeof(*ss->field) * (size_t)(count)); if (_r) return _r; } while (0)

extern "C" {

int desmo_session_destroy(desmo_session* ss) {
    if (!ss) return DESMO_OK;
    for (void* p : ss->allocs) cudaFree(p);
    if (ss->pinned_losses) cudaFreeHost(ss->pinned_losses);
    if (ss->st) cudaStreamDestroy(ss->st);
    delete ss;
    return DESMO_OK;
}

static int session_build(desmo_session* ss) {
    const desmo_shape& s = ss->s;
    const int K = ss->d.K, Kp = ss->d.Kp;
    DESMO_CUDA(cudaStreamCreateWithFlags(&ss->st, cudaStreamNonBlocking));
    DALLOC(U, (size_t)s.m * s.ld);
    DALLOC(P, (size_t)s.r * s.ld);
    DALLOC(phi, (size_t)s.r * s.ld); DALLOC(phi_m, (size_t)s.r * s.ld); DALLOC(phi_u, (size_t)s.r * s.ld); DALLOC(dphi, (size_t)s.r * s.ld);
    DALLOC(gates, K); DALLOC(gates_m, K); DALLOC(gates_u, K);
    DALLOC(rows, (size_t)K * s.mld); DALLOC(rows_m, (size_t)K * s.mld); DALLOC(rows_u, (size_t)K * s.mld);
    if (s.nF > 0) {
        const int w = 2 * s.nF + 1;
        DALLOC(coefs, (size_t)K * w); DALLOC(coefs_m, (size_t)K * w); DALLOC(coefs_u, (size_t)K * w);
        DALLOC(periods, K); DALLOC(periods_m, K); DALLOC(periods_u, K);
    }
    DALLOC(omega, 3 * s.r); DALLOC(omega_m, 3 * s.r); DALLOC(omega_u, 3 * s.r);
    DALLOC(W, (size_t)Kp * s.mld);
    DALLOC(red, (size_t)desmo_red_count(&s));
    DALLOC(hyper, DESMO_HYP_COUNT);
    DALLOC(losses, 4);
    DALLOC(step, 1);
    size_t wsb = 0;
    int rc = desmo_workspace_bytes(&s, &wsb);
    if (rc) return rc;
    rc = dalloc(ss, &ss->ws, wsb);
    if (rc) return rc;
    DESMO_CUDA(cudaMallocHost(reinterpret_cast<void**>(&ss->pinned_losses), 4 * sizeof(float)));
    return DESMO_OK;
}

int desmo_session_create(int64_t n, int32_t m, int32_t r, int32_t polyorder, int32_t nF, int32_t path, desmo_session** out) {
    return desmo_session_create_sharded(n, n, m, r, polyorder, nF, path, out);
}

int desmo_session_set_allreduce(desmo_session* ss, desmo_allreduce_fn fn, void* user) {
    if (!ss) { set_error("desmo_session_set_allreduce: null"); return DESMO_ERR_ARG; }
    ss->allreduce = fn;
    ss->allreduce_user = user;
    return DESMO_OK;
}

int desmo_session_create_sharded(int64_t n, int64_t n_global, int32_t m, int32_t r, int32_t polyorder, int32_t nF, int32_t path,
                                 desmo_session** out) {
    if (!out) { set_error("desmo_session_create: null out"); return DESMO_ERR_ARG; }
    *out = nullptr;
    desmo_session* ss = new desmo_session();
    ss->s.n = n; ss->s.n_global = n_global; ss->s.ld = (n + 255) / 256 * 256; ss->s.m = m; ss->s.mld = (m + 15) / 16 * 16;
    ss->s.r = r; ss->s.polyorder = polyorder; ss->s.nF = nF; ss->s.path = path;
    int rc = validate_shape(&ss->s, &ss->d);
    if (!rc) rc = device_ok();
    if (!rc) rc = session_build(ss);
    if (rc) { desmo_session_destroy(ss); return rc; }
    ss->width = nF > 0 ? 2 * nF + 1 : m;
    *out = ss;
    return DESMO_OK;
}

int desmo_session_set_pod_host(desmo_session* ss, const double* pod_host) {  // [n][r] fp64 as POD_analysis returns (CYL:204)
    if (!ss || !pod_host) { set_error("desmo_session_set_pod_host: null"); return DESMO_ERR_ARG; }
    std::vector<float> tmp((size_t)ss->s.r * ss->s.ld, 0.0f);
    for (int64_t x = 0; x < ss->s.n; ++x)
        for (int i = 0; i < ss->s.r; ++i) tmp[(size_t)i * ss->s.ld + x] = (float)pod_host[(size_t)x * ss->s.r + i];  // .type(FloatTensor), CYL:539
    DESMO_CUDA(cudaMemcpyAsync(ss->P, tmp.data(), tmp.size() * sizeof(float), cudaMemcpyHostToDevice, ss->st));
    DESMO_CUDA(cudaStreamSynchronize(ss->st));
    return DESMO_OK;
}

int desmo_session_set_params_host(desmo_session* ss, const float* phi, const float* gates, const float* rows_or_coefs,
                                  const float* periods, const float* omega) {
    if (!ss || !phi || !gates || !rows_or_coefs || !omega || (ss->s.nF > 0 && !periods)) { set_error("desmo_session_set_params_host: null"); return DESMO_ERR_ARG; }
    const desmo_shape& s = ss->s;
    const int K = ss->d.K;
    DESMO_CUDA(cudaMemsetAsync(ss->phi, 0, sizeof(float) * (size_t)s.r * s.ld, ss->st));
    DESMO_CUDA(cudaMemcpy2DAsync(ss->phi, s.ld * sizeof(float), phi, s.n * sizeof(float), s.n * sizeof(float), s.r, cudaMemcpyHostToDevice, ss->st));
    DESMO_CUDA(cudaMemcpyAsync(ss->gates, gates, K * sizeof(float), cudaMemcpyHostToDevice, ss->st));
    DESMO_CUDA(cudaMemcpyAsync(ss->omega, omega, 3 * s.r * sizeof(float), cudaMemcpyHostToDevice, ss->st));
    if (s.nF > 0) {
        DESMO_CUDA(cudaMemcpyAsync(ss->coefs, rows_or_coefs, sizeof(float) * (size_t)K * ss->width, cudaMemcpyHostToDevice, ss->st));
        DESMO_CUDA(cudaMemcpyAsync(ss->periods, periods, K * sizeof(float), cudaMemcpyHostToDevice, ss->st));
    } else {
        DESMO_CUDA(cudaMemcpy2DAsync(ss->rows, s.mld * sizeof(float), rows_or_coefs, s.m * sizeof(float), s.m * sizeof(float), K, cudaMemcpyHostToDevice, ss->st));
    }
    // fresh optimizer (Adamax state starts at zero, step 0)
    float* zero[] = {ss->phi_m, ss->phi_u, ss->gates_m, ss->gates_u, ss->rows_m, ss->rows_u, ss->omega_m, ss->omega_u};
    size_t zb[] = {(size_t)s.r * s.ld, (size_t)s.r * s.ld, (size_t)K, (size_t)K, (size_t)K * s.mld, (size_t)K * s.mld, (size_t)3 * s.r, (size_t)3 * s.r};
    for (int i = 0; i < 8; ++i) DESMO_CUDA(cudaMemsetAsync(zero[i], 0, zb[i] * sizeof(float), ss->st));
    if (s.nF > 0) {
        DESMO_CUDA(cudaMemsetAsync(ss->coefs_m, 0, sizeof(float) * (size_t)K * ss->width, ss->st));
        DESMO_CUDA(cudaMemsetAsync(ss->coefs_u, 0, sizeof(float) * (size_t)K * ss->width, ss->st));
        DESMO_CUDA(cudaMemsetAsync(ss->periods_m, 0, sizeof(float) * K, ss->st));
        DESMO_CUDA(cudaMemsetAsync(ss->periods_u, 0, sizeof(float) * K, ss->st));
    }
    DESMO_CUDA(cudaMemsetAsync(ss->step, 0, sizeof(int32_t), ss->st));
    DESMO_CUDA(cudaStreamSynchronize(ss->st));
    return DESMO_OK;
}

int desmo_session_set_hyper(desmo_session* ss, const float* lrs /*[5]*/, float beta, float l1_lambda) {
    if (!ss || !lrs) { set_error("desmo_session_set_hyper: null"); return DESMO_ERR_ARG; }
    float h[DESMO_HYP_COUNT];
    for (int i = 0; i < 5; ++i) h[i] = lrs[i];
    h[DESMO_HYP_BETA] = beta;
    h[DESMO_HYP_L1_LAMBDA] = l1_lambda;
    DESMO_CUDA(cudaMemcpyAsync(ss->hyper, h, sizeof(h), cudaMemcpyHostToDevice, ss->st));
    DESMO_CUDA(cudaStreamSynchronize(ss->st));
    return DESMO_OK;
}

// Upload the reference's (m, n) fp32 batch (CYL:708 `.type(FloatTensor).to(device)`) into the padded device layout.
int desmo_session_upload_snapshot_host(desmo_session* ss, const float* snapshot_host) {
    if (!ss || !snapshot_host) { set_error("desmo_session_upload_snapshot_host: null"); return DESMO_ERR_ARG; }
    const desmo_shape& s = ss->s;
    if (s.n == s.ld) {  // no padding between snapshots: one contiguous transfer
        DESMO_CUDA(cudaMemcpyAsync(ss->U, snapshot_host, sizeof(float) * (size_t)s.n * s.m, cudaMemcpyHostToDevice, ss->st));
    } else {
        DESMO_CUDA(cudaMemcpy2DAsync(ss->U, s.ld * sizeof(float), snapshot_host, s.n * sizeof(float), s.n * sizeof(float), s.m,
                                     cudaMemcpyHostToDevice, ss->st));
    }
    return DESMO_OK;
}

static int session_step_device(desmo_session* ss) {
    const desmo_shape* s = &ss->s;
    int rc = desmo_build_w(s, ss->gates, ss->rows, ss->coefs, ss->periods, ss->W, ss->step, ss->ws, ss->st);
    if (rc) return rc;
    rc = desmo_fused_residual_grad(s, ss->U, ss->P, ss->phi, ss->omega, ss->W, ss->dphi, ss->red, ss->ws, ss->st);
    if (rc) return rc;
    if (s->n_global != s->n) {  // this rank owns a slab of the points: the one exchange of the step
        if (!ss->allreduce) { set_error("sharded session (n_global != n) without an all-reduce hook: call desmo_session_set_allreduce"); return DESMO_ERR_ARG; }
        if (ss->allreduce(ss->red, desmo_red_count(s), ss->st, ss->allreduce_user)) { set_error("all-reduce hook failed"); return DESMO_ERR_CUDA; }
    }
    return desmo_adamax_update(s, ss->red, ss->dphi, ss->P, ss->phi, ss->phi_m, ss->phi_u, ss->gates, ss->gates_m, ss->gates_u, ss->rows,
                               ss->rows_m, ss->rows_u, ss->coefs, ss->coefs_m, ss->coefs_u, ss->periods, ss->periods_m, ss->periods_u,
                               ss->omega, ss->omega_m, ss->omega_u, ss->hyper, ss->step, ss->losses, ss->ws, ss->st);
}

// One epoch of the reference loop through host memory: (optional) upload of the batch, fused step, losses back on the host.
int desmo_session_step_host(desmo_session* ss, const float* snapshot_host_or_null, float* losses_host /*[4]*/) {
    if (!ss) { set_error("desmo_session_step_host: null"); return DESMO_ERR_ARG; }
    int rc = DESMO_OK;
    if (snapshot_host_or_null && (rc = desmo_session_upload_snapshot_host(ss, snapshot_host_or_null))) return rc;
    if ((rc = session_step_device(ss))) return rc;
    DESMO_CUDA(cudaMemcpyAsync(ss->pinned_losses, ss->losses, 4 * sizeof(float), cudaMemcpyDeviceToHost, ss->st));
    DESMO_CUDA(cudaStreamSynchronize(ss->st));
    if (losses_host) memcpy(losses_host, ss->pinned_losses, 4 * sizeof(float));
    return DESMO_OK;
}

int desmo_session_get_params_host(desmo_session* ss, float* phi, float* gates, float* rows_or_coefs, float* periods, float* omega) {
    if (!ss) { set_error("desmo_session_get_params_host: null"); return DESMO_ERR_ARG; }
    const desmo_shape& s = ss->s;
    const int K = ss->d.K;
    if (phi) DESMO_CUDA(cudaMemcpy2DAsync(phi, s.n * sizeof(float), ss->phi, s.ld * sizeof(float), s.n * sizeof(float), s.r, cudaMemcpyDeviceToHost, ss->st));
    if (gates) DESMO_CUDA(cudaMemcpyAsync(gates, ss->gates, K * sizeof(float), cudaMemcpyDeviceToHost, ss->st));
    if (omega) DESMO_CUDA(cudaMemcpyAsync(omega, ss->omega, 3 * s.r * sizeof(float), cudaMemcpyDeviceToHost, ss->st));
    if (s.nF > 0) {
        if (rows_or_coefs) DESMO_CUDA(cudaMemcpyAsync(rows_or_coefs, ss->coefs, sizeof(float) * (size_t)K * ss->width, cudaMemcpyDeviceToHost, ss->st));
        if (periods) DESMO_CUDA(cudaMemcpyAsync(periods, ss->periods, K * sizeof(float), cudaMemcpyDeviceToHost, ss->st));
    } else if (rows_or_coefs) {
        DESMO_CUDA(cudaMemcpy2DAsync(rows_or_coefs, s.m * sizeof(float), ss->rows, s.mld * sizeof(float), s.m * sizeof(float), K, cudaMemcpyDeviceToHost, ss->st));
    }
    DESMO_CUDA(cudaStreamSynchronize(ss->st));
    return DESMO_OK;
}

int desmo_train_host(int64_t n, int32_t m, int32_t r, int32_t polyorder, int32_t nF, const float* snapshot_host, const double* pod_host,
                     float* phi_host, float* gates_host, float* rows_or_coefs_host, float* periods_host, float* omega_host,
                     const float* lrs, float beta, float l1_lambda, int32_t steps, float* losses_host, int32_t path) {
    desmo_session* ss = nullptr;
    int rc = desmo_session_create(n, m, r, polyorder, nF, path, &ss);
    if (rc) return rc;
    if (!(rc = desmo_session_set_pod_host(ss, pod_host)) &&
        !(rc = desmo_session_set_params_host(ss, phi_host, gates_host, rows_or_coefs_host, periods_host, omega_host)) &&
        !(rc = desmo_session_set_hyper(ss, lrs, beta, l1_lambda)) && !(rc = desmo_session_upload_snapshot_host(ss, snapshot_host))) {
        for (int i = 0; i < steps && !rc; ++i) rc = desmo_session_step_host(ss, nullptr, losses_host ? losses_host + 4 * i : nullptr);
        if (!rc) rc = desmo_session_get_params_host(ss, phi_host, gates_host, rows_or_coefs_host, periods_host, omega_host);
    }
    desmo_session_destroy(ss);
    return rc;
}

}  // extern "C"
