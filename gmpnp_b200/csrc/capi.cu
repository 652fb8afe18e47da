// C-ABI entry points of libgmpnp.so (see include/gmpnp.h for the contract and the
// reference call sites each function replaces).
#include "common.cuh"

int edl1d_launch_newton(gmpnp_handle*, int, double*, double*, const double*, const gmpnp_newton_opts*, int,
                        const double*, double*, int*, double*, double*, double*, int*, int*, cudaStream_t);
int edl1d_launch_assemble(gmpnp_handle*, const double*, const double*, double*, double*, cudaStream_t);
int edl1d_launch_field(gmpnp_handle*, const double*, double*, cudaStream_t);

extern "C" {

const char* gmpnp_strerror(int code) {
    switch (code) {
        case GMPNP_OK: return "ok";
        case GMPNP_ERR_ARG: return "invalid argument";
        case GMPNP_ERR_CUDA: return "CUDA error (see gmpnp_last_cuda_error)";
        case GMPNP_ERR_ALLOC: return "allocation failed";
        case GMPNP_ERR_STATE: return "handle not ready (parameters / Dirichlet values not set, or wrong dimension)";
        default: return "unknown error";
    }
}

const char* gmpnp_last_cuda_error(const gmpnp_handle* h) { return h ? h->last_cuda_error.c_str() : ""; }

int gmpnp_version(void) { return 100; }

long long gmpnp_launch_count(const gmpnp_handle* h) { return h ? h->launches : 0; }

int gmpnp_create_1d(gmpnp_handle** out, int device, const double* h_x, int n_nodes, int n_species, int batch) {
    if (!out || !h_x || n_nodes < 2 || n_species != 6 || batch < 1) return GMPNP_ERR_ARG;
    for (int k = 1; k < n_nodes; ++k)
        if (!(h_x[k] > h_x[k - 1])) return GMPNP_ERR_ARG;       // cell k = nodes (k, k+1), sorted
    gmpnp_handle* h = new gmpnp_handle();
    h->dim = 1; h->device = device; h->batch = batch; h->ns = n_species; h->nc = n_species + 1;
    h->n_nodes = n_nodes;
    *out = h;
    GMPNP_CUDA_TRY(h, cudaSetDevice(device));
    GMPNP_CUDA_TRY(h, cudaMalloc(&h->d_x, sizeof(double) * n_nodes));
    GMPNP_CUDA_TRY(h, cudaMemcpy(h->d_x, h_x, sizeof(double) * n_nodes, cudaMemcpyHostToDevice));
    GMPNP_CUDA_TRY(h, cudaMalloc(&h->d_params, sizeof(double) * GMPNP_NPAR * batch));
    GMPNP_CUDA_TRY(h, cudaMalloc(&h->d_ws, sizeof(double) * 56 * (size_t)n_nodes * batch));
    return GMPNP_OK;
}

int gmpnp_set_params(gmpnp_handle* h, const double* h_params, int batch) {
    if (!h || !h_params || batch != h->batch) return GMPNP_ERR_ARG;
    GMPNP_CUDA_TRY(h, cudaSetDevice(h->device));
    GMPNP_CUDA_TRY(h, cudaMemcpy(h->d_params, h_params, sizeof(double) * GMPNP_NPAR * batch, cudaMemcpyHostToDevice));
    h->params_set = true;
    return GMPNP_OK;
}

static int check_1d(gmpnp_handle* h) {
    if (!h) return GMPNP_ERR_ARG;
    if (h->dim != 1 || !h->params_set) return GMPNP_ERR_STATE;
    return GMPNP_OK;
}

int gmpnp_assemble_1d(gmpnp_handle* h, const double* d_u, const double* d_un, double* d_F, double* d_J, void* stream) {
    int rc = check_1d(h); if (rc) return rc;
    if (!d_u || !d_un) return GMPNP_ERR_ARG;
    GMPNP_CUDA_TRY(h, cudaSetDevice(h->device));
    return edl1d_launch_assemble(h, d_u, d_un, d_F, d_J, (cudaStream_t)stream);
}

int gmpnp_newton_1d(gmpnp_handle* h, double* d_u, const double* d_un, const gmpnp_newton_opts* opts, int* d_iters,
                    double* d_r0, double* d_r, int* d_status, void* stream) {
    int rc = check_1d(h); if (rc) return rc;
    if (!d_u || !d_un || !opts) return GMPNP_ERR_ARG;
    GMPNP_CUDA_TRY(h, cudaSetDevice(h->device));
    return edl1d_launch_newton(h, 0, d_u, nullptr, d_un, opts, 1, nullptr, nullptr, d_iters, d_r0, d_r, nullptr,
                               nullptr, d_status, (cudaStream_t)stream);
}

int gmpnp_march_1d(gmpnp_handle* h, double* d_u, double* d_un, int n_steps, const gmpnp_newton_opts* opts,
                   double* d_hist, int* d_iters, double* d_hfrac, int* d_status, void* stream) {
    int rc = check_1d(h); if (rc) return rc;
    if (!d_u || !d_un || !opts || n_steps < 1) return GMPNP_ERR_ARG;
    GMPNP_CUDA_TRY(h, cudaSetDevice(h->device));
    return edl1d_launch_newton(h, 1, d_u, d_un, nullptr, opts, n_steps, nullptr, d_hist, d_iters, nullptr, nullptr,
                               d_hfrac, nullptr, d_status, (cudaStream_t)stream);
}

int gmpnp_steady_continuation_1d(gmpnp_handle* h, double* d_u, const double* d_Vpath, int n_V,
                                 const gmpnp_newton_opts* opts, int* d_iters, int* d_stage, int* d_status,
                                 void* stream) {
    int rc = check_1d(h); if (rc) return rc;
    if (!d_u || !d_Vpath || !opts || n_V < 1) return GMPNP_ERR_ARG;
    GMPNP_CUDA_TRY(h, cudaSetDevice(h->device));
    return edl1d_launch_newton(h, 2, d_u, nullptr, nullptr, opts, n_V, d_Vpath, nullptr, d_iters, nullptr, nullptr,
                               nullptr, d_stage, d_status, (cudaStream_t)stream);
}

int gmpnp_field_1d(gmpnp_handle* h, const double* d_u, double* d_field, void* stream) {
    int rc = check_1d(h); if (rc) return rc;
    if (!d_u || !d_field) return GMPNP_ERR_ARG;
    GMPNP_CUDA_TRY(h, cudaSetDevice(h->device));
    return edl1d_launch_field(h, d_u, d_field, (cudaStream_t)stream);
}

}  // extern "C"

void pore3d_free_ext(gmpnp_handle* h);

extern "C" void gmpnp_destroy(gmpnp_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->dim == 3) pore3d_free_ext(h);
    void* bufs[] = {h->d_x, h->d_params, h->d_ws, h->d_tets, h->d_geom, h->d_row_ptr, h->d_col_idx, h->d_diag_idx,
                    h->d_blk_ptr, h->d_blk_src, h->d_node_ptr, h->d_node_src, h->d_dir_dof, h->d_dir_flag,
                    h->d_dir_val, h->d_mom, h->d_Fe, h->d_J, h->d_Dinv, h->d_F, h->d_krylov, h->d_small,
                    h->d_ismall, h->d_sort};
    for (void* b : bufs)
        if (b) cudaFree(b);
    if (h->h_pinned) cudaFreeHost(h->h_pinned);
    delete h;
}
