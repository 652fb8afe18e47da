// Shared declarations of the GMPNP CUDA library (sm_100a, fp64 throughout).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <nvtx3/nvToolsExt.h>
#include "../../include/gmpnp.h"

// NVTX range of one library phase (assemble / preconditioner / linear solve / Newton / march step): shows up on an
// Nsight Systems timeline, costs a few nanoseconds when no tool is attached (header-only NVTX3, no extra library).
struct GmpnpRange {
    explicit GmpnpRange(const char* name) { nvtxRangePushA(name); }
    ~GmpnpRange() { nvtxRangePop(); }
    GmpnpRange(const GmpnpRange&) = delete;
    GmpnpRange& operator=(const GmpnpRange&) = delete;
};

#define GMPNP_CUDA_TRY(h, expr)                                              \
    do {                                                                     \
        cudaError_t _e = (expr);                                             \
        if (_e != cudaSuccess) {                                             \
            (h)->last_cuda_error = std::string(#expr) + ": " + cudaGetErrorString(_e); \
            return GMPNP_ERR_CUDA;                                           \
        }                                                                    \
    } while (0)

struct gmpnp_handle {
    int dim = 0;            // 1 or 3
    int device = 0;
    int batch = 0;
    int ns = 0;             // species
    int nc = 0;             // components = ns + 1
    int n_nodes = 0;        // vertices
    long long launches = 0;
    std::string last_cuda_error;
    // device buffers owned by the handle
    double* d_x = nullptr;        // 1D: [n]  3D: [n][3]
    double* d_params = nullptr;   // [batch][GMPNP_NPAR]
    bool params_set = false;
    // ---- 1D ----
    double* d_ws = nullptr;       // elimination workspace [batch][n][56]
    double* d_ws2 = nullptr;      // partitioned elimination: spike workspace [batch][n][56] (allocated on first use)
    // ---- 3D ----
    int n_tet = 0, n_dir = 0, n_blocks = 0;
    int* d_tets = nullptr;        // [T][4]
    double* d_geom = nullptr;     // [T][13]: grad lambda (4x3) + volume
    int* d_row_ptr = nullptr;     // [V+1]
    int* d_col_idx = nullptr;     // [nb]
    int* d_diag_idx = nullptr;    // [V] index of the diagonal block of each row
    int* d_blk_ptr = nullptr;     // [nb+1] -> contributions of (tet, a, b) to each block
    int* d_blk_src = nullptr;     // [16 T] packed tet*16 + a*4 + b
    int* d_node_ptr = nullptr;    // [V+1] -> incident (tet, a) of each vertex
    int* d_node_src = nullptr;    // [4 T] packed tet*4 + a
    int* d_dir_dof = nullptr;     // [n_dir]
    int* d_dir_flag = nullptr;    // [V*9] index into dir list or -1
    double* d_dir_val = nullptr;  // [batch][n_dir]
    bool dir_set = false;
    std::vector<int> h_row_ptr, h_col_idx;
    // 3D work buffers (allocated lazily)
    double* d_mom = nullptr;      // per-tet Jacobian moments [batch][T][NMOM]
    double* d_Fe = nullptr;       // per-tet residual vectors [batch][T][36]
    double* d_J = nullptr;        // BSR values [batch][nb][81]
    double* d_Dinv = nullptr;     // inverted diagonal blocks [batch][V][81]
    double* d_F = nullptr;        // residual [batch][V*9]
    double* d_krylov = nullptr;   // GMRES basis etc.
    size_t krylov_doubles = 0;
    double* d_small = nullptr;    // small per-problem scalars
    int* d_ismall = nullptr;
    void* h_pinned = nullptr;     // pinned host scratch for control read-backs
    double* d_sort = nullptr;     // median scratch
    void* ext3d = nullptr;        // 3D-only state (pore3d.cu: Host3D), owned by the handle
};
