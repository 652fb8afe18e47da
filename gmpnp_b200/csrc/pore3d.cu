// 3D pore path -- placeholder entry points until the kernels land (return GMPNP_ERR_STATE).
#include "common.cuh"
extern "C" {
int gmpnp_create_3d(gmpnp_handle**, int, const double*, int, const int*, int, const int*, int, int, int) { return GMPNP_ERR_STATE; }
int gmpnp_set_dirichlet_3d(gmpnp_handle*, const double*, int) { return GMPNP_ERR_STATE; }
int gmpnp_pattern_3d(const gmpnp_handle*, int*, int*, int*) { return GMPNP_ERR_STATE; }
int gmpnp_assemble_3d(gmpnp_handle*, const double*, const double*, double*, double*, void*) { return GMPNP_ERR_STATE; }
int gmpnp_spmv_3d(gmpnp_handle*, const double*, const double*, double*, void*) { return GMPNP_ERR_STATE; }
int gmpnp_newton_3d(gmpnp_handle*, double*, const double*, const gmpnp_newton_opts*, int*, double*, double*, int*, int*, void*) { return GMPNP_ERR_STATE; }
int gmpnp_median_3d(gmpnp_handle*, const double*, int, double*, void*) { return GMPNP_ERR_STATE; }
}
