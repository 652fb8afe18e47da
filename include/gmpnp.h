/*
 * gmpnp.h -- C-ABI of the B200-native GMPNP hot path (libgmpnp.so).
 *
 * The reference (divyabohra/GMPNP) has no FFI: its hot path is the Python call
 *     solve(F == 0, u, bcs, solver_parameters=...)
 * at 1D/MPNP_CO2ER_EDL.py:737-742 and 3D/MPNP_CO2ER_pore.py:789-799, executed inside the
 * pseudo-time loops 1D:633-796 / 3D:782-858.  This library is what a maintainer would
 * bind (ctypes) at exactly those call sites, with the loop hoisted
 * inside.  Everything is fp64.  Conventions:
 *   - plain C, no C++/torch types; `int` status returns (0 = ok, <0 = API/CUDA error,
 *     see gmpnp_strerror); numerical outcomes are reported per problem in `status[]`;
 *   - all `d_*` pointers are caller-owned DEVICE buffers (e.g. torch tensors' data_ptr());
 *     `h_*` pointers are host buffers; `stream` is a cudaStream_t passed as void*;
 *   - a handle owns its mesh copy, parameter records and work buffers; handles are independent
 *     (thread-compatible per handle); the library has no process-wide state.
 *
 * Unknown layout: node-major interleaved, u[problem][node][comp], comp = species in the
 * reference's order (1D:117 / 3D:138) with the potential last (1D:303, 3D:407).
 */
#ifndef GMPNP_H
#define GMPNP_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gmpnp_handle gmpnp_handle;

/* ---- packed per-problem parameter record: double[GMPNP_NPAR] ------------------------
 * Host code (gmpnp_b200/params.py, mirroring 1D:81-213,368-375 and 3D:115-324) fills it. */
#define GMPNP_NPAR 64
#define GMPNP_P_NS      0   /* number of species (6 in 1D, 8 in 3D)                          */
#define GMPNP_P_Z       1   /* [8] charges z_i                       (1D:158, 3D:233)         */
#define GMPNP_P_NU      9   /* [8] scale_vol_i = a_i^3 c0_i N_A      (1D:200, 3D:287); 0=PNP  */
#define GMPNP_P_ZC0    17   /* [8] z_i * c0_i  (Poisson source, 1D:422-427)                   */
#define GMPNP_P_S      25   /* [5] scale_R of H, OH, HCO3, CO32, CO2 (1D:190, 3D:277)         */
#define GMPNP_P_KW     30   /* kw2*c0_H*c0_OH      (1D:384)                                   */
#define GMPNP_P_KA     31   /* ka1*c0_OH*c0_HCO3   (1D:389)                                   */
#define GMPNP_P_KB     32   /* kb1*c0_CO2*c0_OH    (1D:390)                                   */
#define GMPNP_P_KA2    33   /* ka2*c0_CO32         (1D:391)                                   */
#define GMPNP_P_KB2    34   /* kb2*c0_HCO3         (1D:391-392)                               */
#define GMPNP_P_KW1    35   /* kw1                                                            */
#define GMPNP_P_EPSW   36   /* eps_rel of water    (1D:413)                                   */
#define GMPNP_P_EPSH   37   /* n_water_H  *c0_H  *1e-3 (1D:414-415)                           */
#define GMPNP_P_EPSC   38   /* n_water_cat*c0_cat*1e-3                                        */
#define GMPNP_P_KAPPA  39   /* coefficient of (u-u_n) v dx: 1/(dt*L_D) 1D:458, 1/dt 3D:534; 0=steady */
#define GMPNP_P_V      40   /* scaled potential at the OHP (1D:354) / pore wall (3D:462)      */
#define GMPNP_P_JFLUX  41   /* [8] 1D point fluxes J_i of `J_i v_i ds` (1D:553, 738)          */
#define GMPNP_P_ICAT   49   /* index of the cation species (= ns-1)                           */
#define GMPNP_P_Q      50   /* q = F^2 L^2/(eps0 R T)   (1D:193, 3D:280)                      */
#define GMPNP_P_JOHPRE 51   /* J_OH_prefactor*current_OHP_ss  (H_OHP controller, 1D:789-791)  */
#define GMPNP_P_JHPRE  52   /* J_H_prefactor *current_OHP_ss  (1D:793)                        */
#define GMPNP_P_HOHP   53   /* H_OHP target; < 0 disables the controller (1D:770)             */
#define GMPNP_P_HFRAC  54   /* current_H_frac initial value (1D:167-170)                      */

/* ---- Newton options (dolfin NewtonSolver semantics) ------------------ */
typedef struct gmpnp_newton_opts {
    double rtol;        /* relative_tolerance  (1e-4 at 1D:361, 3D:794)                    */
    double atol;        /* absolute_tolerance  (1e-4 at 1D:362, 3D:795)                    */
    double relax;       /* relaxation_parameter (1.0 in 1D, 0.9 at 3D:796)                 */
    double xtol;        /* increment criterion: ||dx||_inf <= xtol*max(1,||x||_inf); nothing else
                           counts as GMPNP_CONVERGED under criterion 1                        */
    int    maxit;       /* maximum_iterations  (50)                                        */
    int    criterion;   /* 0 = residual (reference), 1 = increment (steady mode)           */
    int    pivot;       /* 1 = partial pivoting inside the 7x7 blocks (default), 0 = none  */
    int    lin_maxit;   /* 3D: max GMRES iterations per Newton step                        */
    int    lin_restart; /* 3D: GMRES restart length                                        */
    double lin_rtol;    /* 3D: GMRES relative residual tolerance                           */
    double xtol_path;   /* continuation: increment tolerance of the intermediate stages
                           (<= 0: use xtol everywhere); the final stage always uses xtol    */
    int    jac_rule;    /* quadrature of the Jacobian: 0 = FFC's choice (degree 4: 3-point Gauss in 1D,
                           14-point Keast in 3D -- the reference's iteration path), 1 = the residual's
                           rule (exact derivative of the discrete F: quadratic convergence; same
                           converged solution, which depends on F's rule only)              */
    int    partitions;  /* 1D: sweeps per problem.  2 = two-sided elimination (top half downwards, bottom half upwards:
                           the throughput form); 4 / 8 = partitioned elimination -- the chain is cut at 2 / 4 separator
                           nodes, interior sub-domains are swept from both ends carrying a 7-column spike, the
                           separators form a small reduced system: ~2.5x shorter critical path at 8 for ~2x the work,
                           for batches that leave SMs idle (strong scaling, single-problem latency); 0 = automatic by
                           batch size.  Same results to round-off.                                                  */
    double xtol_floor;  /* increment criterion only, 0 = off (the default: strict contract above).  > 0: an
                           increment that has stopped contracting (||dx|| >= 0.25 ||dx_prev||, after >= 3
                           iterations) with xtol*s < ||dx||_inf <= xtol_floor*s, s = max(1,||x||_inf), ends the
                           iteration with status GMPNP_STAGNATED (round-off floor of the linear solve, e.g. of
                           the pivot-free elimination); the iterate is kept and the final relative increment is
                           returned (d_dx), so the caller decides: accept, or re-run with pivoting.           */
} gmpnp_newton_opts;

/* per-problem status codes written to status[] */
#define GMPNP_CONVERGED        0
#define GMPNP_MAXIT            1   /* dolfin would raise RuntimeError here                   */
#define GMPNP_NOT_FINITE       2   /* NaN/Inf in residual or singular pivot                  */
#define GMPNP_LINEAR_FAILED    3   /* 3D: GMRES did not reach lin_rtol                       */
#define GMPNP_STAGNATED        4   /* increment stalled between xtol and xtol_floor (see opts) */

/* API error codes */
#define GMPNP_OK               0
#define GMPNP_ERR_ARG         -1
#define GMPNP_ERR_CUDA        -2
#define GMPNP_ERR_ALLOC       -3
#define GMPNP_ERR_STATE       -4

/* fp64 FMA micro-benchmark of the device (TFLOP/s, DFMA = 2 flops): the measured denominator for fp64 fractions. */
int gmpnp_fp64_peak(int device, double* tflops);

const char* gmpnp_strerror(int code);
const char* gmpnp_last_cuda_error(const gmpnp_handle* h);
int  gmpnp_version(void);
void gmpnp_destroy(gmpnp_handle* h);

/* ------------------------------------------------------------------------------------
 * 1D planar EDL (replaces the FEniCS work behind 1D/MPNP_CO2ER_EDL.py:737-742)
 * ---------------------------------------------------------------------------------- */

/* `h_x`: the n_nodes sorted vertex coordinates of the interval mesh (Mesh(), 1D:231-234;
 * cell k = nodes (k,k+1)).  All `batch` problems of a handle share the mesh.            */
int gmpnp_create_1d(gmpnp_handle** out, int device, const double* h_x, int n_nodes,
                    int n_species, int batch);

/* Packed parameter records, h_params[batch][GMPNP_NPAR] (host pointer; copied).           */
int gmpnp_set_params(gmpnp_handle* h, const double* h_params, int batch);

/* Residual and block-tridiagonal Jacobian at (u, u_n), Dirichlet rows applied exactly as
 * dolfin does (identity row, residual x-g; 1D:350-355).  Testable alone.
 *   d_F [batch][n][7], d_J [batch][n][3][7][7] (sub-, main-, super-diagonal block of each
 *   node row, row-major inside a block).  Either output may be NULL.                      */
int gmpnp_assemble_1d(gmpnp_handle* h, const double* d_u, const double* d_un,
                      double* d_F, double* d_J, void* stream);

/* One reference `solve(F == 0, u, bcs)` per problem: Newton from the incoming d_u.
 * Outputs (device, may be NULL): iters[batch], r0[batch], r[batch], status[batch].        */
int gmpnp_newton_1d(gmpnp_handle* h, double* d_u, const double* d_un,
                    const gmpnp_newton_opts* opts, int* d_iters, double* d_r0, double* d_r,
                    int* d_status, void* stream);

/* The reference's pseudo-time loop (1D:633-796): n_steps times { solve; H_OHP controller
 * (1D:766-793); u_n <- u }.  d_un is updated in place.  d_hist (optional)
 * [batch][n_steps][n][7] receives u after every step; d_iters (optional) [batch][n_steps];
 * d_hfrac (optional) [batch] final current_H_frac.  A failed step stops that problem only. */
int gmpnp_march_1d(gmpnp_handle* h, double* d_u, double* d_un, int n_steps,
                   const gmpnp_newton_opts* opts, double* d_hist, int* d_iters,
                   double* d_hfrac, int* d_status, void* stream);

/* Steady equations (kappa forced to 0) with voltage continuation: for s < n_V the OHP
 * potential is d_Vpath[problem][s] and Newton restarts from the previous stage's solution.
 * A NaN entry ends that problem's path early (ragged paths in one batch).
 * d_iters (optional) [batch][n_V]; d_stage (optional) [batch] = number of stages completed;
 * d_dx (optional) [batch] = ||dx||_inf / max(1,||u||_inf) of the last Newton update of the last stage
 * (what the increment criterion saw).  GMPNP_STAGNATED at an intermediate stage does not end a path. */
int gmpnp_steady_continuation_1d(gmpnp_handle* h, double* d_u, const double* d_Vpath, int n_V,
                                 const gmpnp_newton_opts* opts, int* d_iters, int* d_stage,
                                 int* d_status, double* d_dx, void* stream);

/* L2 projection of -d(phi)/dx onto P1 (dolfin project(-grad(u_p), W), 1D:802-803) for
 * every problem: d_field[batch][n].                                                       */
int gmpnp_field_1d(gmpnp_handle* h, const double* d_u, double* d_field, void* stream);

/* The same projection evaluated at the OHP only (node 0: the field_OHP of 1D:940-954 and of the result table
 * 1D/Stern_CO2ER.py:66-68), d_out[batch] -- for per-point sweep summaries: a one-sweep elimination over the first 256
 * nodes (the mass matrix is diagonally dominant: the truncation error is below 0.5^256).                              */
int gmpnp_field_ohp_1d(gmpnp_handle* h, const double* d_u, double* d_out, void* stream);

/* number of kernels this handle has launched so far (for bench.py's gpu_launches)          */
long long gmpnp_launch_count(const gmpnp_handle* h);

/* ------------------------------------------------------------------------------------
 * 3D pore (replaces the FEniCS work behind 3D/MPNP_CO2ER_pore.py:789-799)
 * ---------------------------------------------------------------------------------- */

/* Mesh (Mesh(), 3D:329-332) and Dirichlet DOF list (DirichletBC, 3D:460-467, already
 * de-duplicated so that the last BC wins).  Values are per problem (gmpnp_set_dirichlet). */
int gmpnp_create_3d(gmpnp_handle** out, int device, const double* h_xyz, int n_vert,
                    const int* h_tets, int n_tet, const int* h_dir_dof, int n_dir,
                    int n_species, int batch);

/* Dirichlet values, h_vals[batch][n_dir] (host pointer; copied).                          */
int gmpnp_set_dirichlet_3d(gmpnp_handle* h, const double* h_vals, int batch);

/* Sparsity of the BSR-9 Jacobian: number of block rows / blocks, and copies of row_ptr
 * (n_vert+1) / col_idx (n_blocks) for the caller (host pointers, may be NULL).            */
int gmpnp_pattern_3d(const gmpnp_handle* h, int* n_blocks, int* h_row_ptr, int* h_col_idx);

/* Residual d_F[batch][n_vert][9] and BSR values d_J[batch][n_blocks][9][9] at (u, u_n),
 * Dirichlet rows applied.  Either output may be NULL.  Handles created with batch >= 24 use
 * the problem-per-lane kernels (same results to round-off, same buffers); the
 * environment variable GMPNP_ASM_LANES=0|1, read by gmpnp_create_3d, forces either layout. */
int gmpnp_assemble_3d(gmpnp_handle* h, const double* d_u, const double* d_un,
                      double* d_F, double* d_J, void* stream);

/* y = J x with the BSR values of the last assembly kept in the handle, or with d_J if
 * given: d_x, d_y [batch][n_vert][9].                                                     */
int gmpnp_spmv_3d(gmpnp_handle* h, const double* d_J, const double* d_x, double* d_y,
                  void* stream);

/* One reference `solve(F == 0, u, bcs)` per problem: damped Newton, each step solved by
 * restarted GMRES with per-node 9x9 block-Jacobi.  d_lin_iters (optional) [batch] = total
 * GMRES iterations.                                                                       */
int gmpnp_newton_3d(gmpnp_handle* h, double* d_u, const double* d_un,
                    const gmpnp_newton_opts* opts, int* d_iters, double* d_r0, double* d_r,
                    int* d_lin_iters, int* d_status, void* stream);

/* ---- the reference's loop inside the library (3D/MPNP_CO2ER_pore.py:782-858) -------------------------------------
 * gmpnp_set_march_data_3d: what the loop needs besides the parameter records:
 *   h_kind[n_dir]   kind of every Dirichlet DOF of gmpnp_create_3d's list: 0 -> 0 (potential at pore entry / exit,
 *                   3D:460-461), 1 -> wall potential (3D:462), 2/3/4 -> CO2 / CO / H2 entry value (3D:463-465);
 *   h_tab[batch][4] (wall potential, initial CO2 entry value, CO entry value, H2 entry value), scaled;
 *   h_sech[batch][8] Sechenov record of the per-step update of the CO2 entry value from the nodal MEDIANS (3D:817-838,
 *                   CO2_conc 3D:70-93): co2 = s[0] * 10^-(s[1] m_OH + s[2] m_HCO3 + s[3] m_CO32 + s[4] m_cat) (s[5] = 0;
 *                   s[5] = 1: the reaction-diffusion script's electroneutral cation, 3D/rxn_diff_CO2ER_pore.py:575-601:
 *                   medians of (H, OH, HCO3, CO32), exponent s[6] m_H + s[1] m_OH + s[2] m_HCO3 + s[3] m_CO32).
 * gmpnp_march_3d: n_steps times { Dirichlet values with the current CO2 entry value (3D:835-838); solve (3D:789-799);
 *   medians + Sechenov update (3D:817-834); u_n <- u (3D:856) }.  A problem whose Newton solve fails stops (its status is
 *   the Newton status; dolfin would raise), the others continue.  Optional outputs (device): d_hist
 *   [batch][n_steps][n_vert][9], d_iters / d_lin_iters [batch][n_steps], d_co2 [batch][n_steps] (entry value USED in the
 *   step), d_steps [batch] (steps completed), d_status [batch].  Control flow is device-resident: the host reads one
 *   counter per Newton iteration and one per step; the linear solves run in one launch each.
 * gmpnp_steady_3d: the same loop run to the steady state, with the wall potential ramped linearly over the first
 *   n_ramp steps (voltage continuation).  From step n_ramp on, a problem whose last increment satisfies max|u - u_n|
 *   <= tol * max(1, max|u|) stops marching (d_converged[p] = 1, d_steps[p] = its number of steps); the loop ends when
 *   no problem is marching any more or after max_steps.  d_iters [batch][max_steps], d_inc_hist [max_steps][batch]
 *   (relative increment per step, 0 for stopped problems), d_co2 [batch] (final entry value), *h_steps_run = steps run.  */
int gmpnp_set_march_data_3d(gmpnp_handle* h, const signed char* h_kind, const double* h_tab, const double* h_sech,
                            int batch);
int gmpnp_march_3d(gmpnp_handle* h, double* d_u, double* d_un, int n_steps, const gmpnp_newton_opts* opts,
                   double* d_hist, int* d_iters, int* d_lin_iters, double* d_co2, int* d_steps, int* d_status,
                   void* stream);
int gmpnp_steady_3d(gmpnp_handle* h, double* d_u, double* d_un, const gmpnp_newton_opts* opts, double tol, int max_steps,
                    int n_ramp, int* d_iters, double* d_inc_hist, double* d_co2, int* d_steps, int* d_status,
                    int* d_converged, int* h_steps_run, void* stream);

/* ---- mesh-partitioned mode (one problem spans the GPUs of a box; SURVEY 8e (2), BASELINE config 5) -----------
 * The reference has no distributed path; these are the per-rank building blocks of the
 * distributed GMRES that replaces the MUMPS solve of 3D/MPNP_CO2ER_pore.py:792 on a partitioned mesh.  A rank's
 * handle is created on its LOCAL mesh (owned vertices first, ghost vertices after); the halo exchange and the
 * all-reduces of the dot products are done by the host with torch.distributed between these calls.
 *   gmpnp_vec_multi_dot:  d_out[k] = sum_{i<n} V_k[i] w[i], V_k = d_V + k*vstride  (two-stage, deterministic)
 *   gmpnp_vec_lincomb:    d_out[i] = beta*d_y[i] + sum_k d_coef[k] V_k[i]  (d_y may be NULL, d_out may alias d_y;
 *                         d_coef is a device array)
 *   gmpnp_bjacobi_setup_3d: invert the diagonal 9x9 blocks of d_J (NULL: the handle's last Jacobian)
 *   gmpnp_bjacobi_apply_3d: d_z = D^-1 d_r on the first n_rows block rows (batch 1)                         */
/* y = J x restricted to the block rows [row0, row1): the partitioned mode multiplies the interior rows (no ghost
 * column) while the halo exchange is in flight and the boundary rows after it.                                */
int gmpnp_spmv_rows_3d(gmpnp_handle* h, const double* d_J, const double* d_x, double* d_y, int row0, int row1,
                       void* stream);
int gmpnp_vec_multi_dot(gmpnp_handle* h, const double* d_V, long long vstride, int nvec, const double* d_w,
                        long long n, double* d_out, void* stream);
int gmpnp_vec_lincomb(gmpnp_handle* h, const double* d_V, long long vstride, int nvec, const double* d_coef,
                      double beta, const double* d_y, double* d_out, long long n, void* stream);
int gmpnp_bjacobi_setup_3d(gmpnp_handle* h, const double* d_J, void* stream);
int gmpnp_bjacobi_apply_3d(gmpnp_handle* h, const double* d_r, double* d_z, int n_rows, void* stream);

/* Distributed z-slab coarse correction (the Galerkin space of the single-mesh preconditioner, 16 slabs x 9
 * components = 144 unknowns, split where the ranks exchange): set the GLOBAL slab id of every local vertex;
 * accumulate this rank's P^T J P over its owned rows into d_Ac[144*144] (host all-reduces it), invert the reduced
 * matrix into the handle; per application restrict d_rc[144] += P^T r over the owned rows (host all-reduces it)
 * and prolong d_z += P (A_c^-1 d_rc).  Batch 1.                                                                */
int gmpnp_set_aggregates_3d(gmpnp_handle* h, const int* h_agg);
int gmpnp_coarse_accumulate_3d(gmpnp_handle* h, const double* d_J, int n_rows, double* d_Ac, void* stream);
int gmpnp_coarse_invert_3d(gmpnp_handle* h, const double* d_Ac, void* stream);
int gmpnp_coarse_restrict_3d(gmpnp_handle* h, const double* d_r, int n_rows, double* d_rc, void* stream);
int gmpnp_coarse_prolong_3d(gmpnp_handle* h, const double* d_rc, double* d_z, int n_rows, void* stream);

/* The INTENDED boundary physics of the 3D script (3D/MPNP_CO2ER_pore.py:474-499 and the `+ J_... * ds(k)` lines
 * 560-750, which Python discards as executed -- SURVEY finding 3; the same terms are live in
 * 3D/rxn_diff_CO2ER_pore.py:480-511): constant wall fluxes `J_wall_i v_i ds(2)` and Robin exit terms
 * `k_exit_i (u_i - 1) v_i ds(3)`.  h_wall_w[n_vert] = sum over the exterior wall facets of a vertex of area/3;
 * h_exit_facets[n_exit][3] / h_exit_area[n_exit] = the exterior exit facets; h_jwall, h_kexit [batch][8].
 * h_wall_w == NULL switches back to the as-executed form (the default).                                        */
int gmpnp_set_facet_terms_3d(gmpnp_handle* h, const double* h_wall_w, const int* h_exit_facets,
                             const double* h_exit_area, int n_exit, const double* h_jwall, const double* h_kexit,
                             int batch);

/* L2 projection of grad(u_i) onto P1 vectors for all 9 components (dolfin project(grad(u), W) and
 * project(-grad(u_p), W), 3D:884-909): d_g[batch][n_vert][9][3] = M^-1 b with the consistent P1 mass matrix,
 * Jacobi-preconditioned CG with n_iter iterations (40 reach round-off on the reference meshes).  Post-processing. */
int gmpnp_grad_project_3d(gmpnp_handle* h, const double* d_u, double* d_g, int n_iter, void* stream);

/* Median over the vertices of component `comp` for every problem (np.median, 3D:817-820):
 * d_med[batch].                                                                            */
int gmpnp_median_3d(gmpnp_handle* h, const double* d_u, int comp, double* d_med, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GMPNP_H */
