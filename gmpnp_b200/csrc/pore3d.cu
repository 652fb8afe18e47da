// 3D cylindrical pore: P1 tetrahedral assembly into fixed-pattern BSR-9, BSR SpMV, block-Jacobi +
// z-slab coarse correction, restarted GMRES, damped Newton.
//
// Replaces the FEniCS work behind `solve(F == 0, u, bcs, {newton, mumps, relaxation 0.9})`
// (3D/MPNP_CO2ER_pore.py:789-799): FFC element kernels for the forms 3D:503-769 (volume terms
// only, as executed -- SURVEY finding 3), SystemAssembler scatter, DirichletBC.apply (3D:460-467),
// the MUMPS LU and dolfin's NewtonSolver loop.
//
// Assembly is atomic-free and deterministic: a per-tet pass computes the quadrature moments of the
// rational (steric) coefficient fields once, then one warp per BSR block GATHERS the contributions
// of the tets sharing that vertex pair in a fixed order and expands its 9x9 entries.
//
// HBM layout:  u[problem][vertex][9];  J[problem][block][9][9] (BSR, row-major blocks, blocks of a
// row sorted by column);  per-tet moment records mom[problem][tet][60], Fe[problem][tet][4][9].
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <numeric>
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace pore3d {

constexpr int NS = 8;
constexpr int NC = 9;
constexpr int NMOM = 64;     // mD[4] | mUD2[8][4] | iUD[8] | Ga[4] | gpa[4] | SU[8] | eps_r(mean) | pad[3]  (512-B records)
constexpr int M_MD = 0, M_UD2 = 4, M_IUD = 36, M_GA = 44, M_GPA = 48, M_SU = 52, M_EPS = 60;
#ifndef TET_MIN_BLOCKS
#define TET_MIN_BLOCKS 2
#endif
constexpr int NZ = 16;       // z-slabs of the coarse space
constexpr int NCO = NZ * NC; // coarse dimension (144)
constexpr size_t MEDIAN_SORT_BYTES = 200 * 1024;   // largest shared-memory sort (16384 vertices); radix select beyond

// FIAT default tetrahedron schemes (oracle/quadrature.py): degree 3 -> 5-point
// Zienkiewicz-Taylor rule for the residual, degree 4 -> 14-point Keast rule for the Jacobian.
// Barycentric coordinates (lambda0 = 1 - x - y - z) and weights summing to 1.
__constant__ double QF_L[5][4];
__constant__ double QF_W[5];
__constant__ double QJ_L[14][4];
__constant__ double QJ_W[14];

static void upload_rules() {       // per handle creation: constant memory belongs to the current device's context
    double fx[5][3] = {{0.25, 0.25, 0.25}, {0.5, 1.0 / 6.0, 1.0 / 6.0}, {1.0 / 6.0, 0.5, 1.0 / 6.0},
                       {1.0 / 6.0, 1.0 / 6.0, 0.5}, {1.0 / 6.0, 1.0 / 6.0, 1.0 / 6.0}};
    double fw[5] = {-0.8, 0.45, 0.45, 0.45, 0.45};
    const double a1 = 0.6984197043243866, b1 = 0.1005267652252045;
    const double a2 = 0.0568813795204234, b2 = 0.3143728734931922;
    double jx[14][3] = {{0.0, 0.5, 0.5}, {0.5, 0.0, 0.5}, {0.5, 0.5, 0.0}, {0.5, 0.0, 0.0}, {0.0, 0.5, 0.0},
                        {0.0, 0.0, 0.5}, {a1, b1, b1}, {b1, b1, b1}, {b1, b1, a1}, {b1, a1, b1},
                        {a2, b2, b2}, {b2, b2, b2}, {b2, b2, a2}, {b2, a2, b2}};
    double jw[14];
    for (int q = 0; q < 6; ++q) jw[q] = 0.0190476190476190;
    for (int q = 6; q < 10; ++q) jw[q] = 0.0885898247429807;
    for (int q = 10; q < 14; ++q) jw[q] = 0.1328387466855907;
    double fl[5][4], jl[14][4];
    for (int q = 0; q < 5; ++q) {
        fl[q][0] = 1.0 - (fx[q][0] + fx[q][1] + fx[q][2]);
        for (int d = 0; d < 3; ++d) fl[q][d + 1] = fx[q][d];
    }
    for (int q = 0; q < 14; ++q) {
        jl[q][0] = 1.0 - (jx[q][0] + jx[q][1] + jx[q][2]);
        for (int d = 0; d < 3; ++d) jl[q][d + 1] = jx[q][d];
    }
    cudaMemcpyToSymbol(QF_L, fl, sizeof(fl));
    cudaMemcpyToSymbol(QF_W, fw, sizeof(fw));
    cudaMemcpyToSymbol(QJ_L, jl, sizeof(jl));
    cudaMemcpyToSymbol(QJ_W, jw, sizeof(jw));
}

// ---------------------------------------------------------------------------------------
// Kernel A: per (problem, tet) quadrature moments (14-point rule) and element residual (5-point)
// ---------------------------------------------------------------------------------------
// The per-tet arithmetic is shared by the two layouts of the assembly (one problem per CTA / one problem per lane, see
// "batch-lane assembly" below); only the strides differ: nodal values at up[(v * NC + i) * US], parameters through the
// accessor P(k), moments at mo[k * MS], element residual at fe[(a * NC + i) * FS].
struct ParamsCta {           // parameter record of the CTA's problem in shared memory
    const double* P;
    __device__ __forceinline__ double operator()(int k) const { return P[k]; }
};
struct ParamsLane {          // records of 32 problems in shared memory, transposed: [k][lane], already offset by the lane
    const double* P;
    __device__ __forceinline__ double operator()(int k) const { return P[k * 32]; }
};

template <int US, int MS, int FS, class PA>
__device__ __forceinline__ void tet_core(const PA P, const int* __restrict__ tet, const double* __restrict__ geomt,
                                         const double* __restrict__ up, const double* __restrict__ unp,
                                         double* __restrict__ mo, double* __restrict__ fe, int want_jac, int want_res) {
    int v[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) v[a] = tet[a];
    double g[4][3];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int d = 0; d < 3; ++d) g[a][d] = geomt[a * 3 + d];
    const double vol = geomt[12];
    double U[4][NC];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int i = 0; i < NC; ++i) U[a][i] = up[((long)v[a] * NC + i) * US];
    // gradients: G = sum_i nu_i grad u_i, gp = grad p
    double G[3] = {0, 0, 0}, gp[3] = {0, 0, 0};
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < NS; ++i) s += P(GMPNP_P_NU + i) * U[a][i];
#pragma unroll
        for (int d = 0; d < 3; ++d) { G[d] += s * g[a][d]; gp[d] += U[a][NS] * g[a][d]; }
    }
    double Ga[4], gpa[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        Ga[a] = G[0] * g[a][0] + G[1] * g[a][1] + G[2] * g[a][2];
        gpa[a] = gp[0] * g[a][0] + gp[1] * g[a][1] + gp[2] * g[a][2];
    }
    double SU[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) SU[i] = U[0][i] + U[1][i] + U[2][i] + U[3][i];

    if (want_jac) {
        double mD[4] = {0, 0, 0, 0}, iUD[NS], m2[NS][4];
#pragma unroll
        for (int i = 0; i < NS; ++i) { iUD[i] = 0.0; m2[i][0] = m2[i][1] = m2[i][2] = m2[i][3] = 0.0; }
        for (int q = 0; q < 14; ++q) {
            const double l0 = QJ_L[q][0], l1 = QJ_L[q][1], l2 = QJ_L[q][2], l3 = QJ_L[q][3];
            const double W = QJ_W[q] * vol;
            double uq[NS], S = 0.0;
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                uq[i] = l0 * U[0][i] + l1 * U[1][i] + l2 * U[2][i] + l3 * U[3][i];
                S += P(GMPNP_P_NU + i) * uq[i];
            }
            const double D = 1.0 / (1.0 - S);
            const double WD = W * D, WD2 = WD * D;
            mD[0] += WD * l0; mD[1] += WD * l1; mD[2] += WD * l2; mD[3] += WD * l3;
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                const double tq = uq[i];
                iUD[i] += WD * tq;
                const double w2 = WD2 * tq;
                m2[i][0] += w2 * l0; m2[i][1] += w2 * l1; m2[i][2] += w2 * l2; m2[i][3] += w2 * l3;
            }
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) mo[(M_MD + b) * MS] = mD[b];
#pragma unroll
        for (int i = 0; i < NS; ++i)
#pragma unroll
            for (int b = 0; b < 4; ++b) mo[(M_UD2 + i * 4 + b) * MS] = m2[i][b];
#pragma unroll
        for (int i = 0; i < NS; ++i) mo[(M_IUD + i) * MS] = iUD[i];
#pragma unroll
        for (int a = 0; a < 4; ++a) { mo[(M_GA + a) * MS] = Ga[a]; mo[(M_GPA + a) * MS] = gpa[a]; }
#pragma unroll
        for (int i = 0; i < NS; ++i) mo[(M_SU + i) * MS] = SU[i];
        {
            const double wm = (P(GMPNP_P_EPSC) * SU[NS - 1] + P(GMPNP_P_EPSH) * SU[0]) * 0.25;
            mo[M_EPS * MS] = P(GMPNP_P_EPSW) * ((55.0 - wm) / 55.0) + 6.0 * (wm / 55.0);
        }
    }
    if (want_res) {
        // ---- element residual, 5-point rule (one negative weight) ---------------------------
        double sUD[NS], R[5][4];
#pragma unroll
        for (int i = 0; i < NS; ++i) sUD[i] = 0.0;
#pragma unroll
        for (int i = 0; i < 5; ++i) R[i][0] = R[i][1] = R[i][2] = R[i][3] = 0.0;
        const double kW = P(GMPNP_P_KW), kA = P(GMPNP_P_KA), kB = P(GMPNP_P_KB);
        const double kA2 = P(GMPNP_P_KA2), kB2 = P(GMPNP_P_KB2), kw1 = P(GMPNP_P_KW1);
        for (int q = 0; q < 5; ++q) {
            const double l[4] = {QF_L[q][0], QF_L[q][1], QF_L[q][2], QF_L[q][3]};
            const double W = QF_W[q] * vol;
            double uq[NS], S = 0.0;
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                uq[i] = l[0] * U[0][i] + l[1] * U[1][i] + l[2] * U[2][i] + l[3] * U[3][i];
                S += P(GMPNP_P_NU + i) * uq[i];
            }
            const double WD = W / (1.0 - S);
#pragma unroll
            for (int i = 0; i < NS; ++i) sUD[i] += WD * uq[i];
            const double w = kW * uq[0] * uq[1], a = kA * uq[1] * uq[2], b = kB * uq[4] * uq[1];
            const double a2 = kA2 * uq[3], b2 = kB2 * uq[2];
            double mr[5];
            mr[0] = P(GMPNP_P_S) * (w - kw1);
            mr[1] = P(GMPNP_P_S + 1) * (w + a + b - kw1 - a2 - b2);
            mr[2] = P(GMPNP_P_S + 2) * (a + b2 - a2 - b);
            mr[3] = P(GMPNP_P_S + 3) * (a2 - a);
            mr[4] = P(GMPNP_P_S + 4) * (b - b2);
#pragma unroll
            for (int i = 0; i < 5; ++i)
#pragma unroll
                for (int aa = 0; aa < 4; ++aa) R[i][aa] += W * l[aa] * mr[i];
        }
        const double kappa = P(GMPNP_P_KAPPA);
        // K_ac = vol g_a.g_c
        double K[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) K[a][c] = vol * (g[a][0] * g[c][0] + g[a][1] * g[c][1] + g[a][2] * g[c][2]);
        double rho[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            double dn[4], dsum = 0.0;
            if (kappa != 0.0) {
#pragma unroll
                for (int a = 0; a < 4; ++a) { dn[a] = U[a][i] - unp[((long)v[a] * NC + i) * US]; dsum += dn[a]; }
            } else {
#pragma unroll
                for (int a = 0; a < 4; ++a) dn[a] = 0.0;
            }
            const double zi = P(GMPNP_P_Z + i);
            const double iU = 0.25 * vol * SU[i];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                double f = kappa * (vol * 0.05) * (dsum + dn[a]);
                f += K[a][0] * U[0][i] + K[a][1] * U[1][i] + K[a][2] * U[2][i] + K[a][3] * U[3][i];
                f += zi * gpa[a] * iU + Ga[a] * sUD[i];
                if (i < 5) f += R[i][a];
                fe[(a * NC + i) * FS] = f;
                rho[a] += P(GMPNP_P_ZC0 + i) * U[a][i];
            }
        }
        const double wm = (P(GMPNP_P_EPSC) * SU[NS - 1] + P(GMPNP_P_EPSH) * SU[0]) * 0.25;
        const double epsm = P(GMPNP_P_EPSW) * ((55.0 - wm) / 55.0) + 6.0 * (wm / 55.0);
        const double rs = rho[0] + rho[1] + rho[2] + rho[3];
#pragma unroll
        for (int a = 0; a < 4; ++a)
            fe[(a * NC + NS) * FS] = -gpa[a] * vol * epsm + P(GMPNP_P_Q) * (vol * 0.05) * (rs + rho[a]);
    }
}

__global__ void __launch_bounds__(128, TET_MIN_BLOCKS)
tet_moments_kernel(int n_tet, int n_vert, const int* __restrict__ tets, const double* __restrict__ geom,
                   const double* __restrict__ params, const double* __restrict__ u, const double* __restrict__ un,
                   double* __restrict__ mom, double* __restrict__ Fe, int want_jac, int want_res) {
    __shared__ double P[GMPNP_NPAR];
    const int prob = blockIdx.y;
    for (int i = threadIdx.x; i < GMPNP_NPAR; i += blockDim.x) P[i] = params[(long)prob * GMPNP_NPAR + i];
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tet) return;
    tet_core<1, 1, 1>(ParamsCta{P}, tets + (long)t * 4, geom + (long)t * 13, u + (long)prob * n_vert * NC,
                      un + (long)prob * n_vert * NC, mom + ((long)prob * n_tet + t) * NMOM,
                      Fe + ((long)prob * n_tet + t) * 36, want_jac, want_res);
}

// ---------------------------------------------------------------------------------------
// Kernel B: residual gather per (problem, vertex) + Dirichlet rows
// ---------------------------------------------------------------------------------------
// Facet terms of the INTENDED boundary physics (3D:474-499, 560-750; dead code as executed -- SURVEY finding 3;
// live in 3D/rxn_diff_CO2ER_pore.py:480-511): `J_wall_i v_i ds(2)` = J_wall_i * (lumped wall area of the vertex) and the
// Robin exit term `k_i (u_i - 1) v_i ds(3)` = k_i sum_w E_vw (u_w,i - 1) with the exit-facet mass matrix E on the
// BSR pattern.  wall_w == nullptr: as executed (no facet terms).
__global__ void residual_gather_kernel(int n_vert, int n_tet, int n_dir, const int* __restrict__ node_ptr,
                                       const int* __restrict__ node_src, const int* __restrict__ dir_flag,
                                       const double* __restrict__ dir_val, const double* __restrict__ Fe,
                                       const double* __restrict__ u, double* __restrict__ F,
                                       const double* __restrict__ wall_w, const double* __restrict__ exit_m,
                                       const int* __restrict__ exit_flag, const double* __restrict__ bc,
                                       const int* __restrict__ row_ptr, const int* __restrict__ col_idx) {
    const int prob = blockIdx.y;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)n_vert * NC) return;
    const int v = (int)(idx / NC), i = (int)(idx % NC);
    const int df = dir_flag[idx];
    double f;
    if (df >= 0) {
        f = u[(long)prob * n_vert * NC + idx] - dir_val[(long)prob * n_dir + df];
    } else {
        f = 0.0;
        const double* fe = Fe + (long)prob * n_tet * 36;
        for (int s = node_ptr[v]; s < node_ptr[v + 1]; ++s) {
            const int src = node_src[s];
            f += fe[(long)(src >> 2) * 36 + (src & 3) * NC + i];
        }
        if (wall_w != nullptr && i < NS) {
            const double* bcp = bc + (long)prob * 16;
            f += bcp[i] * wall_w[v];
            if (exit_flag[v]) {
                const double* up = u + (long)prob * n_vert * NC;
                double e = 0.0;
                for (int s = row_ptr[v]; s < row_ptr[v + 1]; ++s) e += exit_m[s] * (up[(long)col_idx[s] * NC + i] - 1.0);
                f += bcp[8 + i] * e;
            }
        }
    }
    F[(long)prob * n_vert * NC + idx] = f;
}

// ---------------------------------------------------------------------------------------
// Kernel C: BSR gather assembly, one warp per (problem, block)
// ---------------------------------------------------------------------------------------
// Every entry of a block is linear in a few sums over the tets that share the vertex pair, so a warp first
// ACCUMULATES those sums over the block's contributions (tet, a, b) -- lane-parallel, one sum per lane -- and
// only then EXPANDS the 81 entries:
//   lanes  0..7  Q_i   = sum Ga_a int(u_i phi_b D^2) + (g_a.g_b) int(u_i D)        (steric, row i)
//   lanes  8..15 Pc_i  = sum (g_a.g_b) (vol/4) SU_i                                 (dF_i/dp before z_i)
//   lanes 16..23 T_s   = sum cT SU_s,  cT = vol/60 (a == b) or vol/120              (reaction moments; the nodal
//                        parts (u_a,s + u_b,s) sum cT are added after the loop, va/vb are fixed per block)
//   all lanes:   dsum = sum kappa M_ab + K_ab + Ga_a mD_b,  zsum = sum gpa_a vol/4,  ppsum = -sum K_ab eps_r,
//                sMab = sum M_ab,  scT = sum cT
//   entry(i,j) = cQ Q_i + cD dsum + cDz zsum + sum_r rc_r T_{rs_r} + cP Pc_i + cE zsum + cM sMab + c8 ppsum
struct EntryConst {          // per-lane constants of entry (i, j) of a 9x9 block
    int i, i7, valid;
    int rs[3];               // 0..7 species selector, 8 = constant (-> sMab)
    double cQ, cD, cDz, cP, cE, cM, c8;
    double rc[3];            // reaction derivative coefficients
};

__device__ void entry_consts(const double* P, int e, EntryConst& E) {
    E.valid = e < 81;
    const int i = E.valid ? e / 9 : 0, j = E.valid ? e % 9 : 0;
    E.i = i; E.i7 = i & 7;
    const bool ss = (i < NS && j < NS);
    E.cQ = ss ? P[GMPNP_P_NU + j] : 0.0;
    E.cD = (ss && i == j) ? 1.0 : 0.0;
    E.cDz = (ss && i == j) ? P[GMPNP_P_Z + i] : 0.0;
    E.cP = (i < NS && j == NS) ? P[GMPNP_P_Z + i] : 0.0;
    double depsj = 0.0;
    if (j == 0) depsj = (6.0 - P[GMPNP_P_EPSW]) / 55.0 * P[GMPNP_P_EPSH];
    if (j == NS - 1) depsj = (6.0 - P[GMPNP_P_EPSW]) / 55.0 * P[GMPNP_P_EPSC];
    E.cE = (i == NS && j < NS) ? -depsj : 0.0;
    E.cM = (i == NS && j < NS) ? P[GMPNP_P_Q] * P[GMPNP_P_ZC0 + j] : 0.0;
    E.c8 = (i == NS && j == NS) ? 1.0 : 0.0;
    for (int t = 0; t < 3; ++t) { E.rc[t] = 0.0; E.rs[t] = 8; }
    if (i >= 5 || j >= 5) return;
    const double kW = P[GMPNP_P_KW], kA = P[GMPNP_P_KA], kB = P[GMPNP_P_KB];
    const double kA2 = P[GMPNP_P_KA2], kB2 = P[GMPNP_P_KB2];
    const double s = P[GMPNP_P_S + i];
    // species indices: H 0, OH 1, HCO3 2, CO32 3, CO2 4 ; selector 8 = constant
    auto set = [&](int t, double c, int sel) { E.rc[t] = s * c; E.rs[t] = sel; };
    switch (i * 5 + j) {
        case 0 * 5 + 0: set(0, kW, 1); break;
        case 0 * 5 + 1: set(0, kW, 0); break;
        case 1 * 5 + 0: set(0, kW, 1); break;
        case 1 * 5 + 1: set(0, kW, 0); set(1, kA, 2); set(2, kB, 4); break;
        case 1 * 5 + 2: set(0, kA, 1); set(1, -kB2, 8); break;
        case 1 * 5 + 3: set(0, -kA2, 8); break;
        case 1 * 5 + 4: set(0, kB, 1); break;
        case 2 * 5 + 1: set(0, kA, 2); set(1, -kB, 4); break;
        case 2 * 5 + 2: set(0, kA, 1); set(1, kB2, 8); break;
        case 2 * 5 + 3: set(0, -kA2, 8); break;
        case 2 * 5 + 4: set(0, -kB, 1); break;
        case 3 * 5 + 1: set(0, -kA, 2); break;
        case 3 * 5 + 2: set(0, -kA, 1); break;
        case 3 * 5 + 3: set(0, kA2, 8); break;
        case 4 * 5 + 1: set(0, kB, 4); break;
        case 4 * 5 + 2: set(0, -kB2, 8); break;
        case 4 * 5 + 4: set(0, kB, 1); break;
        default: break;
    }
}

constexpr int ASM_WARPS = 8;

__global__ void __launch_bounds__(ASM_WARPS * 32, 4)
assemble_bsr_kernel(int n_blocks, int n_vert, int n_tet, const int* __restrict__ blk_ptr,
                    const int* __restrict__ blk_src, const double2* __restrict__ blk_geo,
                    const int* __restrict__ blk_row, const int* __restrict__ col_idx,
                    const int* __restrict__ dir_flag, const double* __restrict__ params, const double* __restrict__ u,
                    const double* __restrict__ mom, double* __restrict__ J, const double* __restrict__ exit_m,
                    const double* __restrict__ bc) {
    __shared__ double P[GMPNP_NPAR];
    __shared__ double sums[ASM_WARPS][32];   // per warp: Q[0..7] | Pc[8..15] | Tt[16..23] | sMab [24]
    // per-entry constants of the 81 block entries, shared by the CTA (kept out of registers: the kernel is bound by
    // the latency of its dependent gather loads, so occupancy matters more than a few shared-memory reads)
    __shared__ double Ed[10][96];            // cQ, cD, cZ (= cDz + cE), cP, cM, c8, rc0, rc1, rc2, cX (exit Robin k_i)
    __shared__ int Ei[5][96];                // i7, i, rs0, rs1, rs2
    const int prob = blockIdx.y;
    for (int i = threadIdx.x; i < GMPNP_NPAR; i += blockDim.x) P[i] = params[(long)prob * GMPNP_NPAR + i];
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x < 96) {
        EntryConst E;
        entry_consts(P, threadIdx.x, E);
        const int e = threadIdx.x;
        Ed[0][e] = E.cQ; Ed[1][e] = E.cD; Ed[2][e] = E.cDz + E.cE; Ed[3][e] = E.cP; Ed[4][e] = E.cM; Ed[5][e] = E.c8;
        Ed[6][e] = E.rc[0]; Ed[7][e] = E.rc[1]; Ed[8][e] = E.rc[2];
        Ed[9][e] = (exit_m != nullptr && E.cD != 0.0) ? bc[(long)prob * 16 + 8 + E.i] : 0.0;
        Ei[0][e] = E.i7; Ei[1][e] = E.i; Ei[2][e] = E.rs[0]; Ei[3][e] = E.rs[1]; Ei[4][e] = E.rs[2];
    }
    __syncthreads();
    const double kappa = P[GMPNP_P_KAPPA];
    const double* up = u + (long)prob * n_vert * NC;
    const double* mo = mom + (long)prob * n_tet * NMOM;
    double* Jp = J + (long)prob * n_blocks * 81;
    // accumulation phase: the warp takes FOUR contributions (tet, a, b) of the block per pass, one per 8-lane group;
    // lane li of a group accumulates the three moment sums of species li (Q_li, Pc_li, T_li) and, redundantly within
    // the group, the five scalar sums; the four groups are combined with two xor-shuffle steps per block.
    const int cs = lane >> 3, li = lane & 7;
    double* sw = sums[w];
    for (int blk = blockIdx.x * ASM_WARPS + w; blk < n_blocks; blk += gridDim.x * ASM_WARPS) {
        const int va = blk_row[blk], vb = col_idx[blk];
        double accQ = 0.0, accP = 0.0, accT = 0.0, dsum = 0.0, zsum = 0.0, ppsum = 0.0, sMab = 0.0, scT = 0.0;
        const int s0 = blk_ptr[blk], s1 = blk_ptr[blk + 1];
        for (int sb = s0; sb < s1; sb += 32) {
            const int cnt = min(32, s1 - sb);
            int my_src = 0;
            double2 my_kv = make_double2(0.0, 0.0);
            if (lane < cnt) { my_src = blk_src[sb + lane]; my_kv = blk_geo[sb + lane]; }
            for (int q0 = 0; q0 < cnt; q0 += 4) {
                const int q = q0 + cs;                              // this group's contribution (may be past the end)
                const int src = __shfl_sync(0xffffffffu, my_src, q & 31);
                double kab = __shfl_sync(0xffffffffu, my_kv.x, q & 31), vol = __shfl_sync(0xffffffffu, my_kv.y, q & 31);
                const bool live = q < cnt;
                if (!live) { kab = 0.0; vol = 0.0; }                // a dead slot contributes exact zeros
                const int t = live ? (src >> 4) : 0, a = (src >> 2) & 3, b = src & 3;
                const double* m = mo + (long)t * NMOM;
                const double Ga = live ? m[M_GA + a] : 0.0, gpa = m[M_GPA + a], mDb = m[M_MD + b], eps = m[M_EPS];
                const double mq = m[M_UD2 + 4 * li + b], mi = m[M_IUD + li], ms = m[M_SU + li];
                const double Kab = kab * vol, mb = 0.25 * vol;
                const double Mab = vol * ((a == b) ? 0.1 : 0.05);
                const double cT = vol * ((a == b) ? (1.0 / 60.0) : (1.0 / 120.0));
                dsum += kappa * Mab + Kab + Ga * mDb;
                zsum += gpa * mb;
                ppsum -= Kab * eps;
                sMab += Mab;
                scT += cT;
                accQ += Ga * mq + kab * mi;
                accP += (kab * mb) * ms;
                accT += cT * ms;
            }
        }
        // combine the four groups (every lane ends up with the block totals)
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
            accQ += __shfl_xor_sync(0xffffffffu, accQ, o); accP += __shfl_xor_sync(0xffffffffu, accP, o);
            accT += __shfl_xor_sync(0xffffffffu, accT, o); dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
            zsum += __shfl_xor_sync(0xffffffffu, zsum, o); ppsum += __shfl_xor_sync(0xffffffffu, ppsum, o);
            sMab += __shfl_xor_sync(0xffffffffu, sMab, o); scT += __shfl_xor_sync(0xffffffffu, scT, o);
        }
        accT += (up[(long)va * NC + li] + up[(long)vb * NC + li]) * scT;      // nodal part of T_s
        __syncwarp();                                    // previous block's expansion reads are done
        if (cs == 0) { sw[li] = accQ; sw[8 + li] = accP; sw[16 + li] = accT; }
        if (lane == 24) sw[24] = sMab;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int e = lane + 32 * k;
            if (e >= 81) continue;
            const int i7 = Ei[0][e], ei = Ei[1][e];
            double val = Ed[0][e] * sw[i7] + Ed[3][e] * sw[8 + i7];
            val += Ed[1][e] * dsum + Ed[2][e] * zsum + Ed[4][e] * sMab + Ed[5][e] * ppsum;
            val += Ed[6][e] * sw[16 + Ei[2][e]] + Ed[7][e] * sw[16 + Ei[3][e]] + Ed[8][e] * sw[16 + Ei[4][e]];
            if (exit_m != nullptr) val += Ed[9][e] * exit_m[blk];
            // Dirichlet rows: identity
            if (dir_flag[(long)va * NC + ei] >= 0) val = (va == vb && e == ei * 10) ? 1.0 : 0.0;
            Jp[(long)blk * 81 + e] = val;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Batch-lane assembly (batches of >= LANES_MIN_BATCH problems on one mesh): one problem per LANE
// ---------------------------------------------------------------------------------------
// All problems of a batch share the mesh, so the gather lists, the geometry and the Dirichlet flags are the same for
// every problem.  The kernels above spend ~90 % of their instructions on that shared index work (warp per (problem, block):
// list loads, shuffles, address arithmetic, cross-lane reductions, 524 warp-instructions per block for ~30 fp64
// operations per contribution).  Here a warp takes one tet / one BSR block for 32 PROBLEMS at once, lane = problem:
// index work is warp-uniform and done once per 32 problems, every sum is lane-local (no shuffles, no reductions, fixed
// ascending tet order -> deterministic and independent of the position in the batch), and the intermediate arrays are
// stored problem-minor so that every load and store of a warp is one contiguous 256-byte segment:
//     uT  [group][vertex][component][32]     nodal values   (lanes_transpose_kernel, from the caller's [problem][v][c])
//     momT[group][tet][NMT = 61][32]         moments        (tet_moments_lanes_kernel)
// The caller-facing layouts do not change: Fe [problem][tet][36] and J [problem][block][81] are written through a
// shared-memory transposition (rows padded to 33 doubles: conflict-free both ways) as contiguous 288- / 648-byte rows.
// Lanes past the batch size compute on a copy of the last problem and store nothing.
constexpr int NMT = 61;                 // moments per tet (the record of NMOM without its padding)
constexpr int LANES_MIN_BATCH = 24;     // below this the warp-per-(problem, block) kernels waste fewer lanes
constexpr int TL_WARPS = 4;             // tets per CTA of tet_moments_lanes_kernel
constexpr int BL_WARPS = 4;             // blocks in flight per CTA of assemble_bsr_lanes_kernel
constexpr int BL_UNROLL = 3;            // contributions in flight per warp
constexpr int TR_LD = 33;               // leading dimension of the transposition tiles

__global__ void __launch_bounds__(256)
lanes_transpose_kernel(int batch, long n, const double* __restrict__ u, double* __restrict__ uT) {
    __shared__ double tile[32][TR_LD];
    const int g = blockIdx.y, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long i0 = (long)blockIdx.x * 32;
    for (int r = ty; r < 32; r += 8) {
        const int p = min(g * 32 + r, batch - 1);
        tile[r][tx] = (i0 + tx < n) ? u[(long)p * n + i0 + tx] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8)
        if (i0 + r < n) uT[((long)g * n + i0 + r) * 32 + tx] = tile[tx][r];
}

// dynamic shared memory: Ps[GMPNP_NPAR][32] | Tr[TL_WARPS][36][TR_LD]
constexpr size_t TL_SMEM = sizeof(double) * (GMPNP_NPAR * 32 + TL_WARPS * 36 * TR_LD);

__global__ void __launch_bounds__(TL_WARPS * 32, TET_MIN_BLOCKS)
tet_moments_lanes_kernel(int batch, int n_tet, int n_vert, const int* __restrict__ tets, const double* __restrict__ geom,
                         const double* __restrict__ params, const double* __restrict__ uT,
                         const double* __restrict__ unT, double* __restrict__ momT, double* __restrict__ Fe,
                         int want_jac, int want_res) {
    extern __shared__ double sm_tl[];
    double* Ps = sm_tl;
    const int g = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < GMPNP_NPAR * 32; i += blockDim.x)    // [k][problem]: conflict-free stores
        Ps[i] = params[(long)min(g * 32 + (i & 31), batch - 1) * GMPNP_NPAR + (i >> 5)];
    __syncthreads();
    double* Tr = sm_tl + GMPNP_NPAR * 32 + w * 36 * TR_LD;
    const long goff = (long)g * n_vert * NC * 32 + lane;
    for (int t = blockIdx.x * TL_WARPS + w; t < n_tet; t += gridDim.x * TL_WARPS) {
        tet_core<32, 32, TR_LD>(ParamsLane{Ps + lane}, tets + (long)t * 4, geom + (long)t * 13, uT + goff, unT + goff,
                                momT + ((long)g * n_tet + t) * NMT * 32 + lane, Tr + lane, want_jac, want_res);
        if (want_res) {
            __syncwarp();
            int q = 0, e = lane;                                      // flat index q * 36 + e over the warp's 32 rows
#pragma unroll 4
            for (int it = 0; it < 36; ++it) {
                if (e >= 36) { e -= 36; ++q; }
                const int p = g * 32 + q;
                if (p < batch) Fe[((long)p * n_tet + t) * 36 + e] = Tr[e * TR_LD + q];
                e += 32;
            }
            __syncwarp();
        }
    }
}

// dynamic shared memory: Ps[GMPNP_NPAR][32] | kx[8][32] | St[WARPS][27][TR_LD]
constexpr size_t bl_smem(int warps) { return sizeof(double) * (GMPNP_NPAR * 32 + 8 * 32 + warps * 27 * TR_LD); }

// Block descriptors in PROCESSING order (host, gmpnp_create_3d): the blocks sorted along a space-filling curve of
// their row vertex, so that the warps resident at one time work on one neighbourhood of the mesh and the 15.6-KB moment
// records of its tets are re-read from L2, not from HBM (one launch per lane group for the same reason).
// meta[2k] = (block, va, vb, first contribution), meta[2k+1] = (number of contributions, Dirichlet mask of the 9 rows
// of va, 0, 0).  The loop is software-pipelined by one block: descriptors and the (tet, a, b) list of block k+1 are
// requested while block k is accumulated / expanded, so the only exposed latency per block is that of the moment loads.
// The 81 entries leave through a 27-entry (three block rows) staging tile per warp: 7 KB instead of 21 KB, which is
// what lets more than 8 warps live on an SM.
template <int UNR, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 2)
assemble_bsr_lanes_kernel(int g, int batch, int n_blocks, int n_vert, int n_tet, const int4* __restrict__ meta,
                          const int* __restrict__ blk_src, const double2* __restrict__ blk_geo,
                          const double* __restrict__ params, const double* __restrict__ uT,
                          const double* __restrict__ momT, double* __restrict__ J,
                          const double* __restrict__ exit_m, const double* __restrict__ bc) {
    extern __shared__ double sm_bl[];
    double* Ps = sm_bl;
    double* kx = sm_bl + GMPNP_NPAR * 32;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < GMPNP_NPAR * 32; i += blockDim.x)
        Ps[i] = params[(long)min(g * 32 + (i & 31), batch - 1) * GMPNP_NPAR + (i >> 5)];
    for (int i = threadIdx.x; i < 8 * 32; i += blockDim.x)
        kx[i] = (exit_m != nullptr) ? bc[(long)min(g * 32 + (i & 31), batch - 1) * 16 + 8 + (i >> 5)] : 0.0;
    __syncthreads();
    const ParamsLane P{Ps + lane};
    double* St = sm_bl + GMPNP_NPAR * 32 + 8 * 32 + w * 27 * TR_LD;
    const double* mo = momT + (long)g * n_tet * NMT * 32 + lane;
    const double* up = uT + (long)g * n_vert * NC * 32 + lane;
    const long jstride = (long)n_blocks * 81;
    const int nq = min(32, batch - g * 32);              // live problems of this lane group
    const int stride = gridDim.x * WARPS;
    int k = blockIdx.x * WARPS + w;
    int4 ma = make_int4(0, 0, 0, 0), mb = make_int4(0, 0, 0, 0);
    int my_src = 0;
    double2 my_kv = make_double2(0.0, 0.0);
    if (k < n_blocks) {
        ma = __ldg(meta + 2 * k); mb = __ldg(meta + 2 * k + 1);
        if (lane < mb.x) { my_src = __ldg(blk_src + ma.w + lane); my_kv = __ldg(blk_geo + ma.w + lane); }
    }
    for (; k < n_blocks; k += stride) {
        const int blk = ma.x, va = ma.y, vb = ma.z, s0 = ma.w, cnt = mb.x, pin = mb.y;
        int cur_src = my_src;
        double2 cur_kv = my_kv;
        const int kn = k + stride;
        if (kn < n_blocks) { ma = __ldg(meta + 2 * kn); mb = __ldg(meta + 2 * kn + 1); }     // descriptor of the next block
        const double* ua = up + (long)va * NC * 32;
        const double* ub = up + (long)vb * NC * 32;
        const double n0 = ua[0] + ub[0], n1 = ua[32] + ub[32], n2 = ua[64] + ub[64], n4 = ua[128] + ub[128];
        const double xm = (exit_m != nullptr) ? __ldg(exit_m + blk) : 0.0;
        double Q[NS], Pc[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) { Q[i] = 0.0; Pc[i] = 0.0; }
        double T0 = 0.0, T1 = 0.0, T2 = 0.0, T4 = 0.0;
        double dG = 0.0, sK = 0.0, zsum = 0.0, ppsum = 0.0, sMab = 0.0, scT = 0.0;
        for (int c0 = 0; c0 < cnt; c0 += 32) {
            if (c0 > 0) {                                    // rare: more than 32 contributions (high-valence diagonal)
                cur_src = 0; cur_kv = make_double2(0.0, 0.0);
                if (c0 + lane < cnt) { cur_src = __ldg(blk_src + s0 + c0 + lane); cur_kv = __ldg(blk_geo + s0 + c0 + lane); }
            }
            const int nc = min(32, cnt - c0);
#pragma unroll UNR
            for (int c = 0; c < nc; ++c) {
                const int src = __shfl_sync(0xffffffffu, cur_src, c);
                const double kab = __shfl_sync(0xffffffffu, cur_kv.x, c), vol = __shfl_sync(0xffffffffu, cur_kv.y, c);
                const int t = src >> 4, a = (src >> 2) & 3, b = src & 3;
                const double* m = mo + (long)t * (NMT * 32);
                const double* ma_ = m + a * 32;
                const double* mb_ = m + b * 32;
                const double Ga = ma_[M_GA * 32], gpa = ma_[M_GPA * 32], mDb = mb_[M_MD * 32], eps = m[M_EPS * 32];
                const double Kab = kab * vol, mq = 0.25 * vol;
                const double Mab = vol * ((a == b) ? 0.1 : 0.05);
                const double cT = vol * ((a == b) ? (1.0 / 60.0) : (1.0 / 120.0));
                const double kmq = kab * mq;
                dG += Ga * mDb;
                sK += Kab;
                zsum += gpa * mq;
                ppsum -= Kab * eps;
                sMab += Mab;
                scT += cT;
#pragma unroll
                for (int i = 0; i < NS; ++i) {
                    const double su = m[(M_SU + i) * 32];
                    Q[i] += Ga * mb_[(M_UD2 + 4 * i) * 32] + kab * m[(M_IUD + i) * 32];
                    Pc[i] += kmq * su;
                    if (i == 0) T0 += cT * su;
                    if (i == 1) T1 += cT * su;
                    if (i == 2) T2 += cT * su;
                    if (i == 4) T4 += cT * su;
                }
            }
        }
        if (kn < n_blocks) {                                 // (tet, a, b) list of the next block: lands during the expansion
            my_src = 0; my_kv = make_double2(0.0, 0.0);
            if (lane < mb.x) { my_src = __ldg(blk_src + ma.w + lane); my_kv = __ldg(blk_geo + ma.w + lane); }
        }
        T0 += n0 * scT; T1 += n1 * scT; T2 += n2 * scT; T4 += n4 * scT;          // nodal part of the reaction moments
        const double dsum = P(GMPNP_P_KAPPA) * sMab + sK + dG;
        const double kW = P(GMPNP_P_KW), kA = P(GMPNP_P_KA), kB = P(GMPNP_P_KB);
        const double r_wT1 = kW * T1, r_wT0 = kW * T0, r_aT2 = kA * T2, r_bT4 = kB * T4, r_aT1 = kA * T1, r_bT1 = kB * T1;
        const double r_b2 = P(GMPNP_P_KB2) * sMab, r_a2 = P(GMPNP_P_KA2) * sMab;
        const double deps = (6.0 - P(GMPNP_P_EPSW)) / 55.0;
        double* jp0 = J + ((long)g * 32 * n_blocks + blk) * 81 + lane;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {                 // three block rows per pass through the staging tile
            __syncwarp();                                // the previous pass has been read out of St
#pragma unroll
            for (int ii = 0; ii < 3; ++ii) {
                const int i = ch * 3 + ii;
                if ((pin >> i) & 1) {                    // Dirichlet row: identity (warp-uniform branch)
#pragma unroll
                    for (int j = 0; j < NC; ++j) St[(ii * NC + j) * TR_LD + lane] = (va == vb && i == j) ? 1.0 : 0.0;
                    continue;
                }
#pragma unroll
                for (int j = 0; j < NC; ++j) {
                    double val;
                    if (i < NS && j < NS) {
                        val = P(GMPNP_P_NU + j) * Q[i];
                        if (i == j) val += dsum + P(GMPNP_P_Z + i) * zsum + kx[(i & 7) * 32 + lane] * xm;
                        if (i < 5 && j < 5) {
                            double rr = 0.0;
                            switch (i * 5 + j) {                     // d(reaction source of i)/d u_j, SURVEY App. A.2
                                case 0: rr = r_wT1; break;
                                case 1: rr = r_wT0; break;
                                case 5: rr = r_wT1; break;
                                case 6: rr = r_wT0 + r_aT2 + r_bT4; break;
                                case 7: rr = r_aT1 - r_b2; break;
                                case 8: rr = -r_a2; break;
                                case 9: rr = r_bT1; break;
                                case 11: rr = r_aT2 - r_bT4; break;
                                case 12: rr = r_aT1 + r_b2; break;
                                case 13: rr = -r_a2; break;
                                case 14: rr = -r_bT1; break;
                                case 16: rr = -r_aT2; break;
                                case 17: rr = -r_aT1; break;
                                case 18: rr = r_a2; break;
                                case 21: rr = r_bT4; break;
                                case 22: rr = -r_b2; break;
                                case 24: rr = r_bT1; break;
                                default: break;
                            }
                            val += P(GMPNP_P_S + (i < 5 ? i : 0)) * rr;
                        }
                    } else if (i < NS) {
                        val = P(GMPNP_P_Z + (i & 7)) * Pc[i & 7];
                    } else if (j < NS) {
                        val = P(GMPNP_P_Q) * P(GMPNP_P_ZC0 + j) * sMab;
                        if (j == 0) val -= deps * P(GMPNP_P_EPSH) * zsum;
                        if (j == NS - 1) val -= deps * P(GMPNP_P_EPSC) * zsum;
                    } else {
                        val = ppsum;
                    }
                    St[(ii * NC + j) * TR_LD + lane] = val;
                }
            }
            __syncwarp();
            if (lane < 27) {                             // 27 contiguous doubles of problem q's block per store
                double* jp = jp0 + ch * 27;
                const double* sp = St + lane * TR_LD;
#pragma unroll 4
                for (int q = 0; q < nq; ++q) {
                    __stcs(jp, sp[q]);                   // streaming: J must not displace the moment records in L2
                    jp += jstride;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// BSR-9 SpMV: one warp per (problem, block row)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bsr_spmv_kernel(int n_vert, int n_blocks, const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                const double* __restrict__ J, const double* __restrict__ x, double* __restrict__ y, int row0, int row1) {
    __shared__ double part[8][96];
    const int prob = blockIdx.y;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int row = row0 + blockIdx.x * 8 + w;
    if (row >= row1) return;
    const double* Jp = J + (long)prob * n_blocks * 81;
    const double* xp = x + (long)prob * n_vert * NC;
    const int j0 = lane % 9, j1 = (lane + 32) % 9, j2 = (lane + 64) % 9;
    const bool third = lane + 64 < 81;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    const int s0 = row_ptr[row], s1 = row_ptr[row + 1];
    // The 648-byte blocks of a row are contiguous and read exactly once (streaming loads, no L1 allocation); the
    // column indices of up to 32 blocks are fetched with one coalesced load and handed out by shuffle, and four
    // blocks (12 independent 256-byte requests per warp) are in flight before the first multiply.
    for (int base = s0; base < s1; base += 32) {
        const int cnt = min(32, s1 - base);
        const int mycol = (lane < cnt) ? col_idx[base + lane] : 0;
        const double* blk = Jp + (long)base * 81;
        int t = 0;
        for (; t + 4 <= cnt; t += 4, blk += 4 * 81) {
            double v[4][3];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                v[q][0] = __ldcs(blk + q * 81 + lane);
                v[q][1] = __ldcs(blk + q * 81 + lane + 32);
                v[q][2] = third ? __ldcs(blk + q * 81 + lane + 64) : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double* xc = xp + (long)__shfl_sync(0xffffffffu, mycol, t + q) * NC;
                a0 += v[q][0] * xc[j0];
                a1 += v[q][1] * xc[j1];
                a2 += v[q][2] * xc[j2];
            }
        }
        for (; t < cnt; ++t, blk += 81) {
            const double v0 = __ldcs(blk + lane), v1 = __ldcs(blk + lane + 32);
            const double v2 = third ? __ldcs(blk + lane + 64) : 0.0;
            const double* xc = xp + (long)__shfl_sync(0xffffffffu, mycol, t) * NC;
            a0 += v0 * xc[j0];
            a1 += v1 * xc[j1];
            a2 += v2 * xc[j2];
        }
    }
    part[w][lane] = a0; part[w][lane + 32] = a1; part[w][lane + 64] = a2;
    __syncwarp();
    if (lane < NC) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 9; ++j) s += part[w][lane * 9 + j];
        y[(long)prob * n_vert * NC + (long)row * NC + lane] = s;
    }
}

// ---------------------------------------------------------------------------------------
// Block-Jacobi: invert the diagonal 9x9 blocks (Gauss-Jordan, partial pivoting)
// ---------------------------------------------------------------------------------------
__global__ void bjacobi_invert_kernel(int n_vert, int n_blocks, const int* __restrict__ diag_idx,
                                      const double* __restrict__ J, double* __restrict__ Dinv) {
    const int prob = blockIdx.y;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_vert) return;
    const double* blk = J + ((long)prob * n_blocks + diag_idx[v]) * 81;
    double A[9][9], I[9][9];
    for (int i = 0; i < 9; ++i)
        for (int j = 0; j < 9; ++j) { A[i][j] = blk[i * 9 + j]; I[i][j] = (i == j) ? 1.0 : 0.0; }
    for (int c = 0; c < 9; ++c) {
        int p = c;
        double best = fabs(A[c][c]);
        for (int r = c + 1; r < 9; ++r) if (fabs(A[r][c]) > best) { best = fabs(A[r][c]); p = r; }
        if (p != c)
            for (int j = 0; j < 9; ++j) {
                double t = A[c][j]; A[c][j] = A[p][j]; A[p][j] = t;
                t = I[c][j]; I[c][j] = I[p][j]; I[p][j] = t;
            }
        const double inv = 1.0 / A[c][c];
        for (int j = 0; j < 9; ++j) { A[c][j] *= inv; I[c][j] *= inv; }
        for (int r = 0; r < 9; ++r) {
            if (r == c) continue;
            const double f = A[r][c];
            for (int j = 0; j < 9; ++j) { A[r][j] -= f * A[c][j]; I[r][j] -= f * I[c][j]; }
        }
    }
    double* o = Dinv + ((long)prob * n_vert + v) * 81;
    for (int i = 0; i < 9; ++i)
        for (int j = 0; j < 9; ++j) o[i * 9 + j] = I[i][j];
}

// ---------------------------------------------------------------------------------------
// Krylov vector kernels (one CTA per problem for reductions: deterministic)
// ---------------------------------------------------------------------------------------
__device__ double block_sum(double v, double* sh) {
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double s = 0.0;
    const int nw = (blockDim.x + 31) >> 5;
    for (int k = 0; k < nw; ++k) s += sh[k];
    return s;
}

// nrm[prob] = ||w|| ; vout[prob * vstride + :] = w / ||w||   (one CTA per problem)
__global__ void __launch_bounds__(1024)
norm_scale_kernel(long n, const double* __restrict__ w, double* __restrict__ vout, long vstride,
                  double* __restrict__ nrm) {
    __shared__ double sh[32];
    const int prob = blockIdx.x;
    const double* wp = w + (long)prob * n;
    double s = 0.0;
    for (long i = threadIdx.x; i < n; i += blockDim.x) s += wp[i] * wp[i];
    s = block_sum(s, sh);
    const double nr = sqrt(s);
    if (threadIdx.x == 0) nrm[prob] = nr;
    if (vout) {
        const double inv = (nr > 0.0) ? 1.0 / nr : 0.0;
        for (long i = threadIdx.x; i < n; i += blockDim.x) vout[(long)prob * vstride + i] = wp[i] * inv;
    }
}

// Newton update with per-problem mask: u -= relax * dx where active; dxmax/umax per problem
__global__ void __launch_bounds__(1024)
newton_update_kernel(long n, double relax, const int* __restrict__ active, const double* __restrict__ dx,
                     double* __restrict__ u, double* __restrict__ dxmax, double* __restrict__ umax) {
    __shared__ double sh1[32], sh2[32];
    const int prob = blockIdx.x;
    double m1 = 0.0, m2 = 0.0;
    if (active[prob]) {
        for (long i = threadIdx.x; i < n; i += blockDim.x) {
            const double d = dx[(long)prob * n + i];
            const double v = u[(long)prob * n + i] - relax * d;
            u[(long)prob * n + i] = v;
            m1 = fmax(m1, fabs(d)); m2 = fmax(m2, fabs(v));
        }
    }
    for (int o = 16; o >= 1; o >>= 1) {
        m1 = fmax(m1, __shfl_xor_sync(0xffffffffu, m1, o));
        m2 = fmax(m2, __shfl_xor_sync(0xffffffffu, m2, o));
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { sh1[w] = m1; sh2[w] = m2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { m1 = fmax(m1, sh1[k]); m2 = fmax(m2, sh2[k]); }
        dxmax[prob] = m1; umax[prob] = m2;
    }
}

// median of component `comp` over the vertices (np.median, 3D:817-820): bitonic sort in shared memory
__global__ void __launch_bounds__(1024)
median_kernel(int n_vert, int npow2, int comp, const double* __restrict__ u, double* __restrict__ med) {
    extern __shared__ double sv[];
    const int prob = blockIdx.x;
    for (int i = threadIdx.x; i < npow2; i += blockDim.x)
        sv[i] = (i < n_vert) ? u[((long)prob * n_vert + i) * NC + comp] : INFINITY;
    __syncthreads();
    for (int k = 2; k <= npow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const bool up = ((i & k) == 0);
                    const double a = sv[i], b = sv[ixj];
                    if ((a > b) == up) { sv[i] = b; sv[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0)
        med[prob] = (n_vert & 1) ? sv[n_vert / 2] : 0.5 * (sv[n_vert / 2 - 1] + sv[n_vert / 2]);
}

// ---------------------------------------------------------------------------------------
// Vector kernels of the mesh-partitioned mode (one problem spans several GPUs: SURVEY 8e (2)).
// Reductions are two-stage (grid-wide partial sums in a fixed chunk order, then one small CTA per
// output), so a single large vector uses the whole GPU and the result is deterministic.
// ---------------------------------------------------------------------------------------
constexpr int VEC_CHUNK = 8192;        // entries per CTA of the first reduction stage

// partial[k][chunk] = sum over the chunk of V_k[i] * w[i]      grid = (nchunks, nvec)
__global__ void __launch_bounds__(256)
vec_dot_partial_kernel(long n, long vstride, const double* __restrict__ V, const double* __restrict__ w,
                       double* __restrict__ partial, int nchunks) {
    __shared__ double sh[8];
    const int chunk = blockIdx.x, k = blockIdx.y;
    const double* vk = V + (long)k * vstride;
    const long i0 = (long)chunk * VEC_CHUNK, i1 = min(n, i0 + VEC_CHUNK);
    double s = 0.0;
    for (long i = i0 + threadIdx.x; i < i1; i += blockDim.x) s += vk[i] * w[i];
    s = block_sum(s, sh);
    if (threadIdx.x == 0) partial[(long)k * nchunks + chunk] = s;
}

// out[k] = sum_chunk partial[k][chunk]                            grid = nvec
__global__ void __launch_bounds__(256)
vec_dot_final_kernel(const double* __restrict__ partial, int nchunks, double* __restrict__ out) {
    __shared__ double sh[8];
    const int k = blockIdx.x;
    double s = 0.0;
    for (int c = threadIdx.x; c < nchunks; c += blockDim.x) s += partial[(long)k * nchunks + c];
    s = block_sum(s, sh);
    if (threadIdx.x == 0) out[k] = s;
}

// out[i] = beta * y[i] + sum_k coef[k] V_k[i]   (y may be null; out may alias y)
__global__ void __launch_bounds__(256)
vec_lincomb_kernel(long n, long vstride, int nvec, const double* __restrict__ V, const double* __restrict__ coef,
                   double beta, const double* y, double* out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = (y != nullptr) ? beta * y[i] : 0.0;
    for (int k = 0; k < nvec; ++k) s += coef[k] * V[(long)k * vstride + i];
    out[i] = s;
}

// ---- distributed z-slab coarse space (same Galerkin correction as the single-mesh path, split at the points where
// the ranks have to exchange: A_c and P^T r are summed over the ranks by the host with an all-reduce) -------------
// Ac += sum over the block rows [0, n_rows) of P^T J P (shared-memory accumulation per CTA, one flush per CTA)
__global__ void __launch_bounds__(1024)
coarse_accumulate_kernel(int n_rows, const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                         const int* __restrict__ agg, const int* __restrict__ dir_flag, const double* __restrict__ J,
                         double* __restrict__ Ac_out) {
    extern __shared__ double Ac[];          // [NCO][NCO]
    for (int i = threadIdx.x; i < NCO * NCO; i += blockDim.x) Ac[i] = 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int row = blockIdx.x * nw + w; row < n_rows; row += gridDim.x * nw) {
        const int I = agg[row];
        for (int s = row_ptr[row]; s < row_ptr[row + 1]; ++s) {
            const int col = col_idx[s];
            const int Jc = agg[col];
            for (int e = lane; e < 81; e += 32) {
                const int i = e / 9, j = e % 9;
                if (dir_flag[(long)row * NC + i] >= 0 || dir_flag[(long)col * NC + j] >= 0) continue;
                atomicAdd(&Ac[(I * NC + i) * NCO + Jc * NC + j], J[(long)s * 81 + e]);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NCO * NCO; i += blockDim.x)
        if (Ac[i] != 0.0) atomicAdd(&Ac_out[i], Ac[i]);
}

__device__ void gj_invert_pivoted(double* Ac, int* perm, double* s_pivinv, int* s_prow, int* s_nbad);

// Aci = pivoted inverse of the (all-reduced) Galerkin matrix; empty coarse rows (all members Dirichlet) -> identity
__global__ void __launch_bounds__(1024)
coarse_invert_kernel(const double* __restrict__ Ac_in, double* __restrict__ Aci, int* __restrict__ flag) {
    extern __shared__ double Ac[];          // [NCO][NCO]
    __shared__ int perm[NCO];
    __shared__ double pivinv;
    __shared__ int prow, nbad;
    for (int i = threadIdx.x; i < NCO * NCO; i += blockDim.x) Ac[i] = Ac_in[i];
    if (threadIdx.x == 0) nbad = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < NCO; i += blockDim.x) {
        bool empty = true;
        for (int j = 0; j < NCO; ++j) if (Ac[i * NCO + j] != 0.0) { empty = false; break; }
        if (empty) Ac[i * NCO + i] = 1.0;
    }
    __syncthreads();
    gj_invert_pivoted(Ac, perm, &pivinv, &prow, &nbad);
    for (int i = threadIdx.x; i < NCO * NCO; i += blockDim.x) Aci[i] = Ac[i];
    if (threadIdx.x == 0 && flag) flag[0] = nbad;
}

// rc += P^T r over the block rows [0, n_rows) (Dirichlet DOFs excluded)
__global__ void __launch_bounds__(256)
coarse_restrict_kernel(int n_rows, const int* __restrict__ agg, const int* __restrict__ dir_flag,
                       const double* __restrict__ r, double* __restrict__ rc) {
    __shared__ double sh[NCO];
    for (int i = threadIdx.x; i < NCO; i += blockDim.x) sh[i] = 0.0;
    __syncthreads();
    const long total = (long)n_rows * NC;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int v = (int)(idx / NC), i = (int)(idx % NC);
        if (dir_flag[idx] < 0) atomicAdd(&sh[agg[v] * NC + i], r[idx]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NCO; i += blockDim.x)
        if (sh[i] != 0.0) atomicAdd(&rc[i], sh[i]);
}

// z += P (Aci rc) on the block rows [0, n_rows) (Dirichlet DOFs excluded); every CTA recomputes the 144-vector
__global__ void __launch_bounds__(256)
coarse_prolong_kernel(int n_rows, const int* __restrict__ agg, const int* __restrict__ dir_flag,
                      const double* __restrict__ Aci, const double* __restrict__ rc, double* __restrict__ z) {
    __shared__ double src[NCO], yc[NCO];
    for (int i = threadIdx.x; i < NCO; i += blockDim.x) src[i] = rc[i];
    __syncthreads();
    for (int t = threadIdx.x; t < NCO; t += blockDim.x) {
        const double* A = Aci + (long)t * NCO;
        double y = 0.0;
        for (int j = 0; j < NCO; ++j) y += A[j] * src[j];
        yc[t] = y;
    }
    __syncthreads();
    const long total = (long)n_rows * NC;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int v = (int)(idx / NC), i = (int)(idx % NC);
        if (dir_flag[idx] < 0) z[idx] += yc[agg[v] * NC + i];
    }
}

// ---------------------------------------------------------------------------------------
// L2 projection of grad(u_i) onto P1 vectors (dolfin project(grad(u), W), 3D:884-909): consistent P1 mass
// matrix M (scalar, same pattern as the BSR blocks, geometry only), right-hand side b_v = sum_{t in v} vol_t/4
// grad(u)_t, Jacobi-preconditioned CG for all 27 columns (9 components x 3 directions) of a problem at once with
// per-column scalars, fixed iteration count, no host synchronisation.  Post-processing, off the hot path.
// ---------------------------------------------------------------------------------------
constexpr int NG = NC * 3;              // columns k = comp * 3 + d

// b[prob][v][k]; also x = 0, r = b, z = r / M_vv, p = z
__global__ void __launch_bounds__(256)
gradproj_rhs_kernel(int n_vert, const int* __restrict__ node_ptr, const int* __restrict__ node_src,
                    const int* __restrict__ tets, const double* __restrict__ geom, const double* __restrict__ mass,
                    const int* __restrict__ diag_idx, const double* __restrict__ u, double* __restrict__ x,
                    double* __restrict__ r, double* __restrict__ z, double* __restrict__ pvec) {
    const int prob = blockIdx.y;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)n_vert * NG) return;
    const int v = (int)(idx / NG), k = (int)(idx % NG), comp = k / 3, d = k % 3;
    const double* up = u + (long)prob * n_vert * NC;
    double b = 0.0;
    for (int s = node_ptr[v]; s < node_ptr[v + 1]; ++s) {
        const int t = node_src[s] >> 2;
        const double* ge = geom + (long)t * 13;
        double gsum = 0.0;
#pragma unroll
        for (int a = 0; a < 4; ++a) gsum += ge[a * 3 + d] * up[(long)tets[t * 4 + a] * NC + comp];
        b += 0.25 * ge[12] * gsum;
    }
    const long o = (long)prob * n_vert * NG + idx;
    const double zi = b / mass[diag_idx[v]];
    x[o] = 0.0; r[o] = b; z[o] = zi; pvec[o] = zi;
}

// q = M p for the 27 columns
__global__ void __launch_bounds__(256)
gradproj_spmv_kernel(int n_vert, const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                     const double* __restrict__ mass, const double* __restrict__ pvec, double* __restrict__ q) {
    const int prob = blockIdx.y;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)n_vert * NG) return;
    const int v = (int)(idx / NG), k = (int)(idx % NG);
    const double* pp = pvec + (long)prob * n_vert * NG;
    double s = 0.0;
    for (int e = row_ptr[v]; e < row_ptr[v + 1]; ++e) s += mass[e] * pp[(long)col_idx[e] * NG + k];
    q[(long)prob * n_vert * NG + idx] = s;
}

// dots[prob][k] = sum_v a[v][k] b[v][k]          grid = (NG, batch)
__global__ void __launch_bounds__(256)
gradproj_dot_kernel(int n_vert, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ dots) {
    __shared__ double sh[8];
    const int k = blockIdx.x, prob = blockIdx.y;
    const double* ap = a + (long)prob * n_vert * NG;
    const double* bp = b + (long)prob * n_vert * NG;
    double s = 0.0;
    for (int v = threadIdx.x; v < n_vert; v += blockDim.x) s += ap[(long)v * NG + k] * bp[(long)v * NG + k];
    s = block_sum(s, sh);
    if (threadIdx.x == 0) dots[(long)prob * NG + k] = s;
}

// x += alpha p, r -= alpha q, z = r / M_vv   with alpha = rz / pq per column
__global__ void __launch_bounds__(256)
gradproj_update_kernel(int n_vert, const double* __restrict__ mass, const int* __restrict__ diag_idx,
                       const double* __restrict__ rz, const double* __restrict__ pq, const double* __restrict__ pvec,
                       const double* __restrict__ q, double* __restrict__ x, double* __restrict__ r, double* __restrict__ z) {
    const int prob = blockIdx.y;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)n_vert * NG) return;
    const int v = (int)(idx / NG), k = (int)(idx % NG);
    const double den = pq[(long)prob * NG + k];
    const double alpha = (den > 0.0) ? rz[(long)prob * NG + k] / den : 0.0;
    const long o = (long)prob * n_vert * NG + idx;
    x[o] += alpha * pvec[o];
    const double rn = r[o] - alpha * q[o];
    r[o] = rn;
    z[o] = rn / mass[diag_idx[v]];
}

// p = z + beta p with beta = rz_new / rz_old per column
__global__ void __launch_bounds__(256)
gradproj_dir_kernel(int n_vert, const double* __restrict__ rz_new, const double* __restrict__ rz_old,
                    const double* __restrict__ z, double* __restrict__ pvec) {
    const int prob = blockIdx.y;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)n_vert * NG) return;
    const int k = (int)(idx % NG);
    const double den = rz_old[(long)prob * NG + k];
    const double beta = (den > 0.0) ? rz_new[(long)prob * NG + k] / den : 0.0;
    const long o = (long)prob * n_vert * NG + idx;
    pvec[o] = z[o] + beta * pvec[o];
}

// z = D^{-1} r on the first n_rows block rows (Dirichlet rows are identity rows of J, so their D^{-1} is I)
__global__ void __launch_bounds__(256)
bjacobi_apply_kernel(int n_rows, const double* __restrict__ Dinv, const double* __restrict__ r, double* __restrict__ z) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)n_rows * NC) return;
    const int v = (int)(idx / NC), i = (int)(idx % NC);
    const double* d = Dinv + (long)v * 81 + i * 9;
    const double* rv = r + (long)v * NC;
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < NC; ++j) s += d[j] * rv[j];
    z[idx] = s;
}

// ---------------------------------------------------------------------------------------
// Coarse-space set-up, deterministic (ordered partial sums, no atomics) with a pivoted inverse.
// Host side (gmpnp_create_3d) sorts the BSR blocks by (slab of row, slab of column) pair and cuts every
// pair's list into chunks of <= CP_CHUNK blocks; pass 1 sums the non-Dirichlet entries of a chunk in list
// order, pass 2 adds the chunk sums of a pair in chunk order, inverts A_c = P^T J P by Gauss-Jordan with
// partial pivoting and stores the TRANSPOSED inverse (coalesced reads in the solve).
// ---------------------------------------------------------------------------------------
constexpr int CP_CHUNK = 128;

__global__ void __launch_bounds__(96)
coarse_partial_kernel(int n_blocks, const int* __restrict__ chunk_ptr, const int* __restrict__ cp_blk,
                      const int* __restrict__ blk_row, const int* __restrict__ col_idx,
                      const int* __restrict__ dir_flag, const double* __restrict__ J, double* __restrict__ part,
                      int n_chunk) {
    const int chunk = blockIdx.x, prob = blockIdx.y, e = threadIdx.x;
    if (e >= 81) return;
    const int i = e / 9, j = e % 9;
    const double* Jp = J + (long)prob * n_blocks * 81;
    double s = 0.0;
    for (int c = chunk_ptr[chunk]; c < chunk_ptr[chunk + 1]; ++c) {
        const int b = cp_blk[c];
        if (dir_flag[(long)blk_row[b] * NC + i] >= 0 || dir_flag[(long)col_idx[b] * NC + j] >= 0) continue;
        s += Jp[(long)b * 81 + e];
    }
    part[((long)prob * n_chunk + chunk) * 81 + e] = s;
}

// In-place Gauss-Jordan inversion of the NCO x NCO matrix Ac (shared memory) with partial pivoting, by one CTA.  A
// vanishing pivot (singular coarse operator) is replaced by 1 and counted in *nbad.  perm / pivinv / prow: shared scratch.
__device__ void gj_invert_pivoted(double* Ac, int* perm, double* s_pivinv, int* s_prow, int* s_nbad) {
    const int tid = threadIdx.x;
    // in-place Gauss-Jordan inversion with partial pivoting (row swaps recorded in perm)
    for (int c = 0; c < NCO; ++c) {
        if (tid < 32) {
            double best = -1.0; int bi = c;
            for (int r = c + tid; r < NCO; r += 32) {
                const double a = fabs(Ac[r * NCO + c]);
                if (a > best) { best = a; bi = r; }
            }
            for (int o = 16; o >= 1; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (tid == 0) { *s_prow = bi; perm[c] = bi; }
        }
        __syncthreads();
        const int p = *s_prow;
        if (p != c) {
            for (int j = tid; j < NCO; j += blockDim.x) {
                const double t = Ac[c * NCO + j]; Ac[c * NCO + j] = Ac[p * NCO + j]; Ac[p * NCO + j] = t;
            }
        }
        __syncthreads();
        if (tid == 0) {
            double pv = Ac[c * NCO + c];
            if (!(fabs(pv) > 1e-300) || !isfinite(pv)) { pv = 1.0; (*s_nbad)++; }
            *s_pivinv = 1.0 / pv;
        }
        __syncthreads();
        const double pi = *s_pivinv;
        for (int j = tid; j < NCO; j += blockDim.x)
            if (j != c) Ac[c * NCO + j] *= pi;
        __syncthreads();
        for (int idx = tid; idx < NCO * NCO; idx += blockDim.x) {
            const int r = idx / NCO, j = idx % NCO;
            if (r == c || j == c) continue;
            Ac[idx] -= Ac[r * NCO + c] * Ac[c * NCO + j];
        }
        __syncthreads();
        for (int r = tid; r < NCO; r += blockDim.x)
            Ac[r * NCO + c] = (r == c) ? pi : -Ac[r * NCO + c] * pi;
        __syncthreads();
    }
    // (P A)^-1 = A^-1 P^T: undo the row swaps as column swaps, last swap first
    for (int c = NCO - 1; c >= 0; --c) {
        const int p = perm[c];
        if (p != c) {
            for (int r = tid; r < NCO; r += blockDim.x) {
                const double t = Ac[r * NCO + c]; Ac[r * NCO + c] = Ac[r * NCO + p]; Ac[r * NCO + p] = t;
            }
        }
        __syncthreads();
    }
}

// one CTA per problem; Ac in shared memory [NCO][NCO]; flag[prob] = number of vanishing pivots replaced by 1
__global__ void __launch_bounds__(1024)
coarse_invert_pivoted_kernel(const int* __restrict__ chunk_pair, const double* __restrict__ part, int n_chunk,
                             double* __restrict__ AciT, int* __restrict__ flag) {
    extern __shared__ double Ac[];          // [NCO][NCO]
    __shared__ int perm[NCO];
    __shared__ double pivinv;
    __shared__ int prow, nbad;
    const int prob = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < NCO * NCO; i += blockDim.x) Ac[i] = 0.0;
    if (tid == 0) nbad = 0;
    __syncthreads();
    // chunks are sorted by pair, so the chunks of a pair are consecutive: 12 groups of 81 threads walk the chunk
    // list; group g owns the pairs with (pair % 12 == g), and adds their chunk sums in chunk order
    const int grp = tid / 81, e = tid % 81;
    if (grp < 12) {
        const int i = e / 9, j = e % 9;
        for (int c = 0; c < n_chunk; ++c) {
            const int pr = chunk_pair[c];
            if (pr % 12 != grp) continue;
            const int I = pr / NZ, Jc = pr % NZ;
            Ac[(I * NC + i) * NCO + Jc * NC + j] += part[((long)prob * n_chunk + c) * 81 + e];
        }
    }
    __syncthreads();
    // empty coarse rows (all members Dirichlet) -> identity
    for (int i = tid; i < NCO; i += blockDim.x) {
        bool empty = true;
        for (int j = 0; j < NCO; ++j) if (Ac[i * NCO + j] != 0.0) { empty = false; break; }
        if (empty) Ac[i * NCO + i] = 1.0;
    }
    __syncthreads();
    gj_invert_pivoted(Ac, perm, &pivinv, &prow, &nbad);
    for (int idx = tid; idx < NCO * NCO; idx += blockDim.x) {
        const int r = idx / NCO, j = idx % NCO;
        AciT[(long)prob * NCO * NCO + (long)j * NCO + r] = Ac[idx];      // transposed
    }
    if (tid == 0 && flag) flag[prob] = nbad;
}

// ---------------------------------------------------------------------------------------
// Persistent GMRES: one thread-block CLUSTER per problem runs the whole right-preconditioned restarted
// GMRES(m) solve  J dx = b  (block-Jacobi + z-slab coarse correction, CGS2) in ONE launch, with no host
// synchronisation: the G CTAs of a cluster own contiguous chunks of block rows / vector entries, exchange their
// partial dot products through distributed shared memory (fixed rank order: deterministic, every CTA gets the
// bitwise-identical sum, so the control flow stays cluster-uniform) and meet at cluster barriers.  The small
// Hessenberg / Givens recurrences are done redundantly by thread 0 of every CTA.
// ---------------------------------------------------------------------------------------
#ifndef GMPNP_GM_THREADS
#define GMPNP_GM_THREADS 512
#endif
constexpr int GM_THREADS = GMPNP_GM_THREADS;
constexpr int GM_WARPS = GM_THREADS / 32;
constexpr int GM_KT = 8;                 // dot products per pass over the chunk

struct GmresArgs {
    int n_vert, n_blocks, m, maxit;
    double rtol;
    const int *row_ptr, *col_idx, *agg, *agg_ptr, *agg_nodes, *dir_flag;
    const double *J, *Dinv, *AciT, *b;
    double *x, *V, *w, *z;
    const int* enabled;                  // [batch] or nullptr: problems with 0 are skipped
    int* its;                            // [batch] GMRES iterations done
    double* relres;                      // [batch] final true relative residual ||b - J x|| / ||b||
};

struct GmresSmem {
    double* red;      // [2][RED_N]  partial sums of this CTA, double-buffered (read by the cluster through DSMEM)
    double* sums;     // [RED_N]     cluster-wide sums
    double* wpart;    // [GM_WARPS][GM_KT]
    double* spart;    // [GM_WARPS][96] SpMV segment sums
    double* rc;       // [3][NCO] restriction slices, then rc in slice 0
    double* yc;       // [3][NCO]
    double* H;        // [(m+1)*m] column-major columns of length m+1
    double* cs; double* sn; double* g; double* y;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sums[k] (k < K) = sum over the cluster's CTAs (rank order) of red[buf][k]; red[buf][0..K) must be written by
// this CTA (by any thread) before the call.  Contains two __syncthreads and one cluster barrier.
__device__ __forceinline__ void cluster_reduce(cg::cluster_group& cl, const GmresSmem& S, int& buf, int K, int red_n) {
    cl.sync();                                        // partials of every CTA complete (also a CTA barrier)
    const int G = cl.num_blocks();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        double s = 0.0;
        for (int q = 0; q < G; ++q) {
            const double* remote = cl.map_shared_rank(S.red + buf * red_n, q);
            s += remote[k];
        }
        S.sums[k] = s;
    }
    buf ^= 1;
    __syncthreads();
}

// partial dot products of w with V_k (k0 <= k < k0+kn, kn <= GM_KT) over [i0, i1) -> red[buf][k]
__device__ __forceinline__ void chunk_dots(const GmresSmem& S, int buf, int red_n, const double* __restrict__ Vp, long n,
                                           const double* __restrict__ w, long i0, long i1, int k0, int kn) {
    double acc[GM_KT];
#pragma unroll
    for (int q = 0; q < GM_KT; ++q) acc[q] = 0.0;
    for (long i = i0 + threadIdx.x; i < i1; i += GM_THREADS) {
        const double wi = w[i];
#pragma unroll
        for (int q = 0; q < GM_KT; ++q)
            if (q < kn) acc[q] += Vp[(long)(k0 + q) * n + i] * wi;
    }
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < GM_KT; ++q) {
        const double v = warp_sum(acc[q]);
        if (lane == 0) S.wpart[wp * GM_KT + q] = v;
    }
    __syncthreads();
    if (threadIdx.x < kn) {
        double s = 0.0;
        for (int ww = 0; ww < GM_WARPS; ++ww) s += S.wpart[ww * GM_KT + threadIdx.x];
        S.red[buf * red_n + k0 + threadIdx.x] = s;
    }
    __syncthreads();
}

// yc = A_c^-1 P^T r for the FULL vector r (every CTA redundantly; r must be visible cluster-wide)
__device__ __forceinline__ void coarse_correction(const GmresSmem& S, const GmresArgs& a, const double* __restrict__ AciT,
                                                  const double* __restrict__ r) {
    const int tid = threadIdx.x;
    if (tid < 3 * NCO) {
        const int sl = tid / NCO, t = tid % NCO, I = t / NC, i = t % NC;
        double s = 0.0;
        for (int k = a.agg_ptr[I] + sl; k < a.agg_ptr[I + 1]; k += 3) {
            const int v = a.agg_nodes[k];
            if (a.dir_flag[(long)v * NC + i] < 0) s += r[(long)v * NC + i];
        }
        S.rc[sl * NCO + t] = s;
    }
    __syncthreads();
    if (tid < NCO) S.rc[tid] = S.rc[tid] + S.rc[NCO + tid] + S.rc[2 * NCO + tid];
    __syncthreads();
    if (tid < 3 * NCO) {
        const int sl = tid / NCO, t = tid % NCO;
        double s = 0.0;
        for (int j = sl * (NCO / 3); j < (sl + 1) * (NCO / 3); ++j) s += AciT[(long)j * NCO + t] * S.rc[j];
        S.yc[sl * NCO + t] = s;
    }
    __syncthreads();
    if (tid < NCO) S.yc[tid] = S.yc[tid] + S.yc[NCO + tid] + S.yc[2 * NCO + tid];
    __syncthreads();
}

// z[i0,i1) = D^-1 r + P yc on this CTA's chunk
__device__ __forceinline__ void precond_chunk(const GmresSmem& S, const GmresArgs& a, const double* __restrict__ Dinv,
                                              const double* __restrict__ r, double* __restrict__ z, long i0, long i1) {
    for (long idx = i0 + threadIdx.x; idx < i1; idx += GM_THREADS) {
        const int v = (int)(idx / NC), i = (int)(idx % NC);
        const double* D = Dinv + (long)v * 81 + i * 9;
        const double* rp = r + (long)v * NC;
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 9; ++j) s += D[j] * rp[j];
        if (a.dir_flag[idx] < 0) s += S.yc[a.agg[v] * NC + i];
        z[idx] = s;
    }
}

// y[v0..v1) = J x (own block rows; x must be visible cluster-wide); mode 0: y = Jx, mode 1: y = b - Jx
__device__ __forceinline__ void spmv_chunk(const GmresSmem& S, const GmresArgs& a, const double* __restrict__ Jp,
                                           const double* __restrict__ xp, double* __restrict__ y,
                                           const double* __restrict__ bp, int v0, int v1) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int j0 = lane % 9, j1 = (lane + 32) % 9, j2 = (lane + 64) % 9;
    const bool third = lane + 64 < 81;
    double* part = S.spart + w * 96;
    for (int row = v0 + w; row < v1; row += GM_WARPS) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        const int s0 = a.row_ptr[row], s1 = a.row_ptr[row + 1];
        for (int base = s0; base < s1; base += 32) {
            const int cnt = min(32, s1 - base);
            const int mycol = (lane < cnt) ? a.col_idx[base + lane] : 0;
            const double* blk = Jp + (long)base * 81;
            int t = 0;
            for (; t + 4 <= cnt; t += 4, blk += 4 * 81) {
                double v[4][3];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    v[q][0] = __ldcs(blk + q * 81 + lane);
                    v[q][1] = __ldcs(blk + q * 81 + lane + 32);
                    v[q][2] = third ? __ldcs(blk + q * 81 + lane + 64) : 0.0;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double* xc = xp + (long)__shfl_sync(0xffffffffu, mycol, t + q) * NC;
                    a0 += v[q][0] * xc[j0];
                    a1 += v[q][1] * xc[j1];
                    a2 += v[q][2] * xc[j2];
                }
            }
            for (; t < cnt; ++t, blk += 81) {
                const double v0_ = __ldcs(blk + lane), v1_ = __ldcs(blk + lane + 32);
                const double v2_ = third ? __ldcs(blk + lane + 64) : 0.0;
                const double* xc = xp + (long)__shfl_sync(0xffffffffu, mycol, t) * NC;
                a0 += v0_ * xc[j0];
                a1 += v1_ * xc[j1];
                a2 += v2_ * xc[j2];
            }
        }
        part[lane] = a0; part[lane + 32] = a1; part[lane + 64] = a2;
        __syncwarp();
        if (lane < NC) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < 9; ++j) s += part[lane * 9 + j];
            const long o = (long)row * NC + lane;
            y[o] = bp ? bp[o] - s : s;
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(GM_THREADS, 1)
gmres_cluster_kernel(GmresArgs a) {
    cg::cluster_group cl = cg::this_cluster();
    const int G = (int)cl.num_blocks(), rank = (int)cl.block_rank();
    const int prob = blockIdx.x / G;
    if (a.enabled && !a.enabled[prob]) return;          // cluster-uniform
    const int tid = threadIdx.x;
    const int V = a.n_vert, m = a.m, ld = m + 1;
    const long n = (long)V * NC;
    const int v0 = (int)((long)V * rank / G), v1 = (int)((long)V * (rank + 1) / G);
    const long i0 = (long)v0 * NC, i1 = (long)v1 * NC;
    extern __shared__ double smem[];
    const int red_n = ld + 3;
    GmresSmem S;
    {
        double* p = smem;
        S.red = p; p += 2 * red_n;
        S.sums = p; p += red_n;
        S.wpart = p; p += GM_WARPS * GM_KT;
        S.spart = p; p += GM_WARPS * 96;
        S.rc = p; p += 3 * NCO;
        S.yc = p; p += 3 * NCO;
        S.H = p; p += (long)ld * m;
        S.cs = p; p += m; S.sn = p; p += m; S.g = p; p += ld; S.y = p; p += ld;
    }
    __shared__ int s_flag;
    const double* Jp = a.J + (long)prob * a.n_blocks * 81;
    const double* Dinv = a.Dinv + (long)prob * V * 81;
    const double* AciT = a.AciT + (long)prob * NCO * NCO;
    const double* b = a.b + (long)prob * n;
    double* x = a.x + (long)prob * n;
    double* Vp = a.V + (long)prob * ld * n;
    double* w = a.w + (long)prob * n;
    double* z = a.z + (long)prob * n;
    int buf = 0;

    auto chunk_norm2 = [&](const double* vec) {          // -> red[buf][0]
        double s = 0.0;
        for (long i = i0 + tid; i < i1; i += GM_THREADS) { const double t = vec[i]; s += t * t; }
        s = warp_sum(s);
        if ((tid & 31) == 0) S.wpart[(tid >> 5) * GM_KT] = s;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int ww = 0; ww < GM_WARPS; ++ww) t += S.wpart[ww * GM_KT];
            S.red[buf * red_n] = t;
        }
        __syncthreads();
    };

    for (long i = i0 + tid; i < i1; i += GM_THREADS) x[i] = 0.0;
    chunk_norm2(b);
    cluster_reduce(cl, S, buf, 1, red_n);
    const double beta0 = sqrt(S.sums[0]);
    const double tol = a.rtol * beta0;
    const double* rcur = b;
    double beta = beta0;
    int total = 0;
    bool fin = isfinite(beta0) && beta0 > 0.0;
    while (fin && beta > tol && total < a.maxit) {
        // ---- cycle start: V_0 = r / beta, g = (beta, 0, ...) --------------------------------------
        const double ib = 1.0 / beta;
        for (long i = i0 + tid; i < i1; i += GM_THREADS) Vp[i] = rcur[i] * ib;
        if (tid == 0) { for (int k = 0; k <= m; ++k) S.g[k] = 0.0; S.g[0] = beta; }
        cl.sync();
        const int mcyc = min(m, a.maxit - total);
        int jd = mcyc;
        for (int j = 0; j < mcyc; ++j) {
            const double* vj = Vp + (long)j * n;
            // w = J M^-1 v_j
            coarse_correction(S, a, AciT, vj);
            precond_chunk(S, a, Dinv, vj, z, i0, i1);
            cl.sync();                                   // z complete
            spmv_chunk(S, a, Jp, z, w, nullptr, v0, v1);
            __syncthreads();
            // CGS2: two classical Gram-Schmidt passes; h = d1 + d2
            for (int pass = 0; pass < 2; ++pass) {
                for (int k0 = 0; k0 <= j; k0 += GM_KT) chunk_dots(S, buf, red_n, Vp, n, w, i0, i1, k0, min(GM_KT, j + 1 - k0));
                cluster_reduce(cl, S, buf, j + 1, red_n);
                double nrm2 = 0.0;
                for (long i = i0 + tid; i < i1; i += GM_THREADS) {
                    // eight independent basis loads in flight per thread (the loop over k has a run-time bound, so
                    // without the explicit batch every load would wait for the previous FMA: latency-bound)
                    double s = w[i];
                    int k = 0;
                    for (; k + GM_KT <= j + 1; k += GM_KT) {
                        double v[GM_KT];
#pragma unroll
                        for (int q = 0; q < GM_KT; ++q) v[q] = Vp[(long)(k + q) * n + i];
#pragma unroll
                        for (int q = 0; q < GM_KT; ++q) s -= S.sums[k + q] * v[q];
                    }
                    for (; k <= j; ++k) s -= S.sums[k] * Vp[(long)k * n + i];
                    w[i] = s;
                    nrm2 += s * s;
                }
                // Hessenberg column accumulates both passes (column j of H used as scratch until the Givens step)
                if (tid == 0) {
                    double* h = S.H + (long)j * ld;
                    for (int k = 0; k <= j; ++k) h[k] = (pass == 0 ? 0.0 : h[k]) + S.sums[k];
                }
                if (pass == 1) {
                    nrm2 = warp_sum(nrm2);
                    if ((tid & 31) == 0) S.wpart[(tid >> 5) * GM_KT] = nrm2;
                    __syncthreads();
                    if (tid == 0) {
                        double t = 0.0;
                        for (int ww = 0; ww < GM_WARPS; ++ww) t += S.wpart[ww * GM_KT];
                        S.red[buf * red_n] = t;
                    }
                }
                __syncthreads();
            }
            cluster_reduce(cl, S, buf, 1, red_n);
            const double nrm = sqrt(S.sums[0]);
            const double inrm = (nrm > 0.0) ? 1.0 / nrm : 0.0;
            for (long i = i0 + tid; i < i1; i += GM_THREADS) Vp[(long)(j + 1) * n + i] = w[i] * inrm;
            if (tid == 0) {
                double* h = S.H + (long)j * ld;
                h[j + 1] = nrm;
                for (int k = 0; k < j; ++k) {
                    const double t = S.cs[k] * h[k] + S.sn[k] * h[k + 1];
                    h[k + 1] = -S.sn[k] * h[k] + S.cs[k] * h[k + 1];
                    h[k] = t;
                }
                const double aa = h[j], bb = h[j + 1];
                const double d = hypot(aa, bb);
                const double cj = (d > 0.0) ? aa / d : 1.0, sj = (d > 0.0) ? bb / d : 0.0;
                S.cs[j] = cj; S.sn[j] = sj;
                h[j] = d; h[j + 1] = 0.0;
                S.g[j + 1] = -sj * S.g[j];
                S.g[j] = cj * S.g[j];
                s_flag = (fabs(S.g[j + 1]) <= tol || !(nrm > 0.0) || !isfinite(nrm)) ? 1 : 0;
            }
            cl.sync();                                   // V_{j+1} visible cluster-wide; s_flag visible in the CTA
            if (s_flag) { jd = j + 1; break; }
        }
        total += jd;
        // ---- x += M^-1 (V y),  y = H^-1 g -----------------------------------------------------------
        if (tid == 0) {
            for (int k = jd - 1; k >= 0; --k) {
                double s = S.g[k];
                for (int l = k + 1; l < jd; ++l) s -= S.H[(long)l * ld + k] * S.y[l];
                const double d = S.H[(long)k * ld + k];
                S.y[k] = (d != 0.0) ? s / d : 0.0;
            }
        }
        __syncthreads();
        for (long i = i0 + tid; i < i1; i += GM_THREADS) {
            double s = 0.0;
            int k = 0;
            for (; k + GM_KT <= jd; k += GM_KT) {
                double v[GM_KT];
#pragma unroll
                for (int q = 0; q < GM_KT; ++q) v[q] = Vp[(long)(k + q) * n + i];
#pragma unroll
                for (int q = 0; q < GM_KT; ++q) s += S.y[k + q] * v[q];
            }
            for (; k < jd; ++k) s += S.y[k] * Vp[(long)k * n + i];
            w[i] = s;
        }
        cl.sync();                                       // t = V y complete (in w)
        coarse_correction(S, a, AciT, w);
        precond_chunk(S, a, Dinv, w, z, i0, i1);
        for (long i = i0 + tid; i < i1; i += GM_THREADS) x[i] += z[i];
        cl.sync();                                       // x complete; nobody reads w (= t) any more
        // true residual r = b - J x -> w
        spmv_chunk(S, a, Jp, x, w, b, v0, v1);
        __syncthreads();
        chunk_norm2(w);
        cluster_reduce(cl, S, buf, 1, red_n);
        beta = sqrt(S.sums[0]);
        rcur = w;
        fin = isfinite(beta);
    }
    if (rank == 0 && tid == 0) {
        a.its[prob] = total;
        a.relres[prob] = (beta0 > 0.0) ? beta / beta0 : (isfinite(beta0) ? 0.0 : NAN);
    }
    cl.sync();                                           // no CTA exits while its shared memory may still be read
}

// ---------------------------------------------------------------------------------------
// Device-resident Newton / march control (per-problem state, no host arithmetic)
// ---------------------------------------------------------------------------------------
struct NewtonCtl {      // per-problem arrays, device
    int* active;        // 1: still iterating in this solve
    int* status;        // -1 while running, then GMPNP_* code
    int* iters;         // Newton iterations of this solve
    int* lin_total;     // GMRES iterations of this solve
    double* r0;         // ||F(u0)||
    double* r;          // ||F(u_k)||
};

// after the first residual evaluation of a solve
__global__ void newton_begin_kernel(int B, const int* __restrict__ enabled, const double* __restrict__ nrm, NewtonCtl c,
                                    int criterion, double atol, int* __restrict__ n_active) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    c.iters[p] = 0; c.lin_total[p] = 0;
    const double r0 = nrm[p];
    c.r0[p] = r0; c.r[p] = r0;
    int act = 1, st = -1;
    if (enabled && !enabled[p]) { act = 0; st = GMPNP_CONVERGED; c.r0[p] = 0.0; c.r[p] = 0.0; }
    else if (!isfinite(r0)) { act = 0; st = GMPNP_NOT_FINITE; }
    else if (criterion == 0 && r0 < atol) { act = 0; st = GMPNP_CONVERGED; }
    c.active[p] = act; c.status[p] = st;
    if (act) atomicAdd(n_active, 1);
}

// after the update and the residual evaluation of Newton iteration k (0-based)
__global__ void newton_step_kernel(int B, int k, int maxit, const double* __restrict__ nrm, const double* __restrict__ dxmax,
                                   const double* __restrict__ umax, const int* __restrict__ lin_its,
                                   const double* __restrict__ relres, NewtonCtl c, int criterion, double rtol,
                                   double atol, double xtol, double lin_rtol, int* __restrict__ n_active) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    if (!c.active[p]) return;
    c.iters[p] = k + 1;
    c.lin_total[p] += lin_its[p];
    const double rn = nrm[p];
    c.r[p] = rn;
    const bool linfail = !(relres[p] <= 1e3 * lin_rtol);
    int act = 1, st = -1;
    if (!isfinite(rn) || !isfinite(dxmax[p])) { act = 0; st = GMPNP_NOT_FINITE; }
    else {
        const bool conv = (criterion == 0) ? ((rn / c.r0[p] < rtol) || (rn < atol))
                                           : (dxmax[p] <= xtol * fmax(1.0, umax[p]));
        if (conv) { act = 0; st = GMPNP_CONVERGED; }
        else if (linfail) { act = 0; st = GMPNP_LINEAR_FAILED; }
        else if (k + 1 >= maxit) { act = 0; st = GMPNP_MAXIT; }
    }
    c.active[p] = act; c.status[p] = st;
    if (act) atomicAdd(n_active, 1);
}

// Dirichlet values from the kind table (3D:460-467, 835-838): kind 0 -> 0, 1 -> wall potential V * ramp,
// 2 -> CO2 entry value (per problem, updated by the Sechenov feedback), 3/4 -> CO / H2 entry values
__global__ void dirichlet_fill_kernel(int B, int n_dir, const signed char* __restrict__ kind, const double* __restrict__ tab,
                                      const double* __restrict__ co2, double ramp, double* __restrict__ vals,
                                      double* __restrict__ params) {
    const int p = blockIdx.y;
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    const double Vw = tab[p * 4 + 0] * ramp;
    if (d == 0 && params) params[(long)p * GMPNP_NPAR + GMPNP_P_V] = Vw;
    if (d >= n_dir) return;
    const int k = kind[d];
    double v = 0.0;
    if (k == 1) v = Vw;
    else if (k == 2) v = co2[p];
    else if (k == 3) v = tab[p * 4 + 2];
    else if (k == 4) v = tab[p * 4 + 3];
    vals[(long)p * n_dir + d] = v;
}

// medians of up to four components per problem in one launch (grid = (batch, ncomp)); bitonic sort in shared memory
__global__ void __launch_bounds__(1024)
median4_kernel(int n_vert, int npow2, int c0, int c1, int c2, int c3, const double* __restrict__ u, double* __restrict__ med) {
    extern __shared__ double sv[];
    const int prob = blockIdx.x;
    const int comps[4] = {c0, c1, c2, c3};
    const int comp = comps[blockIdx.y];
    for (int i = threadIdx.x; i < npow2; i += blockDim.x)
        sv[i] = (i < n_vert) ? u[((long)prob * n_vert + i) * NC + comp] : INFINITY;
    __syncthreads();
    for (int k = 2; k <= npow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const bool up = ((i & k) == 0);
                    const double a = sv[i], b = sv[ixj];
                    if ((a > b) == up) { sv[i] = b; sv[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0)
        med[prob * 4 + blockIdx.y] = (n_vert & 1) ? sv[n_vert / 2] : 0.5 * (sv[n_vert / 2 - 1] + sv[n_vert / 2]);
}

// The same medians for meshes that do not fit the shared-memory sort (> 16384 vertices): exact radix SELECT of the two
// middle order statistics straight from global memory -- eight passes of eight bits over order-preserving 64-bit keys,
// a 256-bin histogram of the elements that match the prefix found so far, integer atomics only (deterministic, and
// the result is an element of the input, so it equals np.median bit for bit).  grid = (batch, ncomp).
__device__ __forceinline__ unsigned long long median_key(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void __launch_bounds__(1024)
median_select_kernel(int n_vert, int ncomp_out, int c0, int c1, int c2, int c3, const double* __restrict__ u,
                     double* __restrict__ med) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned int s_rank;
    const int prob = blockIdx.x;
    const int comps[4] = {c0, c1, c2, c3};
    const double* up = u + (long)prob * n_vert * NC + comps[blockIdx.y];
    double picked[2];
    for (int which = 0; which < 2; ++which) {
        if (threadIdx.x == 0) { s_prefix = 0ull; s_rank = (unsigned)(which == 0 ? (n_vert - 1) / 2 : n_vert / 2); }
        unsigned long long mask = 0ull;
        for (int pass = 7; pass >= 0; --pass) {
            for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0u;
            __syncthreads();
            const unsigned long long prefix = s_prefix;
            for (int i = threadIdx.x; i < n_vert; i += blockDim.x) {
                const unsigned long long k = median_key(up[(long)i * NC]);
                if ((k & mask) == prefix) atomicAdd(&hist[(unsigned)(k >> (8 * pass)) & 255u], 1u);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned int r = s_rank, cum = 0u;
                int bin = 0;
                for (; bin < 255; ++bin) {
                    if (cum + hist[bin] > r) break;
                    cum += hist[bin];
                }
                s_rank = r - cum;
                s_prefix = prefix | ((unsigned long long)bin << (8 * pass));
            }
            mask |= 0xffull << (8 * pass);
            __syncthreads();
        }
        const unsigned long long k = s_prefix;
        const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
        picked[which] = __longlong_as_double((long long)b);
        __syncthreads();
    }
    if (threadIdx.x == 0) med[prob * ncomp_out + blockIdx.y] = (n_vert & 1) ? picked[0] : 0.5 * (picked[0] + picked[1]);
}

// Sechenov CO2 entry value from the nodal medians (3D:817-838; CO2_conc 3D:70-93).  sech[p][8] =
// {A = fugacity * K_H * 1000 / c0_CO2, (h_OH + h_CO2) c0_OH / 1000, (h_HCO3 + h_CO2) c0_HCO3 / 1000,
//  (h_CO32 + h_CO2) c0_CO32 / 1000, (h_cat + h_CO2) c0_cat / 1000, mode, c0_H, -}.
// mode 0 (GMPNP): med = (OH, HCO3, CO32, cat).  mode 1 (rxn-diff, RD3:575-601): med = (H, OH, HCO3, CO32) and the
// cation is the electroneutral estimate c_cat = c_HCO3 + 2 c_CO32 + c_OH - c_H.
__global__ void sechenov_kernel(int B, const int* __restrict__ alive, const double* __restrict__ med,
                                const double* __restrict__ sech, double* __restrict__ co2) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    if (alive && !alive[p]) return;
    const double* s = sech + (long)p * 8;
    const double* m = med + (long)p * 4;
    double e;
    if (s[5] == 0.0) {
        e = s[1] * m[0];
        e += s[2] * m[1];
        e += s[3] * m[2];
        e += s[4] * m[3];
    } else {
        // coefficients are per scaled unit: s[1..3] for OH, HCO3, CO32; cation from electroneutrality in mol/m3:
        // s[4] holds (h_cat + h_CO2) / 1000 and s[6], s[7].. are unused; the c0 factors are folded by the host into
        // s[1..3] (own term + cation share) and s[6] (H share, negative)
        e = s[1] * m[1];
        e += s[2] * m[2];
        e += s[3] * m[3];
        e += s[6] * m[0];
    }
    co2[p] = s[0] * pow(10.0, -e);
}

// per problem: inc = max|u - un| / max(1, max|u|); optional un <- u and history row
__global__ void __launch_bounds__(1024)
march_advance_kernel(long n, const int* __restrict__ alive, const double* __restrict__ u, double* __restrict__ un,
                     double* __restrict__ hist, long hist_stride, double* __restrict__ inc) {
    __shared__ double sh1[32], sh2[32];
    const int prob = blockIdx.x;
    if (alive && !alive[prob]) { if (threadIdx.x == 0 && inc) inc[prob] = 0.0; return; }
    double m1 = 0.0, m2 = 0.0;
    for (long i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = u[(long)prob * n + i];
        m1 = fmax(m1, fabs(v - un[(long)prob * n + i]));
        m2 = fmax(m2, fabs(v));
        un[(long)prob * n + i] = v;
        if (hist) hist[(long)prob * hist_stride + i] = v;
    }
    for (int o = 16; o >= 1; o >>= 1) {
        m1 = fmax(m1, __shfl_xor_sync(0xffffffffu, m1, o));
        m2 = fmax(m2, __shfl_xor_sync(0xffffffffu, m2, o));
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { sh1[w] = m1; sh2[w] = m2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { m1 = fmax(m1, sh1[k]); m2 = fmax(m2, sh2[k]); }
        if (inc) inc[prob] = m1 / fmax(1.0, m2);
    }
}

// steady-state test of the march: a problem whose last relative increment is <= tol stops marching (converged)
__global__ void steady_check_kernel(int B, double tol, int* __restrict__ alive, const double* __restrict__ inc,
                                    int* __restrict__ conv, int* __restrict__ n_alive) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    if (!alive[p]) return;
    if (inc[p] <= tol) { alive[p] = 0; conv[p] = 1; }
    else atomicAdd(n_alive, 1);
}

// march bookkeeping after a Newton solve of step `step`: per-problem iteration counts, alive flags, final status
__global__ void march_record_kernel(int B, int step, int n_steps, NewtonCtl c, int* __restrict__ alive,
                                    int* __restrict__ iters_out, int* __restrict__ lin_out, int* __restrict__ status_out,
                                    int* __restrict__ steps_done, const double* __restrict__ co2, double* __restrict__ co2_out,
                                    int* __restrict__ n_alive) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    if (!alive[p]) return;
    if (iters_out) iters_out[(long)p * n_steps + step] = c.iters[p];
    if (lin_out) lin_out[(long)p * n_steps + step] = c.lin_total[p];
    if (co2_out) co2_out[(long)p * n_steps + step] = co2[p];
    if (c.status[p] != GMPNP_CONVERGED) { alive[p] = 0; status_out[p] = c.status[p]; }
    else { steps_done[p] = step + 1; atomicAdd(n_alive, 1); }
}

}  // namespace pore3d

// =======================================================================================
// host side
// =======================================================================================
using namespace pore3d;

template <class T>
static int dev_upload(gmpnp_handle* h, T** dptr, const std::vector<T>& v) {
    GMPNP_CUDA_TRY(h, cudaMalloc((void**)dptr, sizeof(T) * std::max<size_t>(1, v.size())));
    if (!v.empty()) GMPNP_CUDA_TRY(h, cudaMemcpy(*dptr, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
    return GMPNP_OK;
}

struct Host3D {   // device arrays that only the 3D path needs and common.cuh does not name
    int* d_blk_row = nullptr;
    double2* d_blk_geo = nullptr;   // per gather-list entry (tet, a, b): (grad lambda_a . grad lambda_b, volume)
    int* d_agg = nullptr;
    int* d_agg_ptr = nullptr;
    int* d_agg_nodes = nullptr;
    double* d_Aci = nullptr;
    double* d_yc = nullptr;
    double* d_mass = nullptr;       // scalar P1 mass matrix on the BSR pattern [nb] (gradient projection)
    double* d_wall_w = nullptr;     // intended BCs: lumped wall-facet area per vertex [V]
    double* d_exit_m = nullptr;     // intended BCs: exit-facet mass matrix on the BSR pattern [nb]
    int* d_exit_flag = nullptr;     // [V] vertex lies on an exit facet
    double* d_bc = nullptr;         // [batch][16]: J_wall[8] | k_exit[8]
    bool facet_terms = false;
    double* d_gp = nullptr;         // gradient-projection work vectors [4][batch][V][27] + scalars
    double* d_partial = nullptr;    // first-stage partial sums of the partitioned-mode reductions
    size_t partial_doubles = 0;
    // coarse set-up: blocks sorted by (row slab, column slab) pair, cut into chunks (deterministic ordered sums)
    int* d_cp_blk = nullptr; int* d_chunk_ptr = nullptr; int* d_chunk_pair = nullptr; int n_chunk = 0;
    double* d_cpart = nullptr;      // [batch][n_chunk][81]
    double* d_AciT = nullptr;       // [batch][NCO][NCO] transposed pivoted inverse (batched solver)
    int* d_cflag = nullptr;         // [batch] vanishing pivots met by the coarse inverse
    // persistent-GMRES / device-resident Newton workspace (ensure_solver)
    int restart_alloc = 0;
    double *d_V = nullptr, *d_w = nullptr, *d_z = nullptr, *d_dx = nullptr, *d_nrm = nullptr, *d_dxmax = nullptr;
    double *d_umax = nullptr, *d_relres = nullptr, *d_r0 = nullptr, *d_r = nullptr;
    int *d_active = nullptr, *d_status = nullptr, *d_iters = nullptr, *d_lin_total = nullptr, *d_lin_its = nullptr;
    int* d_nact = nullptr;
    // march / steady drivers (gmpnp_set_march_data_3d)
    signed char* d_kind = nullptr;  // [n_dir] Dirichlet kind 0..4
    double *d_tab = nullptr, *d_sech = nullptr, *d_co2 = nullptr, *d_med = nullptr, *d_inc = nullptr;
    int *d_alive = nullptr, *d_steps = nullptr, *d_mstatus = nullptr, *d_conv = nullptr;
    bool march_set = false, sech_mode1 = false;
    // batch-lane assembly (batches of >= LANES_MIN_BATCH problems): one problem per lane, problem-minor work arrays
    bool lanes = false;
    int n_groups = 0;               // ceil(batch / 32)
    double *d_uT = nullptr, *d_unT = nullptr;   // [n_groups][V][NC][32]
    int4* d_lane_meta = nullptr;    // [2 * n_blocks] block descriptors in processing order (assemble_bsr_lanes_kernel)
};
// owned by the handle (gmpnp_handle::ext3d): no process-wide state, handles are independent
static Host3D* ext(gmpnp_handle* h) { return static_cast<Host3D*>(h->ext3d); }

void pore3d_free_ext(gmpnp_handle* h) {
    Host3D* e = ext(h);
    if (!e) return;
    h->ext3d = nullptr;
    void* bufs[] = {e->d_wall_w, e->d_exit_m, e->d_exit_flag, e->d_bc, e->d_mass, e->d_gp, e->d_partial, e->d_blk_row, e->d_blk_geo, e->d_agg, e->d_agg_ptr, e->d_agg_nodes, e->d_Aci, e->d_yc, e->d_V, e->d_w, e->d_z,
                    e->d_dx, e->d_nrm, e->d_dxmax, e->d_umax, e->d_relres, e->d_r0, e->d_r, e->d_active, e->d_status,
                    e->d_iters, e->d_lin_total, e->d_lin_its, e->d_nact, e->d_cp_blk, e->d_chunk_ptr, e->d_chunk_pair,
                    e->d_cpart, e->d_AciT, e->d_cflag, e->d_kind, e->d_tab, e->d_sech, e->d_co2, e->d_med, e->d_inc,
                    e->d_alive, e->d_steps, e->d_mstatus, e->d_conv, e->d_uT, e->d_unT, e->d_lane_meta};
    for (void* b : bufs) if (b) cudaFree(b);
    delete e;
}

extern "C" {

int gmpnp_create_3d(gmpnp_handle** out, int device, const double* h_xyz, int n_vert, const int* h_tets, int n_tet,
                    const int* h_dir_dof, int n_dir, int n_species, int batch) {
    if (!out || !h_xyz || !h_tets || n_vert < 4 || n_tet < 1 || n_species != 8 || batch < 1 || n_dir < 0 ||
        (n_dir > 0 && !h_dir_dof))
        return GMPNP_ERR_ARG;
    for (int t = 0; t < n_tet * 4; ++t)
        if (h_tets[t] < 0 || h_tets[t] >= n_vert) return GMPNP_ERR_ARG;
    for (int d = 0; d < n_dir; ++d)
        if (h_dir_dof[d] < 0 || h_dir_dof[d] >= n_vert * NC) return GMPNP_ERR_ARG;
    gmpnp_handle* h = new gmpnp_handle();
    h->dim = 3; h->device = device; h->batch = batch; h->ns = 8; h->nc = 9;
    h->n_nodes = n_vert; h->n_tet = n_tet; h->n_dir = n_dir;
    *out = h;
    Host3D* e = new Host3D();
    h->ext3d = e;
    GMPNP_CUDA_TRY(h, cudaSetDevice(device));
    upload_rules();
    // ---- geometry: grad lambda_a and volume (same formulas as oracle/forms.py:geometry) -------
    std::vector<double> geom((size_t)n_tet * 13);
    for (int t = 0; t < n_tet; ++t) {
        const double* p0 = h_xyz + 3 * (size_t)h_tets[4 * t];
        double Jm[3][3];
        for (int r = 0; r < 3; ++r) {
            const double* pr = h_xyz + 3 * (size_t)h_tets[4 * t + r + 1];
            for (int d = 0; d < 3; ++d) Jm[r][d] = pr[d] - p0[d];
        }
        const double det = Jm[0][0] * (Jm[1][1] * Jm[2][2] - Jm[1][2] * Jm[2][1]) -
                           Jm[0][1] * (Jm[1][0] * Jm[2][2] - Jm[1][2] * Jm[2][0]) +
                           Jm[0][2] * (Jm[1][0] * Jm[2][1] - Jm[1][1] * Jm[2][0]);
        if (det == 0.0) { return GMPNP_ERR_ARG; }
        // inverse of Jm; grad lambda_{r+1} = column r of inv(Jm)
        double inv[3][3];
        inv[0][0] = (Jm[1][1] * Jm[2][2] - Jm[1][2] * Jm[2][1]) / det;
        inv[0][1] = (Jm[0][2] * Jm[2][1] - Jm[0][1] * Jm[2][2]) / det;
        inv[0][2] = (Jm[0][1] * Jm[1][2] - Jm[0][2] * Jm[1][1]) / det;
        inv[1][0] = (Jm[1][2] * Jm[2][0] - Jm[1][0] * Jm[2][2]) / det;
        inv[1][1] = (Jm[0][0] * Jm[2][2] - Jm[0][2] * Jm[2][0]) / det;
        inv[1][2] = (Jm[0][2] * Jm[1][0] - Jm[0][0] * Jm[1][2]) / det;
        inv[2][0] = (Jm[1][0] * Jm[2][1] - Jm[1][1] * Jm[2][0]) / det;
        inv[2][1] = (Jm[0][1] * Jm[2][0] - Jm[0][0] * Jm[2][1]) / det;
        inv[2][2] = (Jm[0][0] * Jm[1][1] - Jm[0][1] * Jm[1][0]) / det;
        double* ge = &geom[(size_t)t * 13];
        for (int d = 0; d < 3; ++d) {
            double s = 0.0;
            for (int r = 0; r < 3; ++r) { ge[(r + 1) * 3 + d] = inv[d][r]; s += inv[d][r]; }
            ge[d] = -s;
        }
        ge[12] = fabs(det) / 6.0;
    }
    // ---- BSR pattern and gather lists -------------------------------------------------------
    std::vector<std::vector<int>> adj(n_vert);
    for (int t = 0; t < n_tet; ++t)
        for (int a = 0; a < 4; ++a)
            for (int b = 0; b < 4; ++b) adj[h_tets[4 * t + a]].push_back(h_tets[4 * t + b]);
    h->h_row_ptr.assign(n_vert + 1, 0);
    for (int v = 0; v < n_vert; ++v) {
        auto& r = adj[v];
        if (r.empty()) r.push_back(v);              // isolated vertex: keep a diagonal block
        std::sort(r.begin(), r.end());
        r.erase(std::unique(r.begin(), r.end()), r.end());
        h->h_row_ptr[v + 1] = h->h_row_ptr[v] + (int)r.size();
    }
    const int nb = h->h_row_ptr[n_vert];
    h->n_blocks = nb;
    h->h_col_idx.resize(nb);
    std::vector<int> blk_row(nb), diag_idx(n_vert);
    for (int v = 0; v < n_vert; ++v) {
        std::copy(adj[v].begin(), adj[v].end(), h->h_col_idx.begin() + h->h_row_ptr[v]);
        for (int s = h->h_row_ptr[v]; s < h->h_row_ptr[v + 1]; ++s) {
            blk_row[s] = v;
            if (h->h_col_idx[s] == v) diag_idx[v] = s;
        }
    }
    auto find_blk = [&](int va, int vb) {
        const int* b0 = h->h_col_idx.data() + h->h_row_ptr[va];
        const int* b1 = h->h_col_idx.data() + h->h_row_ptr[va + 1];
        return (int)(std::lower_bound(b0, b1, vb) - h->h_col_idx.data());
    };
    std::vector<int> blk_cnt(nb + 1, 0), node_cnt(n_vert + 1, 0);
    for (int t = 0; t < n_tet; ++t)
        for (int a = 0; a < 4; ++a) {
            node_cnt[h_tets[4 * t + a] + 1]++;
            for (int b = 0; b < 4; ++b) blk_cnt[find_blk(h_tets[4 * t + a], h_tets[4 * t + b]) + 1]++;
        }
    std::partial_sum(blk_cnt.begin(), blk_cnt.end(), blk_cnt.begin());
    std::partial_sum(node_cnt.begin(), node_cnt.end(), node_cnt.begin());
    std::vector<int> blk_src((size_t)16 * n_tet), node_src((size_t)4 * n_tet);
    std::vector<double2> blk_geo((size_t)16 * n_tet);
    {
        std::vector<int> bpos(blk_cnt.begin(), blk_cnt.end() - 1), npos(node_cnt.begin(), node_cnt.end() - 1);
        for (int t = 0; t < n_tet; ++t)            // ascending tet order => deterministic summation order
            for (int a = 0; a < 4; ++a) {
                node_src[npos[h_tets[4 * t + a]]++] = t * 4 + a;
                for (int b = 0; b < 4; ++b) {
                    const int pos = bpos[find_blk(h_tets[4 * t + a], h_tets[4 * t + b])]++;
                    blk_src[pos] = t * 16 + a * 4 + b;
                    const double* ge = &geom[(size_t)t * 13];
                    blk_geo[pos] = make_double2(ge[a * 3] * ge[b * 3] + ge[a * 3 + 1] * ge[b * 3 + 1] +
                                                ge[a * 3 + 2] * ge[b * 3 + 2], ge[12]);
                }
            }
    }
    // ---- Dirichlet flags, z-slab aggregates -------------------------------------------------
    std::vector<int> dir_flag((size_t)n_vert * NC, -1);
    for (int d = 0; d < n_dir; ++d) dir_flag[h_dir_dof[d]] = d;
    double zmin = h_xyz[2], zmax = h_xyz[2];
    for (int v = 0; v < n_vert; ++v) { zmin = std::min(zmin, h_xyz[3 * v + 2]); zmax = std::max(zmax, h_xyz[3 * v + 2]); }
    std::vector<int> agg(n_vert), agg_ptr(NZ + 1, 0), agg_nodes(n_vert);
    for (int v = 0; v < n_vert; ++v) {
        int b = (int)((h_xyz[3 * v + 2] - zmin) / (zmax - zmin > 0 ? zmax - zmin : 1.0) * NZ);
        agg[v] = std::min(std::max(b, 0), NZ - 1);
        agg_ptr[agg[v] + 1]++;
    }
    std::partial_sum(agg_ptr.begin(), agg_ptr.end(), agg_ptr.begin());
    {
        std::vector<int> pos(agg_ptr.begin(), agg_ptr.end() - 1);
        for (int v = 0; v < n_vert; ++v) agg_nodes[pos[agg[v]]++] = v;
    }
    std::vector<double> xyz(h_xyz, h_xyz + (size_t)3 * n_vert);
    std::vector<int> tets(h_tets, h_tets + (size_t)4 * n_tet);
    std::vector<int> dir_dof(h_dir_dof, h_dir_dof + n_dir);
    int rc;
    if ((rc = dev_upload(h, &h->d_x, xyz))) return rc;
    if ((rc = dev_upload(h, &h->d_tets, tets))) return rc;
    if ((rc = dev_upload(h, &h->d_geom, geom))) return rc;
    if ((rc = dev_upload(h, &h->d_row_ptr, h->h_row_ptr))) return rc;
    if ((rc = dev_upload(h, &h->d_col_idx, h->h_col_idx))) return rc;
    if ((rc = dev_upload(h, &h->d_diag_idx, diag_idx))) return rc;
    if ((rc = dev_upload(h, &h->d_blk_ptr, blk_cnt))) return rc;
    if ((rc = dev_upload(h, &h->d_blk_src, blk_src))) return rc;
    if ((rc = dev_upload(h, &h->d_node_ptr, node_cnt))) return rc;
    if ((rc = dev_upload(h, &h->d_node_src, node_src))) return rc;
    if ((rc = dev_upload(h, &h->d_dir_dof, dir_dof))) return rc;
    if ((rc = dev_upload(h, &h->d_dir_flag, dir_flag))) return rc;
    if ((rc = dev_upload(h, &e->d_blk_row, blk_row))) return rc;
    if ((rc = dev_upload(h, &e->d_blk_geo, blk_geo))) return rc;
    {   // scalar P1 mass matrix on the block pattern: M_vw = sum_t vol_t (1 + delta_vw) / 20
        std::vector<double> mass(nb, 0.0);
        for (int s = 0; s < nb; ++s)
            for (int c = blk_cnt[s]; c < blk_cnt[s + 1]; ++c) {
                const int src = blk_src[c], a = (src >> 2) & 3, b = src & 3;
                mass[s] += blk_geo[c].y * ((a == b) ? 0.1 : 0.05);
            }
        if ((rc = dev_upload(h, &e->d_mass, mass))) return rc;
    }
    if ((rc = dev_upload(h, &e->d_agg, agg))) return rc;
    if ((rc = dev_upload(h, &e->d_agg_ptr, agg_ptr))) return rc;
    if ((rc = dev_upload(h, &e->d_agg_nodes, agg_nodes))) return rc;
    const size_t B = batch;
    {   // coarse set-up lists: blocks sorted by (slab of row, slab of column), stable in the block index, cut in chunks
        std::vector<int> order(nb);
        std::iota(order.begin(), order.end(), 0);
        auto pair_of = [&](int s_) { return agg[blk_row[s_]] * NZ + agg[h->h_col_idx[s_]]; };
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return pair_of(x) < pair_of(y); });
        std::vector<int> chunk_ptr(1, 0), chunk_pair;
        for (int c = 0; c < nb;) {
            const int pr = pair_of(order[c]);
            int end = c;
            while (end < nb && end - c < CP_CHUNK && pair_of(order[end]) == pr) ++end;
            chunk_pair.push_back(pr);
            chunk_ptr.push_back(end);
            c = end;
        }
        e->n_chunk = (int)chunk_pair.size();
        if ((rc = dev_upload(h, &e->d_cp_blk, order))) return rc;
        if ((rc = dev_upload(h, &e->d_chunk_ptr, chunk_ptr))) return rc;
        if ((rc = dev_upload(h, &e->d_chunk_pair, chunk_pair))) return rc;
        GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_cpart, sizeof(double) * 81 * (size_t)e->n_chunk * B));
        GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_AciT, sizeof(double) * NCO * NCO * B));
        GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_cflag, sizeof(int) * B));
    }
    GMPNP_CUDA_TRY(h, cudaMalloc(&h->d_params, sizeof(double) * GMPNP_NPAR * B));
    GMPNP_CUDA_TRY(h, cudaMalloc(&h->d_dir_val, sizeof(double) * std::max(1, n_dir) * B));
    {   // assembly layout: one problem per lane for real batches (GMPNP_ASM_LANES=0/1 forces the choice, for tests)
        const char* env = getenv("GMPNP_ASM_LANES");
        e->lanes = env ? (atoi(env) != 0) : (batch >= LANES_MIN_BATCH);
        e->n_groups = (batch + 31) / 32;
    }
    size_t mom_doubles = NMOM * (size_t)n_tet * B;
    if (e->lanes) {
        mom_doubles = std::max(mom_doubles, (size_t)NMT * n_tet * e->n_groups * 32);
        const size_t ut = (size_t)e->n_groups * n_vert * NC * 32;
        GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_uT, sizeof(double) * ut));
        GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_unT, sizeof(double) * ut));
        {   // block descriptors, ordered along a Morton curve of the row vertex (one common scale for the three axes)
            double lo[3] = {h_xyz[0], h_xyz[1], h_xyz[2]}, ext_max = 0.0;
            double hi[3] = {h_xyz[0], h_xyz[1], h_xyz[2]};
            for (int v = 0; v < n_vert; ++v)
                for (int d = 0; d < 3; ++d) { lo[d] = std::min(lo[d], h_xyz[3 * v + d]); hi[d] = std::max(hi[d], h_xyz[3 * v + d]); }
            for (int d = 0; d < 3; ++d) ext_max = std::max(ext_max, hi[d] - lo[d]);
            auto spread = [](unsigned long long x) {       // 21 bits -> every third bit
                x &= 0x1fffffull;
                x = (x | x << 32) & 0x1f00000000ffffull; x = (x | x << 16) & 0x1f0000ff0000ffull;
                x = (x | x << 8) & 0x100f00f00f00f00full; x = (x | x << 4) & 0x10c30c30c30c30c3ull;
                x = (x | x << 2) & 0x1249249249249249ull;
                return x;
            };
            std::vector<unsigned long long> code(n_vert);
            for (int v = 0; v < n_vert; ++v) {
                unsigned long long c = 0;
                for (int d = 0; d < 3; ++d) {
                    const double f = ext_max > 0 ? (h_xyz[3 * v + d] - lo[d]) / ext_max : 0.0;
                    c |= spread((unsigned long long)(f * 2097151.0)) << d;
                }
                code[v] = c;
            }
            std::vector<int> order(nb);
            std::iota(order.begin(), order.end(), 0);
            std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return code[blk_row[x]] < code[blk_row[y]]; });
            std::vector<int4> meta(2 * (size_t)nb);
            for (int k = 0; k < nb; ++k) {
                const int b_ = order[k], va = blk_row[b_];
                int pin = 0;
                for (int i = 0; i < NC; ++i) pin |= (dir_flag[(size_t)va * NC + i] >= 0) << i;
                meta[2 * k] = make_int4(b_, va, h->h_col_idx[b_], blk_cnt[b_]);
                meta[2 * k + 1] = make_int4(blk_cnt[b_ + 1] - blk_cnt[b_], pin, 0, 0);
            }
            if ((rc = dev_upload(h, &e->d_lane_meta, meta))) return rc;
        }
        GMPNP_CUDA_TRY(h, cudaFuncSetAttribute(tet_moments_lanes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)TL_SMEM));
        GMPNP_CUDA_TRY(h, cudaFuncSetAttribute(assemble_bsr_lanes_kernel<BL_UNROLL, BL_WARPS>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bl_smem(BL_WARPS)));
    }
    GMPNP_CUDA_TRY(h, cudaMalloc(&h->d_mom, sizeof(double) * mom_doubles));
    GMPNP_CUDA_TRY(h, cudaMalloc(&h->d_Fe, sizeof(double) * 36 * (size_t)n_tet * B));
    GMPNP_CUDA_TRY(h, cudaMalloc(&h->d_J, sizeof(double) * 81 * (size_t)nb * B));
    GMPNP_CUDA_TRY(h, cudaMalloc(&h->d_Dinv, sizeof(double) * 81 * (size_t)n_vert * B));
    GMPNP_CUDA_TRY(h, cudaMalloc(&h->d_F, sizeof(double) * NC * (size_t)n_vert * B));
    GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_Aci, sizeof(double) * NCO * NCO * B));
    GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_yc, sizeof(double) * NCO * B));
    GMPNP_CUDA_TRY(h, cudaMallocHost(&h->h_pinned, sizeof(double) * 8 * B + 64));
    return GMPNP_OK;
}

int gmpnp_set_dirichlet_3d(gmpnp_handle* h, const double* h_vals, int batch) {
    if (!h || h->dim != 3 || batch != h->batch || (h->n_dir > 0 && !h_vals)) return GMPNP_ERR_ARG;
    GMPNP_CUDA_TRY(h, cudaSetDevice(h->device));
    if (h->n_dir > 0)
        GMPNP_CUDA_TRY(h, cudaMemcpy(h->d_dir_val, h_vals, sizeof(double) * h->n_dir * batch, cudaMemcpyHostToDevice));
    h->dir_set = true;
    return GMPNP_OK;
}

int gmpnp_pattern_3d(const gmpnp_handle* h, int* n_blocks, int* h_row_ptr, int* h_col_idx) {
    if (!h || h->dim != 3) return GMPNP_ERR_ARG;
    if (n_blocks) *n_blocks = h->n_blocks;
    if (h_row_ptr) std::copy(h->h_row_ptr.begin(), h->h_row_ptr.end(), h_row_ptr);
    if (h_col_idx) std::copy(h->h_col_idx.begin(), h->h_col_idx.end(), h_col_idx);
    return GMPNP_OK;
}

}  // extern "C"

static int check_3d(gmpnp_handle* h) {
    if (!h) return GMPNP_ERR_ARG;
    if (h->dim != 3 || !h->params_set || !h->dir_set) return GMPNP_ERR_STATE;
    return GMPNP_OK;
}

static int launch_assemble(gmpnp_handle* h, const double* d_u, const double* d_un, double* d_F, double* d_J,
                           cudaStream_t st) {
    GmpnpRange nvtx_range("gmpnp:assemble_3d");
    const int T = h->n_tet, V = h->n_nodes, B = h->batch;
    Host3D* e = ext(h);
    const double* exit_m = e->facet_terms ? e->d_exit_m : nullptr;
    if (e->lanes) {
        const int G = e->n_groups;
        const long n = (long)V * NC;
        dim3 gT((unsigned)((n + 31) / 32), G);
        lanes_transpose_kernel<<<gT, 256, 0, st>>>(B, n, d_u, e->d_uT);
        h->launches++;
        const double* unT = e->d_uT;                 // without u_n the time term must be off (kappa = 0), as above
        if (d_F && d_un) {
            lanes_transpose_kernel<<<gT, 256, 0, st>>>(B, n, d_un, e->d_unT);
            h->launches++;
            unT = e->d_unT;
        }
        const int per_group = std::max(1, (148 * TET_MIN_BLOCKS + G - 1) / G);  // one wave
        dim3 gA(std::min((T + TL_WARPS - 1) / TL_WARPS, per_group), G);
        tet_moments_lanes_kernel<<<gA, TL_WARPS * 32, TL_SMEM, st>>>(B, T, V, h->d_tets, h->d_geom, h->d_params, e->d_uT,
                                                                    unT, h->d_mom, h->d_Fe, d_J != nullptr,
                                                                    d_F != nullptr);
        h->launches++;
    } else {
        dim3 gA((T + 127) / 128, B);
        tet_moments_kernel<<<gA, 128, 0, st>>>(T, V, h->d_tets, h->d_geom, h->d_params, d_u, d_un, h->d_mom, h->d_Fe,
                                               d_J != nullptr, d_F != nullptr);
        h->launches++;
    }
    if (d_F) {
        dim3 gB(((long)V * NC + 255) / 256, B);
        residual_gather_kernel<<<gB, 256, 0, st>>>(V, T, h->n_dir, h->d_node_ptr, h->d_node_src, h->d_dir_flag,
                                                   h->d_dir_val, h->d_Fe, d_u, d_F,
                                                   e->facet_terms ? e->d_wall_w : nullptr, e->d_exit_m,
                                                   e->d_exit_flag, e->d_bc, h->d_row_ptr, h->d_col_idx);
        h->launches++;
    }
    if (d_J && e->lanes) {
        // one launch (one wave of 2 CTAs per SM) per lane group: the whole GPU works on ONE group's moment records at a
        // time, so the distance between the four rows that re-read a tet's record stays within the L2
        // measured on a B200 (batch 128, config 3): 4 warps per CTA, 2 CTAs per SM, three contributions in flight per warp;
        // more warps (6 or 8 per CTA) or more CTAs are slower (3.2-3.6 vs 2.9 ms), see DESIGN 3.3
        const int gx = std::min((h->n_blocks + BL_WARPS - 1) / BL_WARPS, 148 * 2);
        for (int g = 0; g < e->n_groups; ++g) {
            assemble_bsr_lanes_kernel<BL_UNROLL, BL_WARPS><<<gx, BL_WARPS * 32, bl_smem(BL_WARPS), st>>>(
                g, B, h->n_blocks, V, T, e->d_lane_meta, h->d_blk_src, e->d_blk_geo, h->d_params, e->d_uT, h->d_mom, d_J,
                exit_m, e->d_bc);
            h->launches++;
        }
    } else if (d_J) {
        int gx = std::min((h->n_blocks + ASM_WARPS - 1) / ASM_WARPS, 148 * 16);
        dim3 gC(gx, B);
        assemble_bsr_kernel<<<gC, ASM_WARPS * 32, 0, st>>>(h->n_blocks, V, T, h->d_blk_ptr, h->d_blk_src,
                                                           e->d_blk_geo, e->d_blk_row, h->d_col_idx,
                                                           h->d_dir_flag, h->d_params, d_u, h->d_mom, d_J, exit_m, e->d_bc);
        h->launches++;
    }
    GMPNP_CUDA_TRY(h, cudaGetLastError());
    return GMPNP_OK;
}

static int launch_spmv(gmpnp_handle* h, const double* d_J, const double* d_x, double* d_y, cudaStream_t st,
                       int row0 = 0, int row1 = -1) {
    GmpnpRange nvtx_range("gmpnp:spmv_3d");
    if (row1 < 0) row1 = h->n_nodes;
    if (row1 <= row0) return GMPNP_OK;
    dim3 g((row1 - row0 + 7) / 8, h->batch);
    bsr_spmv_kernel<<<g, 256, 0, st>>>(h->n_nodes, h->n_blocks, h->d_row_ptr, h->d_col_idx, d_J, d_x, d_y, row0, row1);
    h->launches++;
    return GMPNP_OK;
}

// ---- solver workspace (GMRES basis, Newton control arrays): allocated into temporaries and swapped in on success ----
static int ensure_solver(gmpnp_handle* h, int m) {
    Host3D* e = ext(h);
    if (e->restart_alloc >= m && e->d_V) return GMPNP_OK;
    const size_t B = h->batch, n = (size_t)h->n_nodes * NC, ld = m + 1;
    struct Req { void** dst; size_t bytes; };
    double *V = nullptr, *w = nullptr, *z = nullptr, *dx = nullptr, *nrm = nullptr, *dxmax = nullptr, *umax = nullptr;
    double *relres = nullptr, *r0 = nullptr, *r = nullptr;
    int *active = nullptr, *status = nullptr, *iters = nullptr, *lin_total = nullptr, *lin_its = nullptr, *nact = nullptr;
    Req reqs[] = {{(void**)&V, sizeof(double) * B * ld * n}, {(void**)&w, sizeof(double) * B * n},
                  {(void**)&z, sizeof(double) * B * n}, {(void**)&dx, sizeof(double) * B * n},
                  {(void**)&nrm, sizeof(double) * B}, {(void**)&dxmax, sizeof(double) * B},
                  {(void**)&umax, sizeof(double) * B}, {(void**)&relres, sizeof(double) * B},
                  {(void**)&r0, sizeof(double) * B}, {(void**)&r, sizeof(double) * B},
                  {(void**)&active, sizeof(int) * B}, {(void**)&status, sizeof(int) * B},
                  {(void**)&iters, sizeof(int) * B}, {(void**)&lin_total, sizeof(int) * B},
                  {(void**)&lin_its, sizeof(int) * B}, {(void**)&nact, sizeof(int) * 4}};
    // free the old set first (a longer restart replaces the basis; peak memory stays one basis), and forget it
    void* old[] = {e->d_V, e->d_w, e->d_z, e->d_dx, e->d_nrm, e->d_dxmax, e->d_umax, e->d_relres, e->d_r0, e->d_r,
                   e->d_active, e->d_status, e->d_iters, e->d_lin_total, e->d_lin_its, e->d_nact};
    for (void* b : old) if (b) cudaFree(b);
    e->d_V = e->d_w = e->d_z = e->d_dx = e->d_nrm = e->d_dxmax = e->d_umax = e->d_relres = e->d_r0 = e->d_r = nullptr;
    e->d_active = e->d_status = e->d_iters = e->d_lin_total = e->d_lin_its = e->d_nact = nullptr;
    e->restart_alloc = 0;
    for (auto& q : reqs) {
        if (cudaMalloc(q.dst, q.bytes) != cudaSuccess) {
            cudaGetLastError();
            for (auto& f : reqs) if (*f.dst) { cudaFree(*f.dst); *f.dst = nullptr; }
            h->last_cuda_error = "cudaMalloc of the GMRES workspace failed";
            return GMPNP_ERR_ALLOC;
        }
    }
    e->d_V = V; e->d_w = w; e->d_z = z; e->d_dx = dx; e->d_nrm = nrm; e->d_dxmax = dxmax; e->d_umax = umax;
    e->d_relres = relres; e->d_r0 = r0; e->d_r = r; e->d_active = active; e->d_status = status; e->d_iters = iters;
    e->d_lin_total = lin_total; e->d_lin_its = lin_its; e->d_nact = nact;
    e->restart_alloc = m;
    return GMPNP_OK;
}

// cluster size of the persistent GMRES kernel: fill the 148 SMs when the batch is small
static int gmres_cluster_size(int batch) {
    // experiments / tests only: GMPNP_GMRES_CLUSTER=1|2|4|8 forces the cluster size (the result of a solve does not
    // depend on it beyond the summation order of the dot products)
    if (const char* env = getenv("GMPNP_GMRES_CLUSTER")) {
        const int g = atoi(env);
        if (g == 1 || g == 2 || g == 4 || g == 8) return g;
    }
    int g = 1;
    while (g < 8 && batch * g * 2 <= 148) g *= 2;
    return g;
}

static size_t gmres_smem_bytes(int m) {
    const size_t ld = m + 1, red_n = ld + 3;
    return sizeof(double) * (3 * red_n + GM_WARPS * GM_KT + GM_WARPS * 96 + 6 * NCO + ld * m + 2 * m + 2 * ld);
}

// Jacobian of the current iterate is in h->d_J: block-Jacobi inverses + coarse inverse (deterministic, pivoted)
static int launch_precond_setup(gmpnp_handle* h, cudaStream_t st) {
    GmpnpRange nvtx_range("gmpnp:precond_setup_3d");
    Host3D* e = ext(h);
    const int V = h->n_nodes, B = h->batch;
    dim3 gj((V + 63) / 64, B);
    bjacobi_invert_kernel<<<gj, 64, 0, st>>>(V, h->n_blocks, h->d_diag_idx, h->d_J, h->d_Dinv);
    dim3 gp(e->n_chunk, B);
    coarse_partial_kernel<<<gp, 96, 0, st>>>(h->n_blocks, e->d_chunk_ptr, e->d_cp_blk, e->d_blk_row, h->d_col_idx,
                                             h->d_dir_flag, h->d_J, e->d_cpart, e->n_chunk);
    cudaFuncSetAttribute(coarse_invert_pivoted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(sizeof(double) * NCO * NCO));
    coarse_invert_pivoted_kernel<<<B, 1024, sizeof(double) * NCO * NCO, st>>>(e->d_chunk_pair, e->d_cpart, e->n_chunk,
                                                                              e->d_AciT, e->d_cflag);
    h->launches += 3;
    GMPNP_CUDA_TRY(h, cudaGetLastError());
    return GMPNP_OK;
}

// J dx = b for every enabled problem: ONE launch, no host synchronisation (gmres_cluster_kernel)
static int launch_gmres(gmpnp_handle* h, const double* d_b, double* d_x, int m, int maxit, double rtol,
                        const int* d_enabled, cudaStream_t st) {
    GmpnpRange nvtx_range("gmpnp:gmres_3d");
    Host3D* e = ext(h);
    GmresArgs a;
    a.n_vert = h->n_nodes; a.n_blocks = h->n_blocks; a.m = m; a.maxit = maxit; a.rtol = rtol;
    a.row_ptr = h->d_row_ptr; a.col_idx = h->d_col_idx; a.agg = e->d_agg; a.agg_ptr = e->d_agg_ptr;
    a.agg_nodes = e->d_agg_nodes; a.dir_flag = h->d_dir_flag;
    a.J = h->d_J; a.Dinv = h->d_Dinv; a.AciT = e->d_AciT; a.b = d_b;
    a.x = d_x; a.V = e->d_V; a.w = e->d_w; a.z = e->d_z;
    a.enabled = d_enabled; a.its = e->d_lin_its; a.relres = e->d_relres;
    const int G = gmres_cluster_size(h->batch);
    const size_t smem = gmres_smem_bytes(m);
    GMPNP_CUDA_TRY(h, cudaFuncSetAttribute(gmres_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(h->batch * G, 1, 1);
    cfg.blockDim = dim3(GM_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    GMPNP_CUDA_TRY(h, cudaLaunchKernelEx(&cfg, gmres_cluster_kernel, a));
    h->launches++;
    return GMPNP_OK;
}

// One reference `solve(F == 0, u, bcs)` per enabled problem (dolfin NewtonSolver semantics), control
// state on the device (NewtonCtl arrays in the handle); the host reads ONE counter (problems still active) per Newton
// iteration.  d_enabled (device, may be NULL): problems with 0 are left untouched and report GMPNP_CONVERGED, 0 iterations.
static int newton_run(gmpnp_handle* h, double* d_u, const double* d_un, const gmpnp_newton_opts* o, const int* d_enabled,
                      cudaStream_t st) {
    GmpnpRange nvtx_range("gmpnp:newton_3d");
    Host3D* e = ext(h);
    const int V = h->n_nodes, B = h->batch;
    const long n = (long)V * NC;
    const int m = o->lin_restart > 0 ? o->lin_restart : 50;
    const int lin_maxit = o->lin_maxit > 0 ? o->lin_maxit : 1000;
    const double lin_rtol = o->lin_rtol > 0 ? o->lin_rtol : 1e-10;
    if (gmres_smem_bytes(m) > 220 * 1024) return GMPNP_ERR_ARG;          // restart length too long for shared memory
    int rc = ensure_solver(h, m); if (rc) return rc;
    NewtonCtl c{e->d_active, e->d_status, e->d_iters, e->d_lin_total, e->d_r0, e->d_r};
    int* hp = (int*)h->h_pinned;
    const int gB = (B + 127) / 128;
    auto residual_norms = [&]() -> int {
        int rc2 = launch_assemble(h, d_u, d_un, h->d_F, nullptr, st); if (rc2) return rc2;
        norm_scale_kernel<<<B, 1024, 0, st>>>(n, h->d_F, nullptr, 0, e->d_nrm);
        h->launches++;
        return GMPNP_OK;
    };
    auto read_active = [&](int& n_act) -> int {
        GMPNP_CUDA_TRY(h, cudaMemcpyAsync(hp, e->d_nact, sizeof(int), cudaMemcpyDeviceToHost, st));
        GMPNP_CUDA_TRY(h, cudaStreamSynchronize(st));
        n_act = hp[0];
        return GMPNP_OK;
    };
    rc = residual_norms(); if (rc) return rc;
    GMPNP_CUDA_TRY(h, cudaMemsetAsync(e->d_nact, 0, sizeof(int), st));
    newton_begin_kernel<<<gB, 128, 0, st>>>(B, d_enabled, e->d_nrm, c, o->criterion, o->atol, e->d_nact);
    h->launches++;
    int n_act = 0;
    rc = read_active(n_act); if (rc) return rc;
    for (int k = 0; k < o->maxit && n_act > 0; ++k) {
        // Jacobian at the current iterate, preconditioner, linear solve, masked update, new residual, control
        rc = launch_assemble(h, d_u, d_un, nullptr, h->d_J, st); if (rc) return rc;
        rc = launch_precond_setup(h, st); if (rc) return rc;
        rc = launch_gmres(h, h->d_F, e->d_dx, m, lin_maxit, lin_rtol, e->d_active, st); if (rc) return rc;
        newton_update_kernel<<<B, 1024, 0, st>>>(n, o->relax, e->d_active, e->d_dx, d_u, e->d_dxmax, e->d_umax);
        h->launches++;
        rc = residual_norms(); if (rc) return rc;
        GMPNP_CUDA_TRY(h, cudaMemsetAsync(e->d_nact, 0, sizeof(int), st));
        newton_step_kernel<<<gB, 128, 0, st>>>(B, k, o->maxit, e->d_nrm, e->d_dxmax, e->d_umax, e->d_lin_its, e->d_relres, c,
                                               o->criterion, o->rtol, o->atol, o->xtol, lin_rtol, e->d_nact);
        h->launches++;
        rc = read_active(n_act); if (rc) return rc;
    }
    GMPNP_CUDA_TRY(h, cudaGetLastError());
    return GMPNP_OK;
}

extern "C" {

int gmpnp_assemble_3d(gmpnp_handle* h, const double* d_u, const double* d_un, double* d_F, double* d_J, void* stream) {
    int rc = check_3d(h); if (rc) return rc;
    if (!d_u || !d_un) return GMPNP_ERR_ARG;
    GMPNP_CUDA_TRY(h, cudaSetDevice(h->device));
    return launch_assemble(h, d_u, d_un, d_F, d_J, (cudaStream_t)stream);
}

int gmpnp_spmv_3d(gmpnp_handle* h, const double* d_J, const double* d_x, double* d_y, void* stream) {
    if (!h || h->dim != 3 || !d_x || !d_y) return GMPNP_ERR_ARG;
    GMPNP_CUDA_TRY(h, cudaSetDevice(h->device));
    launch_spmv(h, d_J ? d_J : h->d_J, d_x, d_y, (cudaStream_t)stream);
    GMPNP_CUDA_TRY(h, cudaGetLastError());
    return GMPNP_OK;
}

int gmpnp_spmv_rows_3d(gmpnp_handle* h, const double* d_J, const double* d_x, double* d_y, int row0, int row1,
                       void* stream) {
    if (!h || h->dim != 3 || !d_x || !d_y || row0 < 0 || row1 > h->n_nodes || row0 > row1) return GMPNP_ERR_ARG;
    GMPNP_CUDA_TRY(h, cudaSetDevice(h->device));
    launch_spmv(h, d_J ? d_J : h->d_J, d_x, d_y, (cudaStream_t)stream, row0, row1);
    GMPNP_CUDA_TRY(h, cudaGetLastError());
    return GMPNP_OK;
}

int gmpnp_newton_3d(gmpnp_handle* h, double* d_u, const double* d_un, const gmpnp_newton_opts* o, int* d_iters,
                    double* d_r0, double* d_r, int* d_lin_iters, int* d_status, void* stream) {
    int rc = check_3d(h); if (rc) return rc;
    if (!d_u || !d_un || !o) return GMPNP_ERR_ARG;
    GMPNP_CUDA_TRY(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    rc = newton_run(h, d_u, d_un, o, nullptr, st); if (rc) return rc;
    Host3D* e = ext(h);
    const size_t B = h->batch;
    const cudaMemcpyKind dd = cudaMemcpyDeviceToDevice;
    if (d_iters) GMPNP_CUDA_TRY(h, cudaMemcpyAsync(d_iters, e->d_iters, sizeof(int) * B, dd, st));
    if (d_status) GMPNP_CUDA_TRY(h, cudaMemcpyAsync(d_status, e->d_status, sizeof(int) * B, dd, st));
    if (d_lin_iters) GMPNP_CUDA_TRY(h, cudaMemcpyAsync(d_lin_iters, e->d_lin_total, sizeof(int) * B, dd, st));
    if (d_r0) GMPNP_CUDA_TRY(h, cudaMemcpyAsync(d_r0, e->d_r0, sizeof(double) * B, dd, st));
    if (d_r) GMPNP_CUDA_TRY(h, cudaMemcpyAsync(d_r, e->d_r, sizeof(double) * B, dd, st));
    GMPNP_CUDA_TRY(h, cudaStreamSynchronize(st));
    return GMPNP_OK;
}

// ---- the reference's loop inside the library (3D/MPNP_CO2ER_pore.py:782-858) ----------------------------------
int gmpnp_set_march_data_3d(gmpnp_handle* h, const signed char* h_kind, const double* h_tab, const double* h_sech,
                            int batch) {
    if (!h || h->dim != 3 || batch != h->batch || !h_tab || !h_sech || (h->n_dir > 0 && !h_kind)) return GMPNP_ERR_ARG;
    for (int d = 0; d < h->n_dir; ++d)
        if (h_kind[d] < 0 || h_kind[d] > 4) return GMPNP_ERR_ARG;
    GMPNP_CUDA_TRY(h, cudaSetDevice(h->device));
    Host3D* e = ext(h);
    const size_t B = batch;
    if (!e->d_kind) {
        GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_kind, std::max(1, h->n_dir)));
        GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_tab, sizeof(double) * 4 * B));
        GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_sech, sizeof(double) * 8 * B));
        GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_co2, sizeof(double) * B));
        GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_med, sizeof(double) * 4 * B));
        GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_inc, sizeof(double) * B));
        GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_alive, sizeof(int) * B));
        GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_steps, sizeof(int) * B));
        GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_mstatus, sizeof(int) * B));
        GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_conv, sizeof(int) * B));
    }
    if (h->n_dir > 0) GMPNP_CUDA_TRY(h, cudaMemcpy(e->d_kind, h_kind, h->n_dir, cudaMemcpyHostToDevice));
    GMPNP_CUDA_TRY(h, cudaMemcpy(e->d_tab, h_tab, sizeof(double) * 4 * B, cudaMemcpyHostToDevice));
    GMPNP_CUDA_TRY(h, cudaMemcpy(e->d_sech, h_sech, sizeof(double) * 8 * B, cudaMemcpyHostToDevice));
    e->sech_mode1 = (h_sech[5] != 0.0);
    e->march_set = true;
    h->dir_set = true;                  // the march fills the Dirichlet values itself
    return GMPNP_OK;
}

// one pseudo-time step for all alive problems: Dirichlet values (wall voltage * ramp, current CO2 entry value), one
// Newton solve, bookkeeping, Sechenov update from the medians, u_n <- u
static int march_step(gmpnp_handle* h, double* d_u, double* d_un, const gmpnp_newton_opts* o, int step, int n_steps,
                      double ramp, double* d_hist_row, long hist_stride, int* d_iters, int* d_lin_iters, double* d_co2_out,
                      int* n_alive, cudaStream_t st, double steady_tol = 0.0) {
    GmpnpRange nvtx_range("gmpnp:march_step_3d");
    Host3D* e = ext(h);
    const int V = h->n_nodes, B = h->batch;
    const long n = (long)V * NC;
    dim3 gd((std::max(1, h->n_dir) + 255) / 256, B);
    dirichlet_fill_kernel<<<gd, 256, 0, st>>>(B, h->n_dir, e->d_kind, e->d_tab, e->d_co2, ramp, h->d_dir_val, h->d_params);
    h->launches++;
    int rc = newton_run(h, d_u, d_un, o, e->d_alive, st); if (rc) return rc;
    NewtonCtl c{e->d_active, e->d_status, e->d_iters, e->d_lin_total, e->d_r0, e->d_r};
    GMPNP_CUDA_TRY(h, cudaMemsetAsync(e->d_nact, 0, sizeof(int), st));
    march_record_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, step, n_steps, c, e->d_alive, d_iters, d_lin_iters, e->d_mstatus,
                                                        e->d_steps, e->d_co2, d_co2_out, e->d_nact);
    int np2 = 1;
    while (np2 < V) np2 <<= 1;
    dim3 gm(B, 4);
    const int mc0 = e->sech_mode1 ? 0 : 1, mc1 = e->sech_mode1 ? 1 : 2, mc2 = e->sech_mode1 ? 2 : 3, mc3 = e->sech_mode1 ? 3 : 7;
    if ((size_t)np2 * sizeof(double) > MEDIAN_SORT_BYTES) {
        median_select_kernel<<<gm, 1024, 0, st>>>(V, 4, mc0, mc1, mc2, mc3, d_u, e->d_med);
    } else {
        cudaFuncSetAttribute(median4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(np2 * sizeof(double)));
        median4_kernel<<<gm, 1024, np2 * sizeof(double), st>>>(V, np2, mc0, mc1, mc2, mc3, d_u, e->d_med);
    }
    sechenov_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, e->d_alive, e->d_med, e->d_sech, e->d_co2);
    march_advance_kernel<<<B, 1024, 0, st>>>(n, e->d_alive, d_u, d_un, d_hist_row, hist_stride, e->d_inc);
    h->launches += 4;
    if (steady_tol > 0.0) {
        GMPNP_CUDA_TRY(h, cudaMemsetAsync(e->d_nact, 0, sizeof(int), st));
        steady_check_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, steady_tol, e->d_alive, e->d_inc, e->d_conv, e->d_nact);
        h->launches++;
    }
    int* hp = (int*)h->h_pinned;
    GMPNP_CUDA_TRY(h, cudaMemcpyAsync(hp, e->d_nact, sizeof(int), cudaMemcpyDeviceToHost, st));
    GMPNP_CUDA_TRY(h, cudaStreamSynchronize(st));
    *n_alive = hp[0];
    GMPNP_CUDA_TRY(h, cudaGetLastError());
    return GMPNP_OK;
}

static int march_begin(gmpnp_handle* h, cudaStream_t st) {
    Host3D* e = ext(h);
    const int B = h->batch;
    if (!e->march_set) return GMPNP_ERR_STATE;
    std::vector<int> ones(B, 1), zeros(B, 0);
    std::vector<double> co2(B);
    std::vector<double> tab(4 * (size_t)B);
    GMPNP_CUDA_TRY(h, cudaMemcpy(tab.data(), e->d_tab, sizeof(double) * 4 * B, cudaMemcpyDeviceToHost));
    for (int p = 0; p < B; ++p) co2[p] = tab[4 * (size_t)p + 1];
    GMPNP_CUDA_TRY(h, cudaMemcpyAsync(e->d_co2, co2.data(), sizeof(double) * B, cudaMemcpyHostToDevice, st));
    GMPNP_CUDA_TRY(h, cudaMemcpyAsync(e->d_alive, ones.data(), sizeof(int) * B, cudaMemcpyHostToDevice, st));
    GMPNP_CUDA_TRY(h, cudaMemcpyAsync(e->d_steps, zeros.data(), sizeof(int) * B, cudaMemcpyHostToDevice, st));
    GMPNP_CUDA_TRY(h, cudaMemcpyAsync(e->d_mstatus, zeros.data(), sizeof(int) * B, cudaMemcpyHostToDevice, st));
    GMPNP_CUDA_TRY(h, cudaMemcpyAsync(e->d_conv, zeros.data(), sizeof(int) * B, cudaMemcpyHostToDevice, st));
    GMPNP_CUDA_TRY(h, cudaStreamSynchronize(st));
    return GMPNP_OK;
}

int gmpnp_march_3d(gmpnp_handle* h, double* d_u, double* d_un, int n_steps, const gmpnp_newton_opts* opts,
                   double* d_hist, int* d_iters, int* d_lin_iters, double* d_co2, int* d_steps, int* d_status,
                   void* stream) {
    if (!h || h->dim != 3 || !h->params_set) return h ? GMPNP_ERR_STATE : GMPNP_ERR_ARG;
    if (!d_u || !d_un || !opts || n_steps < 1) return GMPNP_ERR_ARG;
    GMPNP_CUDA_TRY(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = march_begin(h, st); if (rc) return rc;
    Host3D* e = ext(h);
    const long n = (long)h->n_nodes * NC;
    const size_t B = h->batch;
    for (int s = 0; s < n_steps; ++s) {
        int n_alive = 0;
        rc = march_step(h, d_u, d_un, opts, s, n_steps, 1.0, d_hist ? d_hist + (long)s * n : nullptr, (long)n_steps * n,
                        d_iters, d_lin_iters, d_co2, &n_alive, st);
        if (rc) return rc;
        if (n_alive == 0) break;
    }
    if (d_steps) GMPNP_CUDA_TRY(h, cudaMemcpyAsync(d_steps, e->d_steps, sizeof(int) * B, cudaMemcpyDeviceToDevice, st));
    if (d_status) GMPNP_CUDA_TRY(h, cudaMemcpyAsync(d_status, e->d_mstatus, sizeof(int) * B, cudaMemcpyDeviceToDevice, st));
    GMPNP_CUDA_TRY(h, cudaStreamSynchronize(st));
    return GMPNP_OK;
}

int gmpnp_steady_3d(gmpnp_handle* h, double* d_u, double* d_un, const gmpnp_newton_opts* opts, double tol, int max_steps,
                    int n_ramp, int* d_iters, double* d_inc_hist, double* d_co2, int* d_steps, int* d_status,
                    int* d_converged, int* h_steps_run, void* stream) {
    if (!h || h->dim != 3 || !h->params_set) return h ? GMPNP_ERR_STATE : GMPNP_ERR_ARG;
    if (!d_u || !d_un || !opts || max_steps < 1 || n_ramp < 1 || !(tol > 0.0)) return GMPNP_ERR_ARG;
    GMPNP_CUDA_TRY(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = march_begin(h, st); if (rc) return rc;
    Host3D* e = ext(h);
    const size_t B = h->batch;
    int steps = 0;
    for (int s = 0; s < max_steps; ++s) {
        const double ramp = std::min(1.0, (double)(s + 1) / (double)n_ramp);
        int n_alive = 0;
        // the steady-state test (per problem, on the device) starts once the ramp is complete
        rc = march_step(h, d_u, d_un, opts, s, max_steps, ramp, nullptr, 0, d_iters, nullptr, nullptr, &n_alive, st,
                        (s + 1 >= n_ramp) ? tol : 0.0);
        if (rc) return rc;
        steps = s + 1;
        if (d_inc_hist)
            GMPNP_CUDA_TRY(h, cudaMemcpyAsync(d_inc_hist + (size_t)s * B, e->d_inc, sizeof(double) * B, cudaMemcpyDeviceToDevice, st));
        if (n_alive == 0) break;
    }
    if (h_steps_run) *h_steps_run = steps;
    const cudaMemcpyKind dd = cudaMemcpyDeviceToDevice;
    if (d_co2) GMPNP_CUDA_TRY(h, cudaMemcpyAsync(d_co2, e->d_co2, sizeof(double) * B, dd, st));
    if (d_steps) GMPNP_CUDA_TRY(h, cudaMemcpyAsync(d_steps, e->d_steps, sizeof(int) * B, dd, st));
    if (d_status) GMPNP_CUDA_TRY(h, cudaMemcpyAsync(d_status, e->d_mstatus, sizeof(int) * B, dd, st));
    if (d_converged) GMPNP_CUDA_TRY(h, cudaMemcpyAsync(d_converged, e->d_conv, sizeof(int) * B, dd, st));
    GMPNP_CUDA_TRY(h, cudaStreamSynchronize(st));
    return GMPNP_OK;
}

// ---- mesh-partitioned mode: building blocks of the distributed GMRES (see gmpnp_b200/dist3d.py) -------------
int gmpnp_vec_multi_dot(gmpnp_handle* h, const double* d_V, long long vstride, int nvec, const double* d_w,
                        long long n, double* d_out, void* stream) {
    if (!h || h->dim != 3 || !d_V || !d_w || !d_out || nvec < 1 || n < 1) return GMPNP_ERR_ARG;
    GMPNP_CUDA_TRY(h, cudaSetDevice(h->device));
    Host3D* e = ext(h);
    cudaStream_t st = (cudaStream_t)stream;
    const int nchunks = (int)((n + VEC_CHUNK - 1) / VEC_CHUNK);
    const size_t need = (size_t)nchunks * nvec;
    if (e->partial_doubles < need) {
        if (e->d_partial) GMPNP_CUDA_TRY(h, cudaFree(e->d_partial));
        e->partial_doubles = std::max(need, (size_t)nchunks * 128);
        GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_partial, sizeof(double) * e->partial_doubles));
    }
    dim3 g(nchunks, nvec);
    vec_dot_partial_kernel<<<g, 256, 0, st>>>(n, vstride, d_V, d_w, e->d_partial, nchunks);
    vec_dot_final_kernel<<<nvec, 256, 0, st>>>(e->d_partial, nchunks, d_out);
    h->launches += 2;
    GMPNP_CUDA_TRY(h, cudaGetLastError());
    return GMPNP_OK;
}

int gmpnp_vec_lincomb(gmpnp_handle* h, const double* d_V, long long vstride, int nvec, const double* d_coef,
                      double beta, const double* d_y, double* d_out, long long n, void* stream) {
    if (!h || h->dim != 3 || !d_V || !d_coef || !d_out || nvec < 1 || n < 1) return GMPNP_ERR_ARG;
    GMPNP_CUDA_TRY(h, cudaSetDevice(h->device));
    vec_lincomb_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, vstride, nvec, d_V, d_coef, beta,
                                                                                   d_y, d_out);
    h->launches++;
    GMPNP_CUDA_TRY(h, cudaGetLastError());
    return GMPNP_OK;
}

int gmpnp_bjacobi_setup_3d(gmpnp_handle* h, const double* d_J, void* stream) {
    if (!h || h->dim != 3) return GMPNP_ERR_ARG;
    GMPNP_CUDA_TRY(h, cudaSetDevice(h->device));
    dim3 gj((h->n_nodes + 63) / 64, h->batch);
    bjacobi_invert_kernel<<<gj, 64, 0, (cudaStream_t)stream>>>(h->n_nodes, h->n_blocks, h->d_diag_idx, d_J ? d_J : h->d_J,
                                                               h->d_Dinv);
    h->launches++;
    GMPNP_CUDA_TRY(h, cudaGetLastError());
    return GMPNP_OK;
}

int gmpnp_bjacobi_apply_3d(gmpnp_handle* h, const double* d_r, double* d_z, int n_rows, void* stream) {
    if (!h || h->dim != 3 || !d_r || !d_z || h->batch != 1 || n_rows < 1 || n_rows > h->n_nodes) return GMPNP_ERR_ARG;
    GMPNP_CUDA_TRY(h, cudaSetDevice(h->device));
    const long total = (long)n_rows * NC;
    bjacobi_apply_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n_rows, h->d_Dinv, d_r, d_z);
    h->launches++;
    GMPNP_CUDA_TRY(h, cudaGetLastError());
    return GMPNP_OK;
}

int gmpnp_set_facet_terms_3d(gmpnp_handle* h, const double* h_wall_w, const int* h_exit_facets,
                             const double* h_exit_area, int n_exit, const double* h_jwall, const double* h_kexit, int batch) {
    if (!h || h->dim != 3) return GMPNP_ERR_ARG;
    Host3D* e = ext(h);
    if (!h_wall_w) { e->facet_terms = false; return GMPNP_OK; }          // back to the as-executed form
    if (n_exit < 0 || (n_exit > 0 && (!h_exit_facets || !h_exit_area)) || !h_jwall || !h_kexit || batch != h->batch)
        return GMPNP_ERR_ARG;
    GMPNP_CUDA_TRY(h, cudaSetDevice(h->device));
    const int V = h->n_nodes, nb = h->n_blocks;
    std::vector<double> em(nb, 0.0);
    std::vector<int> flag(V, 0);
    for (int f = 0; f < n_exit; ++f) {
        for (int a = 0; a < 3; ++a) {
            const int va = h_exit_facets[3 * f + a];
            if (va < 0 || va >= V) return GMPNP_ERR_ARG;
            flag[va] = 1;
            for (int b = 0; b < 3; ++b) {
                const int vb = h_exit_facets[3 * f + b];
                const int* b0 = h->h_col_idx.data() + h->h_row_ptr[va];
                const int* b1 = h->h_col_idx.data() + h->h_row_ptr[va + 1];
                const int* it = std::lower_bound(b0, b1, vb);
                if (it == b1 || *it != vb) return GMPNP_ERR_ARG;        // a facet edge is always a tet edge
                em[it - h->h_col_idx.data()] += h_exit_area[f] * ((a == b) ? (1.0 / 6.0) : (1.0 / 12.0));
            }
        }
    }
    std::vector<double> bc((size_t)batch * 16);
    for (int p = 0; p < batch; ++p)
        for (int i = 0; i < 8; ++i) { bc[(size_t)p * 16 + i] = h_jwall[p * 8 + i]; bc[(size_t)p * 16 + 8 + i] = h_kexit[p * 8 + i]; }
    void* old[] = {e->d_wall_w, e->d_exit_m, e->d_exit_flag, e->d_bc};
    for (void* b : old) if (b) cudaFree(b);
    e->d_wall_w = nullptr; e->d_exit_m = nullptr; e->d_exit_flag = nullptr; e->d_bc = nullptr;
    std::vector<double> ww(h_wall_w, h_wall_w + V);
    int rc;
    if ((rc = dev_upload(h, &e->d_wall_w, ww))) return rc;
    if ((rc = dev_upload(h, &e->d_exit_m, em))) return rc;
    if ((rc = dev_upload(h, &e->d_exit_flag, flag))) return rc;
    if ((rc = dev_upload(h, &e->d_bc, bc))) return rc;
    e->facet_terms = true;
    return GMPNP_OK;
}

int gmpnp_grad_project_3d(gmpnp_handle* h, const double* d_u, double* d_g, int n_iter, void* stream) {
    if (!h || h->dim != 3 || !d_u || !d_g || n_iter < 1) return GMPNP_ERR_ARG;
    GMPNP_CUDA_TRY(h, cudaSetDevice(h->device));
    Host3D* e = ext(h);
    cudaStream_t st = (cudaStream_t)stream;
    const int V = h->n_nodes, B = h->batch;
    const size_t nvec = (size_t)B * V * NG;
    if (!e->d_gp) GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_gp, sizeof(double) * (4 * nvec + 3 * (size_t)B * NG)));
    double *r = e->d_gp, *z = r + nvec, *pv = z + nvec, *q = pv + nvec;
    double *rz = q + nvec, *pq = rz + (size_t)B * NG, *rz2 = pq + (size_t)B * NG;
    dim3 g((unsigned)(((long)V * NG + 255) / 256), B), gd(NG, B);
    gradproj_rhs_kernel<<<g, 256, 0, st>>>(V, h->d_node_ptr, h->d_node_src, h->d_tets, h->d_geom, e->d_mass,
                                           h->d_diag_idx, d_u, d_g, r, z, pv);
    gradproj_dot_kernel<<<gd, 256, 0, st>>>(V, r, z, rz);
    h->launches += 2;
    for (int it = 0; it < n_iter; ++it) {
        gradproj_spmv_kernel<<<g, 256, 0, st>>>(V, h->d_row_ptr, h->d_col_idx, e->d_mass, pv, q);
        gradproj_dot_kernel<<<gd, 256, 0, st>>>(V, pv, q, pq);
        gradproj_update_kernel<<<g, 256, 0, st>>>(V, e->d_mass, h->d_diag_idx, rz, pq, pv, q, d_g, r, z);
        gradproj_dot_kernel<<<gd, 256, 0, st>>>(V, r, z, rz2);
        gradproj_dir_kernel<<<g, 256, 0, st>>>(V, rz2, rz, z, pv);
        std::swap(rz, rz2);
        h->launches += 5;
    }
    GMPNP_CUDA_TRY(h, cudaGetLastError());
    return GMPNP_OK;
}

int gmpnp_set_aggregates_3d(gmpnp_handle* h, const int* h_agg) {
    if (!h || h->dim != 3 || !h_agg) return GMPNP_ERR_ARG;
    for (int v = 0; v < h->n_nodes; ++v)
        if (h_agg[v] < 0 || h_agg[v] >= NZ) return GMPNP_ERR_ARG;
    GMPNP_CUDA_TRY(h, cudaSetDevice(h->device));
    GMPNP_CUDA_TRY(h, cudaMemcpy(ext(h)->d_agg, h_agg, sizeof(int) * h->n_nodes, cudaMemcpyHostToDevice));
    return GMPNP_OK;
}

int gmpnp_coarse_accumulate_3d(gmpnp_handle* h, const double* d_J, int n_rows, double* d_Ac, void* stream) {
    if (!h || h->dim != 3 || !d_Ac || h->batch != 1 || n_rows < 1 || n_rows > h->n_nodes) return GMPNP_ERR_ARG;
    GMPNP_CUDA_TRY(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int smem = (int)(sizeof(double) * NCO * NCO);
    GMPNP_CUDA_TRY(h, cudaFuncSetAttribute(coarse_accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    GMPNP_CUDA_TRY(h, cudaMemsetAsync(d_Ac, 0, sizeof(double) * NCO * NCO, st));
    const int grid = std::max(1, std::min(148, (n_rows + 31) / 32));
    coarse_accumulate_kernel<<<grid, 1024, smem, st>>>(n_rows, h->d_row_ptr, h->d_col_idx, ext(h)->d_agg, h->d_dir_flag,
                                                        d_J ? d_J : h->d_J, d_Ac);
    h->launches++;
    GMPNP_CUDA_TRY(h, cudaGetLastError());
    return GMPNP_OK;
}

int gmpnp_coarse_invert_3d(gmpnp_handle* h, const double* d_Ac, void* stream) {
    if (!h || h->dim != 3 || !d_Ac || h->batch != 1) return GMPNP_ERR_ARG;
    GMPNP_CUDA_TRY(h, cudaSetDevice(h->device));
    const int smem = (int)(sizeof(double) * NCO * NCO);
    GMPNP_CUDA_TRY(h, cudaFuncSetAttribute(coarse_invert_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    coarse_invert_kernel<<<1, 1024, smem, (cudaStream_t)stream>>>(d_Ac, ext(h)->d_Aci, ext(h)->d_cflag);
    h->launches++;
    GMPNP_CUDA_TRY(h, cudaGetLastError());
    return GMPNP_OK;
}

int gmpnp_coarse_restrict_3d(gmpnp_handle* h, const double* d_r, int n_rows, double* d_rc, void* stream) {
    if (!h || h->dim != 3 || !d_r || !d_rc || h->batch != 1 || n_rows < 1 || n_rows > h->n_nodes) return GMPNP_ERR_ARG;
    GMPNP_CUDA_TRY(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    GMPNP_CUDA_TRY(h, cudaMemsetAsync(d_rc, 0, sizeof(double) * NCO, st));
    const long total = (long)n_rows * NC;
    const int grid = (int)std::max<long>(1, std::min<long>(148 * 4, (total + 2047) / 2048));
    coarse_restrict_kernel<<<grid, 256, 0, st>>>(n_rows, ext(h)->d_agg, h->d_dir_flag, d_r, d_rc);
    h->launches++;
    GMPNP_CUDA_TRY(h, cudaGetLastError());
    return GMPNP_OK;
}

int gmpnp_coarse_prolong_3d(gmpnp_handle* h, const double* d_rc, double* d_z, int n_rows, void* stream) {
    if (!h || h->dim != 3 || !d_rc || !d_z || h->batch != 1 || n_rows < 1 || n_rows > h->n_nodes) return GMPNP_ERR_ARG;
    GMPNP_CUDA_TRY(h, cudaSetDevice(h->device));
    const long total = (long)n_rows * NC;
    const int grid = (int)std::max<long>(1, std::min<long>(148 * 4, (total + 2047) / 2048));
    coarse_prolong_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n_rows, ext(h)->d_agg, h->d_dir_flag, ext(h)->d_Aci, d_rc, d_z);
    h->launches++;
    GMPNP_CUDA_TRY(h, cudaGetLastError());
    return GMPNP_OK;
}

int gmpnp_median_3d(gmpnp_handle* h, const double* d_u, int comp, double* d_med, void* stream) {
    if (!h || h->dim != 3 || !d_u || !d_med || comp < 0 || comp >= NC) return GMPNP_ERR_ARG;
    int np2 = 1;
    while (np2 < h->n_nodes) np2 <<= 1;
    GMPNP_CUDA_TRY(h, cudaSetDevice(h->device));
    if ((size_t)np2 * sizeof(double) > MEDIAN_SORT_BYTES) {               // beyond the shared-memory sort: radix select
        median_select_kernel<<<dim3(h->batch, 1), 1024, 0, (cudaStream_t)stream>>>(h->n_nodes, 1, comp, comp, comp, comp,
                                                                                 d_u, d_med);
    } else {
        cudaFuncSetAttribute(median_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(np2 * sizeof(double)));
        median_kernel<<<h->batch, 1024, np2 * sizeof(double), (cudaStream_t)stream>>>(h->n_nodes, np2, comp, d_u, d_med);
    }
    h->launches++;
    GMPNP_CUDA_TRY(h, cudaGetLastError());
    return GMPNP_OK;
}

}  // extern "C"
