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
#include <map>
#include <mutex>
#include <numeric>
#include "common.cuh"

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

// FIAT default tetrahedron schemes (oracle/quadrature.py): degree 3 -> 5-point
// Zienkiewicz-Taylor rule for the residual, degree 4 -> 14-point Keast rule for the Jacobian.
// Barycentric coordinates (lambda0 = 1 - x - y - z) and weights summing to 1.
__constant__ double QF_L[5][4];
__constant__ double QF_W[5];
__constant__ double QJ_L[14][4];
__constant__ double QJ_W[14];

static void upload_rules() {
    static bool done = false;
    if (done) return;
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
    done = true;
}

// ---------------------------------------------------------------------------------------
// Kernel A: per (problem, tet) quadrature moments (14-point rule) and element residual (5-point)
// ---------------------------------------------------------------------------------------
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
    const double* up = u + (long)prob * n_vert * NC;
    const double* unp = un + (long)prob * n_vert * NC;
    int v[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) v[a] = tets[t * 4 + a];
    double g[4][3];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int d = 0; d < 3; ++d) g[a][d] = geom[(long)t * 13 + a * 3 + d];
    const double vol = geom[(long)t * 13 + 12];
    double U[4][NC];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int i = 0; i < NC; ++i) U[a][i] = up[(long)v[a] * NC + i];
    // gradients: G = sum_i nu_i grad u_i, gp = grad p
    double G[3] = {0, 0, 0}, gp[3] = {0, 0, 0};
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < NS; ++i) s += P[GMPNP_P_NU + i] * U[a][i];
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
                S += P[GMPNP_P_NU + i] * uq[i];
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
        double* mo = mom + ((long)prob * n_tet + t) * NMOM;
#pragma unroll
        for (int b = 0; b < 4; ++b) mo[M_MD + b] = mD[b];
#pragma unroll
        for (int i = 0; i < NS; ++i)
#pragma unroll
            for (int b = 0; b < 4; ++b) mo[M_UD2 + i * 4 + b] = m2[i][b];
#pragma unroll
        for (int i = 0; i < NS; ++i) mo[M_IUD + i] = iUD[i];
#pragma unroll
        for (int a = 0; a < 4; ++a) { mo[M_GA + a] = Ga[a]; mo[M_GPA + a] = gpa[a]; }
#pragma unroll
        for (int i = 0; i < NS; ++i) mo[M_SU + i] = SU[i];
        {
            const double wm = (P[GMPNP_P_EPSC] * SU[NS - 1] + P[GMPNP_P_EPSH] * SU[0]) * 0.25;
            mo[M_EPS] = P[GMPNP_P_EPSW] * ((55.0 - wm) / 55.0) + 6.0 * (wm / 55.0);
        }
    }
    if (want_res) {
        // ---- element residual, 5-point rule (one negative weight) ---------------------------
        double sUD[NS], R[5][4];
#pragma unroll
        for (int i = 0; i < NS; ++i) sUD[i] = 0.0;
#pragma unroll
        for (int i = 0; i < 5; ++i) R[i][0] = R[i][1] = R[i][2] = R[i][3] = 0.0;
        const double kW = P[GMPNP_P_KW], kA = P[GMPNP_P_KA], kB = P[GMPNP_P_KB];
        const double kA2 = P[GMPNP_P_KA2], kB2 = P[GMPNP_P_KB2], kw1 = P[GMPNP_P_KW1];
        for (int q = 0; q < 5; ++q) {
            const double l[4] = {QF_L[q][0], QF_L[q][1], QF_L[q][2], QF_L[q][3]};
            const double W = QF_W[q] * vol;
            double uq[NS], S = 0.0;
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                uq[i] = l[0] * U[0][i] + l[1] * U[1][i] + l[2] * U[2][i] + l[3] * U[3][i];
                S += P[GMPNP_P_NU + i] * uq[i];
            }
            const double WD = W / (1.0 - S);
#pragma unroll
            for (int i = 0; i < NS; ++i) sUD[i] += WD * uq[i];
            const double w = kW * uq[0] * uq[1], a = kA * uq[1] * uq[2], b = kB * uq[4] * uq[1];
            const double a2 = kA2 * uq[3], b2 = kB2 * uq[2];
            double mr[5];
            mr[0] = P[GMPNP_P_S] * (w - kw1);
            mr[1] = P[GMPNP_P_S + 1] * (w + a + b - kw1 - a2 - b2);
            mr[2] = P[GMPNP_P_S + 2] * (a + b2 - a2 - b);
            mr[3] = P[GMPNP_P_S + 3] * (a2 - a);
            mr[4] = P[GMPNP_P_S + 4] * (b - b2);
#pragma unroll
            for (int i = 0; i < 5; ++i)
#pragma unroll
                for (int aa = 0; aa < 4; ++aa) R[i][aa] += W * l[aa] * mr[i];
        }
        const double kappa = P[GMPNP_P_KAPPA];
        double* fe = Fe + ((long)prob * n_tet + t) * 36;
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
                for (int a = 0; a < 4; ++a) { dn[a] = U[a][i] - unp[(long)v[a] * NC + i]; dsum += dn[a]; }
            } else {
#pragma unroll
                for (int a = 0; a < 4; ++a) dn[a] = 0.0;
            }
            const double zi = P[GMPNP_P_Z + i];
            const double iU = 0.25 * vol * SU[i];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                double f = kappa * (vol * 0.05) * (dsum + dn[a]);
                f += K[a][0] * U[0][i] + K[a][1] * U[1][i] + K[a][2] * U[2][i] + K[a][3] * U[3][i];
                f += zi * gpa[a] * iU + Ga[a] * sUD[i];
                if (i < 5) f += R[i][a];
                fe[a * NC + i] = f;
                rho[a] += P[GMPNP_P_ZC0 + i] * U[a][i];
            }
        }
        const double wm = (P[GMPNP_P_EPSC] * SU[NS - 1] + P[GMPNP_P_EPSH] * SU[0]) * 0.25;
        const double epsm = P[GMPNP_P_EPSW] * ((55.0 - wm) / 55.0) + 6.0 * (wm / 55.0);
        const double rs = rho[0] + rho[1] + rho[2] + rho[3];
#pragma unroll
        for (int a = 0; a < 4; ++a)
            fe[a * NC + NS] = -gpa[a] * vol * epsm + P[GMPNP_P_Q] * (vol * 0.05) * (rs + rho[a]);
    }
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
// Coarse space (piecewise constants per z-slab and component): A_c = P^T J P, explicit inverse
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
coarse_setup_kernel(int n_vert, int n_blocks, const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                    const int* __restrict__ agg, const int* __restrict__ dir_flag, const double* __restrict__ J,
                    double* __restrict__ Aci) {
    extern __shared__ double Ac[];          // [NCO][NCO]
    const int prob = blockIdx.x;
    for (int i = threadIdx.x; i < NCO * NCO; i += blockDim.x) Ac[i] = 0.0;
    __syncthreads();
    const double* Jp = J + (long)prob * n_blocks * 81;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int row = w; row < n_vert; row += nw) {
        const int I = agg[row];
        for (int s = row_ptr[row]; s < row_ptr[row + 1]; ++s) {
            const int col = col_idx[s];
            const int Jc = agg[col];
            for (int e = lane; e < 81; e += 32) {
                const int i = e / 9, j = e % 9;
                if (dir_flag[(long)row * NC + i] >= 0 || dir_flag[(long)col * NC + j] >= 0) continue;
                atomicAdd(&Ac[(I * NC + i) * NCO + Jc * NC + j], Jp[(long)s * 81 + e]);
            }
        }
    }
    __syncthreads();
    // empty coarse columns (all members Dirichlet) -> identity
    for (int i = threadIdx.x; i < NCO; i += blockDim.x)
        if (Ac[i * NCO + i] == 0.0) Ac[i * NCO + i] = 1.0;
    __syncthreads();
    // in-place Gauss-Jordan inversion without pivoting on the (diagonally dominant-ish) Galerkin matrix;
    // a vanishing pivot is replaced to keep the preconditioner finite
    __shared__ double pivinv;
    for (int c = 0; c < NCO; ++c) {
        if (threadIdx.x == 0) {
            double p = Ac[c * NCO + c];
            if (!(fabs(p) > 1e-300)) p = 1.0;
            pivinv = 1.0 / p;
        }
        __syncthreads();
        const double pi = pivinv;
        // scale pivot row (except pivot), pivot element becomes 1/p
        for (int j = threadIdx.x; j < NCO; j += blockDim.x)
            if (j != c) Ac[c * NCO + j] *= pi;
        __syncthreads();
        for (int idx = threadIdx.x; idx < NCO * NCO; idx += blockDim.x) {
            const int r = idx / NCO, j = idx % NCO;
            if (r == c || j == c) continue;
            Ac[idx] -= Ac[r * NCO + c] * Ac[c * NCO + j];
        }
        __syncthreads();
        for (int r = threadIdx.x; r < NCO; r += blockDim.x)
            Ac[r * NCO + c] = (r == c) ? pi : -Ac[r * NCO + c] * pi;
        __syncthreads();
    }
    for (int i = threadIdx.x; i < NCO * NCO; i += blockDim.x) Aci[(long)prob * NCO * NCO + i] = Ac[i];
}

// yc = A_c^{-1} P^T r   (one CTA per problem; deterministic)
__global__ void __launch_bounds__(NCO)
coarse_solve_kernel(int n_vert, const int* __restrict__ agg_ptr, const int* __restrict__ agg_nodes,
                    const int* __restrict__ dir_flag, const double* __restrict__ Aci, const double* __restrict__ r,
                    long rstride, double* __restrict__ yc) {
    __shared__ double rc[NCO];
    const int prob = blockIdx.x;
    const int t = threadIdx.x, I = t / NC, i = t % NC;
    const double* rp = r + (long)prob * rstride;
    double s = 0.0;
    for (int k = agg_ptr[I]; k < agg_ptr[I + 1]; ++k) {
        const int v = agg_nodes[k];
        if (dir_flag[(long)v * NC + i] < 0) s += rp[(long)v * NC + i];
    }
    rc[t] = s;
    __syncthreads();
    const double* A = Aci + (long)prob * NCO * NCO + (long)t * NCO;
    double y = 0.0;
    for (int j = 0; j < NCO; ++j) y += A[j] * rc[j];
    yc[(long)prob * NCO + t] = y;
}

// z = D^{-1} r + P yc
__global__ void precond_apply_kernel(int n_vert, const int* __restrict__ agg, const int* __restrict__ dir_flag,
                                     const double* __restrict__ Dinv, const double* __restrict__ yc,
                                     const double* __restrict__ r, long rstride, double* __restrict__ z,
                                     int use_coarse) {
    const int prob = blockIdx.y;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)n_vert * NC) return;
    const int v = (int)(idx / NC), i = (int)(idx % NC);
    const double* D = Dinv + ((long)prob * n_vert + v) * 81 + i * 9;
    const double* rp = r + (long)prob * rstride + (long)v * NC;
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 9; ++j) s += D[j] * rp[j];
    if (use_coarse && dir_flag[idx] < 0) s += yc[(long)prob * NCO + agg[v] * NC + i];
    z[(long)prob * n_vert * NC + idx] = s;
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

// dots[prob][k] = V_k . w for k < nvec   (grid = (nvec, batch))
__global__ void __launch_bounds__(256)
multi_dot_kernel(long n, long vstride, const double* __restrict__ V, const double* __restrict__ w,
                 double* __restrict__ dots, int ld) {
    __shared__ double sh[8];
    const int k = blockIdx.x, prob = blockIdx.y;
    const double* vk = V + ((long)prob * ld + k) * vstride;
    const double* wp = w + (long)prob * n;
    double s = 0.0;
    for (long i = threadIdx.x; i < n; i += blockDim.x) s += vk[i] * wp[i];
    s = block_sum(s, sh);
    if (threadIdx.x == 0) dots[(long)prob * ld + k] = s;
}

// w -= sum_k dots[k] V_k ; optionally hacc[k] += dots[k]
__global__ void gs_update_kernel(long n, long vstride, int nvec, const double* __restrict__ V,
                                 const double* __restrict__ dots, double* __restrict__ w, int ld) {
    const int prob = blockIdx.y;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = w[(long)prob * n + i];
    for (int k = 0; k < nvec; ++k) s -= dots[(long)prob * ld + k] * V[((long)prob * ld + k) * vstride + i];
    w[(long)prob * n + i] = s;
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

// Hessenberg column j: h = h1 + h2 (CGS2), h[j+1] = nrm; apply stored Givens, create new one, update g.
// state per problem: H[(m+1) x m] column-major, cs[m], sn[m], gvec[m+1], jdone, beta
__global__ void givens_kernel(int batch, int j, int m, const double* __restrict__ d1, const double* __restrict__ d2,
                              const double* __restrict__ nrm, double* __restrict__ H, double* __restrict__ cs,
                              double* __restrict__ sn, double* __restrict__ gvec, int* __restrict__ jdone,
                              const double* __restrict__ tol_abs, int ld) {
    const int prob = blockIdx.x * blockDim.x + threadIdx.x;
    if (prob >= batch) return;
    if (jdone[prob] >= 0) return;                 // this problem already converged in this cycle
    double* h = H + ((long)prob * m + j) * (m + 1);
    for (int k = 0; k <= j; ++k) h[k] = d1[(long)prob * ld + k] + d2[(long)prob * ld + k];
    h[j + 1] = nrm[prob];
    double* c = cs + (long)prob * m;
    double* s = sn + (long)prob * m;
    for (int k = 0; k < j; ++k) {
        const double t = c[k] * h[k] + s[k] * h[k + 1];
        h[k + 1] = -s[k] * h[k] + c[k] * h[k + 1];
        h[k] = t;
    }
    const double a = h[j], b = h[j + 1];
    const double d = hypot(a, b);
    const double cj = (d > 0.0) ? a / d : 1.0, sj = (d > 0.0) ? b / d : 0.0;
    c[j] = cj; s[j] = sj;
    h[j] = d; h[j + 1] = 0.0;
    double* g = gvec + (long)prob * (m + 1);
    g[j + 1] = -sj * g[j];
    g[j] = cj * g[j];
    if (fabs(g[j + 1]) <= tol_abs[prob] || !(nrm[prob] > 0.0)) jdone[prob] = j + 1;
}

// y = H^{-1} g (upper triangular, first jd columns); coef[prob][k] = y_k (0 beyond jd)
__global__ void hsolve_kernel(int batch, int m, const double* __restrict__ H, const double* __restrict__ gvec,
                              const int* __restrict__ jdone, double* __restrict__ coef, int ld) {
    const int prob = blockIdx.x * blockDim.x + threadIdx.x;
    if (prob >= batch) return;
    const int jd = (jdone[prob] >= 0) ? jdone[prob] : m;
    double* y = coef + (long)prob * ld;
    const double* g = gvec + (long)prob * (m + 1);
    for (int k = 0; k < ld; ++k) y[k] = 0.0;
    for (int k = jd - 1; k >= 0; --k) {
        double s = g[k];
        for (int l = k + 1; l < jd; ++l) s -= H[((long)prob * m + l) * (m + 1) + k] * y[l];
        const double d = H[((long)prob * m + k) * (m + 1) + k];
        y[k] = (d != 0.0) ? s / d : 0.0;
    }
}

// out = sum_k coef[k] V_k
__global__ void lincomb_kernel(long n, long vstride, int nvec, const double* __restrict__ V,
                               const double* __restrict__ coef, double* __restrict__ out, int ld) {
    const int prob = blockIdx.y;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int k = 0; k < nvec; ++k) s += coef[(long)prob * ld + k] * V[((long)prob * ld + k) * vstride + i];
    out[(long)prob * n + i] = s;
}

// y = a*x + b*y (per problem scalars optional)
__global__ void axpby_kernel(long n, double a, const double* __restrict__ x, double b, double* __restrict__ y) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x + (long)blockIdx.y * n;
    if ((long)blockIdx.x * blockDim.x + threadIdx.x >= n) return;
    y[i] = a * x[i] + b * y[i];
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

// Aci = inverse of the (all-reduced) Galerkin matrix; empty coarse columns (all members Dirichlet) -> identity
__global__ void __launch_bounds__(1024)
coarse_invert_kernel(const double* __restrict__ Ac_in, double* __restrict__ Aci) {
    extern __shared__ double Ac[];          // [NCO][NCO]
    for (int i = threadIdx.x; i < NCO * NCO; i += blockDim.x) Ac[i] = Ac_in[i];
    __syncthreads();
    for (int i = threadIdx.x; i < NCO; i += blockDim.x)
        if (Ac[i * NCO + i] == 0.0) Ac[i * NCO + i] = 1.0;
    __syncthreads();
    __shared__ double pivinv;
    for (int c = 0; c < NCO; ++c) {
        if (threadIdx.x == 0) {
            double p = Ac[c * NCO + c];
            if (!(fabs(p) > 1e-300)) p = 1.0;
            pivinv = 1.0 / p;
        }
        __syncthreads();
        const double pi = pivinv;
        for (int j = threadIdx.x; j < NCO; j += blockDim.x)
            if (j != c) Ac[c * NCO + j] *= pi;
        __syncthreads();
        for (int idx = threadIdx.x; idx < NCO * NCO; idx += blockDim.x) {
            const int r = idx / NCO, j = idx % NCO;
            if (r == c || j == c) continue;
            Ac[idx] -= Ac[r * NCO + c] * Ac[c * NCO + j];
        }
        __syncthreads();
        for (int r = threadIdx.x; r < NCO; r += blockDim.x)
            Ac[r * NCO + c] = (r == c) ? pi : -Ac[r * NCO + c] * pi;
        __syncthreads();
    }
    for (int i = threadIdx.x; i < NCO * NCO; i += blockDim.x) Aci[i] = Ac[i];
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
    int restart_alloc = 0;
    double *d_V = nullptr, *d_w = nullptr, *d_z = nullptr, *d_dx = nullptr, *d_d1 = nullptr, *d_d2 = nullptr;
    double *d_nrm = nullptr, *d_H = nullptr, *d_cs = nullptr, *d_sn = nullptr, *d_g = nullptr, *d_tol = nullptr;
    double *d_coef = nullptr, *d_beta = nullptr, *d_dxmax = nullptr, *d_umax = nullptr;
    int *d_jdone = nullptr, *d_active = nullptr;
};
// side table keyed by handle (handles are independent; the table itself is guarded so that different handles can be
// created, used and destroyed from different threads)
static std::map<gmpnp_handle*, Host3D*> g_ext;
static std::mutex g_ext_mutex;

static Host3D* ext(gmpnp_handle* h) {
    std::lock_guard<std::mutex> lk(g_ext_mutex);
    auto it = g_ext.find(h);
    return it == g_ext.end() ? nullptr : it->second;
}

void pore3d_free_ext(gmpnp_handle* h) {
    Host3D* e = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_ext_mutex);
        auto it = g_ext.find(h);
        if (it == g_ext.end()) return;
        e = it->second;
        g_ext.erase(it);
    }
    void* bufs[] = {e->d_wall_w, e->d_exit_m, e->d_exit_flag, e->d_bc, e->d_mass, e->d_gp, e->d_partial, e->d_blk_row, e->d_blk_geo, e->d_agg, e->d_agg_ptr, e->d_agg_nodes, e->d_Aci, e->d_yc, e->d_V, e->d_w, e->d_z,
                    e->d_dx, e->d_d1, e->d_d2, e->d_nrm, e->d_H, e->d_cs, e->d_sn, e->d_g, e->d_tol, e->d_coef,
                    e->d_beta, e->d_dxmax, e->d_umax, e->d_jdone, e->d_active};
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
    {
        std::lock_guard<std::mutex> lk(g_ext_mutex);
        g_ext[h] = e;
    }
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
    GMPNP_CUDA_TRY(h, cudaMalloc(&h->d_params, sizeof(double) * GMPNP_NPAR * B));
    GMPNP_CUDA_TRY(h, cudaMalloc(&h->d_dir_val, sizeof(double) * std::max(1, n_dir) * B));
    GMPNP_CUDA_TRY(h, cudaMalloc(&h->d_mom, sizeof(double) * NMOM * (size_t)n_tet * B));
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
    const int T = h->n_tet, V = h->n_nodes, B = h->batch;
    dim3 gA((T + 127) / 128, B);
    tet_moments_kernel<<<gA, 128, 0, st>>>(T, V, h->d_tets, h->d_geom, h->d_params, d_u, d_un, h->d_mom, h->d_Fe,
                                           d_J != nullptr, d_F != nullptr);
    h->launches++;
    if (d_F) {
        dim3 gB(((long)V * NC + 255) / 256, B);
        residual_gather_kernel<<<gB, 256, 0, st>>>(V, T, h->n_dir, h->d_node_ptr, h->d_node_src, h->d_dir_flag,
                                                   h->d_dir_val, h->d_Fe, d_u, d_F,
                                                   ext(h)->facet_terms ? ext(h)->d_wall_w : nullptr, ext(h)->d_exit_m,
                                                   ext(h)->d_exit_flag, ext(h)->d_bc, h->d_row_ptr, h->d_col_idx);
        h->launches++;
    }
    if (d_J) {
        int gx = std::min((h->n_blocks + ASM_WARPS - 1) / ASM_WARPS, 148 * 16);
        dim3 gC(gx, B);
        assemble_bsr_kernel<<<gC, ASM_WARPS * 32, 0, st>>>(h->n_blocks, V, T, h->d_blk_ptr, h->d_blk_src,
                                                           ext(h)->d_blk_geo, ext(h)->d_blk_row, h->d_col_idx,
                                                           h->d_dir_flag, h->d_params, d_u, h->d_mom, d_J,
                                                           ext(h)->facet_terms ? ext(h)->d_exit_m : nullptr, ext(h)->d_bc);
        h->launches++;
    }
    GMPNP_CUDA_TRY(h, cudaGetLastError());
    return GMPNP_OK;
}

static int launch_spmv(gmpnp_handle* h, const double* d_J, const double* d_x, double* d_y, cudaStream_t st,
                       int row0 = 0, int row1 = -1) {
    if (row1 < 0) row1 = h->n_nodes;
    if (row1 <= row0) return GMPNP_OK;
    dim3 g((row1 - row0 + 7) / 8, h->batch);
    bsr_spmv_kernel<<<g, 256, 0, st>>>(h->n_nodes, h->n_blocks, h->d_row_ptr, h->d_col_idx, d_J, d_x, d_y, row0, row1);
    h->launches++;
    return GMPNP_OK;
}

static int ensure_krylov(gmpnp_handle* h, int m) {
    Host3D* e = ext(h);
    if (e->restart_alloc >= m) return GMPNP_OK;
    void* old[] = {e->d_V, e->d_w, e->d_z, e->d_dx, e->d_d1, e->d_d2, e->d_nrm, e->d_H, e->d_cs, e->d_sn, e->d_g,
                   e->d_tol, e->d_coef, e->d_beta, e->d_dxmax, e->d_umax, e->d_jdone, e->d_active};
    for (void* b : old) if (b) cudaFree(b);
    const size_t B = h->batch, n = (size_t)h->n_nodes * NC, ld = m + 1;
    GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_V, sizeof(double) * B * ld * n));
    GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_w, sizeof(double) * B * n));
    GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_z, sizeof(double) * B * n));
    GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_dx, sizeof(double) * B * n));
    GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_d1, sizeof(double) * B * ld));
    GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_d2, sizeof(double) * B * ld));
    GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_coef, sizeof(double) * B * ld));
    GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_nrm, sizeof(double) * B));
    GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_H, sizeof(double) * B * m * ld));
    GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_cs, sizeof(double) * B * m));
    GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_sn, sizeof(double) * B * m));
    GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_g, sizeof(double) * B * ld));
    GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_tol, sizeof(double) * B));
    GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_beta, sizeof(double) * B));
    GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_dxmax, sizeof(double) * B));
    GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_umax, sizeof(double) * B));
    GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_jdone, sizeof(int) * B));
    GMPNP_CUDA_TRY(h, cudaMalloc(&e->d_active, sizeof(int) * B));
    e->restart_alloc = m;
    return GMPNP_OK;
}

// z = M^{-1} r  (block-Jacobi + additive z-slab coarse correction)
static void launch_precond(gmpnp_handle* h, const double* r, long rstride, double* z, cudaStream_t st) {
    Host3D* e = ext(h);
    const int V = h->n_nodes, B = h->batch;
    coarse_solve_kernel<<<B, NCO, 0, st>>>(V, e->d_agg_ptr, e->d_agg_nodes, h->d_dir_flag, e->d_Aci, r, rstride,
                                           e->d_yc);
    dim3 g(((long)V * NC + 255) / 256, B);
    precond_apply_kernel<<<g, 256, 0, st>>>(V, e->d_agg, h->d_dir_flag, h->d_Dinv, e->d_yc, r, rstride, z, 1);
    h->launches += 2;
}

// Right-preconditioned restarted GMRES on J dx = F for the whole batch.  Returns per-problem iteration
// counts and the final relative residual estimates in host arrays.
static int gmres_solve(gmpnp_handle* h, const double* d_b, double* d_x, int m, int maxit, double rtol,
                       std::vector<int>& its, std::vector<double>& relres, cudaStream_t st) {
    Host3D* e = ext(h);
    const int V = h->n_nodes, B = h->batch;
    const long n = (long)V * NC;
    const int ld = m + 1;
    int rc = ensure_krylov(h, m); if (rc) return rc;
    double* hp = (double*)h->h_pinned;
    const int gridn = (int)((n + 255) / 256);
    its.assign(B, 0); relres.assign(B, 0.0);
    GMPNP_CUDA_TRY(h, cudaMemsetAsync(d_x, 0, sizeof(double) * n * B, st));
    // r0 = b (x0 = 0)
    norm_scale_kernel<<<B, 1024, 0, st>>>(n, d_b, nullptr, 0, e->d_beta);
    h->launches++;
    GMPNP_CUDA_TRY(h, cudaMemcpyAsync(hp, e->d_beta, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
    GMPNP_CUDA_TRY(h, cudaStreamSynchronize(st));
    std::vector<double> beta0(hp, hp + B), tol(B);
    for (int p = 0; p < B; ++p) tol[p] = rtol * beta0[p];
    GMPNP_CUDA_TRY(h, cudaMemcpyAsync(e->d_tol, tol.data(), sizeof(double) * B, cudaMemcpyHostToDevice, st));
    const double* rcur = d_b;
    int total = 0;
    std::vector<char> done(B, 0);
    for (int p = 0; p < B; ++p) if (!(beta0[p] > 0.0)) done[p] = 1;
    std::vector<double> g0((size_t)B * ld);
    std::vector<int> jd(B);
    while (total < maxit) {
        // cycle start: V_0 = r / ||r||, g = (||r||, 0, ...)
        norm_scale_kernel<<<B, 1024, 0, st>>>(n, rcur, e->d_V, (long)ld * n, e->d_beta);
        h->launches++;
        GMPNP_CUDA_TRY(h, cudaMemcpyAsync(hp, e->d_beta, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
        GMPNP_CUDA_TRY(h, cudaStreamSynchronize(st));
        std::fill(g0.begin(), g0.end(), 0.0);
        bool all = true;
        for (int p = 0; p < B; ++p) {
            g0[(size_t)p * ld] = hp[p];
            relres[p] = beta0[p] > 0 ? hp[p] / beta0[p] : 0.0;
            jd[p] = -1;
            if (done[p] || !(hp[p] > tol[p])) { done[p] = 1; jd[p] = 0; }
            else all = false;
        }
        if (all) break;
        GMPNP_CUDA_TRY(h, cudaMemcpyAsync(e->d_g, g0.data(), sizeof(double) * B * ld, cudaMemcpyHostToDevice, st));
        GMPNP_CUDA_TRY(h, cudaMemcpyAsync(e->d_jdone, jd.data(), sizeof(int) * B, cudaMemcpyHostToDevice, st));
        const int mcyc = std::min(m, maxit - total);
        for (int j = 0; j < mcyc; ++j) {
            // w = J M^{-1} v_j
            launch_precond(h, e->d_V + (long)j * n, (long)ld * n, e->d_z, st);
            launch_spmv(h, h->d_J, e->d_z, e->d_w, st);
            // classical Gram-Schmidt, twice (CGS2): two reductions per step, deterministic
            dim3 gd(j + 1, B), gu(gridn, B);
            multi_dot_kernel<<<gd, 256, 0, st>>>(n, n, e->d_V, e->d_w, e->d_d1, ld);
            gs_update_kernel<<<gu, 256, 0, st>>>(n, n, j + 1, e->d_V, e->d_d1, e->d_w, ld);
            multi_dot_kernel<<<gd, 256, 0, st>>>(n, n, e->d_V, e->d_w, e->d_d2, ld);
            gs_update_kernel<<<gu, 256, 0, st>>>(n, n, j + 1, e->d_V, e->d_d2, e->d_w, ld);
            norm_scale_kernel<<<B, 1024, 0, st>>>(n, e->d_w, e->d_V + (long)(j + 1) * n, (long)ld * n, e->d_nrm);
            givens_kernel<<<(B + 63) / 64, 64, 0, st>>>(B, j, m, e->d_d1, e->d_d2, e->d_nrm, e->d_H, e->d_cs, e->d_sn,
                                                        e->d_g, e->d_jdone, e->d_tol, ld);
            h->launches += 6;
        }
        total += mcyc;
        // x += M^{-1} (V y)
        hsolve_kernel<<<(B + 63) / 64, 64, 0, st>>>(B, m, e->d_H, e->d_g, e->d_jdone, e->d_coef, ld);
        dim3 gu(gridn, B);
        lincomb_kernel<<<gu, 256, 0, st>>>(n, n, mcyc, e->d_V, e->d_coef, e->d_w, ld);
        launch_precond(h, e->d_w, n, e->d_z, st);
        axpby_kernel<<<gu, 256, 0, st>>>(n, 1.0, e->d_z, 1.0, d_x);
        // true residual r = b - J x  -> e->d_w
        launch_spmv(h, h->d_J, d_x, e->d_w, st);
        axpby_kernel<<<gu, 256, 0, st>>>(n, 1.0, d_b, -1.0, e->d_w);
        h->launches += 5;
        GMPNP_CUDA_TRY(h, cudaMemcpyAsync(jd.data(), e->d_jdone, sizeof(int) * B, cudaMemcpyDeviceToHost, st));
        GMPNP_CUDA_TRY(h, cudaStreamSynchronize(st));
        for (int p = 0; p < B; ++p)
            if (!done[p]) its[p] += (jd[p] >= 0) ? jd[p] : mcyc;
        rcur = e->d_w;
    }
    // final residual check
    norm_scale_kernel<<<B, 1024, 0, st>>>(n, rcur, nullptr, 0, e->d_beta);
    h->launches++;
    GMPNP_CUDA_TRY(h, cudaMemcpyAsync(hp, e->d_beta, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
    GMPNP_CUDA_TRY(h, cudaStreamSynchronize(st));
    for (int p = 0; p < B; ++p) relres[p] = beta0[p] > 0 ? hp[p] / beta0[p] : 0.0;
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
    Host3D* e = ext(h);
    const int V = h->n_nodes, B = h->batch;
    const long n = (long)V * NC;
    const int m = o->lin_restart > 0 ? o->lin_restart : 50;
    const int lin_maxit = o->lin_maxit > 0 ? o->lin_maxit : 1000;
    rc = ensure_krylov(h, m); if (rc) return rc;
    double* hp = (double*)h->h_pinned;
    std::vector<int> iters(B, 0), status(B, -1), lin_total(B, 0), active(B, 1);
    std::vector<double> r0(B, 0.0), r(B, 0.0);
    cudaFuncSetAttribute(coarse_setup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * NCO * NCO));
    auto residual_norms = [&](std::vector<double>& out) -> int {
        int rc2 = launch_assemble(h, d_u, d_un, h->d_F, nullptr, st); if (rc2) return rc2;
        norm_scale_kernel<<<B, 1024, 0, st>>>(n, h->d_F, nullptr, 0, e->d_nrm);
        h->launches++;
        GMPNP_CUDA_TRY(h, cudaMemcpyAsync(hp, e->d_nrm, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
        GMPNP_CUDA_TRY(h, cudaStreamSynchronize(st));
        out.assign(hp, hp + B);
        return GMPNP_OK;
    };
    rc = residual_norms(r0); if (rc) return rc;
    r = r0;
    for (int p = 0; p < B; ++p) {
        if (!std::isfinite(r0[p])) { status[p] = GMPNP_NOT_FINITE; active[p] = 0; }
        else if (o->criterion == 0 && r0[p] < o->atol) { status[p] = GMPNP_CONVERGED; active[p] = 0; }
    }
    for (int k = 0; k < o->maxit; ++k) {
        bool any = false;
        for (int p = 0; p < B; ++p) any |= (active[p] != 0);
        if (!any) break;
        // Jacobian at the current iterate, preconditioner setup
        rc = launch_assemble(h, d_u, d_un, nullptr, h->d_J, st); if (rc) return rc;
        dim3 gj((V + 63) / 64, B);
        bjacobi_invert_kernel<<<gj, 64, 0, st>>>(V, h->n_blocks, h->d_diag_idx, h->d_J, h->d_Dinv);
        coarse_setup_kernel<<<B, 1024, sizeof(double) * NCO * NCO, st>>>(V, h->n_blocks, h->d_row_ptr, h->d_col_idx,
                                                                         e->d_agg, h->d_dir_flag, h->d_J, e->d_Aci);
        h->launches += 2;
        std::vector<int> lits; std::vector<double> rel;
        rc = gmres_solve(h, h->d_F, e->d_dx, m, lin_maxit, o->lin_rtol > 0 ? o->lin_rtol : 1e-10, lits, rel, st);
        if (rc) return rc;
        GMPNP_CUDA_TRY(h, cudaMemcpyAsync(e->d_active, active.data(), sizeof(int) * B, cudaMemcpyHostToDevice, st));
        newton_update_kernel<<<B, 1024, 0, st>>>(n, o->relax, e->d_active, e->d_dx, d_u, e->d_dxmax, e->d_umax);
        h->launches++;
        GMPNP_CUDA_TRY(h, cudaMemcpyAsync(hp, e->d_dxmax, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
        GMPNP_CUDA_TRY(h, cudaMemcpyAsync(hp + B, e->d_umax, sizeof(double) * B, cudaMemcpyDeviceToHost, st));
        GMPNP_CUDA_TRY(h, cudaStreamSynchronize(st));
        std::vector<double> dxm(hp, hp + B), um(hp + B, hp + 2 * B);
        std::vector<double> rn;
        rc = residual_norms(rn); if (rc) return rc;
        for (int p = 0; p < B; ++p) {
            if (!active[p]) continue;
            iters[p] = k + 1;
            lin_total[p] += lits[p];
            r[p] = rn[p];
            const bool linfail = rel[p] > 1e3 * (o->lin_rtol > 0 ? o->lin_rtol : 1e-10);
            if (!std::isfinite(rn[p]) || !std::isfinite(dxm[p])) { status[p] = GMPNP_NOT_FINITE; active[p] = 0; continue; }
            bool conv;
            if (o->criterion == 0) conv = (rn[p] / r0[p] < o->rtol) || (rn[p] < o->atol);
            else conv = dxm[p] <= o->xtol * std::max(1.0, um[p]);
            if (conv) { status[p] = GMPNP_CONVERGED; active[p] = 0; }
            else if (linfail) { status[p] = GMPNP_LINEAR_FAILED; active[p] = 0; }
        }
    }
    for (int p = 0; p < B; ++p) if (status[p] < 0) status[p] = GMPNP_MAXIT;
    if (d_iters) GMPNP_CUDA_TRY(h, cudaMemcpyAsync(d_iters, iters.data(), sizeof(int) * B, cudaMemcpyHostToDevice, st));
    if (d_status) GMPNP_CUDA_TRY(h, cudaMemcpyAsync(d_status, status.data(), sizeof(int) * B, cudaMemcpyHostToDevice, st));
    if (d_lin_iters) GMPNP_CUDA_TRY(h, cudaMemcpyAsync(d_lin_iters, lin_total.data(), sizeof(int) * B, cudaMemcpyHostToDevice, st));
    if (d_r0) GMPNP_CUDA_TRY(h, cudaMemcpyAsync(d_r0, r0.data(), sizeof(double) * B, cudaMemcpyHostToDevice, st));
    if (d_r) GMPNP_CUDA_TRY(h, cudaMemcpyAsync(d_r, r.data(), sizeof(double) * B, cudaMemcpyHostToDevice, st));
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
    coarse_invert_kernel<<<1, 1024, smem, (cudaStream_t)stream>>>(d_Ac, ext(h)->d_Aci);
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
    if ((size_t)np2 * sizeof(double) > 200 * 1024) return GMPNP_ERR_ARG;      // shared-memory sort only
    GMPNP_CUDA_TRY(h, cudaSetDevice(h->device));
    cudaFuncSetAttribute(median_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(np2 * sizeof(double)));
    median_kernel<<<h->batch, 1024, np2 * sizeof(double), (cudaStream_t)stream>>>(h->n_nodes, np2, comp, d_u, d_med);
    h->launches++;
    GMPNP_CUDA_TRY(h, cudaGetLastError());
    return GMPNP_OK;
}

}  // extern "C"
