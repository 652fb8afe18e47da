// 1D planar EDL: fused P1 assembly + block-tridiagonal (7x7) Thomas elimination + Newton, whole Newton /
// time-march / continuation loop device-resident.  A problem is handled by FOUR 8-lane groups: two-sided
// ("twisted") elimination -- top half downwards, bottom half upwards -- and, per half, a PRODUCER group that
// integrates the cells and assembles block row k+1 while the CONSUMER group eliminates block row k
// (warp-specialised: producers and consumers live in different warps of the CTA and hand block rows over through a
// three-slot shared-memory queue guarded by named barriers).
//
// Replaces the FEniCS work behind `solve(F + J_OH*v_OH*ds + J_H*v_H*ds == 0, u, bcs)`
// (1D/MPNP_CO2ER_EDL.py:737-742): FFC element kernels for the forms 1D:381-595, dolfin's
// SystemAssembler scatter, DirichletBC.apply (1D:350-355), the UMFPACK LU and dolfin's
// NewtonSolver loop.
//
// Data layout (HBM):
//   u, u_n          [problem][node][7]            node-major interleaved (56 B per node)
//   workspace       [problem][node][7][8]         row j = (C'_k[j][0..6], d'_k[j]): the
//                                                 eliminated super-diagonal block and rhs
// Lane mapping inside a group: lane c<7 owns COLUMN c of every 7x7 block (trial component c),
// lane 7 owns the right-hand side / residual.  The Jacobian never exists in HBM: row k is
// assembled from cells k-1,k just in time, eliminated, and only (C'_k, d'_k) is stored.
#include "common.cuh"

namespace edl1d {

constexpr int NS = 6;
constexpr int NC = 7;
// Shared memory of newton1d_kernel, per 8-lane group (doubles):
// Block-row queue between a producer group and its consumer group: NSLOT slots of
//   A[7][8] (sub-diagonal block = block (1,0) of the cell behind, row-major: A[i*8+c]; its unused column 7 carries
//   the right-hand side d[i] = residual row i of the node), B[7][8] (diagonal block), C[7][8] (coupling block ahead;
//   column 7 of row 0 carries the half's residual sum of squares in the closing slot)
// Slot r % NSLOT carries block row r; while integrating the cell ahead of row r the producer already deposits that
// cell's (1,0)/(1,1) blocks in slot (r+1) % NSLOT, so three slots keep producer and consumer one row apart.
constexpr int NSLOT = 3;
constexpr int Q_A = 0, Q_B = 56, Q_C = 112, Q_SLOT = 168;
__device__ __forceinline__ constexpr int Q_D(int i) { return Q_A + i * 8 + 7; }      // d[i] rides in column 7 of A
constexpr int SM_Q = 0;                     // [NSLOT][Q_SLOT]
constexpr int SM_M = NSLOT * Q_SLOT;        // producer: residual rows of the current node [8]
constexpr int SM_X = SM_M + 8;              // consumer, back-substitution: solution of the neighbour row, double-buffered [2][8]
constexpr int SM_PC = SM_X + 16;            // consumer, Gauss-Jordan: pivot column broadcast, double-buffered [2][8]
constexpr int SM_GROUP = SM_PC + 16 + 8;    // doubles per group (== 8 mod 16: the two groups of a half-warp use
                                            // disjoint banks)
static_assert(SM_GROUP % 16 == 8, "bank layout");
// plus, per pair of groups (= problem): the parameter record [64]; the merge area of a pair aliases the
// staging ring of its bottom-half group (idle between the two sweeps).
// Shared memory of assemble1d_kernel, per group: params [64], staged nodal values [2][8], residual rows [8]
constexpr int AS_P = 0, AS_U = 64, AS_M = 80, AS_GROUP = 88;
// Multi-row staging ring (cp.async, RING rows ahead of the row being processed), per group:
//   forward sweep : fu[RING][8] = (u_0..u_6 of a node, 1.0), fn[RING][8] = (u_n,0..6 of the node, x of the node)
//   backward sweep: bw[RING][4][8] double2 = the lane's own 64-B workspace row, bu[RING][8] = the lane's u component
// The two sweeps alias the same memory.  One row of global-load latency (~1 us) is several backward rows long,
// and a register prefetch gets spilled by the 255-register forward body, so the rows are staged through
// shared memory instead of registers.
#ifndef GMPNP_RING
#define GMPNP_RING 4
#endif
constexpr int RING = GMPNP_RING;
constexpr int SM_RING = RING * 72 + 8;  // doubles per group (== 8 mod 16, see SM_GROUP)
constexpr int GROUPS_PER_BLOCK = 16;    // assemble1d_kernel: one group per (problem, node)
constexpr int THREADS = GROUPS_PER_BLOCK * 8;
// newton1d_kernel: 4 warps = 2 consumer (master) warps + 2 producer warps; a warp holds 2 problems x 2 halves
constexpr int NW_PAIRS = 2;                       // consumer/producer warp pairs per CTA
constexpr int NW_THREADS = NW_PAIRS * 2 * 32;
constexpr int NW_GROUPS = NW_PAIRS * 4;           // (problem, half) slots per CTA
constexpr int NW_PROBLEMS = NW_PAIRS * 2;
// named barriers (id 0 is __syncthreads): per warp pair full[NSLOT], empty[NSLOT], command
constexpr int BAR_PER_PAIR = 2 * NSLOT + 1;
static_assert(1 + NW_PAIRS * BAR_PER_PAIR <= 16, "named barriers");
constexpr int CMD_FACTOR = 1, CMD_EXIT = 2;

__device__ __forceinline__ void bar_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }

// 3-point and 2-point Gauss-Legendre on [0,1] (FFC: degree 4 -> 3 points for J, degree 3 ->
// 2 points for F)
__device__ __constant__ double GX3[3] = {0.11270166537925831, 0.5, 0.88729833462074169};
__device__ __constant__ double GW3[3] = {0.27777777777777779, 0.44444444444444442, 0.27777777777777779};
__device__ __constant__ double GX2[2] = {0.21132486540518713, 0.78867513459481287};
__device__ __constant__ double GW2[2] = {0.5, 0.5};

struct LaneConst {
    double coef[5];   // elementary-rate derivative coefficients of this lane's column (w,a,b,a2,b2)
    int sel[5];       // which staged nodal value multiplies it (7 = constant 1)
    double rsig[6];   // -R_c = rsig . (uH uOH, uOH uHCO3, uCO2 uOH, uCO32, uHCO3, 1): this lane's own reaction row
                      // of the residual, rate constants folded in
};

__device__ __forceinline__ void cp_async8(double* dst_smem, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src)
                 : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst_smem, const double* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 1/x to within ~1 ulp: hardware seed (>= 20 correct bits, PTX rcp.approx.ftz.f64) + two Newton steps
// (20 -> 40 -> 80 bits); the IEEE division sequence costs ~4x as many instructions, and pivots and steric
// denominators do not need correct rounding
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

__device__ __forceinline__ void lane_consts(const double* P, int c, LaneConst& L) {
#pragma unroll
    for (int r = 0; r < 5; ++r) { L.coef[r] = 0.0; L.sel[r] = 7; }
    const double kW = P[GMPNP_P_KW], kA = P[GMPNP_P_KA], kB = P[GMPNP_P_KB];
    const double kA2 = P[GMPNP_P_KA2], kB2 = P[GMPNP_P_KB2];
    if (c == 0) { L.coef[0] = kW; L.sel[0] = 1; }
    else if (c == 1) { L.coef[0] = kW; L.sel[0] = 0; L.coef[1] = kA; L.sel[1] = 2; L.coef[2] = kB; L.sel[2] = 4; }
    else if (c == 2) { L.coef[1] = kA; L.sel[1] = 1; L.coef[4] = kB2; L.sel[4] = 7; }
    else if (c == 3) { L.coef[3] = kA2; L.sel[3] = 7; }
    else if (c == 4) { L.coef[2] = kB; L.sel[2] = 1; }
    // own reaction row of the residual (1D:383-410): signs of (w, a, b, a2, b2, kw1), scaled by scale_R
    const double kk[6] = {kW, kA, kB, kA2, kB2, P[GMPNP_P_KW1]};
    const double sg[5][6] = {{1, 0, 0, 0, 0, -1}, {1, 1, 1, -1, -1, -1}, {0, 1, -1, -1, 1, 0},
                             {0, -1, 0, 1, 0, 0}, {0, 0, 1, 0, -1, 0}};
#pragma unroll
    for (int t = 0; t < 6; ++t) L.rsig[t] = 0.0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        if (c == i) {
#pragma unroll
            for (int t = 0; t < 6; ++t) L.rsig[t] = P[GMPNP_P_S + i] * sg[i][t] * kk[t];
        }
    }
}

// Element blocks of one cell for this lane.
//   lane c<7 : cab[i] = dF_e[a,i]/dU[b,c]   (column c of block (a,b)), and f0/f1 = F_e[0,c], F_e[1,c]:
//              every lane integrates the residual row of its own component, so no lane runs a separate
//              residual pass (lane 7 only collects the rows, see forward_sweep)
struct CellCols { double c00[NC], c01[NC], c10[NC], c11[NC]; double f0, f1; };

// STASH: the blocks of local node 1 (c10, c11: needed by the NEXT row) go straight to shared memory
// (s10[i*8+c], s11[i*8+c]) instead of staying in registers across the elimination of the current row.
template <int NQJ, bool STASH>
__device__ __forceinline__ void cell_columns(const double* __restrict__ P, const double* __restrict__ sU0,
                                             const double* __restrict__ sU1, const LaneConst& L, int c, double h,
                                             double prow, const double (&U0)[NC], const double (&U1)[NC],
                                             double myU0, double myU1, double myN0, double myN1, CellCols& o,
                                             double* __restrict__ s10 = nullptr, double* __restrict__ s11 = nullptr) {
    auto put10 = [&](int i, double v) { if (STASH) s10[i * 8 + c] = v; else o.c10[i] = v; };
    auto put11 = [&](int i, double v) { if (STASH) s11[i * 8 + c] = v; else o.c11[i] = v; };
    // prow scales the Poisson row (row NS of every block and the Poisson residual): the solver passes 1/q so
    // that the row is O(1) like the species rows and the in-block pivot is almost always the diagonal
    // (q ~ 1e9 in 1D, 1D:193); the materialising kernel passes 1.
    const double ih = fast_rcp(h);
    double g[NC], dU[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) { dU[i] = U1[i] - U0[i]; g[i] = dU[i] * ih; }
    double G = 0.0;
#pragma unroll
    for (int i = 0; i < NS; ++i) G += P[GMPNP_P_NU + i] * g[i];
    const double gp = g[NS];
    const double kappa = P[GMPNP_P_KAPPA];
    // grad phi_0 = -ih, grad phi_1 = +ih
    const double Ga0 = -G * ih, Ga1 = G * ih;
    const double gpa0 = -gp * ih, gpa1 = gp * ih;
    // lane 7 (right-hand-side lane) has no column: it shadows the potential lane, its blocks are never read
    const double myg = (myU1 - myU0) * ih;
    // quadrature accumulators: a0 = int u_i D, a1 = int u_i phi_0 D^2, a2 = int u_i phi_1 D^2, mD = int phi_b D
    // (Jacobian rule); own-row residual sums a0F = int u_c D, R0/R1 = int (-R_c) phi_a (residual rule)
    double a0[NS], a1[NS], a2[NS], mD0 = 0.0, mD1 = 0.0;
    double a0F = 0.0, R0 = 0.0, R1 = 0.0;
#pragma unroll
    for (int i = 0; i < NS; ++i) { a0[i] = 0.0; a1[i] = 0.0; a2[i] = 0.0; }
    auto species_acc = [&](double l0, double l1, double W, const double (&uq)[NS], double D) {
        const double WD = W * D, WD2 = WD * D;
        mD0 += WD * l0; mD1 += WD * l1;
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const double t = uq[i];
            a0[i] += WD * t; a1[i] += WD2 * l0 * t; a2[i] += WD2 * l1 * t;
        }
    };
    auto resid_acc = [&](double l0, double l1, double W, const double (&uq)[NS], double D) {
        a0F += W * D * (l0 * myU0 + l1 * myU1);
        const double mr = L.rsig[0] * (uq[0] * uq[1]) + L.rsig[1] * (uq[1] * uq[2]) + L.rsig[2] * (uq[4] * uq[1]) +
                          L.rsig[3] * uq[3] + L.rsig[4] * uq[2] + L.rsig[5];
        R0 += W * l0 * mr; R1 += W * l1 * mr;
    };
    if (c < NS) {
        if (NQJ == 2) {
            // consistent Jacobian: J and F share the 2-point rule, so the point evaluations are shared too
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const double l1 = GX2[q], l0 = 1.0 - l1, W = GW2[q] * h;
                double uq[NS], S = 0.0;
#pragma unroll
                for (int i = 0; i < NS; ++i) { uq[i] = fma(l1, dU[i], U0[i]); S += P[GMPNP_P_NU + i] * uq[i]; }
                const double D = fast_rcp(1.0 - S);
                species_acc(l0, l1, W, uq, D);
                resid_acc(l0, l1, W, uq, D);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const double l1 = GX3[q], l0 = 1.0 - l1, W = GW3[q] * h;
                double uq[NS], S = 0.0;
#pragma unroll
                for (int i = 0; i < NS; ++i) { uq[i] = fma(l1, dU[i], U0[i]); S += P[GMPNP_P_NU + i] * uq[i]; }
                species_acc(l0, l1, W, uq, fast_rcp(1.0 - S));
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const double l1 = GX2[q], l0 = 1.0 - l1, W = GW2[q] * h;
                double uq[NS], S = 0.0;
#pragma unroll
                for (int i = 0; i < NS; ++i) { uq[i] = fma(l1, dU[i], U0[i]); S += P[GMPNP_P_NU + i] * uq[i]; }
                resid_acc(l0, l1, W, uq, fast_rcp(1.0 - S));
            }
        }
        // ---- residual row of species c (2-point Gauss; time/source moments in closed form) -----------
        {
            const double d0 = myU0 - myN0, d1 = myU1 - myN1;
            const double sUi = 0.5 * h * (myU0 + myU1);
            const double zi = P[GMPNP_P_Z + c];
            o.f0 = kappa * h * ((1.0 / 3.0) * d0 + (1.0 / 6.0) * d1) - myg + zi * gpa0 * sUi + Ga0 * a0F + R0;
            o.f1 = kappa * h * ((1.0 / 6.0) * d0 + (1.0 / 3.0) * d1) + myg + zi * gpa1 * sUi + Ga1 * a0F + R1;
        }
        // ---- species column j = c --------------------------------------------------
        const double nuj = P[GMPNP_P_NU + c];
        const double zj = P[GMPNP_P_Z + c];
        const double ih2 = ih * ih;
        const double Md = h * (1.0 / 3.0), Mo = h * (1.0 / 6.0), mb = 0.5 * h;
        // reaction block column: int phi_a phi_b u_sel = (h/12) (alpha u_sel(node 0) + beta u_sel(node 1)) with
        // (alpha, beta) = (3,1), (1,1), (1,3) for ab = 00, 01, 11 (closed form), so the five rate derivatives are
        // combined per NODE first (Pn0, Pn1) and the three blocks are formed from those
        double E0[5], E1[5];
        const double h12 = h * (1.0 / 12.0);
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            const double ch = L.coef[r] * h12;
            E0[r] = ch * sU0[L.sel[r]];
            E1[r] = ch * sU1[L.sel[r]];
        }
        const double s0 = P[GMPNP_P_S], s1 = P[GMPNP_P_S + 1], s2 = P[GMPNP_P_S + 2];
        const double s3 = P[GMPNP_P_S + 3], s4 = P[GMPNP_P_S + 4];
        auto rx = [&](const double (&E)[5], double (&R)[5]) {
            R[0] = s0 * E[0];
            R[1] = s1 * (E[0] + E[1] + E[2] - E[3] - E[4]);
            R[2] = s2 * (E[1] + E[4] - E[3] - E[2]);
            R[3] = s3 * (E[3] - E[1]);
            R[4] = s4 * (E[2] - E[4]);
        };
        double Pn0[5], Pn1[5];
        rx(E0, Pn0); rx(E1, Pn1);
        double R00[5], R01[5], R11[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            R01[i] = Pn0[i] + Pn1[i];
            R00[i] = fma(2.0, Pn0[i], R01[i]);
            R11[i] = fma(2.0, Pn1[i], R01[i]);
        }
        // diagonal (i == j) extras per block
        const double d00 = kappa * Md + ih + zj * gpa0 * mb + Ga0 * mD0;
        const double d01 = kappa * Mo - ih + zj * gpa0 * mb + Ga0 * mD1;
        const double d10 = kappa * Mo - ih + zj * gpa1 * mb + Ga1 * mD0;
        const double d11 = kappa * Md + ih + zj * gpa1 * mb + Ga1 * mD1;
        const double nG0 = nuj * Ga0, nG1 = nuj * Ga1, nk = nuj * ih2;
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const double k = nk * a0[i];
            double v00 = fma(nG0, a1[i], k);
            double v01 = fma(nG0, a2[i], -k);
            double v10 = fma(nG1, a1[i], -k);
            double v11 = fma(nG1, a2[i], k);
            if (i < 5) { v00 += R00[i]; v01 += R01[i]; v10 += R01[i]; v11 += R11[i]; }
            const double dg = (i == c) ? 1.0 : 0.0;
            o.c00[i] = fma(dg, d00, v00);
            o.c01[i] = fma(dg, d01, v01);
            put10(i, fma(dg, d10, v10));
            put11(i, fma(dg, d11, v11));
        }
        // Poisson row:  -eps'_j (gp.grad a) m_b + q z_j c0_j M_ab
        double depsj = 0.0;
        if (c == 0) depsj = (6.0 - P[GMPNP_P_EPSW]) * (1.0 / 55.0) * P[GMPNP_P_EPSH];
        if (c == NS - 1) depsj = (6.0 - P[GMPNP_P_EPSW]) * (1.0 / 55.0) * P[GMPNP_P_EPSC];
        const double qz = P[GMPNP_P_Q] * P[GMPNP_P_ZC0 + c] * prow;
        const double dm = depsj * mb * prow;
        o.c00[NS] = -dm * gpa0 + qz * Md;
        o.c01[NS] = -dm * gpa0 + qz * Mo;
        put10(NS, -dm * gpa1 + qz * Mo);
        put11(NS, -dm * gpa1 + qz * Md);
    } else {
        // ---- potential column and the Poisson residual row ------------------------------------
        const double ih2 = ih * ih;
        double rho0 = 0.0, rho1 = 0.0;
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            const double iU = 0.5 * h * (U0[i] + U1[i]);
            const double v = P[GMPNP_P_Z + i] * ih2 * iU;
            o.c00[i] = v; put11(i, v); o.c01[i] = -v; put10(i, -v);
            rho0 += P[GMPNP_P_ZC0 + i] * U0[i];
            rho1 += P[GMPNP_P_ZC0 + i] * U1[i];
        }
        const double wm = P[GMPNP_P_EPSC] * 0.5 * (U0[NS - 1] + U1[NS - 1]) + P[GMPNP_P_EPSH] * 0.5 * (U0[0] + U1[0]);
        const double epsm = P[GMPNP_P_EPSW] * ((55.0 - wm) * (1.0 / 55.0)) + 6.0 * (wm * (1.0 / 55.0));
        const double v = -ih2 * h * epsm * prow;
        o.c00[NS] = v; put11(NS, v); o.c01[NS] = -v; put10(NS, -v);
        const double qh = P[GMPNP_P_Q] * h;
        o.f0 = (-gpa0 * h * epsm + qh * ((1.0 / 3.0) * rho0 + (1.0 / 6.0) * rho1)) * prow;
        o.f1 = (-gpa1 * h * epsm + qh * ((1.0 / 6.0) * rho0 + (1.0 / 3.0) * rho1)) * prow;
    }
}

// ---------------------------------------------------------------------------------------
// group helpers
// ---------------------------------------------------------------------------------------
struct Group {
    int c;            // lane in group (column id)
    bool live;        // newton1d_kernel: the pair works on a real problem (false: padding pair of the last CTA)
    unsigned mask;    // participation mask of the 8 lanes
    int base;         // first lane of the group inside the warp
    const double* P;  // parameter record of the problem (shared memory)
    double* sm;       // per-group shared memory (work area)
    double* su;       // assemble1d_kernel only: staged nodal values [2][8]
    double* ring;     // per-group staging ring (SM_RING doubles, 16-B aligned)
    // two-sided elimination: two groups (a "pair", 16 lanes) share one problem
    int half;         // 0: sweeps down from node 0; 1: sweeps up from node n-1
    unsigned pmask;   // participation mask of the pair
    int pbase;        // first lane of the pair
    double* psm;      // pair-shared merge area [8][8]
    double* q;        // block-row queue shared by the producer group and its consumer group [NSLOT][Q_SLOT]
    int bar;          // first named barrier of this warp pair: full[s] = bar + s, empty[s] = bar + NSLOT + s, command = bar + 2 NSLOT
};

// publish this lane's component of a node (already in a register) in sU[slot] and give every lane
// a register copy of all 7 components
__device__ __forceinline__ void stage_node(const Group& g, double mine, int slot, double (&U)[NC]) {
    double* sU = g.su + slot * 8;
    sU[g.c] = (g.c < NC) ? mine : 1.0;
    __syncwarp(g.mask);
#pragma unroll
    for (int i = 0; i < NC; ++i) U[i] = sU[i];
    __syncwarp(g.mask);
}

__device__ __forceinline__ void load_node(const Group& g, const double* __restrict__ up, long node,
                                          int slot, double (&U)[NC]) {
    stage_node(g, (g.c < NC) ? up[node * NC + g.c] : 1.0, slot, U);
}

// Elimination of one block row: B' = B - A X_prev (X_prev = previous row's eliminated coupling block /
// rhs), then Gauss-Jordan on [B' | C | d'].  Lane c<7 holds column c of A, B' and C (C in Y), lane 7 holds d'
// (in Y).  Step j broadcasts pivot column j from lane j by shuffles; every lane repeats the (cheap) pivot
// search so all agree on the pivot row without a shared-memory round trip; rows are swapped physically so
// all register indexing is static.  On exit Y = eliminated coupling column (lanes<7) / rhs (lane 7).
template <bool PIVOT>
__device__ __forceinline__ void eliminate_row(const Group& g, const double* __restrict__ sA, double (&B)[NC],
                                              double (&Y)[NC], const double (&X)[NC], int& singular) {
    const int c = g.c;
    if (sA != nullptr) {
        // A row-major in shared memory (lane c wrote column c: conflict-free), rows read back as 4 x 128 bit;
        // the writes are ordered before these reads by the group barrier after the residual gather
        double t[NC];
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            const double2* row = reinterpret_cast<const double2*>(sA + i * 8);
            const double2 a0 = row[0], a1 = row[1], a2 = row[2], a3 = row[3];
            t[i] = a0.x * X[0] + a0.y * X[1] + a1.x * X[2] + a1.y * X[3] + a2.x * X[4] + a2.y * X[5] + a3.x * X[6];
        }
        if (c < NC) {
#pragma unroll
            for (int i = 0; i < NC; ++i) B[i] -= t[i];
        } else {
#pragma unroll
            for (int i = 0; i < NC; ++i) Y[i] -= t[i];
        }
    }
#pragma unroll
    for (int j = 0; j < NC; ++j) {
        // pivot column j: published by its owner lane in shared memory (double-buffered: one group
        // barrier per step), read back by every lane as a broadcast
        double pc[NC];
        {
            double2* pcs = reinterpret_cast<double2*>(g.sm + SM_PC + (j & 1) * 8);
            if (c == j) {
                pcs[0] = make_double2(B[0], B[1]); pcs[1] = make_double2(B[2], B[3]);
                pcs[2] = make_double2(B[4], B[5]); pcs[3] = make_double2(B[6], 0.0);
            }
            __syncwarp();
            const double2 p0 = pcs[0], p1 = pcs[1], p2 = pcs[2], p3 = pcs[3];
            pc[0] = p0.x; pc[1] = p0.y; pc[2] = p1.x; pc[3] = p1.y; pc[4] = p2.x; pc[5] = p2.y; pc[6] = p3.x;
        }
        if (PIVOT) {
            int p = j;
            double best = fabs(pc[j]);
#pragma unroll
            for (int i = j + 1; i < NC; ++i) {
                const double a = fabs(pc[i]);
                if (a > best) { best = a; p = i; }
            }
            if (p != j) {                    // group-uniform; the diagonal is the usual pivot
#pragma unroll
                for (int i = j + 1; i < NC; ++i) {
                    const bool sw = (p == i);
                    const double tp = pc[j], tb = B[j], ty = Y[j];
                    pc[j] = sw ? pc[i] : tp; pc[i] = sw ? tp : pc[i];
                    B[j] = sw ? B[i] : tb;   B[i] = sw ? tb : B[i];
                    Y[j] = sw ? Y[i] : ty;   Y[i] = sw ? ty : Y[i];
                }
            }
        }
        const double piv = pc[j];
        const double inv = fast_rcp(piv);
        if (!(fabs(piv) > 0.0) || !isfinite(inv)) singular = 1;
        const double bj = B[j] * inv, yj = Y[j] * inv;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            if (i == j) continue;
            B[i] -= pc[i] * bj;
            Y[i] -= pc[i] * yj;
        }
        B[j] = bj; Y[j] = yj;
    }
}

// PRODUCER side of one elimination sweep over `rows` block rows starting at node `first` and moving in direction
// `dir` (+1: top half, downwards; -1: bottom half, upwards -- a twisted / "burn at both ends" factorisation, so the
// two halves of a problem are eliminated concurrently).  Row by row: integrate the cell ahead (in sweep order), form
// block row k = (A | B | C | d) from the cell behind and the cell ahead, apply boundary rows, hand it to the consumer
// group through queue slot r % NSLOT.  The cell ahead always exists because each half stops short of the other end of
// the domain.  A cell is integrated in the sweep's own orientation (local node 0 = current node): the forms only
// contain products of two gradients, so they are invariant under the reflection.
// After the last row a closing slot carries what the merge needs: the (1,0)/(1,1) blocks and the residual row of the
// last cell, and this half's sum of squares of the residual.
template <int NQJ>
__device__ void producer_sweep(const Group& g, const LaneConst& L, const double* __restrict__ x, int n, int first,
                               int dir, int rows, const double* __restrict__ up, const double* __restrict__ unp) {
    const double* P = g.P;
    const int c = g.c;
    double* sF = g.sm + SM_M;                        // residual rows of the current node, gathered for lane 7
    const bool use_un = (P[GMPNP_P_KAPPA] != 0.0);   // steady equations never read u_n
    const double qscale = P[GMPNP_P_Q], prow = 1.0 / qscale;
    double* fu = g.ring;                             // [RING][8]: u of a node, slot 7 = 1.0
    double* fn = g.ring + RING * 8;                  // [RING][8]: u_n of a node, slot 7 = x of the node
    // stage node first + dir*j into ring slot j % RING (always commits, so group counting stays uniform)
    auto issue = [&](int j) {
        const int node = first + dir * j;
        if (node >= 0 && node < n) {
            const int s8 = (j & (RING - 1)) * 8;
            if (c < NC) {
                cp_async8(fu + s8 + c, up + (long)node * NC + c);
                if (use_un) cp_async8(fn + s8 + c, unp + (long)node * NC + c);
            } else {
                cp_async8(fn + s8 + 7, x + node);
            }
        }
        cp_async_commit();
    };
    if (c == 7) {
#pragma unroll
        for (int j = 0; j < RING; ++j) fu[j * 8 + 7] = 1.0;
    }
#pragma unroll
    for (int j = 0; j < RING; ++j) issue(j);
    double f1_behind = 0.0;       // cell behind: this lane's residual row at the shared node
    double rsq = 0.0;
    int slot = 0;
    for (int r = 0; r < rows; ++r) {
        const int k = first + dir * r;               // current node
        const int s0 = (r & (RING - 1)) * 8, s1 = ((r + 1) & (RING - 1)) * 8;
        const int nslot = (slot + 1 == NSLOT) ? 0 : slot + 1;
        double* qs = g.q + slot * Q_SLOT;            // block row r
        double* qn = g.q + nslot * Q_SLOT;           // receives the (1,0)/(1,1) blocks of the cell ahead (for row r+1)
        if (r >= 2) bar_sync(g.bar + NSLOT + nslot); // the consumer is done with row r-2, which lived in slot nslot
        cp_async_wait<RING - 2>();                   // node r+1 has landed (groups 0 .. r+RING-1 are in flight)
        __syncwarp();
        {
            // nodal values of the cell ahead (local node 0 = current node) straight from the staging ring
            double U0[NC], U1[NC];
            {
                const double2* q = reinterpret_cast<const double2*>(fu + s0);
                const double2 q0 = q[0], q1 = q[1], q2 = q[2];
                U0[0] = q0.x; U0[1] = q0.y; U0[2] = q1.x; U0[3] = q1.y; U0[4] = q2.x; U0[5] = q2.y; U0[6] = fu[s0 + 6];
                const double2* w = reinterpret_cast<const double2*>(fu + s1);
                const double2 w0 = w[0], w1 = w[1], w2 = w[2];
                U1[0] = w0.x; U1[1] = w0.y; U1[2] = w1.x; U1[3] = w1.y; U1[4] = w2.x; U1[5] = w2.y; U1[6] = fu[s1 + 6];
            }
            const double h = fabs(fn[s1 + 7] - fn[s0 + 7]);
            const double myU0 = fu[s0 + c], myU1 = fu[s1 + c];          // lane 7 reads the constant 1.0
            const double myN0 = (c < NC && use_un) ? fn[s0 + c] : 0.0, myN1 = (c < NC && use_un) ? fn[s1 + c] : 0.0;
            CellCols cc;
            cell_columns<NQJ, true>(P, fu + s0, fu + s1, L, c, h, prow, U0, U1, myU0, myU1, myN0, myN1, cc, qn + Q_A, qn + Q_B);
            // ---- row k: A = (1,0) behind (already in the slot), B = (1,1) behind + c00, coupling ahead = c01 ----
            if (r > 0) {                                     // the common case, kept free of the boundary selects
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    qs[Q_B + i * 8 + c] += cc.c00[i];
                    qs[Q_C + i * 8 + c] = cc.c01[i];
                }
            } else {
                // first row of a half = boundary node: nothing behind it; Dirichlet rows become identity rows
                const bool dir_all = (k == n - 1);           // x = 1: all components (1D:350-353)
                const bool dir_pot = (k == 0);               // OHP: potential = V (1D:354)
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    double bv = cc.c00[i], cv = cc.c01[i];
                    if (dir_all || (dir_pot && i == NS)) { bv = (i == c) ? 1.0 : 0.0; cv = 0.0; }
                    qs[Q_B + i * 8 + c] = bv;
                    qs[Q_C + i * 8 + c] = cv;
                }
            }
            sF[c] = f1_behind + cc.f0;               // this lane's residual row of node k
            f1_behind = cc.f1;
        }
        __syncwarp();
        if (c == 7) {
            double Y[NC];
#pragma unroll
            for (int i = 0; i < NC; ++i) Y[i] = sF[i];
            double pscale = qscale;                  // ||b||_2 of the reference's (unscaled) system: undo the Poisson-row
            if (r == 0) {                            // scaling except on Dirichlet rows
                // point fluxes `J_i v_i ds` at both end points (1D:553, 738)
#pragma unroll
                for (int i = 0; i < NS; ++i) Y[i] += P[GMPNP_P_JFLUX + i];
                // Dirichlet rows (1D:350-355): x=1 all components = (1,..,1,0); x=0 potential = V.  Only the first row
                // of a half is a boundary node (the halves start at nodes 0 and n-1).
                if (k == n - 1) {
#pragma unroll
                    for (int i = 0; i < NC; ++i) Y[i] = fu[s0 + i] - ((i < NS) ? 1.0 : 0.0);
                }
                if (k == 0) Y[NS] = fu[s0 + NS] - P[GMPNP_P_V];
                pscale = 1.0;
            }
#pragma unroll
            for (int i = 0; i < NS; ++i) rsq += Y[i] * Y[i];
            const double yp = Y[NS] * pscale;
            rsq += yp * yp;
#pragma unroll
            for (int i = 0; i < NC; ++i) qs[Q_D(i)] = Y[i];
        }
        __syncwarp();                                // every lane is past its last read of ring slot r
        issue(r + RING);
        bar_arrive(g.bar + slot);                    // block row r is complete (each lane after its own writes)
        slot = nslot;
    }
    // drain the staging ring first: the merge area of a pair aliases the ring of its bottom-half group
    cp_async_wait<0>();
    __syncwarp();
    // closing slot: (1,0)/(1,1) blocks of the last cell are in it already; add its residual row and the half's sum
    {
        double* qs = g.q + slot * Q_SLOT;
        if (c < NC) qs[Q_D(c)] = f1_behind;
        if (c == 7) qs[Q_C + 7] = rsq;
        bar_arrive(g.bar + slot);
    }
}

// CONSUMER side of the sweep: takes block row r from queue slot r % NSLOT, eliminates it against the previous row
// (B' = B - A C'_(k-1), Gauss-Jordan inside the block), stores (coupling'_k | d'_k) in ws[node].  Returns the index of
// the closing slot (already waited for); X returns the last row's eliminated block column (lanes<7) / rhs (lane 7).
template <bool PIVOT>
__device__ int consumer_sweep(const Group& g, int first, int dir, int rows, double* __restrict__ ws, double (&X)[NC],
                              int& singular) {
    const int c = g.c;
#pragma unroll
    for (int i = 0; i < NC; ++i) X[i] = 0.0;
    int slot = 0;
    for (int r = 0; r < rows; ++r) {
        const int k = first + dir * r;
        double* qs = g.q + slot * Q_SLOT;
        bar_sync(g.bar + slot);                      // block row r has been assembled
        double B[NC], Y[NC];
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            B[i] = qs[Q_B + i * 8 + c];
            Y[i] = (c < NC) ? qs[Q_C + i * 8 + c] : qs[Q_D(i)];
        }
        eliminate_row<PIVOT>(g, r > 0 ? qs + Q_A : nullptr, B, Y, X, singular);
        if (r + 2 < rows) bar_arrive(g.bar + NSLOT + slot);      // slot free for the producer's row r+2
        // ---- store (coupling'_k | d'_k) ------------------------------------------------------
        double* w = ws + (long)k * 56;
        if (g.live) {
#pragma unroll
            for (int i = 0; i < NC; ++i) w[i * 8 + c] = Y[i];
        }
#pragma unroll
        for (int i = 0; i < NC; ++i) X[i] = Y[i];
        slot = (slot + 1 == NSLOT) ? 0 : slot + 1;
    }
    bar_sync(g.bar + slot);                          // closing slot
    return slot;
}

// Back substitution x_k = d'_k - coupling'_k x_(k-dir') over `rows` rows starting at node `first`, moving in
// direction `dir`, seeded with the already known neighbour solution xn; updates u <- u - relax * x and
// accumulates max|dx|, max|u_new| in this lane.
__device__ void backward_sweep(const Group& g, int first, int dir, int rows, int rows_max, bool store,
                               double (&xn)[NC], double* __restrict__ up, const double* __restrict__ ws, double relax,
                               double& mdx, double& mu) {
    // rows_max: the warp-uniform trip count (the other half may have one row more); rows beyond `rows` are idle
    const int c = g.c;
    double2* bwc = reinterpret_cast<double2*>(g.ring) + c;     // [RING][4][8]: lane c's row of the workspace
    double* buc = g.ring + RING * 64 + c;                      // [RING][8]:    lane c's component of u
    double* sx = g.sm + SM_X;
    // every lane stages what it will read itself, so no cross-lane visibility is needed for the ring;
    // running global pointers, ring slots as compile-time constants (the row loop is unrolled RING times)
    const double* wsrc = ws + (long)first * 56 + c * 8;        // workspace row of the next node to stage
    const double* usrc = up + (long)first * NC + c;
    double* udst = up + (long)first * NC + c;                  // u of the next node to update
    const long wstep = (long)dir * 56, ustep = (long)dir * NC;
    int staged = 0;
    auto issue = [&](int s) {
        if (staged < rows && c < NC) {
#pragma unroll
            for (int v = 0; v < 4; ++v) cp_async16(bwc + (s * 4 + v) * 8, wsrc + 2 * v);
            cp_async8(buc + s * 8, usrc);
        }
        cp_async_commit();
        wsrc += wstep; usrc += ustep; ++staged;
    };
#pragma unroll
    for (int s = 0; s < RING; ++s) issue(s);
    for (int q0 = 0; q0 < rows_max; q0 += RING) {
#pragma unroll
        for (int s = 0; s < RING; ++s) {
            if (q0 + s < rows_max) {                           // warp-uniform
                cp_async_wait<RING - 1>();
                double xi = 0.0;
                if (c < NC && q0 + s < rows) {
                    const double2 r0 = bwc[(s * 4 + 0) * 8], r1 = bwc[(s * 4 + 1) * 8];
                    const double2 r2 = bwc[(s * 4 + 2) * 8], r3 = bwc[(s * 4 + 3) * 8];
                    const double ucur = buc[s * 8];
                    xi = r3.y;
                    xi -= r0.x * xn[0]; xi -= r0.y * xn[1]; xi -= r1.x * xn[2]; xi -= r1.y * xn[3];
                    xi -= r2.x * xn[4]; xi -= r2.y * xn[5]; xi -= r3.x * xn[6];
                    const double un = ucur - relax * xi;
                    if (store) *udst = un;
                    mdx = fmax(mdx, fabs(xi));
                    mu = fmax(mu, fabs(un));
                }
                udst += ustep;
                issue(s);
                // x_k of all components to every lane (double-buffered: one group barrier per row)
                double* sxq = sx + (s & 1) * 8;
                sxq[c] = xi;
                __syncwarp();
                const double2* xs = reinterpret_cast<const double2*>(sxq);
                const double2 a0 = xs[0], a1 = xs[1], a2 = xs[2];
                xn[0] = a0.x; xn[1] = a0.y; xn[2] = a1.x; xn[3] = a1.y; xn[4] = a2.x; xn[5] = a2.y; xn[6] = sxq[6];
            }
        }
    }
    cp_async_wait<0>();
    __syncwarp();
}

struct NewtonOut { int iters; double r0, r; int status; double dx; };

// Factorisation sweep of both halves + merge.  Both halves eliminate m = n/2 rows (top: 0..m-1 downwards, bottom:
// n-1 .. n-m upwards), so the two groups of a pair -- and with the warp-uniform drivers below all four groups of a
// warp -- run the same number of rows.  Even n: the bottom's last row is node m and the merge eliminates the virtual
// row (I - A'_m C'_{m-1}) x_m = d'_m - A'_m d'_{m-1}.  Odd n: node m is left over and is eliminated in the merge from
// the blocks both halves stashed for it: (B_m - A_m C'_{m-1} - C_m A'_{m+1}) x_m = d_m - A_m d'_{m-1} - C_m d'_{m+1}.
// x_m is broadcast to the pair.  Returns ||b||^2 of the whole problem.  Executed by all 32 lanes in lock step.
template <bool PIVOT>
__device__ double factor_problem(const Group& g, int n, double* ws, double (&xm)[NC], int& singular, int* cmd) {
    const int m = n >> 1;
    const bool odd = (n & 1) != 0;
    const int c = g.c;
    double X[NC];
    const int first = g.half ? n - 1 : 0, dir = g.half ? -1 : 1;
    // wake the producer warp: it assembles the m block rows of both halves of the warp's two problems
    if ((threadIdx.x & 31) == 0) *cmd = CMD_FACTOR;
    bar_sync(g.bar + 2 * NSLOT);
    const int cslot = consumer_sweep<PIVOT>(g, first, dir, m, ws, X, singular);
    const double* qc = g.q + cslot * Q_SLOT;         // closing slot of this half
    double rsq = qc[Q_C + 7];
    const double f1_last = (c < NC) ? qc[Q_D(c)] : 0.0;
    double Y[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) Y[i] = 0.0;
    if (!odd) {
        // bottom publishes (A'_m | d'_m) row-major: psm[i*8 + c], column 7 = d'_m
        if (g.half) {
#pragma unroll
            for (int i = 0; i < NC; ++i) g.psm[i * 8 + c] = X[i];
        }
        __syncwarp();
        double B[NC];
#pragma unroll
        for (int i = 0; i < NC; ++i) B[i] = (i == c) ? 1.0 : 0.0;
        if (c == 7) {
#pragma unroll
            for (int i = 0; i < NC; ++i) Y[i] = g.psm[i * 8 + 7];          // rhs d'_m
        }
        int sing_m = 0;
        eliminate_row<PIVOT>(g, g.psm, B, Y, X, sing_m);                    // only the top group's result is used
        if (!g.half) singular |= sing_m;
    } else {
        // each half: its part of row m = (block (1,1) | residual row) of its last cell minus (block (1,0)) * X
        const double* sA_last = qc + Q_A;
        const double* sB_last = qc + Q_B;
        double part[NC];
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            const double2* row = reinterpret_cast<const double2*>(sA_last + i * 8);
            const double2 a0 = row[0], a1 = row[1], a2 = row[2], a3 = row[3];
            const double t = a0.x * X[0] + a0.y * X[1] + a1.x * X[2] + a1.y * X[3] + a2.x * X[4] + a2.y * X[5] + a3.x * X[6];
            part[i] = ((c < NC) ? sB_last[i * 8 + c] : qc[Q_D(i)]) - t;
        }
        if (g.half) {
#pragma unroll
            for (int i = 0; i < NC; ++i) g.psm[i * 8 + c] = part[i];
            if (c < NC) g.psm[56 + c] = f1_last;                            // raw residual row (for ||b||)
        }
        __syncwarp();
        double B[NC];
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            const double tot = part[i] + g.psm[i * 8 + c];
            B[i] = (c < NC) ? tot : 0.0;
            Y[i] = (c < NC) ? 0.0 : tot;
        }
        if (c == 7 && !g.half) {
            // ||b||^2 of row m: raw residual rows of both halves, Poisson row unscaled
            const double qscale = g.P[GMPNP_P_Q];
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                double d = qc[Q_D(i)] + g.psm[56 + i];
                if (i == NS) d *= qscale;
                rsq += d * d;
            }
        }
        int sing_m = 0;
        eliminate_row<PIVOT>(g, nullptr, B, Y, X, sing_m);                  // only the top group's result is used
        if (!g.half) singular |= sing_m;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NC; ++i) xm[i] = __shfl_sync(0xffffffffu, Y[i], g.pbase + 7);
    rsq = __shfl_sync(0xffffffffu, rsq, g.base + 7);                        // lane 7 holds the group's sum
    rsq += __shfl_xor_sync(0xffffffffu, rsq, 8);
    singular |= __shfl_xor_sync(0xffffffffu, singular, 8);
    return rsq;
}

// Back substitution of both halves and the Newton update (stores only if `store`); returns max|dx|, max|u| over the
// problem.  Row m (x_m known from the merge) is updated by the bottom group.
__device__ void solve_problem(const Group& g, int n, const double (&xm)[NC], double* up, const double* ws,
                              double relax, bool store, double& dxmax, double& umax) {
    const int m = n >> 1;
    double xn[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) xn[i] = xm[i];
    double mdx = 0.0, mu = 0.0;
    if (g.half && g.c < NC) {
        double xi = xm[0];
#pragma unroll
        for (int i = 1; i < NC; ++i) if (i == g.c) xi = xm[i];
        const double un = up[(long)m * NC + g.c] - relax * xi;
        if (store) up[(long)m * NC + g.c] = un;
        mdx = fabs(xi); mu = fabs(un);
    }
    const int first = g.half ? m + 1 : m - 1, dir = g.half ? 1 : -1, rows = g.half ? n - 1 - m : m;
    backward_sweep(g, first, dir, rows, m, store, xn, up, ws, relax, mdx, mu);
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) {
        mdx = fmax(mdx, __shfl_xor_sync(0xffffffffu, mdx, o));
        mu = fmax(mu, __shfl_xor_sync(0xffffffffu, mu, o));
    }
    dxmax = mdx; umax = mu;
}

// dolfin NewtonSolver semantics for the problem handled by a pair of groups.  The loop is
// WARP-UNIFORM: the two pairs of a warp iterate together until both are done; a pair that has converged (or is
// not `enabled`: padding pair, finished continuation path, failed earlier) keeps executing the sweeps but does not
// store its u.  This costs nothing (the warp occupies its slot until its slower pair finishes anyway) and lets
// every barrier and shuffle of the sweeps use the full-warp mask: the runtime 8-lane masks expanded to a
// MATCH/VOTE sequence of six instructions per barrier, eleven barriers per block row.
template <bool PIVOT>
__device__ NewtonOut newton_solve(const Group& g, int n, double* up, double* ws, const gmpnp_newton_opts& o,
                                  bool enabled, int* cmd) {
    NewtonOut out;
    int singular = 0;
    double xm[NC];
    double rsq = factor_problem<PIVOT>(g, n, ws, xm, singular, cmd);
    double r = sqrt(rsq);
    out.r0 = r;
    int k = 0;
    bool conv = (o.criterion == 0) ? (r < o.atol) : false;
    bool bad = !isfinite(r) || singular;
    bool stag = false;
    double dx_prev = INFINITY, dx_rel = INFINITY;
    bool active = enabled && !conv && !bad && k < o.maxit;
    while (__any_sync(0xffffffffu, active)) {
        double dxmax, umax;
        solve_problem(g, n, xm, up, ws, o.relax, active, dxmax, umax);
        bool still = active;
        if (active) {
            ++k;
            if (o.criterion == 1) {
                const double scale = fmax(1.0, umax);
                conv = dxmax <= o.xtol * scale;
                // opt-in round-off floor (gmpnp.h, xtol_floor): an increment that is small but no longer contracts
                // sits at cond(J)*eps of the linear solve, above xtol -- reported as GMPNP_STAGNATED, never as converged
                if (!conv && o.xtol_floor > 0.0 && k >= 3 && dxmax <= o.xtol_floor * scale && dxmax >= 0.25 * dx_prev)
                    stag = true;
                dx_prev = dxmax;
                dx_rel = dxmax / scale;
                if (!isfinite(dxmax)) bad = true;
                if (conv || bad || stag) still = false;   // dolfin-like: no re-assembly after an increment stop
            }
        }
        if (!__any_sync(0xffffffffu, still)) break;   // nobody needs the re-assembly (increment stop / failure)
        __syncwarp();               // the other half's u updates must be visible before re-assembly
        int sing2 = 0;
        const double rsq2 = factor_problem<PIVOT>(g, n, ws, xm, sing2, cmd);
        if (still) {
            r = sqrt(rsq2);
            if (!isfinite(r) || sing2) bad = true;
            if (o.criterion == 0) conv = (r / out.r0 < o.rtol) || (r < o.atol);
        }
        active = still && !conv && !bad && k < o.maxit;
    }
    out.iters = k;
    out.r = r;
    out.dx = dx_rel;
    out.status = bad ? GMPNP_NOT_FINITE : (conv ? GMPNP_CONVERGED : (stag ? GMPNP_STAGNATED : GMPNP_MAXIT));
    return out;
}

constexpr int PROBLEMS_PER_BLOCK = NW_PROBLEMS;
constexpr size_t NEWTON_SMEM_DOUBLES = (size_t)NW_GROUPS * (SM_GROUP + SM_RING) + NW_PROBLEMS * GMPNP_NPAR + 8;

// roles: warps [0, NW_PAIRS) are consumers (and run the Newton / march / continuation control), warp NW_PAIRS + w is
// the producer of consumer warp w.  A lane and its twin in the partner warp handle the same (problem, half).
__device__ __forceinline__ void group_setup(Group& g, int batch, int& prob, double* smem, bool& producer, int*& cmd) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // consumers in warps [0, NW_PAIRS), producers above.  (Measured: alternating the roles with the CTA index, so that
    // every scheduler sees a mix of producer and consumer warps, is SLOWER -- 87.9 vs 70.2 ms on the saturated launch.)
    producer = warp >= NW_PAIRS;
    const int wp = producer ? warp - NW_PAIRS : warp;          // warp pair
    g.c = lane & 7;
    g.base = lane & ~7;
    g.mask = 0xFFu << g.base;
    g.half = (lane >> 3) & 1;
    g.pbase = lane & ~15;
    g.pmask = 0xFFFFu << g.pbase;
    const int gid = wp * 4 + (lane >> 3);                      // (problem, half) slot of the CTA
    g.sm = smem + gid * SM_GROUP;
    g.q = g.sm + SM_Q;
    g.ring = smem + NW_GROUPS * SM_GROUP + gid * SM_RING;
    g.psm = smem + NW_GROUPS * SM_GROUP + (gid | 1) * SM_RING;         // ring of the pair's bottom-half group
    g.P = smem + NW_GROUPS * (SM_GROUP + SM_RING) + (gid >> 1) * GMPNP_NPAR;
    g.su = nullptr;
    g.bar = 1 + wp * BAR_PER_PAIR;
    cmd = reinterpret_cast<int*>(smem + NW_GROUPS * (SM_GROUP + SM_RING) + NW_PROBLEMS * GMPNP_NPAR) + wp;
    prob = blockIdx.x * PROBLEMS_PER_BLOCK + wp * 2 + (lane >> 4);
    // padding pairs of the last CTA shadow the last problem (valid memory to read) but never store
    g.live = prob < batch;
    if (!g.live) prob = batch - 1;
}

// parameter record of the pair's problem -> shared memory (both groups of the pair cooperate)
__device__ __forceinline__ void load_params(const Group& g, const double* __restrict__ params, int prob) {
    double* P = const_cast<double*>(g.P);
    for (int i = (threadIdx.x & 15); i < GMPNP_NPAR; i += 16) P[i] = params[(long)prob * GMPNP_NPAR + i];
    __syncwarp();
}

// mode 0: single Newton solve (gmpnp_newton_1d)
// mode 1: pseudo-time march  (gmpnp_march_1d): n_stage steps, H_OHP controller, u_n <- u
// mode 2: steady continuation (gmpnp_steady_continuation_1d): kappa = 0, V from Vpath
// Control flow of the consumer warps is warp-uniform (see newton_solve): per-pair outcomes are flags, never early
// exits; the producer warps only execute commands.
#ifndef GMPNP_NEWTON_MIN_BLOCKS
#define GMPNP_NEWTON_MIN_BLOCKS 3
#endif
template <bool PIVOT, int NQJ>
__global__ void __launch_bounds__(NW_THREADS, GMPNP_NEWTON_MIN_BLOCKS)
newton1d_kernel(int mode, int batch, int n, const double* __restrict__ x, const double* __restrict__ params,
                double* __restrict__ u, double* __restrict__ un_rw, const double* __restrict__ un_ro,
                double* __restrict__ wsall, gmpnp_newton_opts opts, int n_stage,
                const double* __restrict__ Vpath, double* __restrict__ hist, int* __restrict__ iters,
                double* __restrict__ r0out, double* __restrict__ rout, double* __restrict__ hfrac_out,
                int* __restrict__ stage_out, int* __restrict__ status) {
    extern __shared__ double smem[];
    Group g; int prob; bool producer; int* cmd;
    group_setup(g, batch, prob, smem, producer, cmd);
    double* up = u + (long)prob * n * NC;
    if (producer) {
        // ---- producer warp: wait for a command, assemble the block rows of one factorisation sweep ------------------
        // (the parameter record in shared memory is loaded and updated by the consumer warp; the command barrier
        // orders those writes, and the consumer's updates of u, before the reads below)
        const double* unp = (mode == 0) ? un_ro + (long)prob * n * NC : (mode == 1) ? un_rw + (long)prob * n * NC : up;
        const int m = n >> 1;
        const int first = g.half ? n - 1 : 0, dir = g.half ? -1 : 1;
        LaneConst L;
        bool have_L = false;
        while (true) {
            bar_sync(g.bar + 2 * NSLOT);
            if (*cmd == CMD_EXIT) break;
            if (!have_L) { lane_consts(g.P, g.c, L); have_L = true; }     // rate constants do not change during a launch
            producer_sweep<NQJ>(g, L, x, n, first, dir, m, up, unp);
        }
        return;
    }
    // ---- consumer warp: control + elimination + back substitution ---------------------------------------------------
    load_params(g, params, prob);
    double* P = const_cast<double*>(g.P);
    double* ws = wsall + (long)prob * n * 56;
    const bool writer = (g.c == 0 && g.half == 0 && g.live);
    const int lane16 = (threadIdx.x & 15);
    auto finish = [&]() {
        if ((threadIdx.x & 31) == 0) *cmd = CMD_EXIT;
        bar_sync(g.bar + 2 * NSLOT);
    };
    if (mode == 0) {
        NewtonOut o = newton_solve<PIVOT>(g, n, up, ws, opts, g.live, cmd);
        if (writer) {
            if (iters) iters[prob] = o.iters;
            if (r0out) r0out[prob] = o.r0;
            if (rout) rout[prob] = o.r;
            if (status) status[prob] = o.status;
        }
        finish();
        return;
    }
    if (mode == 1) {
        double* unp = un_rw + (long)prob * n * NC;
        double frac = P[GMPNP_P_HFRAC];
        const double hohp = P[GMPNP_P_HOHP];
        int st = GMPNP_CONVERGED, done = 0;
        bool alive = g.live;
        for (int s = 0; s < n_stage; ++s) {
            if (!__any_sync(0xffffffffu, alive)) break;
            NewtonOut o = newton_solve<PIVOT>(g, n, up, ws, opts, alive, cmd);
            if (alive) {
                if (writer && iters) iters[(long)prob * n_stage + s] = o.iters;
                if (o.status != GMPNP_CONVERGED) { st = o.status; alive = false; }
                else ++done;
            }
            __syncwarp();
            if (alive) {
                // u_n <- u (1D:796) and history row
                for (long i = lane16; i < (long)n * NC; i += 16) {
                    const double v = up[i];
                    unp[i] = v;
                    if (hist) hist[((long)prob * n_stage + s) * n * NC + i] = v;
                }
            }
            __syncwarp();
            if (alive && hohp >= 0.0) {
                // proton-current controller, 1D:766-793
                const double f = up[0];
                if (f < 0) frac = frac / 1.1;
                else if (f < (hohp - 0.05)) frac = frac / 1.05;
                else if (f < (hohp - 0.025)) frac = frac / 1.01;
                else if (f > hohp && f <= (hohp + 0.4) && frac <= 1.0) frac = frac * 1.04;
                else if (f > (hohp + 0.4) && frac <= 1.0) frac = frac * 1.15;
            }
            __syncwarp();
            if (writer && alive && hohp >= 0.0) {
                P[GMPNP_P_JFLUX + 1] = -1.0 * P[GMPNP_P_JOHPRE] * (1 - frac);
                P[GMPNP_P_JFLUX + 0] = P[GMPNP_P_JHPRE] * frac;
            }
            __syncwarp();
        }
        if (writer) {
            if (status) status[prob] = st;
            if (hfrac_out) hfrac_out[prob] = frac;
            if (stage_out) stage_out[prob] = done;
        }
        finish();
        return;
    }
    // mode 2: steady continuation
    {
        if (writer) P[GMPNP_P_KAPPA] = 0.0;
        __syncwarp();
        int st = GMPNP_CONVERGED, done = 0;
        const double xtol_final = opts.xtol;
        bool alive = g.live;
        for (int s = 0; s < n_stage; ++s) {
            const double Vs = Vpath[(long)prob * n_stage + s];
            if (isnan(Vs)) alive = false;               // ragged path: this problem is done
            if (!__any_sync(0xffffffffu, alive)) break;
            const bool final_stage = (s + 1 == n_stage) || isnan(Vpath[(long)prob * n_stage + s + 1]);
            gmpnp_newton_opts o2 = opts;
            o2.xtol = (final_stage || !(opts.xtol_path > 0.0)) ? xtol_final : opts.xtol_path;
            __syncwarp();
            if (writer && alive) P[GMPNP_P_V] = Vs;
            __syncwarp();
            NewtonOut o = newton_solve<PIVOT>(g, n, up, ws, o2, alive, cmd);
            if (alive) {
                if (writer && iters) iters[(long)prob * n_stage + s] = o.iters;
                if (writer && rout) rout[prob] = o.r;
                if (writer && hfrac_out) hfrac_out[prob] = o.dx;          // d_dx of the C-ABI
                // a stalled increment ends the iteration of this stage but not the path; the status of the LAST
                // stage is what the caller sees
                if (o.status == GMPNP_CONVERGED || o.status == GMPNP_STAGNATED) { ++done; st = o.status; }
                else { st = o.status; alive = false; }
            }
        }
        if (writer) {
            if (status) status[prob] = st;
            if (stage_out) stage_out[prob] = done;
        }
        finish();
    }
}

// =======================================================================================
// Partitioned elimination: S = 4 or 8 sweeps per problem (strong scaling / single-problem latency)
// =======================================================================================
// The two-sided factorisation above has a critical path of n/2 block rows per Newton iteration whatever the number
// of idle SMs.  Here the chain is cut into Q = S/2 + 1 sub-domains by Q - 1 SEPARATOR nodes.  The two outer sub-domains
// are swept once (from the domain boundary towards their separator, as above); every interior sub-domain is swept
// TWICE, downwards and upwards, each sweep starting next to a separator whose unknown value it carries along as a
// 7-column SPIKE W (x_k = d'_k - C'_k x_next - W'_k x_sep).  All S sweeps run concurrently (one producer group + one
// consumer group each) over ~n/Q rows.  The two sweeps that END at a separator contribute their halves of its block
// row; substituting their last rows gives a block-tridiagonal system in the Q - 1 separator unknowns, solved
// redundantly by every group (<= 4 block rows).  Back substitution then runs concurrently again: an outer sweep
// recovers all its rows, the two sweeps of an interior sub-domain recover the half nearest their end separator
// (only those rows' factors are stored).  Work: ~1.6x the assembly and ~2.6x the elimination at S = 8 for a 2.5x
// shorter chain -- chosen only when the batch leaves SMs idle (gmpnp_newton_opts.partitions).
// NumPy prototype of the algebra: tests/studies/partitioned_thomas_prototype.py.
constexpr int NSEP_MAX = 4;                 // separators per CTA (S = 4: 2 problems x 2; S = 8: 1 problem x 4)
constexpr int SEP_PART = 0, SEP_SPIKE = 112, SEP_FRAW = 224, SEP_RED = 240, SEP_XS = 296, SEP_REC = 304;
constexpr int PART_SCRATCH = 64;            // per CTA: reduction scratch of the problem-wide control values
constexpr size_t PART_SMEM_DOUBLES = (size_t)NW_GROUPS * (SM_GROUP + SM_RING) + 2 * GMPNP_NPAR + NSEP_MAX * SEP_REC +
                                     PART_SCRATCH + 8;
constexpr int BAR_PROBLEM = 15;             // S = 8: the two consumer warps of a problem meet here

struct Sweep {
    int first, dir, rows, rows_max, off;   // off = rows_max - rows (0 or 1): idle iterations at the start of the uniform loop
    int bs0;                               // back substitution covers sweep rows [bs0, rows)
    int pre;                               // 1: starts next to a separator (carries the spike; integrates the separator cell first)
    int si_start, si_end;                  // separator index at the start (-1: domain boundary) and at the end of the sweep
    int half_end;                          // half of the end separator's record this sweep fills: 0 arrives from above, 1 from below
    int sep_end;                           // node index of the end separator
};

template <int S>
__device__ __forceinline__ void sweep_layout(int n, int j, Sweep& w) {
    constexpr int Q = S / 2 + 1;
    const int inner = n - (Q - 1), base = inner / Q, rem = inner % Q;
    int q, dir;
    if (j == 0) { q = 0; dir = 1; }
    else if (j == S - 1) { q = Q - 1; dir = -1; }
    else { q = (j + 1) >> 1; dir = (j & 1) ? 1 : -1; }
    const int rows = base + (q < rem ? 1 : 0);
    const int a = q * base + min(q, rem) + q, b = a + rows - 1;
    w.dir = dir; w.rows = rows; w.first = dir > 0 ? a : b;
    w.rows_max = base + (rem > 0 ? 1 : 0);
    w.off = w.rows_max - rows;
    const bool outer = (j == 0) || (j == S - 1);
    w.pre = outer ? 0 : 1;
    if (dir > 0) { w.si_start = outer ? -1 : q - 1; w.si_end = q; w.half_end = 0; w.sep_end = b + 1; }
    else { w.si_start = outer ? -1 : q; w.si_end = q - 1; w.half_end = 1; w.sep_end = a - 1; }
    const int h = rows >> 1;
    w.bs0 = outer ? 0 : (dir > 0 ? h : rows - h);
}

template <int S>
__device__ __forceinline__ void problem_sync() {
    if (S == 8) bar_sync(BAR_PROBLEM);
    else __syncwarp();
}

// eliminate_row with a spike block: W (lane c < 7: column c) is one more right-hand-side block of the row.
//   spike_init : first row next to a separator -- its sub-diagonal block couples to the separator unknown: W = A
//   otherwise  : W = -A Wp (Wp = eliminated spike of the previous row), B -= A X_C, d -= A X_d as in eliminate_row
template <bool PIVOT>
__device__ __forceinline__ void eliminate_row_w(const Group& g, const double* __restrict__ sA, bool spike_init,
                                                double (&B)[NC], double (&Y)[NC], double (&W)[NC], const double (&X)[NC],
                                                const double (&Wp)[NC], int& singular) {
    const int c = g.c;
#pragma unroll
    for (int i = 0; i < NC; ++i) W[i] = 0.0;
    if (sA != nullptr) {
        if (spike_init) {
            if (c < NC) {
#pragma unroll
                for (int i = 0; i < NC; ++i) W[i] = sA[i * 8 + c];
            }
        } else {
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                const double2* row = reinterpret_cast<const double2*>(sA + i * 8);
                const double2 a0 = row[0], a1 = row[1], a2 = row[2], a3 = row[3];
                const double t = a0.x * X[0] + a0.y * X[1] + a1.x * X[2] + a1.y * X[3] + a2.x * X[4] + a2.y * X[5] + a3.x * X[6];
                const double tw = a0.x * Wp[0] + a0.y * Wp[1] + a1.x * Wp[2] + a1.y * Wp[3] + a2.x * Wp[4] + a2.y * Wp[5] + a3.x * Wp[6];
                if (c < NC) { B[i] -= t; W[i] = -tw; }
                else Y[i] -= t;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NC; ++j) {
        double pc[NC];
        {
            double2* pcs = reinterpret_cast<double2*>(g.sm + SM_PC + (j & 1) * 8);
            if (c == j) {
                pcs[0] = make_double2(B[0], B[1]); pcs[1] = make_double2(B[2], B[3]);
                pcs[2] = make_double2(B[4], B[5]); pcs[3] = make_double2(B[6], 0.0);
            }
            __syncwarp();
            const double2 p0 = pcs[0], p1 = pcs[1], p2 = pcs[2], p3 = pcs[3];
            pc[0] = p0.x; pc[1] = p0.y; pc[2] = p1.x; pc[3] = p1.y; pc[4] = p2.x; pc[5] = p2.y; pc[6] = p3.x;
        }
        if (PIVOT) {
            int p = j;
            double best = fabs(pc[j]);
#pragma unroll
            for (int i = j + 1; i < NC; ++i) {
                const double a = fabs(pc[i]);
                if (a > best) { best = a; p = i; }
            }
            if (p != j) {
#pragma unroll
                for (int i = j + 1; i < NC; ++i) {
                    const bool sw = (p == i);
                    const double tp = pc[j], tb = B[j], ty = Y[j], tw = W[j];
                    pc[j] = sw ? pc[i] : tp; pc[i] = sw ? tp : pc[i];
                    B[j] = sw ? B[i] : tb;   B[i] = sw ? tb : B[i];
                    Y[j] = sw ? Y[i] : ty;   Y[i] = sw ? ty : Y[i];
                    W[j] = sw ? W[i] : tw;   W[i] = sw ? tw : W[i];
                }
            }
        }
        const double piv = pc[j];
        const double inv = fast_rcp(piv);
        if (!(fabs(piv) > 0.0) || !isfinite(inv)) singular = 1;
        const double bj = B[j] * inv, yj = Y[j] * inv, wj = W[j] * inv;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            if (i == j) continue;
            B[i] -= pc[i] * bj;
            Y[i] -= pc[i] * yj;
            W[i] -= pc[i] * wj;
        }
        B[j] = bj; Y[j] = yj; W[j] = wj;
    }
}

// Producer of one sweep of the partitioned factorisation.  The loop counter `it` is uniform over the warp
// (it = -1 .. rows_max-1); the sweep's own row is r = it - off, it integrates the cell ahead of row r for r >= -pre
// (r = -1: the cell between its start separator and its first row, whose (1,0)/(1,1) blocks and residual row open row 0).
template <int NQJ>
__device__ void producer_sweep_part(const Group& g, const LaneConst& L, const Sweep& w, const double* __restrict__ x, int n,
                                    const double* __restrict__ up, const double* __restrict__ unp) {
    const double* P = g.P;
    const int c = g.c;
    double* sF = g.sm + SM_M;
    const bool use_un = (P[GMPNP_P_KAPPA] != 0.0);
    const double qscale = P[GMPNP_P_Q], prow = 1.0 / qscale;
    double* fu = g.ring;
    double* fn = g.ring + RING * 8;
    const int dir = w.dir, node0 = w.first - dir * w.pre;        // first staged node (the start separator when pre)
    auto issue = [&](int j) {
        const int node = node0 + dir * j;
        if (node >= 0 && node < n) {
            const int s8 = (j & (RING - 1)) * 8;
            if (c < NC) {
                cp_async8(fu + s8 + c, up + (long)node * NC + c);
                if (use_un) cp_async8(fn + s8 + c, unp + (long)node * NC + c);
            } else {
                cp_async8(fn + s8 + 7, x + node);
            }
        }
        cp_async_commit();
    };
    if (c == 7) {
#pragma unroll
        for (int j = 0; j < RING; ++j) fu[j * 8 + 7] = 1.0;
    }
#pragma unroll
    for (int j = 0; j < RING; ++j) issue(j);
    double f1_behind = 0.0, rsq = 0.0;
    int t = 0;                                       // staged-sequence index of the current node
    for (int it = -1; it < w.rows_max; ++it) {
        const int r = it - w.off;                    // this sweep's row (r < -pre: idle)
        const int slot = (it < 0) ? 0 : it % NSLOT;
        const int nslot = (it + 1) % NSLOT;
        double* qs = g.q + slot * Q_SLOT;
        double* qn = g.q + nslot * Q_SLOT;
        if (it >= 2) bar_sync(g.bar + NSLOT + nslot);            // the consumers are done with iteration it-2 (slot nslot)
        if (r >= -w.pre) {
            const int s0 = (t & (RING - 1)) * 8, s1 = ((t + 1) & (RING - 1)) * 8;
            const int k = w.first + dir * r;         // current node (the separator itself for r = -1)
            cp_async_wait<RING - 2>();
            __syncwarp(g.mask);
            {
                double U0[NC], U1[NC];
                {
                    const double2* q = reinterpret_cast<const double2*>(fu + s0);
                    const double2 q0 = q[0], q1 = q[1], q2 = q[2];
                    U0[0] = q0.x; U0[1] = q0.y; U0[2] = q1.x; U0[3] = q1.y; U0[4] = q2.x; U0[5] = q2.y; U0[6] = fu[s0 + 6];
                    const double2* ww = reinterpret_cast<const double2*>(fu + s1);
                    const double2 w0 = ww[0], w1 = ww[1], w2 = ww[2];
                    U1[0] = w0.x; U1[1] = w0.y; U1[2] = w1.x; U1[3] = w1.y; U1[4] = w2.x; U1[5] = w2.y; U1[6] = fu[s1 + 6];
                }
                const double h = fabs(fn[s1 + 7] - fn[s0 + 7]);
                const double myU0 = fu[s0 + c], myU1 = fu[s1 + c];
                const double myN0 = (c < NC && use_un) ? fn[s0 + c] : 0.0, myN1 = (c < NC && use_un) ? fn[s1 + c] : 0.0;
                CellCols cc;
                cell_columns<NQJ, true>(P, fu + s0, fu + s1, L, c, h, prow, U0, U1, myU0, myU1, myN0, myN1, cc, qn + Q_A, qn + Q_B);
                if (r >= 0) {
                    if (r > 0 || w.pre) {            // a cell behind exists: its (1,1) block is in the slot already
#pragma unroll
                        for (int i = 0; i < NC; ++i) {
                            qs[Q_B + i * 8 + c] += cc.c00[i];
                            qs[Q_C + i * 8 + c] = cc.c01[i];
                        }
                    } else {
                        const bool dir_all = (k == n - 1), dir_pot = (k == 0);
#pragma unroll
                        for (int i = 0; i < NC; ++i) {
                            double bv = cc.c00[i], cv = cc.c01[i];
                            if (dir_all || (dir_pot && i == NS)) { bv = (i == c) ? 1.0 : 0.0; cv = 0.0; }
                            qs[Q_B + i * 8 + c] = bv;
                            qs[Q_C + i * 8 + c] = cv;
                        }
                    }
                    sF[c] = f1_behind + cc.f0;
                }
                f1_behind = cc.f1;
            }
            __syncwarp(g.mask);
            if (r >= 0 && c == 7) {
                double Y[NC];
#pragma unroll
                for (int i = 0; i < NC; ++i) Y[i] = sF[i];
                double pscale = qscale;
                if (r == 0 && !w.pre) {
#pragma unroll
                    for (int i = 0; i < NS; ++i) Y[i] += P[GMPNP_P_JFLUX + i];
                    if (k == n - 1) {
#pragma unroll
                        for (int i = 0; i < NC; ++i) Y[i] = fu[s0 + i] - ((i < NS) ? 1.0 : 0.0);
                    }
                    if (k == 0) Y[NS] = fu[s0 + NS] - P[GMPNP_P_V];
                    pscale = 1.0;
                }
                if (r >= w.bs0) {                    // every row is counted once: by the sweep that back-substitutes it
#pragma unroll
                    for (int i = 0; i < NS; ++i) rsq += Y[i] * Y[i];
                    const double yp = Y[NS] * pscale;
                    rsq += yp * yp;
                }
#pragma unroll
                for (int i = 0; i < NC; ++i) qs[Q_D(i)] = Y[i];
            }
            __syncwarp(g.mask);
            issue(t + RING);
            ++t;
        }
        if (it >= 0) bar_arrive(g.bar + slot);       // iteration `it` is complete for this group (idle or not)
    }
    cp_async_wait<0>();
    __syncwarp(g.mask);
    {
        double* qs = g.q + (w.rows_max % NSLOT) * Q_SLOT;        // closing slot
        if (c < NC) qs[Q_D(c)] = f1_behind;
        if (c == 7) qs[Q_C + 7] = rsq;
        bar_arrive(g.bar + w.rows_max % NSLOT);
    }
}

// Consumer of one sweep of the partitioned factorisation; returns the closing slot.  X / Wp return the eliminated
// coupling column (lanes < 7) / right-hand side (lane 7) and the eliminated spike column of the sweep's last row.
template <bool PIVOT>
__device__ int consumer_sweep_part(const Group& g, const Sweep& w, double* __restrict__ ws, double* __restrict__ ws2,
                                   double (&X)[NC], double (&Wp)[NC], int& singular) {
    const int c = g.c;
#pragma unroll
    for (int i = 0; i < NC; ++i) { X[i] = 0.0; Wp[i] = 0.0; }
    for (int it = 0; it < w.rows_max; ++it) {
        const int r = it - w.off;
        const int slot = it % NSLOT;
        double* qs = g.q + slot * Q_SLOT;
        bar_sync(g.bar + slot);
        double B[NC], Y[NC], W[NC];
        const bool real = r >= 0;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            B[i] = real ? qs[Q_B + i * 8 + c] : ((i == c) ? 1.0 : 0.0);
            Y[i] = real ? ((c < NC) ? qs[Q_C + i * 8 + c] : qs[Q_D(i)]) : 0.0;
        }
        int sing = 0;
        eliminate_row_w<PIVOT>(g, (real && (r > 0 || w.pre)) ? qs + Q_A : nullptr, real && r == 0 && w.pre, B, Y, W, X, Wp, sing);
        if (it + 2 < w.rows_max) bar_arrive(g.bar + NSLOT + slot);
        if (real) {
            singular |= sing;
            if (r >= w.bs0 && g.live) {
                const int k = w.first + w.dir * r;
                double* o1 = ws + (long)k * 56;
                double* o2 = ws2 + (long)k * 56;
#pragma unroll
                for (int i = 0; i < NC; ++i) { o1[i * 8 + c] = Y[i]; o2[i * 8 + c] = W[i]; }
            }
#pragma unroll
            for (int i = 0; i < NC; ++i) { X[i] = Y[i]; Wp[i] = W[i]; }
        }
    }
    const int cslot = w.rows_max % NSLOT;
    bar_sync(g.bar + cslot);
    return cslot;
}

// Back substitution of one sweep: rows r = rows-1 .. bs0 (from its end separator backwards),
// x_k = d'_k - C'_k x_next - W'_k x_S; updates u and accumulates max|dx|, max|u|.  cnt_max: uniform trip count.
__device__ void backward_sweep_part(const Group& g, const Sweep& w, int cnt_max, bool store, double (&xn)[NC],
                                    const double (&xS)[NC], double* __restrict__ up, const double* __restrict__ ws,
                                    const double* __restrict__ ws2, double relax, double& mdx, double& mu) {
    const int c = g.c;
    const int cnt = w.rows - w.bs0;
    const int start = w.first + w.dir * (w.rows - 1), dir = -w.dir;
    double2* bwc = reinterpret_cast<double2*>(g.ring) + c;     // [RING][4][8]: lane c's row of (C' | d')
    double* buc = g.ring + RING * 64 + c;                      // [RING][8]:    lane c's component of u
    double2* bw2 = reinterpret_cast<double2*>(g.q) + c;        // [RING][4][8]: lane c's row of W' (queue memory, idle now)
    double* sx = g.sm + SM_X;
    const double* wsrc = ws + (long)start * 56 + c * 8;
    const double* w2src = ws2 + (long)start * 56 + c * 8;
    const double* usrc = up + (long)start * NC + c;
    double* udst = up + (long)start * NC + c;
    const long wstep = (long)dir * 56, ustep = (long)dir * NC;
    int staged = 0;
    auto issue = [&](int s) {
        if (staged < cnt && c < NC) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                cp_async16(bwc + (s * 4 + v) * 8, wsrc + 2 * v);
                cp_async16(bw2 + (s * 4 + v) * 8, w2src + 2 * v);
            }
            cp_async8(buc + s * 8, usrc);
        }
        cp_async_commit();
        wsrc += wstep; w2src += wstep; usrc += ustep; ++staged;
    };
#pragma unroll
    for (int s = 0; s < RING; ++s) issue(s);
    for (int q0 = 0; q0 < cnt_max; q0 += RING) {
#pragma unroll
        for (int s = 0; s < RING; ++s) {
            if (q0 + s < cnt_max) {
                cp_async_wait<RING - 1>();
                double xi = 0.0;
                if (c < NC && q0 + s < cnt) {
                    const double2 r0 = bwc[(s * 4 + 0) * 8], r1 = bwc[(s * 4 + 1) * 8];
                    const double2 r2 = bwc[(s * 4 + 2) * 8], r3 = bwc[(s * 4 + 3) * 8];
                    const double2 v0 = bw2[(s * 4 + 0) * 8], v1 = bw2[(s * 4 + 1) * 8];
                    const double2 v2 = bw2[(s * 4 + 2) * 8], v3 = bw2[(s * 4 + 3) * 8];
                    const double ucur = buc[s * 8];
                    xi = r3.y;
                    xi -= r0.x * xn[0]; xi -= r0.y * xn[1]; xi -= r1.x * xn[2]; xi -= r1.y * xn[3];
                    xi -= r2.x * xn[4]; xi -= r2.y * xn[5]; xi -= r3.x * xn[6];
                    xi -= v0.x * xS[0]; xi -= v0.y * xS[1]; xi -= v1.x * xS[2]; xi -= v1.y * xS[3];
                    xi -= v2.x * xS[4]; xi -= v2.y * xS[5]; xi -= v3.x * xS[6];
                    const double un = ucur - relax * xi;
                    if (store) *udst = un;
                    mdx = fmax(mdx, fabs(xi));
                    mu = fmax(mu, fabs(un));
                }
                udst += ustep;
                issue(s);
                double* sxq = sx + (s & 1) * 8;
                sxq[c] = xi;
                __syncwarp();
                const double2* xs = reinterpret_cast<const double2*>(sxq);
                const double2 a0 = xs[0], a1 = xs[1], a2 = xs[2];
                xn[0] = a0.x; xn[1] = a0.y; xn[2] = a1.x; xn[3] = a1.y; xn[4] = a2.x; xn[5] = a2.y; xn[6] = sxq[6];
            }
        }
    }
    cp_async_wait<0>();
    __syncwarp();
}

// per-problem shared data of the partitioned kernel
struct Part {
    double* sep;      // separator records of this problem [Q-1][SEP_REC]
    double* scratch;  // [PART_SCRATCH / problems] reduction scratch
    int j;            // sweep index of this group inside the problem
    int lp, nlp;      // lane index among the problem's consumer lanes, their number
};

// Factorisation of one problem with S sweeps; returns ||b||^2; xE / xS return the solution at the sweep's end / start
// separator.  Executed by all consumer lanes of the problem in lock step.
template <bool PIVOT, int S>
__device__ double factor_problem_part(const Group& g, const Sweep& w, const Part& pt, double* ws, double* ws2,
                                      double (&xE)[NC], double (&xS)[NC], int& singular, int* cmd) {
    constexpr int NSEP = S / 2;
    const int c = g.c;
    double X[NC], Wp[NC];
    problem_sync<S>();          // all consumer groups of the problem are done with the previous phase (u, scratch, records)
    if ((threadIdx.x & 31) == 0) *cmd = CMD_FACTOR;
    bar_sync(g.bar + 2 * NSLOT);
    int sing = 0;
    const int cslot = consumer_sweep_part<PIVOT>(g, w, ws, ws2, X, Wp, sing);
    const double* qc = g.q + cslot * Q_SLOT;
    // this sweep's half of its end separator's block row: (1,1) block and residual row of its last cell, minus the
    // (1,0) block times the last eliminated row; the spike part couples the separator to the sweep's start separator
    {
        double* rec = pt.sep + w.si_end * SEP_REC;
        double* part = rec + SEP_PART + w.half_end * 56;
        double* spk = rec + SEP_SPIKE + w.half_end * 56;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            const double2* row = reinterpret_cast<const double2*>(qc + Q_A + i * 8);
            const double2 a0 = row[0], a1 = row[1], a2 = row[2], a3 = row[3];
            const double t = a0.x * X[0] + a0.y * X[1] + a1.x * X[2] + a1.y * X[3] + a2.x * X[4] + a2.y * X[5] + a3.x * X[6];
            const double tw = a0.x * Wp[0] + a0.y * Wp[1] + a1.x * Wp[2] + a1.y * Wp[3] + a2.x * Wp[4] + a2.y * Wp[5] + a3.x * Wp[6];
            part[i * 8 + c] = ((c < NC) ? qc[Q_B + i * 8 + c] : qc[Q_D(i)]) - t;
            spk[i * 8 + c] = (c < NC) ? -tw : 0.0;
        }
        if (c < NC) rec[SEP_FRAW + w.half_end * 8 + c] = qc[Q_D(c)];
        if (c == 7) { pt.scratch[pt.j] = qc[Q_C + 7]; pt.scratch[S + pt.j] = (double)sing; }
    }
    problem_sync<S>();
    // reduced block-tridiagonal system in the separator unknowns, solved redundantly by every group
    {
        double XX[NC];
#pragma unroll
        for (int i = 0; i < NC; ++i) XX[i] = 0.0;
        int sing_r = 0;
        for (int si = 0; si < NSEP; ++si) {
            const double* rec = pt.sep + si * SEP_REC;
            double B[NC], Y[NC];
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                const double dsum = rec[SEP_PART + i * 8 + c] + rec[SEP_PART + 56 + i * 8 + c];
                B[i] = (c < NC) ? dsum : 0.0;
                Y[i] = (c < NC) ? rec[SEP_SPIKE + 56 + i * 8 + c] : dsum;      // coupling to the NEXT separator / rhs
            }
            eliminate_row<PIVOT>(g, si > 0 ? rec + SEP_SPIKE : nullptr, B, Y, XX, sing_r);
#pragma unroll
            for (int i = 0; i < NC; ++i) XX[i] = Y[i];
            if (pt.j == 0) {
                double* red = pt.sep + si * SEP_REC + SEP_RED;
#pragma unroll
                for (int i = 0; i < NC; ++i) red[i * 8 + c] = Y[i];
            }
        }
        sing |= sing_r;
    }
    problem_sync<S>();
    double xs[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) { xs[i] = 0.0; xE[i] = 0.0; xS[i] = 0.0; }
    for (int si = NSEP - 1; si >= 0; --si) {
        const double* red = pt.sep + si * SEP_REC + SEP_RED;
        double xv[NC];
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            double v = red[i * 8 + 7];
            if (si < NSEP - 1) {
#pragma unroll
                for (int jj = 0; jj < NC; ++jj) v -= red[i * 8 + jj] * xs[jj];
            }
            xv[i] = v;
        }
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            xs[i] = xv[i];
            if (si == w.si_end) xE[i] = xv[i];
            if (si == w.si_start) xS[i] = xv[i];
        }
        if (pt.j == 0 && c == 0) {
            double* o = pt.sep + si * SEP_REC + SEP_XS;
#pragma unroll
            for (int i = 0; i < NC; ++i) o[i] = xv[i];
        }
    }
    // ||b||^2: rows counted once by the sweep that back-substitutes them + the raw separator rows
    double rsq = 0.0, sflag = (double)sing;
#pragma unroll
    for (int jj = 0; jj < S; ++jj) { rsq += pt.scratch[jj]; sflag += pt.scratch[S + jj]; }
    const double qscale = g.P[GMPNP_P_Q];
    for (int si = 0; si < NSEP; ++si) {
        const double* rec = pt.sep + si * SEP_REC;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            double d = rec[SEP_FRAW + i] + rec[SEP_FRAW + 8 + i];
            if (i == NS) d *= qscale;
            rsq += d * d;
        }
    }
    singular |= (sflag != 0.0);
    return rsq;
}

// Back substitution of all sweeps, separator updates, problem-wide max|dx|, max|u|.
template <int S>
__device__ void solve_problem_part(const Group& g, const Sweep& w, const Part& pt, const double (&xE)[NC],
                                   const double (&xS)[NC], double* up, const double* ws, const double* ws2, double relax,
                                   bool store, double& dxmax, double& umax) {
    double xn[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) xn[i] = xE[i];
    double mdx = 0.0, mu = 0.0;
    // the sweep that arrives at a separator from above updates the separator node itself
    if (w.half_end == 0 && g.c < NC) {
        double xi = xE[0];
#pragma unroll
        for (int i = 1; i < NC; ++i) if (i == g.c) xi = xE[i];
        const double un = up[(long)w.sep_end * NC + g.c] - relax * xi;
        if (store) up[(long)w.sep_end * NC + g.c] = un;
        mdx = fabs(xi); mu = fabs(un);
    }
    backward_sweep_part(g, w, w.rows_max, store, xn, xS, up, ws, ws2, relax, mdx, mu);
#pragma unroll
    for (int o = 4; o >= 1; o >>= 1) {
        mdx = fmax(mdx, __shfl_xor_sync(0xffffffffu, mdx, o));
        mu = fmax(mu, __shfl_xor_sync(0xffffffffu, mu, o));
    }
    problem_sync<S>();                               // previous readers of the scratch are done
    if (g.c == 0) { pt.scratch[2 * S + pt.j] = mdx; pt.scratch[3 * S + pt.j] = mu; }
    problem_sync<S>();
    mdx = 0.0; mu = 0.0;
#pragma unroll
    for (int jj = 0; jj < S; ++jj) { mdx = fmax(mdx, pt.scratch[2 * S + jj]); mu = fmax(mu, pt.scratch[3 * S + jj]); }
    dxmax = mdx; umax = mu;
}

template <bool PIVOT, int S>
__device__ NewtonOut newton_solve_part(const Group& g, const Sweep& w, const Part& pt, double* up, double* ws, double* ws2,
                                       const gmpnp_newton_opts& o, bool enabled, int* cmd) {
    NewtonOut out;
    int singular = 0;
    double xE[NC], xS[NC];
    double rsq = factor_problem_part<PIVOT, S>(g, w, pt, ws, ws2, xE, xS, singular, cmd);
    double r = sqrt(rsq);
    out.r0 = r;
    int k = 0;
    bool conv = (o.criterion == 0) ? (r < o.atol) : false;
    bool bad = !isfinite(r) || singular;
    bool stag = false;
    double dx_prev = INFINITY, dx_rel = INFINITY;
    bool active = enabled && !conv && !bad && k < o.maxit;
    while (active) {                                 // every consumer lane of the problem takes the same decisions
        double dxmax, umax;
        solve_problem_part<S>(g, w, pt, xE, xS, up, ws, ws2, o.relax, true, dxmax, umax);
        ++k;
        if (o.criterion == 1) {
            const double scale = fmax(1.0, umax);
            conv = dxmax <= o.xtol * scale;
            if (!conv && o.xtol_floor > 0.0 && k >= 3 && dxmax <= o.xtol_floor * scale && dxmax >= 0.25 * dx_prev)
                stag = true;
            dx_prev = dxmax;
            dx_rel = dxmax / scale;
            if (!isfinite(dxmax)) bad = true;
            if (conv || bad || stag) break;          // dolfin-like: no re-assembly after an increment stop
        }
        problem_sync<S>();                           // all u updates of the problem before the re-assembly
        int sing2 = 0;
        const double rsq2 = factor_problem_part<PIVOT, S>(g, w, pt, ws, ws2, xE, xS, sing2, cmd);
        r = sqrt(rsq2);
        if (!isfinite(r) || sing2) bad = true;
        if (o.criterion == 0) conv = (r / out.r0 < o.rtol) || (r < o.atol);
        active = !conv && !bad && k < o.maxit;
    }
    out.iters = k;
    out.r = r;
    out.dx = dx_rel;
    out.status = bad ? GMPNP_NOT_FINITE : (conv ? GMPNP_CONVERGED : (stag ? GMPNP_STAGNATED : GMPNP_MAXIT));
    return out;
}

// Partitioned variant of newton1d_kernel: same modes and outputs, S sweeps per problem (8 / S problems per CTA).
template <bool PIVOT, int NQJ, int S>
__global__ void __launch_bounds__(NW_THREADS, GMPNP_NEWTON_MIN_BLOCKS)
newton1d_part_kernel(int mode, int batch, int n, const double* __restrict__ x, const double* __restrict__ params,
                     double* __restrict__ u, double* __restrict__ un_rw, const double* __restrict__ un_ro,
                     double* __restrict__ wsall, double* __restrict__ ws2all, gmpnp_newton_opts opts, int n_stage,
                     const double* __restrict__ Vpath, double* __restrict__ hist, int* __restrict__ iters,
                     double* __restrict__ r0out, double* __restrict__ rout, double* __restrict__ hfrac_out,
                     int* __restrict__ stage_out, int* __restrict__ status) {
    static_assert(S == 4 || S == 8, "sweeps per problem");
    constexpr int PPB = 8 / S;                                  // problems per CTA
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool producer = warp >= NW_PAIRS;
    const int wp = producer ? warp - NW_PAIRS : warp;
    Group g;
    g.c = lane & 7; g.base = lane & ~7; g.mask = 0xFFu << g.base;
    g.half = 0; g.pbase = 0; g.pmask = 0; g.psm = nullptr; g.su = nullptr;
    const int gid = wp * 4 + (lane >> 3);
    g.sm = smem + gid * SM_GROUP;
    g.q = g.sm + SM_Q;
    g.ring = smem + NW_GROUPS * SM_GROUP + gid * SM_RING;
    const int pl = (S == 4) ? wp : 0;                           // problem slot inside the CTA
    double* tail = smem + NW_GROUPS * (SM_GROUP + SM_RING);
    g.P = tail + pl * GMPNP_NPAR;
    g.bar = 1 + wp * BAR_PER_PAIR;
    Part pt;
    pt.sep = tail + 2 * GMPNP_NPAR + pl * (S / 2) * SEP_REC;
    pt.scratch = tail + 2 * GMPNP_NPAR + NSEP_MAX * SEP_REC + pl * (PART_SCRATCH / 2);
    pt.j = (S == 4) ? (lane >> 3) : gid;
    pt.lp = (S == 4) ? lane : wp * 32 + lane;
    pt.nlp = (S == 4) ? 32 : 64;
    int* cmd = reinterpret_cast<int*>(tail + 2 * GMPNP_NPAR + NSEP_MAX * SEP_REC + PART_SCRATCH) + wp;
    int prob = blockIdx.x * PPB + pl;
    g.live = prob < batch;
    if (!g.live) prob = batch - 1;
    Sweep w;
    sweep_layout<S>(n, pt.j, w);
    double* up = u + (long)prob * n * NC;
    if (producer) {
        const double* unp = (mode == 0) ? un_ro + (long)prob * n * NC : (mode == 1) ? un_rw + (long)prob * n * NC : up;
        LaneConst L;
        bool have_L = false;
        while (true) {
            bar_sync(g.bar + 2 * NSLOT);
            if (*cmd == CMD_EXIT) break;
            if (!have_L) { lane_consts(g.P, g.c, L); have_L = true; }
            producer_sweep_part<NQJ>(g, L, w, x, n, up, unp);
        }
        return;
    }
    // ---- consumer warps: parameter record of the problem, control, elimination, back substitution ------------------
    {
        double* P = const_cast<double*>(g.P);
        for (int i = pt.lp; i < GMPNP_NPAR; i += pt.nlp) P[i] = params[(long)prob * GMPNP_NPAR + i];
    }
    problem_sync<S>();
    double* P = const_cast<double*>(g.P);
    double* ws = wsall + (long)prob * n * 56;
    double* ws2 = ws2all + (long)prob * n * 56;
    const bool writer = (pt.lp == 0 && g.live);
    auto finish = [&]() {
        if ((threadIdx.x & 31) == 0) *cmd = CMD_EXIT;
        bar_sync(g.bar + 2 * NSLOT);
    };
    if (mode == 0) {
        NewtonOut o = newton_solve_part<PIVOT, S>(g, w, pt, up, ws, ws2, opts, g.live, cmd);
        if (writer) {
            if (iters) iters[prob] = o.iters;
            if (r0out) r0out[prob] = o.r0;
            if (rout) rout[prob] = o.r;
            if (status) status[prob] = o.status;
        }
        finish();
        return;
    }
    if (mode == 1) {
        double* unp = un_rw + (long)prob * n * NC;
        double frac = P[GMPNP_P_HFRAC];
        const double hohp = P[GMPNP_P_HOHP];
        int st = GMPNP_CONVERGED, done = 0;
        bool alive = g.live;
        for (int s = 0; s < n_stage && alive; ++s) {
            NewtonOut o = newton_solve_part<PIVOT, S>(g, w, pt, up, ws, ws2, opts, alive, cmd);
            if (writer && iters) iters[(long)prob * n_stage + s] = o.iters;
            if (o.status != GMPNP_CONVERGED) { st = o.status; alive = false; }
            else ++done;
            problem_sync<S>();
            if (alive) {
                for (long i = pt.lp; i < (long)n * NC; i += pt.nlp) {
                    const double v = up[i];
                    unp[i] = v;
                    if (hist) hist[((long)prob * n_stage + s) * n * NC + i] = v;
                }
            }
            problem_sync<S>();
            if (alive && hohp >= 0.0) {
                const double f = up[0];
                if (f < 0) frac = frac / 1.1;
                else if (f < (hohp - 0.05)) frac = frac / 1.05;
                else if (f < (hohp - 0.025)) frac = frac / 1.01;
                else if (f > hohp && f <= (hohp + 0.4) && frac <= 1.0) frac = frac * 1.04;
                else if (f > (hohp + 0.4) && frac <= 1.0) frac = frac * 1.15;
            }
            problem_sync<S>();
            if (writer && alive && hohp >= 0.0) {
                P[GMPNP_P_JFLUX + 1] = -1.0 * P[GMPNP_P_JOHPRE] * (1 - frac);
                P[GMPNP_P_JFLUX + 0] = P[GMPNP_P_JHPRE] * frac;
            }
            problem_sync<S>();
        }
        if (writer) {
            if (status) status[prob] = st;
            if (hfrac_out) hfrac_out[prob] = frac;
            if (stage_out) stage_out[prob] = done;
        }
        finish();
        return;
    }
    {
        if (pt.lp == 0) P[GMPNP_P_KAPPA] = 0.0;
        problem_sync<S>();
        int st = GMPNP_CONVERGED, done = 0;
        const double xtol_final = opts.xtol;
        bool alive = g.live;
        for (int s = 0; s < n_stage && alive; ++s) {
            const double Vs = Vpath[(long)prob * n_stage + s];
            if (isnan(Vs)) break;                       // ragged path: this problem is done
            const bool final_stage = (s + 1 == n_stage) || isnan(Vpath[(long)prob * n_stage + s + 1]);
            gmpnp_newton_opts o2 = opts;
            o2.xtol = (final_stage || !(opts.xtol_path > 0.0)) ? xtol_final : opts.xtol_path;
            problem_sync<S>();
            if (pt.lp == 0) P[GMPNP_P_V] = Vs;
            problem_sync<S>();
            NewtonOut o = newton_solve_part<PIVOT, S>(g, w, pt, up, ws, ws2, o2, alive, cmd);
            if (writer && iters) iters[(long)prob * n_stage + s] = o.iters;
            if (writer && rout) rout[prob] = o.r;
            if (writer && hfrac_out) hfrac_out[prob] = o.dx;
            if (o.status == GMPNP_CONVERGED || o.status == GMPNP_STAGNATED) { ++done; st = o.status; }
            else { st = o.status; alive = false; }
        }
        if (writer) {
            if (status) status[prob] = st;
            if (stage_out) stage_out[prob] = done;
        }
        finish();
    }
}

// Materialised residual + block-tridiagonal Jacobian (gmpnp_assemble_1d): one group per
// (problem, node row).  Used for kernel-parity tests against the oracle.
__global__ void __launch_bounds__(THREADS)
assemble1d_kernel(int batch, int n, const double* __restrict__ x, const double* __restrict__ params,
                  const double* __restrict__ u, const double* __restrict__ un,
                  double* __restrict__ F, double* __restrict__ J) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31;
    Group g;
    g.c = lane & 7; g.base = lane & ~7; g.mask = 0xFFu << g.base;
    const int gid = threadIdx.x >> 3;
    double* base = smem + gid * AS_GROUP;
    g.sm = base; g.su = base + AS_U; g.P = base + AS_P;
    const long item = (long)blockIdx.x * GROUPS_PER_BLOCK + gid;
    if (item >= (long)batch * n) return;
    const int prob = (int)(item / n), k = (int)(item % n);
    for (int i = g.c; i < GMPNP_NPAR; i += 8) base[AS_P + i] = params[(long)prob * GMPNP_NPAR + i];
    __syncwarp(g.mask);
    const double* P = g.P;
    LaneConst L;
    lane_consts(P, g.c, L);
    const double* up = u + (long)prob * n * NC;
    const double* unp = un + (long)prob * n * NC;
    const int c = g.c;
    double* sF = base + AS_M;
    double A[NC], B[NC], C[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) { A[i] = 0.0; B[i] = 0.0; C[i] = 0.0; }
    double Um[NC], U0[NC], U1[NC];
    const double my0 = (c < NC) ? up[(long)k * NC + c] : 1.0;
    const double myn0 = (c < NC) ? unp[(long)k * NC + c] : 0.0;
    stage_node(g, my0, 1, U0);
    double frow = 0.0;                 // this lane's residual row of node k
    if (k > 0) {
        const double mym = (c < NC) ? up[(long)(k - 1) * NC + c] : 1.0;
        const double mynm = (c < NC) ? unp[(long)(k - 1) * NC + c] : 0.0;
        stage_node(g, mym, 0, Um);
        CellCols cc;
#pragma unroll
        for (int i = 0; i < NC; ++i) { cc.c00[i] = 0; cc.c01[i] = 0; cc.c10[i] = 0; cc.c11[i] = 0; }
        cell_columns<3, false>(P, g.su, g.su + 8, L, c, x[k] - x[k - 1], 1.0, Um, U0, mym, my0, mynm, myn0, cc);
#pragma unroll
        for (int i = 0; i < NC; ++i) { A[i] = cc.c10[i]; B[i] += cc.c11[i]; }
        frow += cc.f1;
    }
    __syncwarp(g.mask);
    if (k + 1 < n) {
        g.su[c] = my0;
        const double my1 = (c < NC) ? up[(long)(k + 1) * NC + c] : 1.0;
        const double myn1 = (c < NC) ? unp[(long)(k + 1) * NC + c] : 0.0;
        stage_node(g, my1, 1, U1);
        CellCols cc;
#pragma unroll
        for (int i = 0; i < NC; ++i) { cc.c00[i] = 0; cc.c01[i] = 0; cc.c10[i] = 0; cc.c11[i] = 0; }
        cell_columns<3, false>(P, g.su, g.su + 8, L, c, x[k + 1] - x[k], 1.0, U0, U1, my0, my1, myn0, myn1, cc);
#pragma unroll
        for (int i = 0; i < NC; ++i) { B[i] += cc.c00[i]; C[i] = cc.c01[i]; }
        frow += cc.f0;
    }
    sF[c] = frow;
    __syncwarp(g.mask);
    if (c == 7) {
#pragma unroll
        for (int i = 0; i < NC; ++i) B[i] = sF[i];
        if (k == 0 || k == n - 1) {
#pragma unroll
            for (int i = 0; i < NS; ++i) B[i] += P[GMPNP_P_JFLUX + i];
        }
        if (k == n - 1) {
#pragma unroll
            for (int i = 0; i < NC; ++i) B[i] = U0[i] - ((i < NS) ? 1.0 : 0.0);
        }
        if (k == 0) B[NS] = U0[NS] - P[GMPNP_P_V];
        if (F) {
#pragma unroll
            for (int i = 0; i < NC; ++i) F[((long)prob * n + k) * NC + i] = B[i];
        }
    } else if (c < NC) {
        if (k == n - 1) {
#pragma unroll
            for (int i = 0; i < NC; ++i) { A[i] = 0.0; C[i] = 0.0; B[i] = (i == c) ? 1.0 : 0.0; }
        }
        if (k == 0) { A[NS] = 0.0; C[NS] = 0.0; B[NS] = (c == NS) ? 1.0 : 0.0; }
        if (J) {
            double* Jr = J + ((long)prob * n + k) * 147;
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                Jr[i * 7 + c] = A[i];
                Jr[49 + i * 7 + c] = B[i];
                Jr[98 + i * 7 + c] = C[i];
            }
        }
    }
}

// L2 projection of -dphi/dx onto P1 (dolfin project(-grad(u_p), W), 1D:802-803): tridiagonal
// consistent-mass solve, one thread per problem (post-processing, off the hot path).
__global__ void field1d_kernel(int batch, int n, const double* __restrict__ x, const double* __restrict__ u,
                               double* __restrict__ field, double* __restrict__ scratch) {
    const int prob = blockIdx.x * blockDim.x + threadIdx.x;
    if (prob >= batch) return;
    const double* up = u + (long)prob * n * NC;
    double* f = field + (long)prob * n;
    double* cp = scratch + (long)prob * n;
    // rows: (h_{k-1}/6) g_{k-1} + ((h_{k-1}+h_k)/3) g_k + (h_k/6) g_{k+1} = b_k
    double cprev = 0.0, dprev = 0.0;
    for (int k = 0; k < n; ++k) {
        const double hm = (k > 0) ? x[k] - x[k - 1] : 0.0;
        const double hp = (k + 1 < n) ? x[k + 1] - x[k] : 0.0;
        double b = 0.0;
        if (k > 0) b += -0.5 * (up[(long)k * NC + NS] - up[(long)(k - 1) * NC + NS]);
        if (k + 1 < n) b += -0.5 * (up[(long)(k + 1) * NC + NS] - up[(long)k * NC + NS]);
        const double a = hm / 6.0, d = (hm + hp) / 3.0, cc = hp / 6.0;
        const double den = d - a * cprev;
        cprev = cc / den;
        dprev = (b - a * dprev) / den;
        cp[k] = cprev;
        f[k] = dprev;
    }
    for (int k = n - 2; k >= 0; --k) f[k] -= cp[k] * f[k + 1];
}

// Projected field at the OHP only (node 0 of project(-grad(u_p), W), 1D:802-805: the quantity the reference's result
// table holds), for per-point sweep summaries.  The P1 mass matrix is strictly diagonally dominant, so eliminating from
// the far end towards node 0 (one sweep, no back substitution) only needs the first K nodes: the influence of node K on
// node 0 is prod_k w_k <= 0.5^K (K = 256: far below round-off).  The elimination weights w_k depend on the mesh only:
// thread 0 of every CTA computes them once into shared memory, then one thread per problem runs K fused multiply-adds.
constexpr int FIELD_K = 256;
__global__ void __launch_bounds__(128)
field_ohp_kernel(int batch, int n, const double* __restrict__ x, const double* __restrict__ u, double* __restrict__ out) {
    __shared__ double w[FIELD_K];      // w_k = c_k / d'_(k+1)
    __shared__ double inv_d0;
    const int K = min(n, FIELD_K);
    if (threadIdx.x == 0) {
        double dprime = 0.0;
        for (int k = K - 1; k >= 0; --k) {
            const double hm = (k > 0) ? x[k] - x[k - 1] : 0.0;
            const double hp = (k + 1 < n) ? x[k + 1] - x[k] : 0.0;
            const double d = (hm + hp) / 3.0, c = hp / 6.0;
            // row k+1 (already reduced): a_(k+1) = hp / 6, diagonal dprime
            w[k] = (k + 1 < K) ? c / dprime : 0.0;
            dprime = d - w[k] * ((k + 1 < K) ? hp / 6.0 : 0.0);
        }
        inv_d0 = 1.0 / dprime;
    }
    __syncthreads();
    const int prob = blockIdx.x * blockDim.x + threadIdx.x;
    if (prob >= batch) return;
    const double* up = u + (long)prob * n * NC + NS;          // potential component
    // b_k = -(1/2) (phi_(k+1) - phi_(k-1)) with one-sided ends (the right-hand side of field1d_kernel)
    double bprime = 0.0;
    double phi_p = (K < n) ? up[(long)K * NC] : 0.0;          // phi_(k+1), starting at k = K-1
    double phi_c = up[(long)(K - 1) * NC];
    for (int k = K - 1; k >= 0; --k) {
        const double phi_m = (k > 0) ? up[(long)(k - 1) * NC] : 0.0;
        double b = 0.0;
        if (k > 0) b += -0.5 * (phi_c - phi_m);
        if (k + 1 < n) b += -0.5 * (phi_p - phi_c);
        bprime = b - w[k] * bprime;
        phi_p = phi_c; phi_c = phi_m;
    }
    out[prob] = bprime * inv_d0;
}

}  // namespace edl1d

// ---------------------------------------------------------------------------------------
// host launchers (called from capi.cu)
// ---------------------------------------------------------------------------------------
template <bool PIV, int NQ>
static cudaError_t launch_newton_variant(gmpnp_handle* h, int blocks, size_t smem, int mode, double* d_u, double* d_un_rw,
                                         const double* d_un_ro, const gmpnp_newton_opts* opts, int n_stage,
                                         const double* d_Vpath, double* d_hist, int* d_iters, double* d_r0, double* d_r,
                                         double* d_hfrac, int* d_stage, int* d_status, cudaStream_t st) {
    using namespace edl1d;
    cudaError_t e = cudaFuncSetAttribute(newton1d_kernel<PIV, NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    newton1d_kernel<PIV, NQ><<<blocks, NW_THREADS, smem, st>>>(mode, h->batch, h->n_nodes, h->d_x, h->d_params, d_u, d_un_rw,
        d_un_ro, h->d_ws, *opts, n_stage, d_Vpath, d_hist, d_iters, d_r0, d_r, d_hfrac, d_stage, d_status);
    return cudaGetLastError();
}

template <bool PIV, int NQ, int S>
static cudaError_t launch_part_variant(gmpnp_handle* h, int mode, double* d_u, double* d_un_rw, const double* d_un_ro,
                                       const gmpnp_newton_opts* opts, int n_stage, const double* d_Vpath, double* d_hist,
                                       int* d_iters, double* d_r0, double* d_r, double* d_hfrac, int* d_stage, int* d_status,
                                       cudaStream_t st) {
    using namespace edl1d;
    const size_t smem = PART_SMEM_DOUBLES * sizeof(double);
    const int ppb = 8 / S, blocks = (h->batch + ppb - 1) / ppb;
    cudaError_t e = cudaFuncSetAttribute(newton1d_part_kernel<PIV, NQ, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    newton1d_part_kernel<PIV, NQ, S><<<blocks, NW_THREADS, smem, st>>>(mode, h->batch, h->n_nodes, h->d_x, h->d_params, d_u,
        d_un_rw, d_un_ro, h->d_ws, h->d_ws2, *opts, n_stage, d_Vpath, d_hist, d_iters, d_r0, d_r, d_hfrac, d_stage, d_status);
    return cudaGetLastError();
}

int edl1d_launch_newton(gmpnp_handle* h, int mode, double* d_u, double* d_un_rw, const double* d_un_ro,
                        const gmpnp_newton_opts* opts, int n_stage, const double* d_Vpath, double* d_hist,
                        int* d_iters, double* d_r0, double* d_r, double* d_hfrac, int* d_stage,
                        int* d_status, cudaStream_t st) {
    using namespace edl1d;
    GmpnpRange nvtx_range(mode == 0 ? "gmpnp:newton_1d" : mode == 1 ? "gmpnp:march_1d" : "gmpnp:steady_continuation_1d");
    const bool consistent = (opts->jac_rule == 1);
    // sweeps per problem: 2 = two-sided elimination (throughput), 4 / 8 = partitioned elimination (latency; needs at
    // least 3 rows per sweep).  0 = automatic: partition when this launch alone cannot fill the 148 SMs.
    int S = opts->partitions;
    if (S == 0) S = (h->batch * 4 <= 148) ? 8 : (h->batch * 2 <= 148) ? 4 : 2;
    if (S != 2 && S != 4 && S != 8) return GMPNP_ERR_ARG;
    while (S > 2 && h->n_nodes < 4 * (S / 2 + 1)) S >>= 1;
    cudaError_t e;
    if (S == 2) {
        const int blocks = (h->batch + PROBLEMS_PER_BLOCK - 1) / PROBLEMS_PER_BLOCK;
        const size_t smem = NEWTON_SMEM_DOUBLES * sizeof(double);
#define GMPNP_ARGS h, blocks, smem, mode, d_u, d_un_rw, d_un_ro, opts, n_stage, d_Vpath, d_hist, d_iters, d_r0, d_r, d_hfrac, d_stage, d_status, st
        if (opts->pivot) e = consistent ? launch_newton_variant<true, 2>(GMPNP_ARGS) : launch_newton_variant<true, 3>(GMPNP_ARGS);
        else             e = consistent ? launch_newton_variant<false, 2>(GMPNP_ARGS) : launch_newton_variant<false, 3>(GMPNP_ARGS);
#undef GMPNP_ARGS
    } else {
        if (!h->d_ws2) GMPNP_CUDA_TRY(h, cudaMalloc(&h->d_ws2, sizeof(double) * 56 * (size_t)h->n_nodes * h->batch));
#define GMPNP_ARGS h, mode, d_u, d_un_rw, d_un_ro, opts, n_stage, d_Vpath, d_hist, d_iters, d_r0, d_r, d_hfrac, d_stage, d_status, st
#define GMPNP_PART(PIV, NQ) (S == 4 ? launch_part_variant<PIV, NQ, 4>(GMPNP_ARGS) : launch_part_variant<PIV, NQ, 8>(GMPNP_ARGS))
        if (opts->pivot) e = consistent ? GMPNP_PART(true, 2) : GMPNP_PART(true, 3);
        else             e = consistent ? GMPNP_PART(false, 2) : GMPNP_PART(false, 3);
#undef GMPNP_PART
#undef GMPNP_ARGS
    }
    h->launches++;
    GMPNP_CUDA_TRY(h, e);
    return GMPNP_OK;
}

int edl1d_launch_assemble(gmpnp_handle* h, const double* d_u, const double* d_un, double* d_F, double* d_J,
                          cudaStream_t st) {
    using namespace edl1d;
    const long items = (long)h->batch * h->n_nodes;
    const int blocks = (int)((items + GROUPS_PER_BLOCK - 1) / GROUPS_PER_BLOCK);
    const size_t smem = (size_t)GROUPS_PER_BLOCK * AS_GROUP * sizeof(double);
    assemble1d_kernel<<<blocks, THREADS, smem, st>>>(h->batch, h->n_nodes, h->d_x, h->d_params, d_u, d_un, d_F, d_J);
    h->launches++;
    GMPNP_CUDA_TRY(h, cudaGetLastError());
    return GMPNP_OK;
}

int edl1d_launch_field_ohp(gmpnp_handle* h, const double* d_u, double* d_out, cudaStream_t st) {
    using namespace edl1d;
    field_ohp_kernel<<<(h->batch + 127) / 128, 128, 0, st>>>(h->batch, h->n_nodes, h->d_x, d_u, d_out);
    h->launches++;
    GMPNP_CUDA_TRY(h, cudaGetLastError());
    return GMPNP_OK;
}

int edl1d_launch_field(gmpnp_handle* h, const double* d_u, double* d_field, cudaStream_t st) {
    using namespace edl1d;
    const int threads = 64;
    const int blocks = (h->batch + threads - 1) / threads;
    // the elimination workspace doubles as scratch (n doubles per problem needed)
    field1d_kernel<<<blocks, threads, 0, st>>>(h->batch, h->n_nodes, h->d_x, d_u, d_field, h->d_ws);
    h->launches++;
    GMPNP_CUDA_TRY(h, cudaGetLastError());
    return GMPNP_OK;
}
