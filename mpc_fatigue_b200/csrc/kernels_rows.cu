// kernels_rows.cu — fused reference-mode OCP node rows: every constraint row and the running cost that the reference's
// scripts build per shooting node around the bridge Functions, in ONE launch per batch, plus the first derivatives of the
// kinematic rows.  One thread per (scenario, node) unit, node-major units u = k * B + b.
//
//   one arm  (python/Pilz_6_DOF/force_optimization_pilz_6DOF.py:130-177, python/Pilz_3_DOF/*.py): rows
//        [0,3)      ee_pos - p_ref                 (the line constraint uses its x, y components, :152-156)
//        [3,3+n)    tau = ID(q, qd, 0) + wsign J^T [F; 0]                                          (:134)
//        then n     Euler defect q + h qd - q_next                                                 (:160,170)
//        then n     thermal defect T+ - T_next  (zero-order hold, mpc_principal.py:296-301,325)
//        cost       w_F F.F + w_qd qd.qd        (w_F = -1: the script's -F^T F, :177)
//   two arms (python/2_pilz_6_DOF/Box_Pilz_6DOF2.py:244-293,454-475; python/Centauro_script/mpc_principal.py:229-327;
//             RepeatedMPCwithThermal_confriction.py:251-273): rows
//        [0,3)      force equilibrium     F_L + F_R - fdes                                         (mpc_principal.py:229)
//        [3,6)      moment equilibrium    (pL - pR) x F_L + (pR - pL) x F_R                        (:234)
//        [6]        |pL - pR|^2 - dist2_ref                                                        (Box_Pilz_6DOF2.py:283-285)
//        [7,10)     relative position     R_L^T (pR - pL) - (the same at the previous node | rel_pos0)   (:241-246)
//        [10,13)    relative orientation  e(R_L R_R^T) - rel_ori0,  e = (S21, S20, S10), S = (Ro - Ro^T)/2   (:248-258)
//        [13,23)    friction cones        A1 (-R_L^T F_L), A2 (-R_R^T F_R)  <= 0                   (confriction.py:251-273)
//        [23,26)    p_box - p_ref,  p_box = (pL + pR) / 2
//        then tau (n), Euler defect (n), thermal defect (n) as above, with tau = ID + wsign (J_L^T [F_L;0] + J_R^T [F_R;0])
//        cost       w_box |p_box - p_ref|^2 + w_qd qd.qd + w_F (F_L.F_L + F_R.F_R)                 (mpc_principal.py:281-284)
//
// The kinematic rows are written once, generic over the scalar (double | Dual); k_rows_kin_jac evaluates them with one
// dual-number seed per launch row (blockIdx.y = direction in (q, F)) and gives d rows / d (q, F).  The torque rows'
// derivatives come from mpcf_node_eval_ref_jvp_batch (analytic) and d tau / d F, d T+ / d tau are emitted here in closed form.
// Static families only (chain3/6/7: one arm; forest12x6 / forest14x7: two arms).  Oracle twin: oracle/core.inc.h: ocp_rows.
#include "launch.cuh"

namespace mpcf {

struct RowsArgs {
    int narm, n, N;
    long B;
    int ee_joint[2];     // end-effector joint, local to its arm
    double ee_p[2][3], ee_R[2][9];
    double wsign, fdes[3], dist2_ref, mu, p_ref[3], w_box, w_qd, w_F, h;
};

// world pose of the end-effector frame of one serial arm and, for T = double callers that ask, the joint axes / origins
template <class T, int L>
struct ArmFk {
    T pos[3], rot[9];
    T z[L][3], o[L][3];
    MPCF_DI void run(const StaticParams<L> &P, const T *q, int ej, const double *ep, const double *eR)
    {
        T R[9], p[3];
#pragma unroll
        for (int i = 0; i < L; ++i) {
            T s, c;
            sincos_t(q[i], s, c);
            T Rl[9];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                Rl[3 * r + 0] = P.Rp[i][3 * r] * c + P.Rp[i][3 * r + 1] * s;
                Rl[3 * r + 1] = P.Rp[i][3 * r + 1] * c - P.Rp[i][3 * r] * s;
                Rl[3 * r + 2] = T(P.Rp[i][3 * r + 2]);
            }
            if (i == 0) {
#pragma unroll
                for (int k = 0; k < 9; ++k) R[k] = Rl[k];
#pragma unroll
                for (int k = 0; k < 3; ++k) p[k] = T(P.pp[i][k]);
            } else {
                T Rn[9];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    p[r] = p[r] + (R[3 * r] * P.pp[i][0] + R[3 * r + 1] * P.pp[i][1] + R[3 * r + 2] * P.pp[i][2]);
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) Rn[3 * r + cc] = R[3 * r] * Rl[cc] + R[3 * r + 1] * Rl[3 + cc] + R[3 * r + 2] * Rl[6 + cc];
                }
#pragma unroll
                for (int k = 0; k < 9; ++k) R[k] = Rn[k];
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) { z[i][k] = R[3 * k + 2]; o[i][k] = p[k]; }
            if (i == ej) {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    pos[r] = p[r] + (R[3 * r] * ep[0] + R[3 * r + 1] * ep[1] + R[3 * r + 2] * ep[2]);
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) rot[3 * r + cc] = R[3 * r] * eR[cc] + R[3 * r + 1] * eR[3 + cc] + R[3 * r + 2] * eR[6 + cc];
                }
            }
        }
    }
};

// the 26 kinematic rows of the two-arm node (header comment) from the two end-effector poses and the two forces
template <class T>
MPCF_DI void box_kin_rows(const RowsArgs &a, const T *pL, const T *RL, const T *pR, const T *RR, const T *FL, const T *FR, const T *relprev,
                          const double *ori0, T *g)
{
    T d[3] = {pL[0] - pR[0], pL[1] - pR[1], pL[2] - pR[2]};
#pragma unroll
    for (int k = 0; k < 3; ++k) g[k] = FL[k] + FR[k] - a.fdes[k];
    T dF[3] = {FL[0] - FR[0], FL[1] - FR[1], FL[2] - FR[2]};  // (pL - pR) x F_L + (pR - pL) x F_R = (pL - pR) x (F_L - F_R)
    cross3(d, dF, g + 3);
    g[6] = d[0] * d[0] + d[1] * d[1] + d[2] * d[2] - a.dist2_ref;
#pragma unroll
    for (int k = 0; k < 3; ++k) g[7 + k] = -(RL[k] * d[0] + RL[3 + k] * d[1] + RL[6 + k] * d[2]) - relprev[k];  // R_L^T (pR - pL)
    // Ro = R_L R_R^T ; e = ((Ro21 - Ro12)/2, (Ro20 - Ro02)/2, (Ro10 - Ro01)/2)
    auto Ro = [&](int r, int c) { return RL[3 * r] * RR[3 * c] + RL[3 * r + 1] * RR[3 * c + 1] + RL[3 * r + 2] * RR[3 * c + 2]; };
    g[10] = 0.5 * (Ro(2, 1) - Ro(1, 2)) - ori0[0];
    g[11] = 0.5 * (Ro(2, 0) - Ro(0, 2)) - ori0[1];
    g[12] = 0.5 * (Ro(1, 0) - Ro(0, 1)) - ori0[2];
    // friction cones on the forces the environment exerts, in the end-effector frames: f1 = -R_L^T F_L, f2 = -R_R^T F_R
    T f1[3], f2[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        f1[k] = -(RL[k] * FL[0] + RL[3 + k] * FL[1] + RL[6 + k] * FL[2]);
        f2[k] = -(RR[k] * FR[0] + RR[3 + k] * FR[1] + RR[6 + k] * FR[2]);
    }
    const double mu = a.mu;
    g[13] = -f1[1];
    g[14] = f1[2] - mu * f1[1];
    g[15] = -f1[2] - mu * f1[1];
    g[16] = f1[0] - mu * f1[1];
    g[17] = -f1[0] - mu * f1[1];
    g[18] = f2[1];
    g[19] = f2[0] + mu * f2[1];
    g[20] = -f2[0] + mu * f2[1];
    g[21] = f2[2] + mu * f2[1];
    g[22] = -f2[2] + mu * f2[1];
#pragma unroll
    for (int k = 0; k < 3; ++k) g[23 + k] = 0.5 * (pL[k] + pR[k]) - a.p_ref[k];
}

template <int L, int NARM>
__global__ void __launch_bounds__(kThreads) k_ocp_rows(const __grid_constant__ StaticParams<L> P0, const __grid_constant__ StaticParams<L> P1,
                                                      RowsArgs a, const double *q, const double *qd, const double *F, const double *T,
                                                      const double *q_last, const double *T_last, const double *rel_pos0,
                                                      const double *rel_ori0, double *rows, double *cost, double *dtau_dF, double *dT_dtau)
{
    const long U = a.B * a.N;
    const long u = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= U) return;
    constexpr int n = NARM * L;
    const long b = u % a.B;
    const int k = (int)(u / a.B);
    constexpr int KIN = NARM == 2 ? 26 : 3;
    double Fv[3 * NARM];
#pragma unroll
    for (int i = 0; i < 3 * NARM; ++i) Fv[i] = F[(long)i * U + u];
    double ckin = 0.0, cqd = 0.0, cF = 0.0;
#pragma unroll
    for (int i = 0; i < 3 * NARM; ++i) cF += Fv[i] * Fv[i];
    double pe[NARM][3], Re[NARM][9];
#pragma unroll
    for (int c = 0; c < NARM; ++c) {
        const StaticParams<L> &P = c == 0 ? P0 : P1;
        const StaticModel<L, L> m{P};
        double qa[L], qda[L], zero[L], tau[L];
#pragma unroll
        for (int i = 0; i < L; ++i) { qa[i] = q[(long)(c * L + i) * U + u]; qda[i] = qd[(long)(c * L + i) * U + u]; zero[i] = 0.0; cqd += qda[i] * qda[i]; }
        ArmFk<double, L> K;
        K.run(P, qa, a.ee_joint[c], a.ee_p[c], a.ee_R[c]);
#pragma unroll
        for (int r = 0; r < 3; ++r) pe[c][r] = K.pos[r];
#pragma unroll
        for (int r = 0; r < 9; ++r) Re[c][r] = K.rot[r];
        Dyn<double, StaticModel<L, L>>::rnea(m, qa, qda, zero, tau);
        const double *Fc = Fv + 3 * c;
#pragma unroll
        for (int i = 0; i < L; ++i) {
            // column i of the frame Jacobian (linear part): z_i x (p_ee - o_i), zero past the end-effector joint
            double col[3] = {0.0, 0.0, 0.0};
            if (i <= a.ee_joint[c]) {
                const double d[3] = {K.pos[0] - K.o[i][0], K.pos[1] - K.o[i][1], K.pos[2] - K.o[i][2]};
                cross3(K.z[i], d, col);
            }
            const double t = tau[i] + a.wsign * (col[0] * Fc[0] + col[1] * Fc[1] + col[2] * Fc[2]);
            const long row = KIN + c * L + i;
            rows[row * U + u] = t;
            if (dtau_dF) {
#pragma unroll
                for (int r = 0; r < 3; ++r) dtau_dF[((long)(c * L + i) * 3 * NARM + 3 * c + r) * U + u] = a.wsign * col[r];
                if (NARM == 2) {
#pragma unroll
                    for (int r = 0; r < 3; ++r) dtau_dF[((long)(c * L + i) * 3 * NARM + 3 * (1 - c) + r) * U + u] = 0.0;
                }
            }
            // Euler defect against the next node's q (last node: the trailing q_N)
            const double qn = qa[i] + a.h * qda[i];
            const double qnext = k + 1 < a.N ? q[(long)(c * L + i) * U + u + a.B] : (q_last ? q_last[(long)(c * L + i) * a.B + b] : qn);
            rows[(row + n) * U + u] = qn - qnext;
            double Tdef = 0.0, dTdt = 0.0;
            if (T) {
                const double lam = P.fat[i][0], kap = P.fat[i][1];
                const double Pl = P.fat[i][2] * t * t + P.fat[i][3] * qda[i] * qda[i];
                const double Ti = T[(long)(c * L + i) * U + u];
                const double za = exp(-lam * a.h);
                const double gain = lam == 0.0 ? a.h * kap : (1.0 - za) * (kap / lam);
                const double Tn = za * Ti + gain * Pl;
                const double Tnext = k + 1 < a.N ? T[(long)(c * L + i) * U + u + a.B] : (T_last ? T_last[(long)(c * L + i) * a.B + b] : Tn);
                Tdef = Tn - Tnext;
                dTdt = gain * 2.0 * P.fat[i][2] * t;
            }
            rows[(row + 2 * n) * U + u] = Tdef;
            if (dT_dtau) dT_dtau[(long)(c * L + i) * U + u] = dTdt;
        }
    }
    if (NARM == 2) {
        double relprev[3], ori0[3], g[26];
        if (k == 0) {
#pragma unroll
            for (int r = 0; r < 3; ++r) relprev[r] = rel_pos0 ? rel_pos0[(long)r * a.B + b] : 0.0;
        } else {  // the same quantity at the previous node (its unit is u - B): frame kinematics only
            double pp[2][3], Rl0[9];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const StaticParams<L> &P = c == 0 ? P0 : P1;
                double qa[L];
#pragma unroll
                for (int i = 0; i < L; ++i) qa[i] = q[(long)(c * L + i) * U + u - a.B];
                ArmFk<double, L> K;
                K.run(P, qa, a.ee_joint[c], a.ee_p[c], a.ee_R[c]);
#pragma unroll
                for (int r = 0; r < 3; ++r) pp[c][r] = K.pos[r];
                if (c == 0) {
#pragma unroll
                    for (int r = 0; r < 9; ++r) Rl0[r] = K.rot[r];
                }
            }
#pragma unroll
            for (int r = 0; r < 3; ++r) relprev[r] = Rl0[r] * (pp[1][0] - pp[0][0]) + Rl0[3 + r] * (pp[1][1] - pp[0][1]) + Rl0[6 + r] * (pp[1][2] - pp[0][2]);
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) ori0[r] = rel_ori0 ? rel_ori0[(long)r * a.B + b] : 0.0;
        box_kin_rows<double>(a, pe[0], Re[0], pe[NARM - 1], Re[NARM - 1], Fv, Fv + 3 * (NARM - 1), relprev, ori0, g);
#pragma unroll
        for (int r = 0; r < 26; ++r) rows[(long)r * U + u] = g[r];
        ckin = g[23] * g[23] + g[24] * g[24] + g[25] * g[25];
    } else {
#pragma unroll
        for (int r = 0; r < 3; ++r) rows[(long)r * U + u] = pe[0][r] - a.p_ref[r];
    }
    cost[u] = a.w_box * ckin + a.w_qd * cqd + a.w_F * cF;
}

// d (kinematic rows) / d (q, F): one dual-number seed per blockIdx.y (d < n: q_d, else F_{d - n}); out[row][n + 3 NARM][U].
// The relative-position row's dependence on the PREVIOUS node's q is the negative of the same block evaluated there.
template <int L, int NARM>
__global__ void __launch_bounds__(kThreads) k_rows_kin_jac(const __grid_constant__ StaticParams<L> P0, const __grid_constant__ StaticParams<L> P1,
                                                          RowsArgs a, const double *q, const double *F, double *out)
{
    const long U = a.B * a.N;
    const long u = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= U) return;
    constexpr int n = NARM * L, ND = n + 3 * NARM;
    constexpr int KIN = NARM == 2 ? 26 : 3;
    const int d = blockIdx.y;
    Dual Fv[3 * NARM], pe[NARM][3], Re[NARM][9];
#pragma unroll
    for (int i = 0; i < 3 * NARM; ++i) Fv[i] = Dual(F[(long)i * U + u], d == n + i ? 1.0 : 0.0);
#pragma unroll
    for (int c = 0; c < NARM; ++c) {
        const StaticParams<L> &P = c == 0 ? P0 : P1;
        Dual qa[L];
#pragma unroll
        for (int i = 0; i < L; ++i) qa[i] = Dual(q[(long)(c * L + i) * U + u], d == c * L + i ? 1.0 : 0.0);
        ArmFk<Dual, L> K;
        K.run(P, qa, a.ee_joint[c], a.ee_p[c], a.ee_R[c]);
#pragma unroll
        for (int r = 0; r < 3; ++r) pe[c][r] = K.pos[r];
#pragma unroll
        for (int r = 0; r < 9; ++r) Re[c][r] = K.rot[r];
    }
    Dual g[KIN];
    if (NARM == 2) {
        const Dual relprev[3] = {Dual(0.0), Dual(0.0), Dual(0.0)};
        const double ori0[3] = {0.0, 0.0, 0.0};
        box_kin_rows<Dual>(a, pe[0], Re[0], pe[NARM - 1], Re[NARM - 1], Fv, Fv + 3 * (NARM - 1), relprev, ori0, g);
    } else {
#pragma unroll
        for (int r = 0; r < 3; ++r) g[r] = pe[0][r];
    }
#pragma unroll
    for (int r = 0; r < KIN; ++r) out[((long)r * ND + d) * U + u] = g[r].d;
}

template <int L, int NARM>
static cudaError_t rows_launch(const StaticParams<L> *cp, const RowsArgs &a, const double *q, const double *qd, const double *F, const double *T,
                               const double *q_last, const double *T_last, const double *rel_pos0, const double *rel_ori0, double *rows,
                               double *cost, double *dtau_dF, double *dT_dtau, double *kin_jac, cudaStream_t s)
{
    const long U = a.B * a.N;
    const unsigned gb = (unsigned)((U + kThreads - 1) / kThreads);
    const StaticParams<L> &P1 = cp[NARM - 1];
    k_ocp_rows<L, NARM><<<gb, kThreads, 0, s>>>(cp[0], P1, a, q, qd, F, T, q_last, T_last, rel_pos0, rel_ori0, rows, cost, dtau_dF, dT_dtau);
    g_launches.fetch_add(1);
    if (kin_jac) {
        k_rows_kin_jac<L, NARM><<<dim3(gb, NARM * L + 3 * NARM), kThreads, 0, s>>>(cp[0], P1, a, q, F, kin_jac);
        g_launches.fetch_add(1);
    }
    return cudaGetLastError();
}

cudaError_t launch_ocp_rows(const LaunchModel &m, const RowsHost &h, const double *q, const double *qd, const double *F, const double *T,
                            const double *q_last, const double *T_last, const double *rel_pos0, const double *rel_ori0, double *rows,
                            double *cost, double *dtau_dF, double *dT_dtau, double *kin_jac, cudaStream_t s)
{
    if (h.B <= 0 || h.N <= 0) return cudaSuccess;
    RowsArgs a;
    a.narm = h.narm; a.n = m.n; a.N = h.N; a.B = h.B;
    for (int c = 0; c < 2; ++c) {
        a.ee_joint[c] = h.ee_joint[c];
        for (int k = 0; k < 3; ++k) a.ee_p[c][k] = h.ee_p[c][k];
        for (int k = 0; k < 9; ++k) a.ee_R[c][k] = h.ee_R[c][k];
    }
    a.wsign = h.wsign; a.dist2_ref = h.dist2_ref; a.mu = h.mu; a.w_box = h.w_box; a.w_qd = h.w_qd; a.w_F = h.w_F; a.h = h.h;
    for (int k = 0; k < 3; ++k) { a.fdes[k] = h.fdes[k]; a.p_ref[k] = h.p_ref[k]; }
#define ROWS_ARGS a, q, qd, F, T, q_last, T_last, rel_pos0, rel_ori0, rows, cost, dtau_dF, dT_dtau, kin_jac, s
    switch (m.fam) {
    case FAM_CHAIN3: return rows_launch<3, 1>(static_cast<const StaticParams<3> *>(m.static_params), ROWS_ARGS);
    case FAM_CHAIN6: return rows_launch<6, 1>(static_cast<const StaticParams<6> *>(m.static_params), ROWS_ARGS);
    case FAM_CHAIN7: return rows_launch<7, 1>(static_cast<const StaticParams<7> *>(m.static_params), ROWS_ARGS);
    case FAM_FOREST12x6: return rows_launch<6, 2>(static_cast<const StaticParams<6> *>(m.chain_params), ROWS_ARGS);
    case FAM_FOREST14x7: return rows_launch<7, 2>(static_cast<const StaticParams<7> *>(m.chain_params), ROWS_ARGS);
    default: return cudaErrorInvalidValue;
    }
#undef ROWS_ARGS
}

}  // namespace mpcf
