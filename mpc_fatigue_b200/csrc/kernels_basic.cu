// kernels_basic.cu — RNEA, frame FK, frame Jacobian, reference-mode node evaluation, cost/residual reduction.
#include "launch.cuh"

namespace mpcf {

std::atomic<long> g_launches{0};
long launch_count() { return g_launches.load(); }

struct RneaBody {
    template <class MP>
    static MPCF_DI void run(const MP &m, long u, long U, const double *q, const double *qd, const double *qdd, double *tau)
    {
        constexpr int UNR = MP::kStatic ? MP::MAXN : 1;
        const int n = m.n();
        double a[MP::MAXN], b[MP::MAXN], c[MP::MAXN], t[MP::MAXN];
#pragma unroll UNR
        for (int i = 0; i < n; ++i) {
            a[i] = q[i * U + u];
            b[i] = qd[i * U + u];
            c[i] = qdd ? qdd[i * U + u] : 0.0;
        }
        Dyn<double, MP>::rnea(m, a, b, c, t);
#pragma unroll UNR
        for (int i = 0; i < n; ++i) tau[i * U + u] = t[i];
    }
};

struct FkBody {
    template <class MP>
    static MPCF_DI void run(const MP &m, long u, long U, FrameArg f, const double *q, double *pos, double *rot)
    {
        constexpr int UNR = MP::kStatic ? MP::MAXN : 1;
        const int n = m.n();
        double a[MP::MAXN], oR[MP::MAXN][9], op[MP::MAXN][3], p[3], R[9];
#pragma unroll UNR
        for (int i = 0; i < n; ++i) a[i] = q[i * U + u];
        Dyn<double, MP>::fk_all(m, a, oR, op);
        Dyn<double, MP>::frame_pose(f.joint, f.R, f.p, oR, op, p, R);
#pragma unroll
        for (int k = 0; k < 3; ++k) pos[k * U + u] = p[k];
#pragma unroll
        for (int k = 0; k < 9; ++k) rot[k * U + u] = R[k];
    }
};

// column i of the LOCAL_WORLD_ALIGNED frame Jacobian: [z_i x (p_f - o_i); z_i] (revolute), [z_i; 0] (prismatic)
template <class MP>
MPCF_DI void jac_column(const MP &m, int i, double (*oR)[9], double (*op)[3], const double *pf, double *col)
{
    const double z[3] = {oR[i][2], oR[i][5], oR[i][8]};
    if (!m.prismatic(i)) {
        const double d[3] = {pf[0] - op[i][0], pf[1] - op[i][1], pf[2] - op[i][2]};
        cross3(z, d, col);
        col[3] = z[0]; col[4] = z[1]; col[5] = z[2];
    } else {
        col[0] = z[0]; col[1] = z[1]; col[2] = z[2];
        col[3] = 0.0; col[4] = 0.0; col[5] = 0.0;
    }
}

struct JacBody {
    template <class MP>
    static MPCF_DI void run(const MP &m, long u, long U, FrameArg f, const double *q, double *J)
    {
        constexpr int UNR = MP::kStatic ? MP::MAXN : 1;
        const int n = m.n();
        double a[MP::MAXN], oR[MP::MAXN][9], op[MP::MAXN][3], p[3], R[9];
#pragma unroll UNR
        for (int i = 0; i < n; ++i) a[i] = q[i * U + u];
        Dyn<double, MP>::fk_all(m, a, oR, op);
        Dyn<double, MP>::frame_pose(f.joint, f.R, f.p, oR, op, p, R);
        int cur = f.joint;
#pragma unroll UNR
        for (int i = n - 1; i >= 0; --i) {
            double col[6] = {0, 0, 0, 0, 0, 0};
            if (i == cur) {
                jac_column(m, i, oR, op, p, col);
                cur = m.parent(i);
            }
#pragma unroll
            for (int r = 0; r < 6; ++r) J[(long)(r * n + i) * U + u] = col[r];
        }
    }
};

// Reference-mode node evaluation, one launch:
//   tau = RNEA(q, qd, qdd) + wsign * sum_e J_e^T W_e ; qnext = q + h qd ; Tnext = ZOH(T; tau, qd, h)
// (python/2_pilz_6_DOF/Box_Pilz_6DOF2.py:290-293,463 ; python/Centauro_script/mpc_principal.py:267-301)
// jtw_only: skip RNEA and the integrators and write only J^T W (mpcf_frame_jac_t_wrench_batch).
struct NodeEvalBody {
    template <class MP>
    static MPCF_DI void run(const MP &m, long u, long U, EeArgs ee, double wsign, const double *q, const double *qd,
                            const double *qdd, const double *W, const double *T, double h, ZohArg zoh, double *tau,
                            double *qnext, double *Tnext, bool jtw_only)
    {
        constexpr int UNR = MP::kStatic ? MP::MAXN : 1;
        const int n = m.n();
        double a[MP::MAXN], b[MP::MAXN], c[MP::MAXN], t[MP::MAXN];
#pragma unroll UNR
        for (int i = 0; i < n; ++i) {
            a[i] = q[i * U + u];
            b[i] = (!jtw_only && qd) ? qd[i * U + u] : 0.0;
            c[i] = (!jtw_only && qdd) ? qdd[i * U + u] : 0.0;
            t[i] = 0.0;
        }
        JointVar<double> jv[MP::MAXN];  // one sincos per joint, shared by the dynamics and the frame kinematics
#pragma unroll UNR
        for (int i = 0; i < n; ++i) Dyn<double, MP>::joint_var(m, i, a[i], jv[i]);
        if (!jtw_only) Dyn<double, MP>::rnea_jv(m, jv, b, c, t);
        if (ee.nee > 0) {
            double oR[MP::MAXN][9], op[MP::MAXN][3];
            Dyn<double, MP>::fk_all_jv(m, jv, oR, op);
            for (int e = 0; e < ee.nee; ++e) {
                double pf[3], Rf[9], w[6];
                Dyn<double, MP>::frame_pose(ee.f[e].joint, ee.f[e].R, ee.f[e].p, oR, op, pf, Rf);
#pragma unroll
                for (int r = 0; r < 6; ++r) w[r] = W[(long)(6 * e + r) * U + u];
                int cur = ee.f[e].joint;
#pragma unroll UNR
                for (int i = n - 1; i >= 0; --i) {
                    if (i == cur) {
                        double col[6];
                        jac_column(m, i, oR, op, pf, col);
                        double acc = 0.0;
#pragma unroll
                        for (int r = 0; r < 6; ++r) acc += col[r] * w[r];
                        t[i] += wsign * acc;
                        cur = m.parent(i);
                    }
                }
            }
        }
#pragma unroll UNR
        for (int i = 0; i < n; ++i) {
            tau[i * U + u] = t[i];
            if (jtw_only) continue;
            if (qnext) qnext[i * U + u] = a[i] + h * b[i];
            if (Tnext) {
                const double P = m.fat(i, 2) * t[i] * t[i] + m.fat(i, 3) * b[i] * b[i];
                const double Ti = T[i * U + u];
                const double lam = m.fat(i, 0);
                Tnext[i * U + u] = (lam == 0.0) ? Ti + h * m.fat(i, 1) * P : zoh.a[i] * Ti + (1.0 - zoh.a[i]) * (m.fat(i, 1) / lam) * P;
            }
        }
    }
};

// First derivatives of the reference-mode joint torque tau = RNEA(q, qd, qdd) + wsign * sum_e J_e^T W_e with respect to
// q and qd: one dual-number sweep per seed (blockIdx.y in [0, 2n)).  d tau/d W_e = wsign * J_e^T comes from the frame
// Jacobian kernel.  These are the Jacobian blocks of the torque-bound rows of the reference's OCPs
// (python/Pilz_6_DOF/force_optimization_pilz_6DOF.py:134-148).
struct NodeEvalJvpBody {
    template <class MP>
    static MPCF_DI void run(const MP &m, long u, long U, EeArgs ee, double wsign, const double *q, const double *qd,
                            const double *qdd, const double *W, double *dtau_dq, double *dtau_dqd)
    {
        constexpr int UNR = MP::kStatic ? MP::MAXN : 1;
        using D = Dyn<Dual, MP>;
        const int n = m.n();
        const int d = blockIdx.y;
        Dual a[MP::MAXN], b[MP::MAXN], c[MP::MAXN], t[MP::MAXN];
#pragma unroll UNR
        for (int i = 0; i < n; ++i) {
            a[i] = Dual(q[i * U + u], d == i ? 1.0 : 0.0);
            b[i] = Dual(qd[i * U + u], d == n + i ? 1.0 : 0.0);
            c[i] = Dual(qdd ? qdd[i * U + u] : 0.0, 0.0);
        }
        JointVar<Dual> jv[MP::MAXN];
#pragma unroll UNR
        for (int i = 0; i < n; ++i) D::joint_var(m, i, a[i], jv[i]);
        D::rnea_jv(m, jv, b, c, t);
        if (ee.nee > 0 && d < n) {  // J^T W depends on q only
            Dual oR[MP::MAXN][9], op[MP::MAXN][3];
            D::fk_all_jv(m, jv, oR, op);
            for (int e = 0; e < ee.nee; ++e) {
                Dual pf[3], Rf[9];
                D::frame_pose(ee.f[e].joint, ee.f[e].R, ee.f[e].p, oR, op, pf, Rf);
                double w[6];
#pragma unroll
                for (int r = 0; r < 6; ++r) w[r] = W[(long)(6 * e + r) * U + u];
                int cur = ee.f[e].joint;
#pragma unroll UNR
                for (int i = n - 1; i >= 0; --i) {
                    if (i == cur) {
                        const Dual z[3] = {oR[i][2], oR[i][5], oR[i][8]};
                        Dual acc;
                        if (!m.prismatic(i)) {
                            const Dual dd[3] = {pf[0] - op[i][0], pf[1] - op[i][1], pf[2] - op[i][2]};
                            Dual lin[3];
                            cross3(z, dd, lin);
                            acc = lin[0] * w[0] + lin[1] * w[1] + lin[2] * w[2] + z[0] * w[3] + z[1] * w[4] + z[2] * w[5];
                        } else {
                            acc = z[0] * w[0] + z[1] * w[1] + z[2] * w[2];
                        }
                        t[i] += wsign * acc;
                        cur = m.parent(i);
                    }
                }
            }
        }
        double *out = d < n ? dtau_dq : dtau_dqd;
        const int col = d % n;
#pragma unroll UNR
        for (int r = 0; r < n; ++r) out[(size_t)(r * n + col) * U + u] = t[r].d;
    }
};

// ---------------------------------------------------------------------------------------------
// per-scenario cost / residual reduction: thread b walks the N nodes of scenario b (u = k*B + b,
// coalesced across b) and writes (cost, defect, torque-bound, fatigue-bound) into out[4][B]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cost_residual_kernel(int n, long B, int N, const double *q, const double *qd, const double *f,
                                                          const double *tau, const double *qn, const double *qdn, const double *fn,
                                                          CostArgs c, double *out)
{
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const long U = B * N;
    double cost = 0.0, r0 = 0.0, r1 = 0.0, r2 = 0.0;
    for (int k = 0; k < N; ++k) {
        const long u = (long)k * B + b;
        const double bound = fmax(c.tau0 * exp(-c.alpha * k * c.dt), c.tau_floor);
        for (int i = 0; i < n; ++i) {
            const long o = (long)i * U + u;
            const double v = qd[o], t = tau[o];
            cost += c.w_qd * v * v + c.w_tau * t * t;
            r1 = fmax(r1, fabs(t) - bound);
            r2 = fmax(r2, fn[o] - c.f_max);
            if (k + 1 < N) {
                const long o1 = o + B;
                r0 = fmax(r0, fabs(qn[o] - q[o1]));
                r0 = fmax(r0, fabs(qdn[o] - qd[o1]));
                r0 = fmax(r0, fabs(fn[o] - f[o1]));
            }
        }
    }
    out[b] = cost;
    out[B + b] = r0;
    out[2 * B + b] = r1;
    out[3 * B + b] = r2;
}

// FP64 pipe probe: 8 independent DFMA chains per thread; the roofline denominator of bench.py
// (MEASURED_PEAKS.json carries no FP64 figure).  flops = blocks * threads * iters * 8 * 2.
__global__ void __launch_bounds__(256) fp64_probe_kernel(long iters, double *out)
{
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    const double a = 0.999999, b = 1e-7;
    for (long i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[(long)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
cudaError_t launch_fp64_probe(long iters, int blocks, double *out, cudaStream_t s)
{
    fp64_probe_kernel<<<blocks, 256, 0, s>>>(iters, out);
    g_launches.fetch_add(1);
    return cudaGetLastError();
}

cudaError_t launch_rnea(const LaunchModel &m, long U, const double *q, const double *qd, const double *qdd, double *tau, cudaStream_t s)
{
    return dispatch<RneaBody>(m, U, 1, s, q, qd, qdd, tau);
}
cudaError_t launch_fk(const LaunchModel &m, const FrameArg &f, long U, const double *q, double *pos, double *rot, cudaStream_t s)
{
    return dispatch<FkBody>(m, U, 1, s, f, q, pos, rot);
}
cudaError_t launch_jac(const LaunchModel &m, const FrameArg &f, long U, const double *q, double *J, cudaStream_t s)
{
    return dispatch<JacBody>(m, U, 1, s, f, q, J);
}
// Contact wrenches as external link forces of the RNEA (serial chain): W_e = [F ; n] is given in world axes at the frame
// point carried by chain joint je[e] (LOCAL_WORLD_ALIGNED, bridge.hpp:144), so in the coordinates of that link it is
// [R^T F ; R^T n + p x R^T F] with R the link's world orientation and p the frame's offset in the link — only the rotation
// chain is needed, and it advances inside the RNEA's own forward sweep.  tau = ID + wsign * sum_e J_e^T W_e.
template <int N, int NEE>
struct WrenchExt {
    double R[9];  // world orientation of the link the sweep is at
    int je[NEE > 0 ? NEE : 1];
    double Wl[NEE > 0 ? NEE : 1][6], pl[NEE > 0 ? NEE : 1][3];
    double wsign;
    MPCF_DI void link(const StaticModel<N, N> &m, int i, const JointVar<double> &jv, double *f)
    {
        if (NEE == 0) return;
        double Rl[9], Rn[9];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            Rl[3 * r + 0] = m.Rp(i, 3 * r) * jv.c + m.Rp(i, 3 * r + 1) * jv.s;
            Rl[3 * r + 1] = m.Rp(i, 3 * r + 1) * jv.c - m.Rp(i, 3 * r) * jv.s;
            Rl[3 * r + 2] = m.Rp(i, 3 * r + 2);
        }
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                Rn[3 * r + c] = (i == 0) ? Rl[3 * r + c] : R[3 * r] * Rl[c] + R[3 * r + 1] * Rl[3 + c] + R[3 * r + 2] * Rl[6 + c];
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = Rn[k];
#pragma unroll
        for (int e = 0; e < NEE; ++e)
            if (je[e] == i) {
                double Fl[3], nl[3], t[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    Fl[c] = R[c] * Wl[e][0] + R[3 + c] * Wl[e][1] + R[6 + c] * Wl[e][2];
                    nl[c] = R[c] * Wl[e][3] + R[3 + c] * Wl[e][4] + R[6 + c] * Wl[e][5];
                }
                cross3(pl[e], Fl, t);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    f[c] += wsign * Fl[c];
                    f[3 + c] += wsign * (nl[c] + t[c]);
                }
            }
    }
};

// Reference-mode node for ONE serial chain of a static family (joints [c0, c0 + N) of the model; the pointers already address
// the chain's planes): tau = ID(q, qd, qdd) + wsign * sum_e J_e^T W_e, q+ = q + h qd, thermal ZOH — one RNEA sweep pair with
// the wrenches injected as external link forces (WrenchExt).  NEE = number of wrenches handled (ee.nee <= NEE).
template <int N, int NEE>
__global__ void __launch_bounds__(kThreads, 3) node_eval_chain_kernel(const __grid_constant__ StaticParams<N> P, long U, EeArgs ee, double wsign,
                                                                     const double *q, const double *qd, const double *qdd, const double *W,
                                                                     const double *T, double h, ZohArg zoh, double *tau, double *qnext,
                                                                     double *Tnext, int c0)
{
    const StaticModel<N, N> m{P};
    using D = Dyn<double, StaticModel<N, N>>;
    const long u = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= U) return;
    double a[N], b[N], c[N], t[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        a[i] = q[i * U + u];
        b[i] = qd ? qd[i * U + u] : 0.0;
        c[i] = qdd ? qdd[i * U + u] : 0.0;
    }
    WrenchExt<N, NEE> ext;
    ext.wsign = wsign;
#pragma unroll
    for (int e = 0; e < NEE; ++e) {
        const int j = e < ee.nee ? ee.f[e].joint - c0 : -1;
        ext.je[e] = (e < ee.nee && ee.f[e].joint >= 0 && j >= 0 && j < N) ? j : -1;  // world-fixed frame / other chain: no term
#pragma unroll
        for (int r = 0; r < 6; ++r) ext.Wl[e][r] = ext.je[e] >= 0 ? W[(long)(6 * e + r) * U + u] : 0.0;
#pragma unroll
        for (int r = 0; r < 3; ++r) ext.pl[e][r] = ee.f[e].p[r];
    }
    JointVar<double> jv[N];
    D::template rnea_impl<false>(m, a, jv, b, c, t, ext);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        tau[i * U + u] = t[i];
        if (qnext) qnext[i * U + u] = a[i] + h * b[i];
        if (Tnext) {
            const double Pl = m.fat(i, 2) * t[i] * t[i] + m.fat(i, 3) * b[i] * b[i];
            const double Ti = T[i * U + u];
            const double lam = m.fat(i, 0), za = zoh.a[c0 + i];
            Tnext[i * U + u] = (lam == 0.0) ? Ti + h * m.fat(i, 1) * Pl : za * Ti + (1.0 - za) * (m.fat(i, 1) / lam) * Pl;
        }
    }
}

template <int L>
static cudaError_t node_eval_chains(const LaunchModel &m, const EeArgs &ee, double wsign, long U, const double *q, const double *qd,
                                    const double *qdd, const double *W, const double *T, double h, const ZohArg &zoh, double *tau,
                                    double *qnext, double *Tnext, cudaStream_t s)
{
    const unsigned gb = (unsigned)((U + kThreads - 1) / kThreads);
    const StaticParams<L> *cp = static_cast<const StaticParams<L> *>(m.n == L ? m.static_params : m.chain_params);
    for (int c = 0; c < m.n / L; ++c) {
        const size_t off = (size_t)c * L * U;
        auto at = [off](auto *p) { return p ? p + off : p; };
        auto go = [&](auto kern) {
            kern<<<gb, kThreads, 0, s>>>(cp[c], U, ee, wsign, q + off, at(qd), at(qdd), W, at(T), h, zoh, tau + off, at(qnext), at(Tnext), c * L);
        };
        if (ee.nee == 0) go(node_eval_chain_kernel<L, 0>);
        else if (ee.nee == 1) go(node_eval_chain_kernel<L, 1>);
        else if (ee.nee == 2) go(node_eval_chain_kernel<L, 2>);
        else go(node_eval_chain_kernel<L, MPCF_MAX_EE>);
        g_launches.fetch_add(1);
    }
    return cudaGetLastError();
}

cudaError_t launch_node_eval(const LaunchModel &m, const EeArgs &ee, double wsign, long U, const double *q, const double *qd,
                             const double *qdd, const double *W, const double *T, double h, const ZohArg &zoh, double *tau,
                             double *qnext, double *Tnext, bool jtw_only, cudaStream_t s)
{
    if (U <= 0) return cudaSuccess;
    if (!jtw_only) {
        switch (family_chain_len(m.fam)) {
        case 3: return node_eval_chains<3>(m, ee, wsign, U, q, qd, qdd, W, T, h, zoh, tau, qnext, Tnext, s);
        case 6: return node_eval_chains<6>(m, ee, wsign, U, q, qd, qdd, W, T, h, zoh, tau, qnext, Tnext, s);
        case 7: return node_eval_chains<7>(m, ee, wsign, U, q, qd, qdd, W, T, h, zoh, tau, qnext, Tnext, s);
        default: break;
        }
    }
    return dispatch<NodeEvalBody>(m, U, 1, s, ee, wsign, q, qd, qdd, W, T, h, zoh, tau, qnext, Tnext, jtw_only);
}
cudaError_t launch_node_eval_jvp_dual(const LaunchModel &m, const EeArgs &ee, double wsign, long U, const double *q, const double *qd,
                                 const double *qdd, const double *W, double *dtau_dq, double *dtau_dqd, cudaStream_t s)
{
    return dispatch<NodeEvalJvpBody>(m, U, 2 * m.n, s, ee, wsign, q, qd, qdd, W, dtau_dq, dtau_dqd);
}
cudaError_t launch_cost_residual(int n, long B, int N, const double *q, const double *qd, const double *f, const double *tau,
                                 const double *qn, const double *qdn, const double *fn, const CostArgs &c, double *out,
                                 cudaStream_t s)
{
    if (B <= 0) return cudaSuccess;
    cost_residual_kernel<<<(unsigned)((B + 255) / 256), 256, 0, s>>>(n, B, N, q, qd, f, tau, qn, qdn, fn, c, out);
    g_launches.fetch_add(1);
    return cudaGetLastError();
}


}  // namespace mpcf
