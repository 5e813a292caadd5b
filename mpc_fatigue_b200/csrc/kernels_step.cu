// kernels_step.cu — forward dynamics and the RK4 step of (q, qd, f) (values only).
#include "launch.cuh"

namespace mpcf {

struct AbaBody {
    static constexpr int kGenericMinBlocks = 3;
    template <class MP>
    static MPCF_DI void run(const MP &m, long u, long U, const double *q, const double *qd, const double *tau, double *qdd)
    {
        constexpr int UNR = MP::kStatic ? MP::MAXN : 1;
        const int n = m.n();
        double a[MP::MAXN], b[MP::MAXN], c[MP::MAXN], t[MP::MAXN];
#pragma unroll UNR
        for (int i = 0; i < n; ++i) {
            a[i] = q[i * U + u];
            b[i] = qd[i * U + u];
            c[i] = tau[i * U + u];
        }
        Dyn<double, MP>::fd(m, a, b, c, t);
#pragma unroll UNR
        for (int i = 0; i < n; ++i) qdd[i * U + u] = t[i];
    }
};

struct StepBody {
    static constexpr int kGenericMinBlocks = 3;
    template <class MP>
    static MPCF_DI void run(const MP &m, long u, long U, const double *q, const double *qd, const double *tau, const double *theat,
                            const double *f, double dt, const double *dt_u, double *qn, double *qdn, double *fn)
    {
        constexpr int UNR = MP::kStatic ? MP::MAXN : 1;
        const int n = m.n();
        double x[3 * MP::MAXN], t[MP::MAXN], xn[3 * MP::MAXN];
#pragma unroll UNR
        for (int i = 0; i < n; ++i) {
            x[i] = q[i * U + u];
            x[n + i] = qd[i * U + u];
            x[2 * n + i] = f[i * U + u];
            t[i] = tau[i * U + u];
        }
        const double h = dt_u ? dt_u[u] : dt;
        // coupled fatigue: the windings heat with theat while the dynamics see tau (one call site: the step body is ~50 KB of code)
        double th[MP::MAXN];
#pragma unroll UNR
        for (int i = 0; i < n; ++i) th[i] = theat ? theat[i * U + u] : t[i];
        Dyn<double, MP>::step_rk4(m, x, t, th, h, xn);
#pragma unroll UNR
        for (int i = 0; i < n; ++i) {
            qn[i * U + u] = xn[i];
            qdn[i * U + u] = xn[n + i];
            fn[i * U + u] = xn[2 * n + i];
        }
    }
};

// Sequential rollout (single shooting): thread = scenario, the state (q, qd, f) stays in registers across the N steps;
// tau is read per step (node-major planes, unit k*B + b) and the trajectory x_1..x_N is written in the same layout.
struct RolloutBody {
    static constexpr int kGenericMinBlocks = 3;
    template <class MP>
    static MPCF_DI void run(const MP &m, long b, long B, int N, const double *q0, const double *qd0, const double *f0, const double *tau,
                            double dt, double *qt, double *qdt, double *ft)
    {
        constexpr int UNR = MP::kStatic ? MP::MAXN : 1;
        const int n = m.n();
        const long U = B * N;
        double x[3 * MP::MAXN], t[MP::MAXN], xn[3 * MP::MAXN];
#pragma unroll UNR
        for (int i = 0; i < n; ++i) {
            x[i] = q0[i * B + b];
            x[n + i] = qd0[i * B + b];
            x[2 * n + i] = f0[i * B + b];
        }
#pragma unroll 1
        for (int k = 0; k < N; ++k) {
            const long u = (long)k * B + b;
#pragma unroll UNR
            for (int i = 0; i < n; ++i) t[i] = tau[i * U + u];
            Dyn<double, MP>::step_rk4(m, x, t, dt, xn);
#pragma unroll UNR
            for (int i = 0; i < n; ++i) {
                qt[i * U + u] = xn[i];
                qdt[i * U + u] = xn[n + i];
                ft[i * U + u] = xn[2 * n + i];
                x[i] = xn[i];
                x[n + i] = xn[n + i];
                x[2 * n + i] = xn[2 * n + i];
            }
        }
    }
};

cudaError_t launch_rollout(const LaunchModel &m, long B, int N, const double *q0, const double *qd0, const double *f0, const double *tau,
                           double dt, double *qt, double *qdt, double *ft, cudaStream_t s)
{
    return dispatch<RolloutBody>(m, B, 1, s, N, q0, qd0, f0, tau, dt, qt, qdt, ft);
}

cudaError_t launch_aba(const LaunchModel &m, long U, const double *q, const double *qd, const double *tau, double *qdd, cudaStream_t s)
{
    return dispatch<AbaBody>(m, U, 1, s, q, qd, tau, qdd);
}
cudaError_t launch_step(const LaunchModel &m, long U, const double *q, const double *qd, const double *tau, const double *f,
                        double dt, const double *dt_u, double *qn, double *qdn, double *fn, cudaStream_t s, const double *theat)
{
    return dispatch<StepBody>(m, U, 1, s, q, qd, tau, theat, f, dt, dt_u, qn, qdn, fn);
}

}  // namespace mpcf
