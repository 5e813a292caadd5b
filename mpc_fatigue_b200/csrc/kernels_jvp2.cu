// kernels_jvp2.cu — analytic Jacobian pipeline of the RK4 dynamics+fatigue step for static chain families
// (chain3 / chain6 / forest12x6).  Three kernels per chunk of units, staged through a caller-provided
// workspace (all planes SoA `[row][chunk]`, coalesced):
//
//   K1 step_stages   thread = unit            primal RK4 (ABA) -> x+, and per stage (q_s, qd_s, qdd_s, fdot_s) -> WS1
//   K2 stage_derivs  thread = (unit, stage)   A_s = dqdd/dq, B_s = dqdd/dqd, C_s = M^-1 (derivs.cuh)        -> WS2
//   K3 chain_rule    thread = (unit, column)  forward accumulation of d x+/d (q, qd, tau, dt) through the four
//                                             stages using A_s, B_s, C_s (block = 32 units x all columns, so the
//                                             columns of one unit share the A/B/C lines through L1)
//
// Cost per unit drops from 19 dual-number sweeps of RK4(ABA) (v1, kernels_jvp.cu) to one primal sweep, four
// derivative evaluations and 19 cheap column recursions.  Generic (run-time tree) models keep the v1 kernel.
#include "derivs.cuh"
#include "launch.cuh"

namespace mpcf {

static inline size_t ws1_rows(int n) { return (size_t)4 * 4 * n; }
static inline size_t ws2_rows(int n) { return (size_t)4 * 3 * n * n; }
size_t jvp_ws_doubles_per_unit(int n) { return ws1_rows(n) + ws2_rows(n); }

// ------------------------------------------------------------------------------------------------ K1
template <int N, int L>
__global__ void __launch_bounds__(kThreads) k_step_stages(const __grid_constant__ StaticParams<N> P, long U, long u0, long cnt, long Uc,
                                                         const double *q, const double *qd, const double *tau, const double *f,
                                                         double dt, const double *dt_u, double *qn, double *qdn, double *fn,
                                                         double *ws1)
{
    const StaticModel<N, L> m{P};
    using D = Dyn<double, StaticModel<N, L>>;
    const long lu = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (lu >= cnt) return;
    const long u = u0 + lu;
    double x[3 * N], t[N], xs[3 * N], xn[3 * N], k[3 * N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        x[i] = q[i * U + u];
        x[N + i] = qd[i * U + u];
        x[2 * N + i] = f[i * U + u];
        t[i] = tau[i * U + u];
    }
    const double h = dt_u ? dt_u[u] : dt;
    const double cs[4] = {0.5, 0.5, 1.0, 0.0}, wt[4] = {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0};
#pragma unroll
    for (int i = 0; i < 3 * N; ++i) { xs[i] = x[i]; xn[i] = x[i]; }
#pragma unroll 1
    for (int s = 0; s < 4; ++s) {
        D::xdot(m, xs, t, k);
        double *w = ws1 + (size_t)s * 4 * N * Uc + lu;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            w[(size_t)i * Uc] = xs[i];
            w[(size_t)(N + i) * Uc] = xs[N + i];
            w[(size_t)(2 * N + i) * Uc] = k[N + i];
            w[(size_t)(3 * N + i) * Uc] = k[2 * N + i];
        }
        const double a = h * wt[s], c = h * cs[s];
#pragma unroll
        for (int i = 0; i < 3 * N; ++i) {
            xn[i] += a * k[i];
            xs[i] = x[i] + c * k[i];
        }
    }
    if (qn) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            qn[i * U + u] = xn[i];
            qdn[i * U + u] = xn[N + i];
            fn[i * U + u] = xn[2 * N + i];
        }
    }
}

// ------------------------------------------------------------------------------------------------ K2
template <int N, int L>
__global__ void __launch_bounds__(kThreads) k_stage_derivs(const __grid_constant__ StaticParams<N> P, long cnt, long Uc, const double *ws1,
                                                          double *ws2)
{
    const StaticModel<N, L> m{P};
    const long lu = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (lu >= cnt) return;
    const int s = blockIdx.y;
    const double *w = ws1 + (size_t)s * 4 * N * Uc + lu;
    double q[N], qd[N], qdd[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        q[i] = w[(size_t)i * Uc];
        qd[i] = w[(size_t)(N + i) * Uc];
        qdd[i] = w[(size_t)(2 * N + i) * Uc];
    }
    double A[N * N], B[N * N], C[N * N];
    FdDerivs<StaticModel<N, L>, L>::run(m, q, qd, qdd, A, B, C);
    double *o = ws2 + (size_t)s * 3 * N * N * Uc + lu;
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
        for (int c = 0; c < N; ++c) {
            if (r / L != c / L) continue;  // other chains: structurally zero, never read
            o[(size_t)(r * N + c) * Uc] = A[r * N + c];
            o[(size_t)(N * N + r * N + c) * Uc] = B[r * N + c];
            o[(size_t)(2 * N * N + r * N + c) * Uc] = C[r * N + c];
        }
}

// public fd-derivs entry for static families: thread = unit, qdd from ABA, then the analytic derivatives
template <int N, int L>
__global__ void __launch_bounds__(kThreads) k_fd_derivs(const __grid_constant__ StaticParams<N> P, long U, const double *q, const double *qd,
                                                       const double *tau, double *Ao, double *Bo, double *Co)
{
    const StaticModel<N, L> m{P};
    const long u = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= U) return;
    double a[N], b[N], c[N], qdd[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { a[i] = q[i * U + u]; b[i] = qd[i * U + u]; c[i] = tau[i * U + u]; }
    Dyn<double, StaticModel<N, L>>::aba(m, a, b, c, qdd);
    double A[N * N], B[N * N], C[N * N];
    FdDerivs<StaticModel<N, L>, L>::run(m, a, b, qdd, A, B, C);
#pragma unroll
    for (int k = 0; k < N * N; ++k) {
        Ao[(size_t)k * U + u] = A[k];
        Bo[(size_t)k * U + u] = B[k];
        Co[(size_t)k * U + u] = C[k];
    }
}

// generic fallback for fd-derivs: one dual-number ABA sweep per seed direction (blockIdx.y in [0, 3n))
struct FdDerivsDualBody {
    template <class MP>
    static MPCF_DI void run(const MP &m, long u, long U, const double *q, const double *qd, const double *tau, double *Ao, double *Bo,
                            double *Co)
    {
        constexpr int UNR = MP::kStatic ? MP::MAXN : 1;
        const int n = m.n();
        const int d = blockIdx.y;
        Dual a[MP::MAXN], b[MP::MAXN], c[MP::MAXN], t[MP::MAXN];
#pragma unroll UNR
        for (int i = 0; i < n; ++i) {
            a[i] = Dual(q[i * U + u], d == i ? 1.0 : 0.0);
            b[i] = Dual(qd[i * U + u], d == n + i ? 1.0 : 0.0);
            c[i] = Dual(tau[i * U + u], d == 2 * n + i ? 1.0 : 0.0);
        }
        Dyn<Dual, MP>::aba(m, a, b, c, t);
        double *out = d < n ? Ao : (d < 2 * n ? Bo : Co);
        const int col = d % n;
#pragma unroll UNR
        for (int r = 0; r < n; ++r) out[(size_t)(r * n + col) * U + u] = t[r].d;
    }
};

// ------------------------------------------------------------------------------------------------ K3
// Column recursion.  With Y_s := dt K_s (+ k_s in the dt column):  X_{s+1} = X_1 + c_s Y_s,  out = X_1 + sum_s w_s Y_s.
template <int N, int L>
__global__ void __launch_bounds__(32 * (3 * N + 1 > 19 ? 19 : 3 * N + 1))
    k_chain_rule(const __grid_constant__ StaticParams<N> P, long U, long u0, long cnt, long Uc, const double *tau, double dt,
                 const double *dt_u, const double *ws1, const double *ws2, double *jac)
{
    const long lu = (long)blockIdx.x * 32 + threadIdx.x;
    if (lu >= cnt) return;
    const long u = u0 + lu;
    const double h = dt_u ? dt_u[u] : dt;
    constexpr int NC = 3 * N + 1;
    constexpr long PC = 4 * N + 1;
    const double cs[4] = {0.5, 0.5, 1.0, 0.0}, wt[4] = {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0};
    for (int col = threadIdx.y; col < NC; col += blockDim.y) {
        const int jq = col < N ? col : -1;                          // q seed
        const int jv = (col >= N && col < 2 * N) ? col - N : -1;    // qd seed
        const int jt = (col >= 2 * N && col < 3 * N) ? col - 2 * N : -1;  // tau seed
        const bool isdt = col == 3 * N;
        double Xq[N], Xv[N], Xf[N], aq[N], av[N], af[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            Xq[i] = (i == jq) ? 1.0 : 0.0;
            Xv[i] = (i == jv) ? 1.0 : 0.0;
            Xf[i] = 0.0;
            aq[i] = Xq[i]; av[i] = Xv[i]; af[i] = 0.0;
        }
        double tauj = 0.0;
        if (jt >= 0) tauj = tau[(size_t)jt * U + u];
#pragma unroll 1
        for (int s = 0; s < 4; ++s) {
            const double *w1 = ws1 + (size_t)s * 4 * N * Uc + lu;
            const double *w2 = ws2 + (size_t)s * 3 * N * N * Uc + lu;
            double Yq[N], Yv[N], Yf[N];
#pragma unroll
            for (int r = 0; r < N; ++r) {
                double kv = 0.0;
#pragma unroll
                for (int c = 0; c < N; ++c) {
                    if (r / L != c / L) continue;
                    kv = fma(w2[(size_t)(r * N + c) * Uc], Xq[c], kv);
                    kv = fma(w2[(size_t)(N * N + r * N + c) * Uc], Xv[c], kv);
                }
                if (jt >= 0 && jt / L == r / L) kv += w2[(size_t)(2 * N * N + r * N + jt) * Uc];
                const double qds = w1[(size_t)(N + r) * Uc];
                double kf = 2.0 * P.fat[r][1] * P.fat[r][3] * qds * Xv[r] - P.fat[r][0] * Xf[r];
                if (r == jt) kf += 2.0 * P.fat[r][1] * P.fat[r][2] * tauj;
                Yq[r] = h * Xv[r];
                Yv[r] = h * kv;
                Yf[r] = h * kf;
                if (isdt) {
                    Yq[r] += qds;
                    Yv[r] += w1[(size_t)(2 * N + r) * Uc];
                    Yf[r] += w1[(size_t)(3 * N + r) * Uc];
                }
            }
#pragma unroll
            for (int i = 0; i < N; ++i) {
                aq[i] = fma(wt[s], Yq[i], aq[i]);
                av[i] = fma(wt[s], Yv[i], av[i]);
                af[i] = fma(wt[s], Yf[i], af[i]);
                Xq[i] = ((i == jq) ? 1.0 : 0.0) + cs[s] * Yq[i];
                Xv[i] = ((i == jv) ? 1.0 : 0.0) + cs[s] * Yv[i];
                Xf[i] = cs[s] * Yf[i];
            }
        }
        const long ocol = isdt ? 4 * N : col;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            jac[((size_t)i * PC + ocol) * U + u] = aq[i];
            jac[((size_t)(N + i) * PC + ocol) * U + u] = av[i];
            jac[((size_t)(2 * N + i) * PC + ocol) * U + u] = af[i];
        }
        if (isdt) {  // the n fatigue columns in closed form (see kernels_jvp.cu)
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const double z = P.fat[j][0] * h;
                const double g = 1.0 + z * (-1.0 + z * (0.5 + z * (-1.0 / 6.0 + z * (1.0 / 24.0))));
#pragma unroll
                for (int r = 0; r < 3 * N; ++r) jac[((size_t)r * PC + 3 * N + j) * U + u] = (r == 2 * N + j) ? g : 0.0;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ launchers
template <int N, int L>
static cudaError_t run_jvp2(const StaticParams<N> &P, long U, const double *q, const double *qd, const double *tau, const double *f,
                            double dt, const double *dt_u, double *qn, double *qdn, double *fn, double *jac, double *ws, long Uc,
                            cudaStream_t s)
{
    double *ws1 = ws, *ws2 = ws + ws1_rows(N) * (size_t)Uc;
    constexpr int NCY = 3 * N + 1 > 19 ? 19 : 3 * N + 1;
    for (long u0 = 0; u0 < U; u0 += Uc) {
        const long cnt = (U - u0) < Uc ? (U - u0) : Uc;
        const unsigned gb = (unsigned)((cnt + kThreads - 1) / kThreads);
        k_step_stages<N, L><<<gb, kThreads, 0, s>>>(P, U, u0, cnt, Uc, q, qd, tau, f, dt, dt_u, qn, qdn, fn, ws1);
        k_stage_derivs<N, L><<<dim3(gb, 4), kThreads, 0, s>>>(P, cnt, Uc, ws1, ws2);
        k_chain_rule<N, L><<<(unsigned)((cnt + 31) / 32), dim3(32, NCY), 0, s>>>(P, U, u0, cnt, Uc, tau, dt, dt_u, ws1, ws2, jac);
        g_launches.fetch_add(3);
    }
    return cudaGetLastError();
}

bool jvp2_supported(const LaunchModel &m) { return m.fam == FAM_CHAIN3 || m.fam == FAM_CHAIN6 || m.fam == FAM_FOREST12x6; }

cudaError_t launch_step_jvp_ws(const LaunchModel &m, long U, const double *q, const double *qd, const double *tau, const double *f,
                               double dt, const double *dt_u, double *qn, double *qdn, double *fn, double *jac, double *ws,
                               size_t ws_bytes, cudaStream_t s)
{
    if (U <= 0) return cudaSuccess;
    const size_t per_unit = jvp_ws_doubles_per_unit(m.n) * sizeof(double);
    long Uc = (long)(ws_bytes / per_unit);
    if (Uc > U) Uc = U;
    Uc -= Uc % 32;  // keep every workspace plane 256-byte aligned
    if (Uc < 32 && U >= 32) return cudaErrorInvalidValue;
    if (Uc < 32) Uc = 32;
    switch (m.fam) {
    case FAM_CHAIN3:
        return run_jvp2<3, 3>(*static_cast<const StaticParams<3> *>(m.static_params), U, q, qd, tau, f, dt, dt_u, qn, qdn, fn, jac, ws, Uc, s);
    case FAM_CHAIN6:
        return run_jvp2<6, 6>(*static_cast<const StaticParams<6> *>(m.static_params), U, q, qd, tau, f, dt, dt_u, qn, qdn, fn, jac, ws, Uc, s);
    case FAM_FOREST12x6:
        return run_jvp2<12, 6>(*static_cast<const StaticParams<12> *>(m.static_params), U, q, qd, tau, f, dt, dt_u, qn, qdn, fn, jac, ws, Uc, s);
    default:
        return cudaErrorInvalidValue;
    }
}

cudaError_t launch_fd_derivs(const LaunchModel &m, long U, const double *q, const double *qd, const double *tau, double *A, double *B,
                             double *C, cudaStream_t s)
{
    if (U <= 0) return cudaSuccess;
    const unsigned gb = (unsigned)((U + kThreads - 1) / kThreads);
    switch (m.fam) {
    case FAM_CHAIN3:
        k_fd_derivs<3, 3><<<gb, kThreads, 0, s>>>(*static_cast<const StaticParams<3> *>(m.static_params), U, q, qd, tau, A, B, C);
        break;
    case FAM_CHAIN6:
        k_fd_derivs<6, 6><<<gb, kThreads, 0, s>>>(*static_cast<const StaticParams<6> *>(m.static_params), U, q, qd, tau, A, B, C);
        break;
    case FAM_FOREST12x6:
        k_fd_derivs<12, 6><<<gb, kThreads, 0, s>>>(*static_cast<const StaticParams<12> *>(m.static_params), U, q, qd, tau, A, B, C);
        break;
    default:
        return dispatch<FdDerivsDualBody>(m, U, 3 * m.n, s, q, qd, tau, A, B, C);
    }
    g_launches.fetch_add(1);
    return cudaGetLastError();
}

}  // namespace mpcf
