// tests/hostcheck/tree_host.cu — TEST INFRASTRUCTURE ONLY.  Compiles the __host__ __device__ arithmetic of the run-time-tree
// Jacobian pipeline (mpc_fatigue_b200/csrc/tree_derivs.cuh, and the column recursion of kernels_tree.cu restated as plain
// loops over the same packed layout, index tables and accumulator identities) for the HOST, so that the formulas can be
// checked against the oracle on a machine without a GPU.  Nothing in the product loads this library.
#include <cmath>
#include <cstring>
#include <vector>

#include "tree_derivs.cuh"

using namespace mpcf;

namespace {
struct HostTree {
    static constexpr int MAXN = 64;
    static constexpr bool kStatic = false;
    int n_;
    const int *par, *jt, *dep, *rp;
    MPCF_HD bool keep(int i) const  // a link other than i + 1 hangs off link i
    {
        for (int c = i + 2; c < n_; ++c)
            if (par[c] == i) return true;
        return false;
    }
    const double *Rp_, *pp_, *mass_, *mc_, *Io_, *arm_, *grav_;
    MPCF_HD int n() const { return n_; }
    MPCF_HD int parent(int i) const { return par[i]; }
    MPCF_HD bool prismatic(int i) const { return jt[i] != 0; }
    MPCF_HD int depth(int i) const { return dep[i]; }
    MPCF_HD int rowptr(int i) const { return rp[i]; }
    MPCF_HD double Rp(int i, int k) const { return Rp_[9 * i + k]; }
    MPCF_HD double pp(int i, int k) const { return pp_[3 * i + k]; }
    MPCF_HD double mass(int i) const { return mass_[i]; }
    MPCF_HD double mc(int i, int k) const { return mc_[3 * i + k]; }
    MPCF_HD double Io(int i, int k) const { return Io_[6 * i + k]; }
    MPCF_HD double arm(int i) const { return arm_[i]; }
    MPCF_HD double grav(int k) const { return grav_[k]; }
};

struct DenseOut {
    int n;
    double *Dq, *Dv;
    void pair(int k, int j, int, double dqkj, double dvkj, double dqjk, double dvjk)
    {
        Dq[k * n + j] = dqkj; Dv[k * n + j] = dvkj;
        Dq[j * n + k] = dqjk; Dv[j * n + k] = dvjk;
    }
};

struct Topo {
    std::vector<int> depth, rowptr;
    Topo(int n, const int *parent) : depth(n), rowptr(n + 1, 0)
    {
        for (int i = 0; i < n; ++i) depth[i] = parent[i] < 0 ? 0 : depth[parent[i]] + 1;
        for (int i = 0; i < n; ++i) rowptr[i + 1] = rowptr[i] + depth[i] + 1;
    }
};
}  // namespace

// Dq, Dv, M dense [n][n]; Lfac dense [n][n] (strictly lower: L_kj; diagonal: D_k) from the packed in-place factorisation
extern "C" int hc_tree_derivs(int n, const int *parent, const int *jtype, const double *Rp, const double *pp, const double *mass, const double *mc,
                              const double *Io, const double *arm, const double *grav, const double *q, const double *qd, const double *qdd,
                              double *Dq, double *Dv, double *M, double *Lfac, double *Cinv)
{
    Topo tp(n, parent);
    HostTree m{n, parent, jtype, tp.depth.data(), tp.rowptr.data(), Rp, pp, mass, mc, Io, arm, grav};
    std::vector<TreeRec> rec(n);
    std::vector<TreeComp> comp(n);
    std::vector<double> Mp(tp.rowptr[n]);
    std::memset(Dq, 0, sizeof(double) * n * n);
    std::memset(Dv, 0, sizeof(double) * n * n);
    std::memset(M, 0, sizeof(double) * n * n);
    std::memset(Lfac, 0, sizeof(double) * n * n);
    DenseOut out{n, Dq, Dv};
    TreeDerivs<HostTree>::forward(m, q, qd, qdd, rec.data(), comp.data());
    TreeDerivs<HostTree>::backward(m, rec.data(), comp.data(), Mp.data(), out);
    for (int k = 0; k < n; ++k)
        for (int j = k; j >= 0; j = parent[j]) M[k * n + j] = M[j * n + k] = Mp[tp.rowptr[k] + tp.depth[j]];
    const bool ok = TreeDerivs<HostTree>::factorize(m, Mp.data());
    for (int k = 0; k < n; ++k)
        for (int j = k; j >= 0; j = parent[j]) Lfac[k * n + j] = Mp[tp.rowptr[k] + tp.depth[j]];
    if (Cinv) {
        std::vector<double> x(n);
        for (int j = 0; j < n; ++j) {
            TreeDerivs<HostTree>::minv_column(m, Mp.data(), j, x.data());
            for (int i = 0; i < n; ++i) Cinv[i * n + j] = x[i];
        }
    }
    return ok ? 0 : -1;
}

// Linv dense [n][n]: the unit lower-triangular inverse of L from invert_unit_factor (diagonal: D_k, as in Lfac)
extern "C" int hc_tree_linv(int n, const int *parent, const int *jtype, const double *Rp, const double *pp, const double *mass, const double *mc,
                            const double *Io, const double *arm, const double *grav, const double *q, const double *qd, const double *qdd, double *Linv)
{
    Topo tp(n, parent);
    HostTree m{n, parent, jtype, tp.depth.data(), tp.rowptr.data(), Rp, pp, mass, mc, Io, arm, grav};
    std::vector<TreeRec> rec(n);
    std::vector<TreeComp> comp(n);
    std::vector<double> Mp(tp.rowptr[n]), Dq(n * n), Dv(n * n), lrow(n);
    std::vector<int> path(n);
    DenseOut out{n, Dq.data(), Dv.data()};
    TreeDerivs<HostTree>::forward(m, q, qd, qdd, rec.data(), comp.data());
    TreeDerivs<HostTree>::backward(m, rec.data(), comp.data(), Mp.data(), out);
    const bool ok = TreeDerivs<HostTree>::factorize(m, Mp.data());
    TreeDerivs<HostTree>::invert_unit_factor(m, Mp.data(), path.data(), lrow.data());
    std::memset(Linv, 0, sizeof(double) * n * n);
    for (int k = 0; k < n; ++k)
        for (int j = k; j >= 0; j = parent[j]) Linv[k * n + j] = Mp[tp.rowptr[k] + tp.depth[j]];
    return ok ? 0 : -1;
}

// The column recursion of kernels_tree.cu (k_tree_chain) as plain loops.  Inputs per stage s = 0..3: dense Dq_s, Dv_s, Lfac_s
// (as returned above), qd_s, qdd_s, fdot_s; fat[n][4] = (lambda, kappa, ctau, cv); tau[n]; step h.  Output jac[3n][4n+1].
extern "C" void hc_tree_chain(int n, const double *Dq, const double *Dv, const double *Lfac, const double *qds, const double *qdds,
                              const double *fdots, const double *fat, const double *tau, double h, double *jac)
{
    const int P = 4 * n + 1, NC = 3 * n + 1;
    const double w[4] = {1.0 / 6, 1.0 / 3, 1.0 / 3, 1.0 / 6}, cs[4] = {0.5, 0.5, 1.0, 0.0};
    std::memset(jac, 0, sizeof(double) * 3 * n * P);
    // per-stage fatigue coefficients exactly as k_tree_stages writes them
    std::vector<double> fnext(4 * n, 0.0), fsum(n, 0.0), gdtsum(n, 0.0), qdbar(n, 0.0);
    for (int i = 0; i < n; ++i) {
        const double z = fat[4 * i] * h;
        const double gam[4] = {1 - z + z * z / 2 - z * z * z / 4, 2 - z + z * z / 2, 2 - z, 1.0};
        for (int s = 0; s < 4; ++s) {
            const double fcoef = gam[s] * (h / 6) * 2 * fat[4 * i + 1] * fat[4 * i + 3] * qds[s * n + i];
            fsum[i] += fcoef;
            if (s > 0) fnext[(s - 1) * n + i] = fcoef * cs[s - 1];
            gdtsum[i] += gam[s] / 6 * fdots[s * n + i];
            qdbar[i] += w[s] * qds[s * n + i];
        }
    }
    for (int t = 0; t < NC; ++t) {
        const int jq = t < n ? t : -1, jv = (t >= n && t < 2 * n) ? t - n : -1, jt = (t >= 2 * n && t < 3 * n) ? t - 2 * n : -1;
        const bool isdt = t == 3 * n;
        std::vector<double> Xq(n), Xv(n), acc(n), Pacc(n, 0.0), AF(n, 0.0), av0(n), Yv(n);
        for (int r = 0; r < n; ++r) { Xq[r] = r == jq; Xv[r] = r == jv; }
        for (int s = 0; s < 4; ++s) {
            const double *dq = Dq + (size_t)s * n * n, *dv = Dv + (size_t)s * n * n, *L = Lfac + (size_t)s * n * n;
            for (int r = 0; r < n; ++r) {
                double a = 0.0;
                for (int c = 0; c < n; ++c) a += dq[r * n + c] * Xq[c] + dv[r * n + c] * Xv[c];
                acc[r] = -a + (r == jt ? 1.0 : 0.0);
            }
            // M k = rhs with M = L^T D L:  L^T y = rhs (i descending), w = y / D, L x = w (column-oriented, j ascending)
            for (int i = n - 1; i >= 0; --i)
                for (int k = i + 1; k < n; ++k) acc[i] -= L[k * n + i] * acc[k];
            for (int i = 0; i < n; ++i) acc[i] /= L[i * n + i];
            for (int j = 0; j < n; ++j)
                for (int i = j + 1; i < n; ++i) acc[i] -= L[i * n + j] * acc[j];
            for (int r = 0; r < n; ++r) {
                Yv[r] = h * acc[r] + (isdt ? qdds[s * n + r] : 0.0);
                if (s == 0) av0[r] = (r == jv) - Yv[r] / 6;
                if (s < 3) { Pacc[r] += Yv[r]; AF[r] += fnext[s * n + r] * Yv[r]; }
                const double yq = h * Xv[r] + (isdt ? qds[s * n + r] : 0.0);
                Xq[r] = (r == jq) + cs[s] * yq;
                Xv[r] = (r == jv) + cs[s] * Yv[r];
            }
        }
        const int col = isdt ? 4 * n : t;
        for (int r = 0; r < n; ++r) {
            jac[(size_t)r * P + col] = (r == jq) + h * (r == jv) + (h / 6) * Pacc[r] + (isdt ? qdbar[r] : 0.0);
            jac[(size_t)(n + r) * P + col] = av0[r] + (2 * Pacc[r] + Yv[r]) / 6;
            double af = AF[r];
            if (r == jv) af += fsum[r];
            if (isdt) af += gdtsum[r];
            if (r == jt) {
                const double z = fat[4 * r] * h;
                af += 2 * fat[4 * r + 1] * fat[4 * r + 2] * tau[r] * h * (1 - z / 2 + z * z / 6 - z * z * z / 24);
            }
            jac[(size_t)(2 * n + r) * P + col] = af;
        }
    }
    for (int j = 0; j < n; ++j) {  // fatigue columns: closed form
        const double z = fat[4 * j] * h;
        jac[(size_t)(2 * n + j) * P + 3 * n + j] = 1 + z * (-1 + z * (0.5 + z * (-1.0 / 6 + z / 24)));
    }
}
