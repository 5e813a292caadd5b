// tree_derivs.cuh — analytic inverse-dynamics derivatives and the factorised joint-space inertia for RUN-TIME TREES
// (branched models with revolute and prismatic 1-DOF joints; config C4's 37-joint tree), the per-(unit, RK4 stage) building
// block of the tree Jacobian pipeline (kernels_tree.cu).  Same world-frame formulation as derivs.cuh, with the
// chain relations replaced by ancestor relations:
//
//   S_j   = [o_j x z_j ; z_j] (revolute)  or  [z_j ; 0] (prismatic)       joint axis as a spatial motion at the world origin
//   xi_j  = S_j x v_p(j),   eta_j = S_j x a_p(j) - xi_j x v_p(j)           p = parent; a includes the -g base acceleration
//   I_c,k, H_c,k, F_c,k, Bs_c,k   composites over the SUBTREE rooted at k
//
//   dID_k/dq_j  = -S_k . (I_c,k eta_j + B_c,k xi_j)                        j an ancestor of k, or k itself
//   dID_j/dq_k  =  S_j . (S_k x* F_c,k - I_c,k eta_k - B_c,k xi_k)         j a strict ancestor of k
//   dID_k/dqd_j =  S_k . (B_c,k S_j - 2 I_c,k xi_j),   dID_j/dqd_k = S_j . (B_c,k S_k - 2 I_c,k xi_k)
//   M_kj = M_jk =  S_k . I_c,k S_j  (+ armature on the diagonal)
//   every entry between joints on different branches is zero.
//
// The joint-space inertia is kept only on its branch-induced sparsity pattern — entry (k, j), j in anc*(k), at
// rowptr(k) + depth(j) — and factorised in place as M = L^T D L (Featherstone, RBDA §6.5: no fill-in outside the pattern).
//
// Everything here is plain arithmetic on caller-provided scratch arrays, `__host__ __device__`: the product compiles it for
// the device only; tests/hostcheck compiles the same source for the host and checks it against the oracle's complex-step
// derivatives (the reference has no code for any of this: north-star addition).
#pragma once

#include "derivs.cuh"

namespace mpcf {

struct TreeRec {   // per link, written by the forward pass
    double R[9], o[3], v[6], a[6];
    LinkFwd K;     // S, xi, eta
};
struct TreeComp {  // per link: composite of its subtree
    RigidInertiaW I;
    double H[6], F[6], B[6];
};

template <class MP>
struct TreeDerivs {
    static constexpr int MAXN = MP::MAXN;

    // comp[i] leaves this pass holding link i's OWN world-frame inertia, momentum, force and B-matrix: the backward pass adds the
    // children's composites to it (no zero-initialisation pass, and the backward pass never reads R, o, v, a again).
    // Along a chain segment (parent(i) == i - 1) the parent's pose and motion are carried in registers; rec[i] receives them
    // only when a link other than i + 1 hangs off link i (m.keep(i)).
    static MPCF_HD void forward(const MP &m, const double *q, const double *qd, const double *qdd, TreeRec *rec, TreeComp *comp)
    {
        const int n = m.n();
        double Rc[9], oc[3], vc[6], ac[6];  // link i - 1, then link i
        for (int i = 0; i < n; ++i) {
            const int par = m.parent(i);
            TreeRec &r = rec[i];
            double Rl[9];
            double s = 0.0, c = 1.0;
            const bool pris = m.prismatic(i);
            if (!pris) sincos(q[i], &s, &c);
            for (int k = 0; k < 3; ++k) {
                Rl[3 * k + 0] = m.Rp(i, 3 * k) * c + m.Rp(i, 3 * k + 1) * s;
                Rl[3 * k + 1] = m.Rp(i, 3 * k + 1) * c - m.Rp(i, 3 * k) * s;
                Rl[3 * k + 2] = m.Rp(i, 3 * k + 2);
            }
            double Rn[9], on[3], vp[6], ap[6];
            if (par < 0) {
                for (int k = 0; k < 9; ++k) Rn[k] = Rl[k];
                for (int k = 0; k < 3; ++k) on[k] = m.pp(i, k);
                for (int k = 0; k < 6; ++k) { vp[k] = 0.0; ap[k] = 0.0; }
                ap[0] = -m.grav(0); ap[1] = -m.grav(1); ap[2] = -m.grav(2);
            } else {
                if (par != i - 1) {
                    const TreeRec &p = rec[par];
                    for (int k = 0; k < 9; ++k) Rc[k] = p.R[k];
                    for (int k = 0; k < 3; ++k) oc[k] = p.o[k];
                    for (int k = 0; k < 6; ++k) { vc[k] = p.v[k]; ac[k] = p.a[k]; }
                }
                for (int a = 0; a < 3; ++a) {
                    for (int b = 0; b < 3; ++b) Rn[3 * a + b] = Rc[3 * a] * Rl[b] + Rc[3 * a + 1] * Rl[3 + b] + Rc[3 * a + 2] * Rl[6 + b];
                    on[a] = oc[a] + Rc[3 * a] * m.pp(i, 0) + Rc[3 * a + 1] * m.pp(i, 1) + Rc[3 * a + 2] * m.pp(i, 2);
                }
                for (int k = 0; k < 6; ++k) { vp[k] = vc[k]; ap[k] = ac[k]; }
            }
            const double z[3] = {Rn[2], Rn[5], Rn[8]};
            if (pris) {  // the joint shifts the link along its axis; the axis is a pure translation
                for (int k = 0; k < 3; ++k) { on[k] += z[k] * q[i]; r.K.S[k] = z[k]; r.K.S[3 + k] = 0.0; }
            } else {
                cross3(on, z, r.K.S);
                r.K.S[3] = z[0]; r.K.S[4] = z[1]; r.K.S[5] = z[2];
            }
            // xi = S x v_parent ; eta = S x a_parent - xi x v_parent
            mxm(r.K.S, vp, r.K.xi);
            double t6[6];
            mxm(r.K.S, ap, r.K.eta);
            mxm(r.K.xi, vp, t6);
            for (int k = 0; k < 6; ++k) r.K.eta[k] -= t6[k];
            for (int k = 0; k < 9; ++k) Rc[k] = Rn[k];
            for (int k = 0; k < 3; ++k) oc[k] = on[k];
            for (int k = 0; k < 6; ++k) {
                vc[k] = vp[k] + r.K.S[k] * qd[i];
                ac[k] = ap[k] + r.K.S[k] * qdd[i] - r.K.xi[k] * qd[i];
            }
            if (m.keep(i)) {
                for (int k = 0; k < 9; ++k) r.R[k] = Rc[k];
                for (int k = 0; k < 3; ++k) r.o[k] = oc[k];
                for (int k = 0; k < 6; ++k) { r.v[k] = vc[k]; r.a[k] = ac[k]; }
            }
            TreeComp &own = comp[i];
            FdDerivs<MP, 1>::link_world(m, i, Rc, oc, vc, ac, own.I, own.H, own.F, own.B);
        }
    }

    // Out: pair(k, j, e, dq_kj, dv_kj, dq_jk, dv_jk) for every link k and every j in anc*(k) (j == k: both orientations coincide);
    // e = rowptr(k) + depth(j) is the packed index of the pair.  Unrelated pairs are not visited (their entries are zero).
    // Mp: packed joint-space inertia (see header), filled here, NOT yet factorised.
    template <class Out>
    static MPCF_HD void backward(const MP &m, const TreeRec *rec, TreeComp *comp, double *Mp, Out &out)
    {
        const int n = m.n();
        TreeComp c;         // composite of the subtree of the link in hand; carried in registers to the parent along chain segments
        c.I.m = 0.0;        // (never read before take() assigns it; this only tells the compiler so)
        for (int e = 0; e < 3; ++e) c.I.h[e] = 0.0;
        for (int e = 0; e < 6; ++e) { c.I.Io[e] = 0.0; c.H[e] = 0.0; c.F[e] = 0.0; c.B[e] = 0.0; }
        bool carried = false;
        auto take = [&](int k) {  // c <- comp[k] (own values + the composites of the children that stored theirs) (+ the carried one)
            const TreeComp &own = comp[k];
            if (carried) {
                c.I.m += own.I.m;
                for (int e = 0; e < 3; ++e) c.I.h[e] += own.I.h[e];
                for (int e = 0; e < 6; ++e) { c.I.Io[e] += own.I.Io[e]; c.H[e] += own.H[e]; c.F[e] += own.F[e]; c.B[e] += own.B[e]; }
            } else {
                c = own;
            }
        };
        auto hand_up = [&](int k) {  // after link k: its composite goes to the parent, in registers when the parent comes next
            const int par = m.parent(k);
            carried = par >= 0 && par == k - 1;
            if (par >= 0 && !carried) {
                TreeComp &p = comp[par];
                p.I.m += c.I.m;
                for (int e = 0; e < 3; ++e) p.I.h[e] += c.I.h[e];
                for (int e = 0; e < 6; ++e) { p.I.Io[e] += c.I.Io[e]; p.H[e] += c.H[e]; p.F[e] += c.F[e]; p.B[e] += c.B[e]; }
            }
        };
        auto emit = [&](int k, int j, int row, const LinkFwd &Kj, const double *rk, const double *sk, const double *gk, const double *gvk) {
            const double dqkj = -(dot6(rk, Kj.eta) + dot3(sk, Kj.xi + 3));
            const double dvkj = dot3(sk, Kj.S + 3) - 2.0 * dot6(rk, Kj.xi);
            const int e = row + m.depth(j);
            Mp[e] = dot6(rk, Kj.S) + (j == k ? m.arm(k) : 0.0);
            out.pair(k, j, e, dqkj, dvkj, j != k ? dot6(Kj.S, gk) : dqkj, j != k ? dot6(Kj.S, gvk) : dvkj);
        };
        for (int k = n - 1; k >= 0;) {
            take(k);
            double rk[6], sk[3], gk[6], gvk[6];
            FdDerivs<MP, 1>::pair_vectors(c.I, c.H, c.F, c.B, rec[k].K, rk, sk, gk, gvk);
            const int row = m.rowptr(k), p = m.parent(k);
            if (p != k - 1 || p < 0) {  // a link whose parent is elsewhere (or a root): its ancestors on their own
                for (int j = k; j >= 0; j = m.parent(j)) emit(k, j, row, rec[j].K, rk, sk, gk, gvk);
                hand_up(k);
                k -= 1;
                continue;
            }
            // link k and its parent p = k - 1 share every ancestor from p upwards: one pass, each (S, xi, eta) record read once
            emit(k, k, row, rec[k].K, rk, sk, gk, gvk);
            hand_up(k);  // carried
            take(p);
            double rp[6], sp[3], gp[6], gvp[6];
            FdDerivs<MP, 1>::pair_vectors(c.I, c.H, c.F, c.B, rec[p].K, rp, sp, gp, gvp);
            const int rowp = m.rowptr(p);
            for (int j = p; j >= 0; j = m.parent(j)) {
                const LinkFwd Kj = rec[j].K;
                emit(k, j, row, Kj, rk, sk, gk, gvk);
                emit(p, j, rowp, Kj, rp, sp, gp, gvp);
            }
            hand_up(p);
            k -= 2;
        }
    }

    // In-place M = L^T D L on the packed pattern (RBDA Table 6.3): afterwards entry (k, j), j a strict ancestor of k, holds
    // L_kj and entry (k, k) holds D_k.  Returns false on a non-positive pivot (the entries are NaN then).
    static MPCF_HD bool factorize(const MP &m, double *Mp)
    {
        const int n = m.n();
        bool ok = true;
        for (int k = n - 1; k >= 0; --k) {
            const int rk = m.rowptr(k);
            double d = Mp[rk + m.depth(k)];
            if (!(d > 0.0)) { ok = false; d = nan_value_hd(); Mp[rk + m.depth(k)] = d; }
            const double dinv = 1.0 / d;
            for (int i = m.parent(k); i >= 0; i = m.parent(i)) {
                const double a = Mp[rk + m.depth(i)] * dinv;
                const int ri = m.rowptr(i);
                for (int j = i; j >= 0; j = m.parent(j)) Mp[ri + m.depth(j)] -= a * Mp[rk + m.depth(j)];
                Mp[rk + m.depth(i)] = a;
            }
        }
        return ok;
    }
    // column j of C = M^-1 from the packed factor (entry (k, a): L_ka, entry (k, k): D_k; M = L^T D L): L^T y = e_j touches only
    // the ancestors of j, then w = D^-1 y, then L x = w row by row along the ancestor chains.  x: caller scratch [n].
    static MPCF_HD void minv_column(const MP &m, const double *Mp, int j, double *x)
    {
        const int n = m.n();
        for (int i = 0; i < n; ++i) x[i] = 0.0;
        x[j] = 1.0;
        for (int k = j; k >= 0; k = m.parent(k)) {
            const double yk = x[k];
            const int rk = m.rowptr(k);
            for (int i = m.parent(k); i >= 0; i = m.parent(i)) x[i] -= Mp[rk + m.depth(i)] * yk;
        }
        for (int k = j; k >= 0; k = m.parent(k)) x[k] /= Mp[m.rowptr(k) + m.depth(k)];
        for (int i = 0; i < n; ++i) {
            double s = x[i];
            const int ri = m.rowptr(i);
            for (int a = m.parent(i); a >= 0; a = m.parent(a)) s -= Mp[ri + m.depth(a)] * x[a];
            x[i] = s;
        }
    }
    // In place, after factorize(): the strictly-lower entries L_kj become those of L^-1 (unit lower triangular on the SAME
    // ancestor pattern: the ancestor relation is transitive, so inverting creates no fill-in); the diagonal keeps D_k.
    // Then M^-1 = L^-1 D^-1 L^-T is two triangular PRODUCTS instead of two solves (k_tree_chain_tc).  Row i, ancestor at depth
    // d:  Linv_id = -L_id - sum_{d < e < depth(i)} L_ie Linv_{anc_e(i), d};  rows in ascending order (ancestors come first).
    // path, lrow: caller scratch [n].
    static MPCF_HD void invert_unit_factor(const MP &m, double *Mp, int *path, double *lrow)
    {
        const int n = m.n();
        for (int i = 0; i < n; ++i) {
            const int di = m.depth(i), ri = m.rowptr(i);
            if (di == 0) continue;
            for (int a = m.parent(i); a >= 0; a = m.parent(a)) path[m.depth(a)] = a;
            for (int d = 0; d < di; ++d) lrow[d] = Mp[ri + d];
            for (int d = di - 1; d >= 0; --d) {
                double s = -lrow[d];
                for (int e = d + 1; e < di; ++e) s -= lrow[e] * Mp[m.rowptr(path[e]) + d];
                Mp[ri + d] = s;
            }
        }
    }
    static MPCF_HD double nan_value_hd()
    {
        const double zero = 0.0;
        return zero / zero;
    }
};

}  // namespace mpcf
