// derivs.cuh — analytic first derivatives of forward dynamics for forests of serial revolute chains:
//   A = d qdd / d q,  B = d qdd / d qd,  C = M^-1 = d qdd / d tau      (all n x n, block-diagonal per chain)
// using  d FD/d u = -M^-1 d ID/d u |_(qdd fixed)  and inverse-dynamics derivatives written in WORLD-frame
// spatial algebra, where no transform appears in the derivative pass:
//
//   S_j  = [o_j x z_j ; z_j]                           joint axis as a spatial motion at the world origin
//   xi_j = S_j x v_p(j)                                (p = parent; v, a are world-frame link velocity / acceleration,
//   eta_j = S_j x a_p(j) - xi_j x v_p(j)                 a includes the -g base acceleration)
//   I_c,k, H_c,k, F_c,k, Bs_c,k                        composites over the sub-chain rooted at k of the world-frame rigid
//                                                      inertia, momentum I v, net force f and the symmetric 3x3
//                                                      Bs = W + W^T - (h v_l^T + v_l h^T) + 2 (v_l.h) 1,  W = [w]x Io
//   B_c xi = [-2 H_l x w(xi) ; Bs w(xi) - H_a x w(xi)] (only the angular part w(xi) of xi enters)
//
//   dID_m/dq_j  = -S_m . (I_c,m eta_j + B_c,m xi_j)                    m >= j
//               =  S_m . (S_j x* F_c,j - I_c,j eta_j - B_c,j xi_j)     m <  j
//   dID_m/dqd_j =  S_m . (B_c,k S_j - 2 I_c,k xi_j),  k = max(m, j)
//   M_mj        =  S_m . I_c,k S_j (+ armature on the diagonal)
//
// Two organisations of the same formulas, both for ONE serial chain (forests run chain by chain) and both streaming — the
// forward pass computes kinematics only, the backward pass rebuilds each link from its child by undoing the joint, nothing
// is accumulated per link or per matrix:  run_cols (forward-dynamics derivatives, C = M^-1 supplied by the caller) and
// run_id_stream (inverse-dynamics derivatives and M, entry by entry).
//
// Derivation in DESIGN.md §4; checked against complex-step differentiation of the oracle's ABA
// (tests/test_gpu_parity.py::test_fd_derivs).  There is no reference code for this (north-star addition).
#pragma once

#include "dyn.cuh"

namespace mpcf {

// spatial motion x motion
MPCF_HD void mxm(const double *a, const double *b, double *o)
{
    double t0[3], t1[3];
    cross3(a + 3, b, t0);
    cross3(a, b + 3, t1);
    cross3(a + 3, b + 3, o + 3);
    o[0] = t0[0] + t1[0]; o[1] = t0[1] + t1[1]; o[2] = t0[2] + t1[2];
}
// spatial motion x* force
MPCF_HD void mxf(const double *a, const double *f, double *o)
{
    double t0[3], t1[3];
    cross3(a + 3, f, o);
    cross3(a + 3, f + 3, t0);
    cross3(a, f, t1);
    o[3] = t0[0] + t1[0]; o[4] = t0[1] + t1[1]; o[5] = t0[2] + t1[2];
}
// one out-of-line copy of sincos (its large-argument slow path is ~150 instructions per inlined call site; the derivative
// kernels are instruction-fetch bound at 110 KB of straight-line code)
static __device__ __noinline__ void sincos_shared(double x, double *s, double *c) { sincos(x, s, c); }

MPCF_HD double dot6(const double *a, const double *b)
{
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}
MPCF_HD double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

struct RigidInertiaW {  // rigid-body inertia about the world origin, world axes
    double m, h[3], Io[6];
    MPCF_HD void apply(const double *x, double *f) const
    {
        double t[3], u[3];
        cross3(h, x + 3, t);
        cross3(h, x, u);
        f[0] = m * x[0] - t[0];
        f[1] = m * x[1] - t[1];
        f[2] = m * x[2] - t[2];
        f[3] = Io[0] * x[3] + Io[1] * x[4] + Io[2] * x[5] + u[0];
        f[4] = Io[1] * x[3] + Io[3] * x[4] + Io[4] * x[5] + u[1];
        f[5] = Io[2] * x[3] + Io[4] * x[4] + Io[5] * x[5] + u[2];
    }
};

struct LinkFwd {
    double S[6], xi[6], eta[6];  // joint axis, xi, eta of one link (world frame): kept for the pairing pass
};
template <int N>
struct LocalLinkStore {
    LinkFwd K[N];
    MPCF_DI void put(int i, const LinkFwd &k) { K[i] = k; }
    MPCF_DI void get(int i, LinkFwd &k) const { k = K[i]; }
};
// slab[(i * 18 + e) * stride + tid]: consecutive threads hit consecutive 8-byte words (no bank conflicts)
struct SharedLinkStore {
    double *slab;
    int stride;
    MPCF_DI void put(int i, const LinkFwd &k)
    {
#pragma unroll
        for (int e = 0; e < 6; ++e) {
            slab[(i * 18 + e) * stride] = k.S[e];
            slab[(i * 18 + 6 + e) * stride] = k.xi[e];
            slab[(i * 18 + 12 + e) * stride] = k.eta[e];
        }
    }
    MPCF_DI void get(int i, LinkFwd &k) const
    {
#pragma unroll
        for (int e = 0; e < 6; ++e) {
            k.S[e] = slab[(i * 18 + e) * stride];
            k.xi[e] = slab[(i * 18 + 6 + e) * stride];
            k.eta[e] = slab[(i * 18 + 12 + e) * stride];
        }
    }
};

// MP must be a static forest policy (kStatic, parent(i) in {i-1, -1}, revolute only).  L = chain length.
template <class MP, int L>
struct FdDerivs {
    static constexpr int N = MP::MAXN;

    // world-frame rigid inertia I of link i (about the world origin), momentum H = I v, net force F = I a + v x* H and the
    // symmetric Bs of the header comment, from the link's world pose (R, o) and spatial velocity / acceleration (v, a)
    static MPCF_HD void link_world(const MP &m, int i, const double *R, const double *o, const double *v, const double *a,
                                   RigidInertiaW &I, double *H, double *F, double *Bsi)
    {
        const double ms = m.mass(i);
        double a3[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) a3[r] = R[3 * r] * m.mc(i, 0) + R[3 * r + 1] * m.mc(i, 1) + R[3 * r + 2] * m.mc(i, 2);
        I.m = ms;
#pragma unroll
        for (int r = 0; r < 3; ++r) I.h[r] = a3[r] + ms * o[r];
        {
            double t[9];  // t = R * Io(local, sym)
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const double r0 = R[3 * r], r1 = R[3 * r + 1], r2 = R[3 * r + 2];
                t[3 * r + 0] = r0 * m.Io(i, 0) + r1 * m.Io(i, 1) + r2 * m.Io(i, 2);
                t[3 * r + 1] = r0 * m.Io(i, 1) + r1 * m.Io(i, 3) + r2 * m.Io(i, 4);
                t[3 * r + 2] = r0 * m.Io(i, 2) + r1 * m.Io(i, 4) + r2 * m.Io(i, 5);
            }
            const int rr[6] = {0, 0, 0, 1, 1, 2}, cc[6] = {0, 1, 2, 1, 2, 2};
            const double ap = dot3(a3, o), pp = dot3(o, o);
            const double dg = 2.0 * ap + ms * pp;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const int r = rr[k], cI = cc[k];
                double e = t[3 * r] * R[3 * cI] + t[3 * r + 1] * R[3 * cI + 1] + t[3 * r + 2] * R[3 * cI + 2];
                e -= a3[r] * o[cI] + o[r] * a3[cI] + ms * o[r] * o[cI];
                if (r == cI) e += dg;
                I.Io[k] = e;
            }
        }
        double Ia[6];
        I.apply(v, H);
        I.apply(a, Ia);
        mxf(v, H, F);
#pragma unroll
        for (int k = 0; k < 6; ++k) F[k] += Ia[k];
        {
            const double *w = v + 3, *vl = v;
            // W = [w]x Io ; Bs = W + W^T - (h vl^T + vl h^T) + 2 (vl.h) 1
            const double Io[3][3] = {{I.Io[0], I.Io[1], I.Io[2]}, {I.Io[1], I.Io[3], I.Io[4]}, {I.Io[2], I.Io[4], I.Io[5]}};
            double W[3][3];
#pragma unroll
            for (int cI = 0; cI < 3; ++cI) {
                W[0][cI] = w[1] * Io[2][cI] - w[2] * Io[1][cI];
                W[1][cI] = w[2] * Io[0][cI] - w[0] * Io[2][cI];
                W[2][cI] = w[0] * Io[1][cI] - w[1] * Io[0][cI];
            }
            const double vh2 = 2.0 * dot3(vl, I.h);
            const int rr[6] = {0, 0, 0, 1, 1, 2}, cc[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const int r = rr[k], cI = cc[k];
                double e = W[r][cI] + W[cI][r] - (I.h[r] * vl[cI] + vl[r] * I.h[cI]);
                if (r == cI) e += vh2;
                Bsi[k] = e;
            }
        }
    }

    // g_k, gv_k, r_k, s_k of the pairing pass from the composites of the sub-chain rooted at k
    static MPCF_HD void pair_vectors(const RigidInertiaW &Ic, const double *Hc, const double *Fc, const double *Bc, const LinkFwd &Kk,
                                     double *rk, double *sk, double *gk, double *gvk)
    {
        const double *Sk = Kk.S;
        Ic.apply(Sk, rk);
        {   // s_k = -2 S_l x H_l + Bs S_a - S_a x H_a
            double t0[3], t1[3];
            cross3(Sk, Hc, t0);
            cross3(Sk + 3, Hc + 3, t1);
            sk[0] = Bc[0] * Sk[3] + Bc[1] * Sk[4] + Bc[2] * Sk[5] - 2.0 * t0[0] - t1[0];
            sk[1] = Bc[1] * Sk[3] + Bc[3] * Sk[4] + Bc[4] * Sk[5] - 2.0 * t0[1] - t1[1];
            sk[2] = Bc[2] * Sk[3] + Bc[4] * Sk[4] + Bc[5] * Sk[5] - 2.0 * t0[2] - t1[2];
        }
        // g_k = S_k x* F_c - I_c eta_k - B_c xi_k ;  gv_k = B_c S_k - 2 I_c xi_k
        double t6[6], u6[6];
        mxf(Sk, Fc, gk);
        Ic.apply(Kk.eta, t6);
        Ic.apply(Kk.xi, u6);
        const double *wx = Kk.xi + 3, *ws = Sk + 3;
        double bx[6], bs[6], t0[3];
        cross3(Hc, wx, t0);  // H_l x w
        bx[0] = -2.0 * t0[0]; bx[1] = -2.0 * t0[1]; bx[2] = -2.0 * t0[2];
        cross3(Hc + 3, wx, t0);
        bx[3] = Bc[0] * wx[0] + Bc[1] * wx[1] + Bc[2] * wx[2] - t0[0];
        bx[4] = Bc[1] * wx[0] + Bc[3] * wx[1] + Bc[4] * wx[2] - t0[1];
        bx[5] = Bc[2] * wx[0] + Bc[4] * wx[1] + Bc[5] * wx[2] - t0[2];
        cross3(Hc, ws, t0);
        bs[0] = -2.0 * t0[0]; bs[1] = -2.0 * t0[1]; bs[2] = -2.0 * t0[2];
        cross3(Hc + 3, ws, t0);
        bs[3] = Bc[0] * ws[0] + Bc[1] * ws[1] + Bc[2] * ws[2] - t0[0];
        bs[4] = Bc[1] * ws[0] + Bc[3] * ws[1] + Bc[4] * ws[2] - t0[1];
        bs[5] = Bc[2] * ws[0] + Bc[4] * ws[1] + Bc[5] * ws[2] - t0[2];
#pragma unroll
        for (int e = 0; e < 6; ++e) {
            gk[e] -= t6[e] + bx[e];
            gvk[e] = bs[e] - 2.0 * u6[e];
        }
    }

    // link kinematics in the world frame: pose (R, o) of link i from its parent's, joint axis S, xi, eta; advances (v, a)
    static MPCF_DI void link_kin(const MP &m, int i, double c, double s, double qdi, double qddi, double *oR, double *o, double *v,
                                 double *a, LinkFwd &Ki)
    {
        const bool root = m.parent(i) < 0;
        double Rl[9], R[9], pos[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            Rl[3 * r + 0] = m.Rp(i, 3 * r) * c + m.Rp(i, 3 * r + 1) * s;
            Rl[3 * r + 1] = m.Rp(i, 3 * r + 1) * c - m.Rp(i, 3 * r) * s;
            Rl[3 * r + 2] = m.Rp(i, 3 * r + 2);
        }
        if (root) {
#pragma unroll
            for (int k = 0; k < 9; ++k) R[k] = Rl[k];
#pragma unroll
            for (int k = 0; k < 3; ++k) pos[k] = m.pp(i, k);
#pragma unroll
            for (int k = 0; k < 6; ++k) { v[k] = 0.0; a[k] = 0.0; }
            a[0] = -m.grav(0); a[1] = -m.grav(1); a[2] = -m.grav(2);
        } else {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) R[3 * r + cc] = oR[3 * r] * Rl[cc] + oR[3 * r + 1] * Rl[3 + cc] + oR[3 * r + 2] * Rl[6 + cc];
                pos[r] = o[r] + oR[3 * r] * m.pp(i, 0) + oR[3 * r + 1] * m.pp(i, 1) + oR[3 * r + 2] * m.pp(i, 2);
            }
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) oR[k] = R[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) o[k] = pos[k];
        const double z[3] = {R[2], R[5], R[8]};
        cross3(o, z, Ki.S);
        Ki.S[3] = z[0]; Ki.S[4] = z[1]; Ki.S[5] = z[2];
        // xi = S x v_parent ; eta = S x a_parent - xi x v_parent   (v, a still hold the parent's values)
        mxm(Ki.S, v, Ki.xi);
        double t6[6];
        mxm(Ki.S, a, Ki.eta);
        mxm(Ki.xi, v, t6);
#pragma unroll
        for (int k = 0; k < 6; ++k) Ki.eta[k] -= t6[k];
        // v_i = v_p + S qd ; a_i = a_p + S qdd - xi qd
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            v[k] += Ki.S[k] * qdi;
            a[k] += Ki.S[k] * qddi - Ki.xi[k] * qdi;
        }
    }

    // adds link k (the last link of the chain starts the sums) to the composites of the sub-chain rooted at k
    static MPCF_DI void accumulate_link(const MP &m, int k, const double *R, const double *o, const double *v, const double *a,
                                        RigidInertiaW &Ic, double *Hc, double *Fc, double *Bc)
    {
        RigidInertiaW Ik;
        double Hk[6], Fk[6], Bk[6];
        link_world(m, k, R, o, v, a, Ik, Hk, Fk, Bk);
        if (k == N - 1) {
            Ic = Ik;
#pragma unroll
            for (int e = 0; e < 6; ++e) { Hc[e] = Hk[e]; Fc[e] = Fk[e]; Bc[e] = Bk[e]; }
        } else {
            Ic.m += Ik.m;
#pragma unroll
            for (int e = 0; e < 3; ++e) Ic.h[e] += Ik.h[e];
#pragma unroll
            for (int e = 0; e < 6; ++e) { Ic.Io[e] += Ik.Io[e]; Hc[e] += Hk[e]; Fc[e] += Fk[e]; Bc[e] += Bk[e]; }
        }
    }

    // (R, o, v, a) of link k -> of link k - 1:  v_p = v - S qd, a_p = a - S qdd + xi qd, R_p = R Rl^T, o_p = o - R_p pp
    static MPCF_DI void undo_joint(const MP &m, int k, double c, double s, double qdk, double qddk, const LinkFwd &Kk, double *R, double *o,
                                   double *v, double *a)
    {
#pragma unroll
        for (int e = 0; e < 6; ++e) {
            v[e] -= Kk.S[e] * qdk;
            a[e] -= Kk.S[e] * qddk - Kk.xi[e] * qdk;
        }
        double Rl[9], Rp[9];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            Rl[3 * r + 0] = m.Rp(k, 3 * r) * c + m.Rp(k, 3 * r + 1) * s;
            Rl[3 * r + 1] = m.Rp(k, 3 * r + 1) * c - m.Rp(k, 3 * r) * s;
            Rl[3 * r + 2] = m.Rp(k, 3 * r + 2);
        }
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) Rp[3 * r + cc] = R[3 * r] * Rl[3 * cc] + R[3 * r + 1] * Rl[3 * cc + 1] + R[3 * r + 2] * Rl[3 * cc + 2];
#pragma unroll
        for (int r = 0; r < 3; ++r) o[r] -= Rp[3 * r] * m.pp(k, 0) + Rp[3 * r + 1] * m.pp(k, 1) + Rp[3 * r + 2] * m.pp(k, 2);
#pragma unroll
        for (int e = 0; e < 9; ++e) R[e] = Rp[e];
    }

    // Column-streamed variant for ONE serial chain (N == L) used by the Jacobian pipeline: C = M^-1 is supplied by the
    // caller (Cs(r, c) returns C[r][c] for c <= r, valid after Cs.ready(); the RK4 stage kernel has it from its LDL^T factors), so
    // column k of A = -C dID/dq and B = -C dID/dqd can be emitted as soon as the backward pass reaches link k:
    //   * the forward pass computes link kinematics only and keeps nothing per link besides (S, xi, eta) in `ks`
    //     (links 0 .. N-2; the last link's stay in registers);
    //   * the backward pass rebuilds each link's inertial quantities from its pose and (v, a), which it recovers from the
    //     child's by undoing the joint (v_p = v - S qd, a_p = a - S qdd + xi qd, R_p = R Rl^T, o_p = o - R_p pp);
    //   * at link k only the entries dID_k'/dq_j with k' > k > j are still pending (at most N^2 / 4 per matrix).
    // emit(mat, row, col, value) is called for mat 0 (A) and 1 (B) only.
    // The m.skip() branches (never taken, see StaticModel::skip) cut the fully unrolled body into one basic block per link.
    template <class Emit, class KS, class CGet>
    static MPCF_DI void run_cols(const MP &m, const double *q, const double *qd, const double *qdd, CGet Cs, Emit emit, KS &ks)
    {
        static_assert(N == L, "run_cols handles a single serial chain");
        double cs[N], sn[N];
        double R[9], o[3], v[6], a[6];
        LinkFwd Klast;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (m.skip(i)) continue;
            sincos_shared(q[i], &sn[i], &cs[i]);
            LinkFwd Ki;
            link_kin(m, i, cs[i], sn[i], qd[i], qdd[i], R, o, v, a, Ki);
            if (i == N - 1) Klast = Ki;  // the backward pass starts at this link: no round trip through the store
            else ks.put(i, Ki);
        }
        Cs.ready();
        double Dq[N][N], Dv[N][N];
        RigidInertiaW Ic;
        double Hc[6], Fc[6], Bc[6];
#pragma unroll
        for (int k = N - 1; k >= 0; --k) {
            if (m.skip(k)) continue;  // never taken: keeps each link's work in its own basic block
            accumulate_link(m, k, R, o, v, a, Ic, Hc, Fc, Bc);
            LinkFwd Kk;
            if (k == N - 1) Kk = Klast; else ks.get(k, Kk);
            double rk[6], sk[3], gk[6], gvk[6];
            pair_vectors(Ic, Hc, Fc, Bc, Kk, rk, sk, gk, gvk);
#pragma unroll
            for (int j = 0; j <= k; ++j) {
                LinkFwd Kj;
                if (j == k) Kj = Kk; else ks.get(j, Kj);
                Dq[k][j] = -(dot6(rk, Kj.eta) + dot3(sk, Kj.xi + 3));
                Dv[k][j] = dot3(sk, Kj.S + 3) - 2.0 * dot6(rk, Kj.xi);
                if (j < k) {
                    Dq[j][k] = dot6(Kj.S, gk);
                    Dv[j][k] = dot6(Kj.S, gvk);
                }
            }
            // column k is complete: rows below k were filled at earlier links
#pragma unroll
            for (int i = 0; i < N; ++i) {
                double xa = 0.0, xb = 0.0;
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    const double c = i >= r ? Cs(i, r) : Cs(r, i);
                    xa -= c * Dq[r][k];
                    xb -= c * Dv[r][k];
                }
                emit(0, i, k, xa);
                emit(1, i, k, xb);
            }
            if (k > 0) undo_joint(m, k, cs[k], sn[k], qd[k], qdd[k], Kk, R, o, v, a);
        }
    }
    // Streaming inverse-dynamics derivatives for ONE serial chain (N == L): nothing is accumulated.  hooks.link(i, R, o) sees
    // every link's world pose in the forward pass; hooks.pair(k, j, Kk, Kj, dq_kj, dq_jk, dv_kj, dv_jk, m_kj) receives, for
    // every j <= k, the entries dID_k/dq_j, dID_j/dq_k, dID_k/dqd_j, dID_j/dqd_k and M_kj = M_jk (armature included on the
    // diagonal; for j == k the two orientations coincide) together with the two joints' (S, xi, eta).
    // Same organisation as run_cols: kinematics-only forward pass, links rebuilt on the way back by undoing joints.
    template <class Hooks, class KS>
    static MPCF_DI void run_id_stream(const MP &m, const double *q, const double *qd, const double *qdd, Hooks &hooks, KS &ks)
    {
        static_assert(N == L, "run_id_stream handles a single serial chain");
        double cs[N], sn[N];
        double R[9], o[3], v[6], a[6];
        LinkFwd Klast;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (m.skip(i)) continue;
            sincos(q[i], &sn[i], &cs[i]);
            LinkFwd Ki;
            link_kin(m, i, cs[i], sn[i], qd[i], qdd[i], R, o, v, a, Ki);
            hooks.link(i, R, o);
            if (i == N - 1) Klast = Ki;
            else ks.put(i, Ki);
        }
        RigidInertiaW Ic;
        double Hc[6], Fc[6], Bc[6];
#pragma unroll
        for (int k = N - 1; k >= 0; --k) {
            if (m.skip(k)) continue;
            accumulate_link(m, k, R, o, v, a, Ic, Hc, Fc, Bc);
            LinkFwd Kk;
            if (k == N - 1) Kk = Klast; else ks.get(k, Kk);
            double rk[6], sk[3], gk[6], gvk[6];
            pair_vectors(Ic, Hc, Fc, Bc, Kk, rk, sk, gk, gvk);
#pragma unroll
            for (int j = 0; j <= k; ++j) {
                LinkFwd Kj;
                if (j == k) Kj = Kk; else ks.get(j, Kj);
                const double dqkj = -(dot6(rk, Kj.eta) + dot3(sk, Kj.xi + 3));
                const double dvkj = dot3(sk, Kj.S + 3) - 2.0 * dot6(rk, Kj.xi);
                const double mkj = dot6(rk, Kj.S) + (j == k ? m.arm(k) : 0.0);
                hooks.pair(k, j, Kk, Kj, dqkj, j < k ? dot6(Kj.S, gk) : dqkj, dvkj, j < k ? dot6(Kj.S, gvk) : dvkj, mkj);
            }
            if (k > 0) undo_joint(m, k, cs[k], sn[k], qd[k], qdd[k], Kk, R, o, v, a);
        }
    }
};

}  // namespace mpcf
