/*
 * oracle/mpcf_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU oracle for the dynamics + fatigue hot path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the product path
 * (mpc_fatigue_b200/) never does.  All batch arrays are HOST pointers in the same SoA layout the
 * CUDA C-ABI uses: component-major planes, `[component][U]`, U contiguous.
 *
 * Parity status: RNEA / frame FK / frame Jacobian / Euler / thermal ZOH are pinned against the
 * reference's stored solutions (tests/golden).  ABA / RK4 / Jacobians are
 * "parity unpinned" (the reference never computes them); they are pinned by self-consistency
 * (RNEA∘ABA = id, CRBA, finite differences, mpmath) in tests/test_oracle_*.py.
 */
#ifndef MPCF_ORACLE_H
#define MPCF_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define MPCFO_MAXN 64

typedef struct {
    int n;               /* nq = nv (1-DOF joints only) */
    int nframes;
    const int *parent;   /* [n], -1 = world */
    const int *jtype;    /* [n], 0 = revolute about local z, 1 = prismatic along local z */
    const double *Rp;    /* [n][9] joint placement rotation in the parent joint frame, row-major */
    const double *pp;    /* [n][3] joint placement translation */
    const double *mass;  /* [n] */
    const double *mc;    /* [n][3] mass * COM (joint frame) */
    const double *Io;    /* [n][6] rotational inertia about the joint origin: xx xy xz yy yz zz */
    const double *arm;   /* [n] armature (rotor inertia) */
    const double *fat;   /* [n][4] lambda, kappa, ctau = Ra/ktau^2, cv = 1/Rh */
    double grav[3];      /* gravity vector, (0,0,-9.81) */
    const int *fparent;  /* [nframes] parent joint, -1 = world */
    const double *fR;    /* [nframes][9] */
    const double *fp;    /* [nframes][3] */
} mpcfo_model;

/* coupled fatigue of a two-arm model carrying one box (core.inc.h: step_rk4_coupled) */
typedef struct {
    const int *chain_of; /* [n] arm index (0 / 1) of every joint */
    int ee_frame[2];     /* end-effector frame of each arm */
    double weight;       /* box weight m g */
} mpcfo_coupling;

/* reference-mode OCP node rows (core.inc.h: ocp_rows; row order in include/mpcf.h: mpcf_ocp_rows_batch) */
typedef struct {
    int narm;
    int ee_frame[2];
    double wsign, fdes[3], dist2_ref, mu, p_ref[3], w_box, w_qd, w_F, h;
} mpcfo_rows_opts;

int mpcfo_set_threads(int nthreads); /* returns the thread count in use (OpenMP) */

int mpcfo_rnea_batch(const mpcfo_model *m, long U, const double *q, const double *qd, const double *qdd,
                     double *tau);
int mpcfo_fk_batch(const mpcfo_model *m, int frame, long U, const double *q, double *pos, double *rot);
int mpcfo_jac_batch(const mpcfo_model *m, int frame, long U, const double *q, double *J);
int mpcfo_crba(const mpcfo_model *m, const double *q, double *M);
int mpcfo_aba_batch(const mpcfo_model *m, long U, const double *q, const double *qd, const double *tau,
                    double *qdd);
int mpcfo_node_eval_ref_batch(const mpcfo_model *m, int nee, const int *ee_frames, double wsign, long U,
                              const double *q, const double *qd, const double *qdd, const double *W,
                              const double *T, double h, double *tau, double *qnext, double *Tnext);
int mpcfo_step_rk4_batch(const mpcfo_model *m, long U, const double *q, const double *qd, const double *tau,
                         const double *f, double dt, const double *dt_u, double *qn, double *qdn, double *fn);
/* jac: [3n][4n+1][U]; rows (q+, qd+, f+), columns (q, qd, tau, f, dt) */
int mpcfo_step_rk4_jvp_batch(const mpcfo_model *m, long U, const double *q, const double *qd,
                             const double *tau, const double *f, double dt, const double *dt_u, double *qn,
                             double *qdn, double *fn, double *jac);

/* the same two entries for the coupled-fatigue step (cp: see mpcfo_coupling) */
int mpcfo_step_rk4_coupled_batch(const mpcfo_model *m, const mpcfo_coupling *cp, long U, const double *q, const double *qd,
                                 const double *tau, const double *f, double dt, const double *dt_u, double *qn, double *qdn,
                                 double *fn);
int mpcfo_step_rk4_coupled_jvp_batch(const mpcfo_model *m, const mpcfo_coupling *cp, long U, const double *q, const double *qd,
                                     const double *tau, const double *f, double dt, const double *dt_u, double *qn, double *qdn,
                                     double *fn, double *jac);

/* rows [nrows][U], cost [U] for B scenarios x N nodes (u = k*B + b); kin_jac [26 | 3][n + 3 narm][U] by complex step */
int mpcfo_ocp_rows_batch(const mpcfo_model *m, const mpcfo_rows_opts *o, long B, int N, const double *q, const double *qd,
                         const double *F, const double *T, const double *q_last, const double *T_last, const double *rel_pos0,
                         const double *rel_ori0, double *rows, double *cost, double *kin_jac);

/* exact zero-order-hold fatigue/thermal map and the ODE right-hand side, element-wise per joint */
int mpcfo_fatigue_zoh_batch(const mpcfo_model *m, long U, const double *T, const double *tau, const double *qd,
                            double h, double *Tnext);
int mpcfo_fatigue_rhs_batch(const mpcfo_model *m, long U, const double *f, const double *tau, const double *qd,
                            double *fdot);

/* complex-step derivatives: forward dynamics (A = dqdd/dq, B = dqdd/dqd, C = M^-1) and inverse dynamics */
int mpcfo_fd_derivs_batch(const mpcfo_model *m, long U, const double *q, const double *qd, const double *tau,
                          double *A, double *B, double *Cm);
int mpcfo_rnea_derivs_batch(const mpcfo_model *m, long U, const double *q, const double *qd, const double *qdd,
                            double *Dq, double *Dv, double *M);

int mpcfo_node_eval_ref_jvp_batch(const mpcfo_model *m, int nee, const int *ee_frames, double wsign, long U, const double *q,
                                  const double *qd, const double *qdd, const double *W, double *Dq, double *Dv);

#ifdef __cplusplus
}
#endif
#endif
