/* vio_oracle.h — plain-C CPU restatement of the reference's backend::Problem LM path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load liboracle.so; nothing under
 * visual-inertial-odometry_b200/ does.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 *   (a) the unmodified reference compiled into oracle/_ref/libref15.so / libref17.so (H, b, S, dx, chi2,
 *       full Solve traces) when that build is present, and
 *   (b) golden vectors committed under tests/golden/ that were generated from those libraries
 *       (tests/golden/make_golden.py), plus the reference's own known answers
 *       (TestMarginalize prior, hessian_nullspace singular values, CurveFitting result).
 *
 * Every function cites the reference file:line it follows; paths are relative to
 * /root/reference/workspace/assignments (A15 = 15-vio-backend, A17 = 17-vins-initialization/vins-mono,
 * EIG = 02-kinematics-in-3D-space/workspace/Eigen).
 */
#ifndef VIO_ORACLE_H
#define VIO_ORACLE_H
#include "../include/vio_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_prior {
    int32_t dim;             /* 0 = none; else must equal P */
    const double *H, *b;
    int32_t err_dim;
    const double *err, *jt_inv;
} orc_prior;

typedef struct orc_result {
    int32_t iterations, linearizations, trial_steps;
    int64_t pcg_iterations;
    double chi2_initial, chi2_final, lambda_initial, lambda_final;
    double ms_total, ms_hessian;
    double chi2_trace[VIO_TRACE_MAX], lambda_trace[VIO_TRACE_MAX];
} orc_result;

/* --- per-factor functions ------------------------------------------------------------------ */
/* EdgeReprojection residual + Jacobians [J_lambda(2x1) J_i(2x6) J_j(2x6)], row-major */
void orc_reproj(double inv_dep, const double *pose_i, const double *pose_j, const double *qic, const double *tic,
                const double *pts_i, const double *pts_j, double r[2], double Jl[2], double Ji[12], double Jj[12]);
void orc_se3prior(const double *pose, const double *pp, const double *qp, double r[6], double J[36]);
void orc_imu(const double *pose_i, const double *sb_i, const double *pose_j, const double *sb_j, double sum_dt,
             const double *dp, const double *dq, const double *dv, const double *lba, const double *lbg,
             const double *jac225, const double *G, double r[15], double *J /* 15 x 30 or NULL */);
void orc_loss(int kind, double delta, double e2, double rho[3]);
void orc_pose_plus(double *pose7, const double *delta6);

/* --- dense Problem (small graphs: (P+M)^2 doubles) ------------------------------------------- */
int orc_dims(const vio_graph *g, int32_t *P, int32_t *M);
/* MakeHessian at the graph's state: Hessian_ ((P+M)^2 row-major) and b_ */
int orc_make_hessian(const vio_graph *g, const orc_prior *prior, int flavour, double *H, double *b);
/* Σ (Robust)Chi2 (+ err_prior norm), with the flavour's ½ */
int orc_chi2(const vio_graph *g, const orc_prior *prior, int flavour, double *chi2);
/* SolveLinearSystem on (H, b): S (damped), bS, dx.  solver: VIO_SOLVER_DENSE_CHOL | VIO_SOLVER_REF_PCG */
int orc_solve_linear(const double *H, const double *b, int P, int M, double lambda, int solver, double *S, double *bS,
                     double *dx, int64_t *pcg_iters);
/* Problem::Solve(iterations); state returned in pose/speedbias/inv_depth (sized like the graph's) */
int orc_solve(const vio_graph *g, const orc_prior *prior, int iterations, const vio_lm_opts *opts, double *pose,
              double *speedbias, double *inv_depth, double *b_prior_out, double *err_prior_out, orc_result *res);

/* 4th Jacobian of the v17 4-vertex EdgeReprojection (extrinsic vertex not fixed), A17/src/backend/edge_reprojection.cc:97-103 */
void orc_reproj_jext(double inv_dep, const double *pose_i, const double *pose_j, const double *qic, const double *tic,
                     const double *pts_i, double Jex[12]);
/* VertexPointXYZ / EdgeReprojectionXYZ (A15/backend/edge_reprojection.cc:113-163): landmark dims are ordered
 * [n_landmark inverse depths | 3 per point]; orc_dims' M counts both. */
void orc_reproj_xyz(const double *X, const double *pose_i, const double qic[4], const double tic[3], const double *obs,
                    double r[2], double *JX /* 2x3 or NULL */, double *JT /* 2x6 or NULL */);
int orc_solve_linear_blocks(const double *H, const double *b, int P, int M1, int Mx, double lambda, int solver, double *S,
                            double *bS, double *dx, int64_t *pcg_iters);
int orc_solve_points(const vio_graph *g, const orc_prior *prior, int iterations, const vio_lm_opts *opts, double *pose,
                     double *speedbias, double *inv_depth, double *point_xyz, double *b_prior_out, double *err_prior_out,
                     orc_result *res);

/* --- block-sparse path for the large synthetic BA (configs 4/5): same per-edge arithmetic, dense
 * containers replaced by 6x6 block storage.  pattern: rowptr (C+1), col (nnzb) as returned by
 * vio_get_schur_bsr; val (nnzb*36) receives the undamped reduced system, bS (6C). ----------------- */
int orc_linearize_bsr(const vio_graph *g, const int32_t *rowptr, const int32_t *col, double *val, double *bS,
                      double *Hll, double *bl, int64_t lm_begin, int64_t lm_end);
/* one MakeHessian-equivalent pass over landmarks [lm_begin, lm_end) WITHOUT storing S: returns a checksum so the
 * work cannot be optimised away; used to time the reference dataflow per edge on a bounded sample */
int orc_linearize_sample(const vio_graph *g, int64_t lm_begin, int64_t lm_end, double *checksum);

/* --- IMU pre-integration: IntegrationBase::push_back over samples 1..n-1 after construction with sample 0
 * (A17/include/factor/integration_base.h:13-158).  noise = {ACC_N, ACC_W, GYR_N, GYR_W}. ------------------------ */
int orc_preintegrate(int32_t n, const double *dt, const double *acc, const double *gyr, const double *ba, const double *bg,
                     const double *noise, double *sum_dt, double *delta_p, double *delta_q_xyzw, double *delta_v,
                     double *jac225, double *cov225);

/* --- Problem::Marginalize({pose[marg_pose], speedbias[marg_sb]}, pose_dim = P) - A17/src/backend/problem.cc:617-795.
 * Outputs sized for P-15: H (dim x dim), b, err (dim), Jt_prior_inv (dim x dim). ----------------------------------- */
int orc_marginalize(const vio_graph *g, const orc_prior *prior, int32_t marg_pose, int32_t marg_sb, int32_t *dim_out,
                    double *H_out, double *b_out, double *err_out, double *jt_inv_out);

#ifdef __cplusplus
}
#endif
#endif
