/* ref_shim.h — C entry points of oracle/_ref/libref15.so and libref17.so.
 *
 * TEST INFRASTRUCTURE ONLY.  These libraries are the UNMODIFIED reference backend
 * (/root/reference/workspace/assignments/15-vio-backend/backend and
 *  .../17-vins-initialization/vins-mono/{include,src}/backend), compiled where the sources lie,
 * plus the shim TU in this directory that builds a reference `Problem` from a flat `vio_graph`
 * through the reference's public API and reads its private members back.  Nothing under
 * visual-inertial-odometry_b200/ may link or load them.
 */
#ifndef REF_SHIM_H
#define REF_SHIM_H
#include "../include/vio_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ref_prior {
    int32_t dim;            /* 0 = none */
    const double *H, *b;    /* dim x dim row-major, dim */
    int32_t err_dim;
    const double *err, *jt_inv; /* err_dim, err_dim x err_dim row-major */
} ref_prior;

typedef struct ref_result {
    int32_t iterations;                 /* "iter:" lines printed by Solve */
    double chi2_trace[VIO_TRACE_MAX];
    double lambda_trace[VIO_TRACE_MAX];
    double chi2_final, lambda_final;
    double ms_solve, ms_hessian;        /* the reference's own two timing lines */
} ref_result;

/* SetOrdering + MakeHessian at the given state: Hessian_ ((P+M)^2 row-major) and b_. */
int REF_FN(hessian)(const vio_graph *g, const ref_prior *prior, double *H, double *b, int32_t *P, int32_t *M);
/* + ComputeLambdaInitLM: initial chi2 and lambda */
int REF_FN(init)(const vio_graph *g, const ref_prior *prior, double *chi2, double *lambda);
/* + one SolveLinearSystem with currentLambda_ = lambda: H_pp_schur_ (damped, as the reference stores it),
 * b_pp_schur_, delta_x_ */
int REF_FN(step)(const vio_graph *g, const ref_prior *prior, double lambda, double *S, double *bS, double *dx);
/* Problem::Solve(iterations); final vertex parameters written to pose/speedbias/inv_depth; prior b/err read back */
int REF_FN(solve)(const vio_graph *g, const ref_prior *prior, int32_t iterations, double *pose, double *speedbias,
                  double *inv_depth, double *b_prior_out, double *err_prior_out, ref_result *res);
/* the same with the VertexPointXYZ estimates (n_point x 3) read back as well */
int REF_FN(solve_points)(const vio_graph *g, const ref_prior *prior, int32_t iterations, double *pose, double *speedbias,
                         double *inv_depth, double *point_xyz, double *b_prior_out, double *err_prior_out, ref_result *res);
/* v17 only (ref_sparse17.cpp): Problem::Solve restated with block-sparse containers around the reference's own
 * Edge / Vertex code - the at-scale CPU baseline and parity target (inverse-depth landmarks, fixed extrinsic vertex,
 * SE3 priors).  timing[4] = seconds in linearise, reduced solve + back-substitution, chi2 passes, total. */
int ref17_sparse_solve(const vio_graph *g, int32_t iterations, int32_t fixed_iterations, double *pose_out, double *inv_depth_out,
                       ref_result *res, double *timing);
#ifdef __cplusplus
}
#endif
#endif
