/* vio_b200.h — C-ABI of the B200-native backend::Problem least-squares hot path.
 *
 * This is the drop-in boundary: plain C, opaque handle, int error codes, caller-owned
 * host arrays borrowed for the duration of a call, no C++/Eigen/torch types.  The C++
 * `myslam::backend::Problem` mirror (include/backend/problem.h) and the ctypes binding
 * (visual-inertial-odometry_b200/capi.py) are both thin clients of exactly these symbols.
 *
 * Reference interfaces replaced (paths relative to /root/reference/workspace/assignments;
 * A15 = 15-vio-backend, A17 = 17-vins-initialization/vins-mono):
 *   vio_set_graph          <- Problem::AddVertex/AddEdge + SetOrdering   A15/backend/problem.cc:40-54,91-103,224-262
 *                                                                         A17/src/backend/problem.cc:44-58,108-121,256-285
 *   vio_set_prior/get_prior<- Set/Get{HessianPrior,bPrior,ErrPrior,JtPrior}, ExtendHessiansPriorSize
 *                                                                         A17/include/backend/problem.h:78-88, A17/src/backend/problem.cc:83-92
 *   vio_linearize          <- Problem::MakeHessian                        A15/backend/problem.cc:280-337, A17/src/backend/problem.cc:303-389
 *   vio_solve_step         <- Problem::SolveLinearSystem                  A15/backend/problem.cc:342-423, A17/src/backend/problem.cc:394-449
 *   vio_solve              <- Problem::Solve(iterations)                  A15/backend/problem.cc:155-222, A17/src/backend/problem.cc:169-250
 *   vio_chi2               <- Σ Edge::Chi2 / RobustChi2 (+ prior)         A15/backend/problem.cc:457-462,501-507, A17/src/backend/problem.cc:501-507,549-556
 *   vio_get_vertices       <- Vertex::Parameters() read-back after Solve  A15/app/TestMonoBA.cpp:193-216
 *
 * All floating point is IEEE double.  Quaternions are stored [x y z w] inside a pose
 * record [tx ty tz qx qy qz qw] exactly like VertexPose::Parameters() (A15/backend/vertex_pose.h:12-14).
 */
#ifndef VIO_B200_H
#define VIO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vio_problem vio_problem; /* opaque; one per host thread + CUDA stream */

/* ---- error codes (0 = success) ------------------------------------------------------- */
enum {
    VIO_OK = 0,
    VIO_ERR_INVALID = 1,     /* bad argument / inconsistent graph description            */
    VIO_ERR_CUDA = 2,        /* a CUDA runtime call failed (see vio_last_error)          */
    VIO_ERR_UNSUPPORTED = 3, /* graph uses a feature outside the device path (documented)*/
    VIO_ERR_EMPTY = 4,       /* Solve on a graph without edges or vertices (reference: Solve returns false) */
    VIO_ERR_NO_DEVICE = 5,   /* no CUDA device: the product path never falls back to CPU  */
    VIO_ERR_STATE = 6        /* call order violated (e.g. solve before set_graph)        */
};

/* ---- enumerations ---------------------------------------------------------------------- */
enum { VIO_LM_V15 = 0, VIO_LM_V17 = 1 };       /* which reference generation's LM constants  */
enum {
    VIO_SOLVER_AUTO = 0,
    VIO_SOLVER_DENSE_CHOL = 1, /* v17: S.ldlt().solve  (A17/src/backend/problem.cc:439)      */
    VIO_SOLVER_REF_PCG = 2,    /* v15: PCGSolver incl. its missing first x update
                                  (A15/backend/problem.cc:530-560)                          */
    VIO_SOLVER_BLOCK_PCG = 3,  /* large BA: 6x6 block-Jacobi PCG on block-sparse S           */
    VIO_SOLVER_BLOCK_CHOL = 5,  /* block-sparse Cholesky on the 6x6 BSR pattern in the natural pose order (symbolic
                                  factorisation on the host once per graph, numeric right-looking factorisation and the
                                  two triangular solves on the device): the exact "block-Cholesky reduced solve" of
                                  BASELINE config 4                                                    */
    VIO_SOLVER_BLOCK_PCG_2L = 4, /* the same PCG with a two-level preconditioner: block-Jacobi + Galerkin coarse
                                  correction over aggregates of consecutive pose blocks (camera chains).
                                  AUTO picks it for block-sparse S with >= 256 pose blocks whose pattern VIO_SOLVER_BCR
                                  does not cover.                                                       */
    VIO_SOLVER_BCR = 6          /* block cyclic reduction: EXACT solve (like S.ldlt().solve, A17/src/backend/problem.cc:439)
                                  for a camera chain / ring, i.e. a block-sparse S that is a cyclic block band of half
                                  bandwidth <= 12 pose blocks in creation order (fixed, uncoupled pose blocks allowed
                                  anywhere).  Nested-dissection Cholesky over dense super-blocks, log2(#cameras / bandwidth)
                                  levels, one persistent kernel.  AUTO picks it whenever the pattern qualifies;
                                  VIO_ERR_UNSUPPORTED when requested explicitly on another pattern.        */
};
enum { VIO_LOSS_TRIVIAL = 0, VIO_LOSS_HUBER = 1, VIO_LOSS_CAUCHY = 2, VIO_LOSS_TUKEY = 3 };
enum { VIO_STORAGE_AUTO = 0, VIO_STORAGE_DENSE = 1, VIO_STORAGE_BSR = 2 };

/* ---- graph description (host SoA, borrowed during vio_set_graph) ------------------------ */
typedef struct vio_graph {
    /* pose-class vertices.  Reference SetOrdering gives every pose-class vertex a slot in
     * creation order, fixed or not (A17/src/backend/problem.cc:256-285). */
    int32_t n_pose;
    const double *pose;          /* n_pose x 7  [t(3) q(xyzw)]                             */
    const uint8_t *pose_fixed;   /* n_pose, may be NULL (= none fixed)                     */
    int32_t n_speedbias;
    const double *speedbias;     /* n_speedbias x 9 [v ba bg]  (v17 VertexSpeedBias)       */
    const uint8_t *speedbias_fixed;
    /* creation order of the pose-class vertices: entry >= 0 -> pose index, entry < 0 ->
     * speed-bias index ~entry.  NULL = all poses first, then all speed-biases.            */
    const int32_t *pclass_order; /* n_pose + n_speedbias                                   */

    /* landmarks: VertexInverseDepth (A15/backend/vertex_inverse_depth.h:12-18)            */
    int32_t n_landmark;
    const double *inv_depth;     /* n_landmark                                             */

    /* EdgeReprojection (A15/backend/edge_reprojection.cc:20-111; A17/src/backend/edge_reprojection.cc:18-108).
     * Preconditions checked by vio_set_graph (all reference drivers satisfy them): every edge of a
     * landmark names the same host pose and the same host observation pts_i; information = rp_info*I2. */
    int64_t n_reproj;
    const int32_t *rp_landmark;  /* n_reproj                                               */
    const int32_t *rp_pose_i;    /* host pose index                                        */
    const int32_t *rp_pose_j;    /* observing pose index                                   */
    const double *rp_pts_i;      /* n_reproj x 3                                           */
    const double *rp_pts_j;      /* n_reproj x 2 (only x,y of pts_j are read by the reference) */
    double rp_info;              /* information = rp_info * I2                             */
    int32_t rp_loss;             /* VIO_LOSS_*  (A17/src/backend/loss_function.cc)          */
    double rp_loss_delta;
    /* camera->body extrinsics.  ext_pose < 0: constants q_ic/t_ic (v15 SetTranslationImuFromCamera,
     * A15/backend/edge_reprojection.cc:42-45).  ext_pose >= 0: index of the extrinsic VertexPose
     * (v17 4-vertex edge).  Fixed (ESTIMATE_EXTRINSIC=0, A17/src/estimator.cpp:915-932): any handle.  Free (the edge's
     * 4th Jacobian, A17/src/backend/edge_reprojection.cc:97-103): dense storage, unsharded, single problem.          */
    int32_t ext_pose;
    double q_ic[4];              /* xyzw                                                   */
    double t_ic[3];

    /* EdgeSE3Prior (A15/backend/edge_prior.cpp:39-80)                                      */
    int32_t n_se3prior;
    const int32_t *sp_pose;
    const double *sp_p;          /* n x 3                                                  */
    const double *sp_q;          /* n x 4 xyzw                                             */
    const double *sp_info;       /* n x 36 row-major 6x6                                   */

    /* EdgeImu (A17/src/backend/edge_imu.cc:13-157) + IntegrationBase constants
     * (A17/include/factor/integration_base.h:160-186)                                      */
    int32_t n_imu;
    const int32_t *imu_pose_i, *imu_sb_i, *imu_pose_j, *imu_sb_j;
    const double *imu_sum_dt;    /* n                                                      */
    const double *imu_delta_p;   /* n x 3                                                  */
    const double *imu_delta_q;   /* n x 4 xyzw                                             */
    const double *imu_delta_v;   /* n x 3                                                  */
    const double *imu_lin_ba;    /* n x 3                                                  */
    const double *imu_lin_bg;    /* n x 3                                                  */
    const double *imu_jacobian;  /* n x 225 row-major 15x15                                */
    const double *imu_covariance;/* n x 225 row-major 15x15                                */
    double gravity[3];           /* global G (A17/include/parameters.h:49)                 */

    int32_t storage;             /* VIO_STORAGE_* for the reduced camera system            */

    /* VertexPointXYZ (A15/backend/vertex_point_xyz.h:12-18) + EdgeReprojectionXYZ, the 2-vertex factor [X_w, T_i]
     * (A15/backend/edge_reprojection.cc:113-163; same code in A17/src/backend/edge_reprojection.cc:130-180).  Points
     * are landmark-class vertices of local dimension 3, ordered after the inverse-depth landmarks:
     * Hessian_ = [P | n_landmark | 3 n_point].  Information = rp_info*I2, loss rp_loss (shared with EdgeReprojection);
     * extrinsics q_ic/t_ic (or the fixed ext_pose vertex).  May be mixed with inverse-depth landmarks.
     * Not supported together with vio_set_shard, vio_marginalize or the lock-step batch.              */
    int32_t n_point;
    int32_t reserved_xyz;
    const double *point_xyz;     /* n_point x 3 world coordinates                          */
    int64_t n_reproj_xyz;
    const int32_t *rx_point;     /* n_reproj_xyz                                           */
    const int32_t *rx_pose;      /* observing pose index                                   */
    const double *rx_obs;        /* n_reproj_xyz x 2 (obs_.head<2>())                      */

    /* Fixed landmark-class vertices (Vertex::SetFixed on a VertexInverseDepth / VertexPointXYZ), may be NULL (= none).
     * MakeHessian skips the Jacobian blocks of ANY fixed vertex (A17/src/backend/problem.cc:325,340): H_ll, b_l and the
     * H_lp rows of a fixed landmark are zero, the pose blocks of its edges are kept.  The reference's own Schur step
     * then inverts that zero H_mm block (problem.cc:421-425: inf, NaN in S - every trial step is rejected); here the
     * block is left out of the Schur complement and of the back-substitution, i.e. the landmark is a constant.        */
    const uint8_t *landmark_fixed; /* n_landmark                                            */
    const uint8_t *point_fixed;    /* n_point                                               */
} vio_graph;

/* ---- LM options / statistics -------------------------------------------------------------- */
typedef struct vio_lm_opts {
    int32_t flavour;        /* VIO_LM_V15 | VIO_LM_V17                                       */
    int32_t solver;         /* VIO_SOLVER_*                                                  */
    int32_t verbose;        /* 1: print the reference's "iter: i , chi= .. , Lambda= .." lines */
    int32_t pcg_max_iter;   /* block PCG cap; <=0: 2*P like the reference call site          */
    double pcg_tol;         /* relative residual; <=0: 1e-6 (A15/backend/problem.cc:544)      */
    int32_t fixed_iterations; /* 1: ignore the reference's convergence stop rules (benchmark) */
    int32_t warm_start;     /* 1: continue the previous vio_solve on this handle (keep its linearisation,
                               lambda, chi2, nu) instead of MakeHessian + ComputeLambdaInitLM            */
} vio_lm_opts;

#define VIO_TRACE_MAX 256
typedef struct vio_stats {
    int32_t iterations;       /* outer LM iterations executed                               */
    int32_t linearizations;   /* MakeHessian-equivalent passes                              */
    int32_t trial_steps;      /* SolveLinearSystem-equivalent solves                        */
    int32_t accepted_steps;
    int64_t pcg_iterations;   /* summed over all reduced solves                             */
    double chi2_initial, chi2_final;
    double lambda_initial, lambda_final;
    double ms_total;          /* CUDA-event time of the whole solve on the handle's stream  */
    double ms_linearize;      /* Σ linearise+accumulate+Schur kernels                       */
    double ms_reduced_solve;  /* Σ reduced-solve kernels (timed launches: PCG / coarse refresh / cyclic reduction) */
    double ms_backsub_update;
    double ms_chi2;
    int32_t n_trace;          /* min(iterations, VIO_TRACE_MAX)                             */
    int32_t solver_used;      /* the VIO_SOLVER_* that VIO_SOLVER_AUTO resolved to            */
    double chi2_trace[VIO_TRACE_MAX];   /* currentChi_ printed at the top of each iteration */
    double lambda_trace[VIO_TRACE_MAX];
} vio_stats;

/* sizes of the assembled system, for the debug taps */
typedef struct vio_dims {
    int32_t P;                /* pose-class dimension (ordering_poses_)                     */
    int32_t M;                /* landmark dimension   (ordering_landmarks_)                 */
    int32_t n_pose_blocks;    /* number of pose-class vertices                              */
    int32_t storage;          /* resolved VIO_STORAGE_*                                     */
    int64_t nnz_blocks;       /* BSR: number of stored 6x6 blocks; dense: 0                 */
    int64_t n_reproj;
    int32_t n_groups;         /* landmark groups (CTA work items) built by the packer       */
    int32_t reserved;
} vio_dims;

/* collective hook for the landmark-sharded multi-GPU path: sum `count` doubles in place at
 * device pointer `dev_ptr` across ranks, ordered on `cuda_stream`.  NULL = single GPU.      */
typedef int (*vio_allreduce_fn)(void *dev_ptr, int64_t count, void *cuda_stream, void *user);

/* ---- lifecycle ------------------------------------------------------------------------------ */
int vio_create(int device, void *cuda_stream /* cudaStream_t or NULL = own stream */, vio_problem **out);
void vio_destroy(vio_problem *p);
const char *vio_last_error(const vio_problem *p);
const char *vio_version(void);
/* sizeof of the ABI structs as this library was compiled (0 vio_graph, 1 vio_lm_opts, 2 vio_stats, 3 vio_dims):
 * lets a foreign-language binding verify its struct layout at load time */
size_t vio_struct_size(int which);
int vio_device_count(void);

/* ---- graph / state -------------------------------------------------------------------------- */
int vio_set_graph(vio_problem *p, const vio_graph *g);
int vio_get_dims(const vio_problem *p, vio_dims *out);
/* the caller's landmark indices this handle (rank) owns, in its packed order (all landmarks when unsharded): at most `cap`
 * entries are written, *count receives their number */
int vio_get_owned_landmarks(const vio_problem *p, int32_t *out, int64_t cap, int64_t *count);
int vio_set_allreduce(vio_problem *p, vio_allreduce_fn fn, void *user);
/* Native multi-GPU path: one process (or thread) per GPU, NCCL over NVLink / NVSwitch, no callback into the host
 * language.  The library loads libnccl.so.2 at run time (dlopen; VIO_ERR_UNSUPPORTED when it is absent) and issues its
 * collectives (one all-reduce of the reduced system per linearisation, three scalar all-reduces per trial step) on the
 * handle's stream.
 *   vio_nccl_unique_id : rank 0 creates the 128-byte ncclUniqueId; the caller ships it to the other ranks (MPI, a file,
 *                        torch.distributed.broadcast_object_list, ...)
 *   vio_nccl_init      : collective over all ranks - creates the communicator (owned by the handle) and sets the
 *                        landmark shard (rank, world); call before vio_set_graph
 *   vio_set_nccl_comm  : adopt an existing ncclComm_t (passed as void*) created elsewhere with the same libnccl      */
int vio_nccl_unique_id(void *id128);
int vio_nccl_init(vio_problem *p, int rank, int world, const void *id128);
int vio_set_nccl_comm(vio_problem *p, void *nccl_comm, int rank, int world);
/* Both calls also set up (collectively) an NVLink peer-memory mailbox between the ranks' handles (CUDA IPC): the small
 * all-reduces of the distributed reduced solve - interface system, pose update, LM scalars - then run as ONE kernel of this
 * library that stores into the peers' memory and sums in rank order (csrc/vio_p2p.cuh) instead of an ncclAllReduce; larger
 * reductions stay on NCCL.  Falls back to NCCL on every rank when any rank cannot export / open a mailbox
 * (VIO_B200_NO_P2P=1 forces the fallback).  vio_p2p_enabled: 1 when the mailbox path is active on this handle. */
int vio_p2p_enabled(const vio_problem *p);
/* landmark sharding (call before vio_set_graph with the FULL graph on every rank): rank r keeps a
 * contiguous, edge-balanced range of landmarks and their observations; pose-class vertices, the
 * reduced-system sparsity pattern and the LM scalars are replicated. Pose-only factors (SE3 prior,
 * IMU, dense prior) are accumulated by rank 0.                                                     */
int vio_set_shard(vio_problem *p, int rank, int world);
/* v17 marginalisation prior; dim must equal P (after ExtendHessiansPriorSize semantics applied by caller).
 * err/jt_inv may be NULL (no err_prior_ => chi2 has no prior term, like an empty err_prior_). err_dim rows of jt_inv (err_dim x err_dim used on head(P-15)). */
int vio_set_prior(vio_problem *p, int32_t dim, const double *H_prior, const double *b_prior,
                  int32_t err_dim, const double *err_prior, const double *Jt_prior_inv);
int vio_get_prior(vio_problem *p, double *b_prior /* P */, double *err_prior /* err_dim */);
/* overwrite the current estimates (used by per-iteration re-synchronised parity tests)       */
int vio_set_vertices(vio_problem *p, const double *pose, const double *speedbias, const double *inv_depth);
int vio_get_vertices(vio_problem *p, double *pose, double *speedbias, double *inv_depth);
/* VertexPointXYZ estimates (n_point x 3) */
int vio_set_points(vio_problem *p, const double *point_xyz);
int vio_get_points(vio_problem *p, double *point_xyz);
/* debug tap on the point blocks of the last linearisation / step: Hmm 3x3 blocks (n_point x 9 row-major), b (n_point x 3),
 * delta_x (n_point x 3); any pointer may be NULL */
int vio_get_point_system(vio_problem *p, double *Hmm, double *b, double *dx);

/* ---- the hot path --------------------------------------------------------------------------- */
int vio_solve(vio_problem *p, int32_t iterations, const vio_lm_opts *opts, vio_stats *stats);

/* step-level entry points (what Solve is made of; also the parity taps)                      */
int vio_linearize(vio_problem *p, const vio_lm_opts *opts);         /* MakeHessian + Schur, at current state  */
int vio_chi2(vio_problem *p, const vio_lm_opts *opts, double *chi2);/* Σ (Robust)Chi2 (+prior) with flavour's ½ */
int vio_solve_step(vio_problem *p, const vio_lm_opts *opts, double lambda, int64_t *pcg_iters); /* SolveLinearSystem */
int vio_apply_step(vio_problem *p, const vio_lm_opts *opts);        /* UpdateStates                           */
int vio_rollback_step(vio_problem *p, const vio_lm_opts *opts);     /* RollbackStates                         */

/* debug taps (host output arrays).  Any pointer may be NULL.                                 */
/* full (P+M)^2 Hessian_ and b_ as the reference holds them after MakeHessian (small graphs only: (P+M) <= 8192) */
int vio_get_hessian(vio_problem *p, const vio_lm_opts *opts, double *H /* (P+M)^2 row-major */, double *b /* P+M */);
/* undamped reduced system: dense P x P (lambda not added) and b_pp_schur_                    */
int vio_get_schur(vio_problem *p, double *S /* P*P row-major */, double *bS /* P */);
/* block-sparse view of the same (BSR storage only)                                          */
int vio_get_schur_bsr(vio_problem *p, int32_t *rowptr /* nb+1 */, int32_t *col /* nnzb */, double *val /* nnzb*36 */, double *bS);
/* debug tap on the two-level PCG preconditioner of the last solve: coarse dimension, block rows per aggregate, the
 * explicit coarse inverse (nc x nc) and the basis Z ([pose block][6][7]); VIO_ERR_STATE if the last solve did not use it. */
int vio_get_coarse(vio_problem *p, int32_t *nc, int32_t *rows_per_aggregate, double *Ainv, double *Z);
int vio_get_delta(vio_problem *p, double *dx_pose /* P */, double *dx_landmark /* M */);
int vio_get_b(vio_problem *p, double *b_pose /* P */, double *b_landmark /* M */);
int vio_get_landmark_diag(vio_problem *p, double *Hmm /* M */);

/* last-kernel timing tap for bench.py: average ms of the linearise kernel over the last solve */
int vio_get_kernel_ms(vio_problem *p, double *ms_linearize_kernel, int64_t *launches);
/* the same for the block-PCG of the last solve: average ms of one k_bpcg_persistent launch (CUDA events on the handle's
 * stream), launches timed, PCG iterations summed over them, average ms of one coarse-preconditioner refresh (basis +
 * Galerkin assembly + inversion) and how many refreshes the lagged-inverse policy made */
int vio_get_solver_ms(vio_problem *p, double *ms_pcg_kernel, int64_t *pcg_launches, double *pcg_iterations,
                      double *ms_coarse_setup, int64_t *coarse_refreshes);
/* total number of kernel launches issued by this handle since creation                      */
int64_t vio_launch_count(const vio_problem *p);

/* ---- Problem::Marginalize(margVertexs = {pose[marg_pose], speedbias[marg_sb]}, pose_dim = P)
 * (A17/src/backend/problem.cc:617-795) on the handle's current graph, state and prior: edges connected to the frame
 * are re-linearised (no vertex treated as fixed), their landmarks and then the frame's pose / speed-bias are
 * eliminated (eigen pseudo-inverse, eps 1e-8), and the new prior of dimension *dim_out = P - 6 - (marg_sb >= 0 ? 9 : 0)
 * is re-factored into H_prior, b_prior, err_prior, Jt_prior_inv (rows ordered by ascending eigenvalue like Eigen's
 * SelfAdjointEigenSolver; eigenvector signs are not defined, so Jt_prior_inv / err_prior match the reference up to a
 * sign per row).  Output arrays may be NULL.                                                                      */
int vio_marginalize(vio_problem *p, int32_t marg_pose, int32_t marg_sb, int32_t *dim_out, double *H_prior,
                    double *b_prior, double *err_prior, double *Jt_prior_inv);

/* ---- batched solve: many independent small problems (BASELINE config 3: 4096 sliding windows) ----------------------
 * n_workers host threads, each with its own handle + CUDA stream, pull items from a shared queue: pack -> H2D ->
 * Solve(iterations) -> D2H.  Kernels of different problems overlap on the device.  Results are bitwise those of
 * vio_set_graph + vio_set_prior + vio_solve + vio_get_vertices on one handle.                                     */
typedef struct vio_batch_item {
    const vio_graph *graph;
    int32_t prior_dim, err_dim;                  /* 0 = no prior                                               */
    const double *H_prior, *b_prior, *err_prior, *Jt_prior_inv;
    double *pose_out, *speedbias_out, *inv_depth_out; /* sized like the graph's arrays; may be NULL            */
    vio_stats *stats;                            /* may be NULL                                                */
    int32_t rc;                                  /* out: VIO_OK or the error of this item                      */
    int32_t reserved;
} vio_batch_item;
int vio_solve_batched(int device, int32_t n_workers, vio_batch_item *items, int64_t n_items, int32_t iterations,
                      const vio_lm_opts *opts);
/* Lock-step variant for windows that share their pose-class structure (same number / order / fixed flags of poses and
 * speed-biases, same IMU-edge count, same reprojection information / loss, same extrinsics, same prior dimensions - what
 * consecutive Estimator::backendOptimization calls of one rig produce, A17/src/estimator.cpp:885-1030); landmark and edge
 * counts may differ per item.  The batch is packed as ONE graph with `batch` stacked P x P reduced systems: every kernel
 * launch covers all items, one CTA per item factorises its reduced system, and the v17 LM control runs per item between
 * launches (items that converge early stop taking steps).  Results agree with vio_solve_batched to rounding (the
 * reductions use a different, still deterministic, order).  v17 flavour + exact reduced solve only.  Batches larger than
 * max_chunk (<= 0: 1024 items) are processed as a two-slot software pipeline: while the LM loop of one chunk runs on the
 * device, the next chunk is packed and uploaded into a second handle.                                                  */
int vio_solve_batched_lockstep(int device, vio_batch_item *items, int64_t n_items, int32_t iterations,
                               const vio_lm_opts *opts, int32_t max_chunk);
/* The lock-step entry keeps one device handle and its host staging alive between calls (a caller that submits batch
 * after batch pays allocations once); this frees them.                                                                */
int vio_lockstep_release(void);

/* ---- IMU pre-integration -------------------------------------------------------------------------
 * IntegrationBase(acc_0, gyr_0, ba, bg) followed by push_back(dt, acc, gyr) for every further sample
 * (A17/include/factor/integration_base.h:13-158: midpoint integration, jacobian = F jacobian,
 * covariance = F covariance F^T + V noise V^T), for a batch of segments at once - what processIMU accumulates
 * between two keyframes (A17/src/estimator.cpp:75-110), and what repropagate() redoes when the bias estimate moved.
 * Segment k owns samples [seg_ptr[k], seg_ptr[k+1]); its first sample is (acc_0, gyr_0) and that sample's dt is not
 * read.  Outputs use the layout of the EdgeImu constants in vio_graph (imu_sum_dt ... imu_covariance).           */
typedef struct vio_imu_segments {
    int32_t n_segments;
    int32_t reserved;
    const int32_t *seg_ptr;  /* n_segments + 1                                                      */
    const double *dt;        /* per sample                                                          */
    const double *acc;       /* per sample x 3                                                      */
    const double *gyr;       /* per sample x 3                                                      */
    const double *ba;        /* per segment x 3: linearized_ba                                      */
    const double *bg;        /* per segment x 3: linearized_bg                                      */
    double acc_n, acc_w, gyr_n, gyr_w; /* ACC_N, ACC_W, GYR_N, GYR_W (A17/include/parameters.h)       */
} vio_imu_segments;
int vio_preintegrate(int device, const vio_imu_segments *in, double *sum_dt, double *delta_p, double *delta_q,
                     double *delta_v, double *jacobian, double *covariance);

/* ---- GENERIC_PROBLEM lane: user-defined host edges --------------------------------------------------------------
 * Problem(GENERIC_PROBLEM) lets callers subclass Vertex/Edge with their own virtual ComputeResidual /
 * ComputeJacobians / Plus (A15/app/CurveFitting.cpp:14-48, A17/test/CurveFitting.cpp:8-45).  Those virtuals are host
 * code by construction, so this lane takes the evaluated factors (stacked dense Jacobian, residuals, per-edge
 * RobustInfo) and does the rest on the device: H = J^T W J, b = -J^T Wb r (MakeHessian, A17/src/backend/problem.cc:
 * 319-358), (H + lambda I) dx = b (generic branch of SolveLinearSystem, :397-404), chi2, scale.                     */
typedef struct vio_dense_system {
    int32_t n;                 /* total local dimension                                                   */
    int32_t rows;              /* stacked residual rows R = sum of the edges' residual dimensions          */
    int32_t dmax;              /* largest residual dimension of an edge                                    */
    int32_t reserved;
    const double *J;           /* R x n row-major, zero where an edge does not touch a vertex              */
    const double *r;           /* R                                                                        */
    const int32_t *row_edge0;  /* R: first stacked row of the edge a row belongs to                        */
    const int32_t *row_dim;    /* R: residual dimension of that edge                                       */
    const double *W;           /* R x dmax: row i = row (i - row_edge0[i]) of the edge's RobustInfo        */
    const double *Wb;          /* R x dmax: same rows of drho * Information (used for b)                   */
} vio_dense_system;
int vio_dense_accumulate(vio_problem *p, const vio_dense_system *s, double *max_abs_diag);
/* chi2 = sum over edges of loss(r^T Information r); info rows like W above, loss per row's edge (VIO_LOSS_*) */
int vio_dense_chi2(vio_problem *p, int32_t rows, int32_t dmax, const double *r, const int32_t *row_edge0,
                   const int32_t *row_dim, const double *info, const int32_t *loss_kind, const double *loss_delta,
                   double *chi2);
/* (H + lambda I) dx = b by Cholesky; also returns dx^T (lambda dx + b) and |dx|^2 for IsGoodStepInLM       */
int vio_dense_solve(vio_problem *p, double lambda, double *dx, double *scale_dot, double *dx_norm2);
int vio_dense_get(vio_problem *p, double *H, double *b);

/* FP64 FMA micro-benchmark (roofline denominator for the FP64-bound kernels): returns TFLOP/s */
int vio_measure_fp64_peak(int device, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* VIO_B200_H */
