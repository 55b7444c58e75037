// vio_dev.h — device-side view of one packed problem (POD, passed to kernels by value).
// HBM layout (all FP64 unless noted), see DESIGN.md §3:
//   pose      [C][7]   AoS  (gathered by index; one 56 B record per pose, as VertexPose::Parameters())
//   poseRT    [C][16]  AoS  R(9) t(3) pad(4): 128 B = one L2 line per gather, rebuilt after every state update
//   landmarks SoA: inv_depth[L], pts_i{x,y,z}[L], host[L] (int32), eptr[L+1] (int32, CSR into the edge arrays)
//   edges     SoA, landmark-sorted: pose_j[E] (int32), pjx[E], pjy[E]
//   outputs   Hll[L], bl[L], wh[L][6] (Hlp host row), wo[E][6] (Hlp observer rows),
//             reduced system "sys" = [S values | bcorr(P) | bp(P) | hdiag(P)] contiguous (one allreduce)
#pragma once
#include <stdint.h>

struct DevView {
    // sizes
    int C, NSB, L, P, NB;  // poses, speed-biases, landmarks, pose-class dim (TOTAL over a lock-step batch), pose-class blocks
    // lock-step batch of structurally identical problems (vio_solve_batched_lockstep): the reduced system is a "tall"
    // dense matrix of `batch` stacked Pper x Pper blocks; pose_off / sb_off are GLOBAL row offsets.  batch = 1, Pper = P otherwise.
    int batch, Pper, Cper, NSBper;
    const int *lm_prob;      // [L] problem of each landmark (batch only)
    const uint8_t *act;      // [batch] per-problem mask for the update / restore kernels, or nullptr
    long long E;           // reprojection edges (local shard)
    int storage;           // 1 dense, 2 bsr
    long long nnzb;
    // state
    double *pose, *pose_bak, *sb, *sb_bak, *invdep, *invdep_bak;
    const uint8_t *pose_fixed, *sb_fixed;
    const int *pose_off, *sb_off, *pose_blk;  // ordering offsets; block position of each pose
    double *poseRT;
    // extrinsics (constant during a solve)
    double Ric[9], tic[3];
    // landmark / edge structure
    const int *lm_host, *lm_eptr;
    const double *lm_pix, *lm_piy, *lm_piz;
    const int *e_pose_j;
    const double *e_pjx, *e_pjy;
    double rp_info;
    int rp_loss;
    double rp_delta;
    // free extrinsic vertex (4-vertex EdgeReprojection with ESTIMATE_EXTRINSIC=1): pose index or -1; its H_lp rows [L][6]
    int ext_pose;
    double *we;
    // linearisation outputs
    double *Hll, *bl, *wh, *wo;
    double *S;         // dense P*P or bsr values
    double *bcorr, *bp, *hdiag, *bS;
    const int *bsr_rowptr, *bsr_col, *bsr_tr;  // tr: id of the transposed block
    // solve outputs
    double *dxp, *dxl;
    // VertexPointXYZ landmarks + EdgeReprojectionXYZ observations (vio_xyz.cuh): CSR by point, caller order
    int Lx;
    long long Ex;
    double *pt, *pt_bak;                 // [Lx][3]
    const int *px_eptr, *ex_pose;        // [Lx+1], [Ex]
    const double *ex_ox, *ex_oy;         // [Ex]
    double *Hxx, *bx, *wx, *dxx;         // [Lx][6] sym 3x3 (xx xy xz yy yz zz), [Lx][3], [Ex][18] H_lp rows 3x6, [Lx][3]
    // fixed landmark-class vertices (nullptr = none): Jacobian blocks skipped, left out of Schur and back-substitution
    const uint8_t *lm_fixed, *pt_fixed;  // [L] (packed order), [Lx]
};
