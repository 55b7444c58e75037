// ref_shim15.cpp — builds the reference v15 backend::Problem from a flat vio_graph.
// TEST INFRASTRUCTURE ONLY (see ref_shim.h).  Compiled by oracle/Makefile against the reference
// sources in /root/reference/workspace/assignments/15-vio-backend (never copied into this repo).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>
#include <Eigen/Dense>

// Access to Hessian_, b_, delta_x_, ... for the parity taps.  Access specifiers do not change
// the object layout, so this TU stays ABI-compatible with the reference TUs.
#define private public
#define protected public
#include "backend/problem.h"
#include "backend/vertex_pose.h"
#include "backend/vertex_inverse_depth.h"
#include "backend/vertex_point_xyz.h"
#include "backend/edge_reprojection.h"
#include "backend/edge_prior.h"
#undef private
#undef protected

#define REF_FN(x) ref15_##x
#include "ref_shim.h"

using namespace myslam::backend;

namespace {
struct Built {
    std::unique_ptr<Problem> problem;
    std::vector<std::shared_ptr<VertexPose>> poses;
    std::vector<std::shared_ptr<VertexInverseDepth>> landmarks;
    std::vector<std::shared_ptr<VertexPointXYZ>> points;
};

struct CoutCapture {
    std::streambuf *old;
    std::streamsize prec;
    std::ostringstream ss;
    CoutCapture() : old(std::cout.rdbuf(ss.rdbuf())), prec(std::cout.precision(17)) {}
    ~CoutCapture() {
        std::cout.rdbuf(old);
        std::cout.precision(prec);
    }
};

bool build(const vio_graph *g, Built &B) {
    if (g->n_speedbias != 0 || g->n_imu != 0) return false;  // v15 has no usable IMU edge
    B.problem.reset(new Problem(Problem::ProblemType::SLAM_PROBLEM));
    for (int i = 0; i < g->n_pose; ++i) {
        std::shared_ptr<VertexPose> v(new VertexPose());
        Eigen::VectorXd x(7);
        for (int k = 0; k < 7; ++k) x[k] = g->pose[7 * i + k];
        v->SetParameters(x);
        if (g->pose_fixed && g->pose_fixed[i]) v->SetFixed();
        B.problem->AddVertex(v);
        B.poses.push_back(v);
    }
    for (int i = 0; i < g->n_se3prior; ++i) {
        Vec3 p(g->sp_p[3 * i], g->sp_p[3 * i + 1], g->sp_p[3 * i + 2]);
        Qd q(g->sp_q[4 * i + 3], g->sp_q[4 * i], g->sp_q[4 * i + 1], g->sp_q[4 * i + 2]);
        std::shared_ptr<EdgeSE3Prior> e(new EdgeSE3Prior(p, q));
        std::vector<std::shared_ptr<Vertex>> vs{B.poses[g->sp_pose[i]]};
        e->SetVertex(vs);
        MatXX info(6, 6);
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6; ++c) info(r, c) = g->sp_info[36 * i + 6 * r + c];
        e->SetInformation(info);
        B.problem->AddEdge(e);
    }
    for (int i = 0; i < g->n_landmark; ++i) {
        std::shared_ptr<VertexInverseDepth> v(new VertexInverseDepth());
        VecX x(1);
        x[0] = g->inv_depth[i];
        v->SetParameters(x);
        if (g->landmark_fixed && g->landmark_fixed[i]) v->SetFixed();
        B.problem->AddVertex(v);
        B.landmarks.push_back(v);
    }
    Eigen::Quaterniond qic(g->q_ic[3], g->q_ic[0], g->q_ic[1], g->q_ic[2]);
    Vec3 tic(g->t_ic[0], g->t_ic[1], g->t_ic[2]);
    for (int64_t i = 0; i < g->n_reproj; ++i) {
        Vec3 pi(g->rp_pts_i[3 * i], g->rp_pts_i[3 * i + 1], g->rp_pts_i[3 * i + 2]);
        Vec3 pj(g->rp_pts_j[2 * i], g->rp_pts_j[2 * i + 1], 1.0);
        std::shared_ptr<EdgeReprojection> e(new EdgeReprojection(pi, pj));
        e->SetTranslationImuFromCamera(qic, tic);
        std::vector<std::shared_ptr<Vertex>> vs{B.landmarks[g->rp_landmark[i]], B.poses[g->rp_pose_i[i]],
                                                B.poses[g->rp_pose_j[i]]};
        e->SetVertex(vs);
        if (g->rp_info != 1.0) {
            MatXX info = MatXX::Identity(2, 2) * g->rp_info;
            e->SetInformation(info);
        }
        B.problem->AddEdge(e);
    }
    // VertexPointXYZ + EdgeReprojectionXYZ, created after the inverse-depth landmarks: Hessian_ = [P | M1 | 3 Mx]
    for (int i = 0; i < g->n_point; ++i) {
        std::shared_ptr<VertexPointXYZ> v(new VertexPointXYZ());
        VecX x(3);
        for (int k = 0; k < 3; ++k) x[k] = g->point_xyz[3 * i + k];
        v->SetParameters(x);
        if (g->point_fixed && g->point_fixed[i]) v->SetFixed();
        B.problem->AddVertex(v);
        B.points.push_back(v);
    }
    for (int64_t i = 0; i < g->n_reproj_xyz; ++i) {
        Vec3 obs(g->rx_obs[2 * i], g->rx_obs[2 * i + 1], 1.0);
        std::shared_ptr<EdgeReprojectionXYZ> e(new EdgeReprojectionXYZ(obs));
        e->SetTranslationImuFromCamera(qic, tic);
        std::vector<std::shared_ptr<Vertex>> vs{B.points[g->rx_point[i]], B.poses[g->rx_pose[i]]};
        e->SetVertex(vs);
        if (g->rp_info != 1.0) {
            MatXX info = MatXX::Identity(2, 2) * g->rp_info;
            e->SetInformation(info);
        }
        B.problem->AddEdge(e);
    }
    return true;
}

void copy_out(const MatXX &A, double *out) {
    if (!out) return;
    for (int r = 0; r < A.rows(); ++r)
        for (int c = 0; c < A.cols(); ++c) out[(size_t)r * A.cols() + c] = A(r, c);
}
void copy_out(const VecX &a, double *out) {
    if (!out) return;
    for (int r = 0; r < a.rows(); ++r) out[r] = a[r];
}
}  // namespace

extern "C" {

int ref15_hessian(const vio_graph *g, const ref_prior *, double *H, double *b, int32_t *P, int32_t *M) {
    Built B;
    if (!build(g, B)) return VIO_ERR_UNSUPPORTED;
    CoutCapture cap;
    B.problem->SetOrdering();
    B.problem->MakeHessian();
    copy_out(B.problem->Hessian_, H);
    copy_out(B.problem->b_, b);
    if (P) *P = (int32_t)B.problem->ordering_poses_;
    if (M) *M = (int32_t)B.problem->ordering_landmarks_;
    return VIO_OK;
}

int ref15_init(const vio_graph *g, const ref_prior *, double *chi2, double *lambda) {
    Built B;
    if (!build(g, B)) return VIO_ERR_UNSUPPORTED;
    CoutCapture cap;
    B.problem->SetOrdering();
    B.problem->MakeHessian();
    B.problem->ComputeLambdaInitLM();
    *chi2 = B.problem->currentChi_;
    *lambda = B.problem->currentLambda_;
    return VIO_OK;
}

int ref15_step(const vio_graph *g, const ref_prior *, double lambda, double *S, double *bS, double *dx) {
    Built B;
    if (!build(g, B)) return VIO_ERR_UNSUPPORTED;
    CoutCapture cap;
    B.problem->SetOrdering();
    B.problem->MakeHessian();
    B.problem->ComputeLambdaInitLM();
    B.problem->currentLambda_ = lambda;
    B.problem->SolveLinearSystem();
    copy_out(B.problem->H_pp_schur_, S);
    copy_out(B.problem->b_pp_schur_, bS);
    copy_out(B.problem->delta_x_, dx);
    return VIO_OK;
}

int ref15_solve_points(const vio_graph *g, const ref_prior *, int32_t iterations, double *pose, double *, double *inv_depth,
                       double *point_xyz, double *, double *, ref_result *res);
int ref15_solve(const vio_graph *g, const ref_prior *pr, int32_t iterations, double *pose, double *sb, double *inv_depth,
                double *bp, double *ep, ref_result *res) {
    return ref15_solve_points(g, pr, iterations, pose, sb, inv_depth, nullptr, bp, ep, res);
}
int ref15_solve_points(const vio_graph *g, const ref_prior *, int32_t iterations, double *pose, double *, double *inv_depth,
                       double *point_xyz, double *, double *, ref_result *res) {
    Built B;
    if (!build(g, B)) return VIO_ERR_UNSUPPORTED;
    std::string log;
    {
        CoutCapture cap;
        B.problem->Solve(iterations);
        log = cap.ss.str();
    }
    if (res) {
        std::memset(res, 0, sizeof(*res));
        std::istringstream in(log);
        std::string line;
        while (std::getline(in, line)) {
            int it;
            double chi, lam;
            if (std::sscanf(line.c_str(), "iter: %d , chi= %lf , Lambda= %lf", &it, &chi, &lam) == 3) {
                if (res->iterations < VIO_TRACE_MAX) {
                    res->chi2_trace[res->iterations] = chi;
                    res->lambda_trace[res->iterations] = lam;
                }
                res->iterations++;
            } else if (std::sscanf(line.c_str(), "problem solve cost: %lf", &chi) == 1) {
                res->ms_solve = chi;
            } else if (std::sscanf(line.c_str(), " makeHessian cost: %lf", &chi) == 1) {
                res->ms_hessian = chi;
            }
        }
        res->chi2_final = B.problem->currentChi_;
        res->lambda_final = B.problem->currentLambda_;
    }
    for (int i = 0; i < g->n_pose; ++i)
        for (int k = 0; k < 7; ++k) pose[7 * i + k] = B.poses[i]->Parameters()[k];
    for (int i = 0; i < g->n_landmark; ++i) inv_depth[i] = B.landmarks[i]->Parameters()[0];
    if (point_xyz)
        for (int i = 0; i < g->n_point; ++i)
            for (int k = 0; k < 3; ++k) point_xyz[3 * i + k] = B.points[i]->Parameters()[k];
    return VIO_OK;
}
}
