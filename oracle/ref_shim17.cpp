// ref_shim17.cpp — builds the reference v17 (vins-mono) backend::Problem from a flat vio_graph.
// TEST INFRASTRUCTURE ONLY (see ref_shim.h).  Compiled by oracle/Makefile against the reference
// sources in /root/reference/workspace/assignments/17-vins-initialization/vins-mono.
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

#define private public
#define protected public
#include "backend/problem.h"
#include "backend/vertex_pose.h"
#include "backend/vertex_speedbias.h"
#include "backend/vertex_inverse_depth.h"
#include "backend/vertex_point_xyz.h"
#include "backend/edge_reprojection.h"
#include "backend/edge_prior.h"
#include "backend/edge_imu.h"
#include "backend/loss_function.h"
#undef private
#undef protected

#define REF_FN(x) ref17_##x
#include "ref_shim.h"

// globals declared in the reference's parameters.h; their definitions live in parameters.cpp,
// which needs OpenCV.  Values are overwritten per call (G) / per preintegration (noise).
double ACC_N = 0.2687, ACC_W = 7.07e-6;
double GYR_N = 0.2121, GYR_W = 7.07e-7;
Eigen::Vector3d G(0.0, 0.0, 9.81);

using namespace myslam::backend;

namespace {
struct Built {
    std::unique_ptr<Problem> problem;
    std::vector<std::shared_ptr<VertexPose>> poses;
    std::vector<std::shared_ptr<VertexSpeedBias>> sbs;
    std::vector<std::shared_ptr<VertexInverseDepth>> landmarks;
    std::vector<std::shared_ptr<VertexPointXYZ>> points;
    std::vector<std::unique_ptr<IntegrationBase>> preint;
    std::unique_ptr<LossFunction> loss;
    ~Built() { problem.reset(); }
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

bool build(const vio_graph *g, const ref_prior *prior, Built &B) {
    if (g->n_reproj > 0 && g->ext_pose < 0) return false;  // v17 EdgeReprojection is 4-vertex
    G = Eigen::Vector3d(g->gravity[0], g->gravity[1], g->gravity[2]);
    B.problem.reset(new Problem(Problem::ProblemType::SLAM_PROBLEM));
    B.poses.resize(g->n_pose);
    B.sbs.resize(g->n_speedbias);
    int npc = g->n_pose + g->n_speedbias;
    for (int k = 0; k < npc; ++k) {
        int ent = g->pclass_order ? g->pclass_order[k] : (k < g->n_pose ? k : ~(k - g->n_pose));
        if (ent >= 0) {
            std::shared_ptr<VertexPose> v(new VertexPose());
            Eigen::VectorXd x(7);
            for (int c = 0; c < 7; ++c) x[c] = g->pose[7 * ent + c];
            v->SetParameters(x);
            if (g->pose_fixed && g->pose_fixed[ent]) v->SetFixed();
            B.problem->AddVertex(v);
            B.poses[ent] = v;
        } else {
            int i = ~ent;
            std::shared_ptr<VertexSpeedBias> v(new VertexSpeedBias());
            Eigen::VectorXd x(9);
            for (int c = 0; c < 9; ++c) x[c] = g->speedbias[9 * i + c];
            v->SetParameters(x);
            if (g->speedbias_fixed && g->speedbias_fixed[i]) v->SetFixed();
            B.problem->AddVertex(v);
            B.sbs[i] = v;
        }
    }
    for (int i = 0; i < g->n_imu; ++i) {
        Eigen::Vector3d z = Eigen::Vector3d::Zero();
        Eigen::Vector3d ba(g->imu_lin_ba[3 * i], g->imu_lin_ba[3 * i + 1], g->imu_lin_ba[3 * i + 2]);
        Eigen::Vector3d bg(g->imu_lin_bg[3 * i], g->imu_lin_bg[3 * i + 1], g->imu_lin_bg[3 * i + 2]);
        std::unique_ptr<IntegrationBase> pi(new IntegrationBase(z, z, ba, bg));
        pi->sum_dt = g->imu_sum_dt[i];
        pi->delta_p = Eigen::Vector3d(g->imu_delta_p[3 * i], g->imu_delta_p[3 * i + 1], g->imu_delta_p[3 * i + 2]);
        pi->delta_q = Eigen::Quaterniond(g->imu_delta_q[4 * i + 3], g->imu_delta_q[4 * i], g->imu_delta_q[4 * i + 1],
                                         g->imu_delta_q[4 * i + 2]);
        pi->delta_v = Eigen::Vector3d(g->imu_delta_v[3 * i], g->imu_delta_v[3 * i + 1], g->imu_delta_v[3 * i + 2]);
        for (int r = 0; r < 15; ++r)
            for (int c = 0; c < 15; ++c) {
                pi->jacobian(r, c) = g->imu_jacobian[225 * i + 15 * r + c];
                pi->covariance(r, c) = g->imu_covariance[225 * i + 15 * r + c];
            }
        std::shared_ptr<EdgeImu> e(new EdgeImu(pi.get()));
        std::vector<std::shared_ptr<Vertex>> vs{B.poses[g->imu_pose_i[i]], B.sbs[g->imu_sb_i[i]],
                                                B.poses[g->imu_pose_j[i]], B.sbs[g->imu_sb_j[i]]};
        e->SetVertex(vs);
        B.problem->AddEdge(e);
        B.preint.push_back(std::move(pi));
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
    switch (g->rp_loss) {
        case VIO_LOSS_HUBER: B.loss.reset(new HuberLoss(g->rp_loss_delta)); break;
        case VIO_LOSS_CAUCHY: B.loss.reset(new CauchyLoss(g->rp_loss_delta)); break;
        case VIO_LOSS_TUKEY: B.loss.reset(new TukeyLoss(g->rp_loss_delta)); break;
        default: break;
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
    for (int64_t i = 0; i < g->n_reproj; ++i) {
        Vec3 pi(g->rp_pts_i[3 * i], g->rp_pts_i[3 * i + 1], g->rp_pts_i[3 * i + 2]);
        Vec3 pj(g->rp_pts_j[2 * i], g->rp_pts_j[2 * i + 1], 1.0);
        std::shared_ptr<EdgeReprojection> e(new EdgeReprojection(pi, pj));
        std::vector<std::shared_ptr<Vertex>> vs{B.landmarks[g->rp_landmark[i]], B.poses[g->rp_pose_i[i]],
                                                B.poses[g->rp_pose_j[i]], B.poses[g->ext_pose]};
        e->SetVertex(vs);
        MatXX info = MatXX::Identity(2, 2) * g->rp_info;
        e->SetInformation(info);
        if (B.loss) e->SetLossFunction(B.loss.get());
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
    if (g->n_reproj_xyz > 0) {
        const double *ex = g->ext_pose >= 0 ? g->pose + 7 * g->ext_pose : nullptr;
        Eigen::Quaterniond qic = ex ? Eigen::Quaterniond(ex[6], ex[3], ex[4], ex[5])
                                    : Eigen::Quaterniond(g->q_ic[3], g->q_ic[0], g->q_ic[1], g->q_ic[2]);
        Vec3 tic = ex ? Vec3(ex[0], ex[1], ex[2]) : Vec3(g->t_ic[0], g->t_ic[1], g->t_ic[2]);
        for (int64_t i = 0; i < g->n_reproj_xyz; ++i) {
            Vec3 obs(g->rx_obs[2 * i], g->rx_obs[2 * i + 1], 1.0);
            std::shared_ptr<EdgeReprojectionXYZ> e(new EdgeReprojectionXYZ(obs));
            e->SetTranslationImuFromCamera(qic, tic);
            std::vector<std::shared_ptr<Vertex>> vs{B.points[g->rx_point[i]], B.poses[g->rx_pose[i]]};
            e->SetVertex(vs);
            MatXX info = MatXX::Identity(2, 2) * g->rp_info;
            e->SetInformation(info);
            if (B.loss) e->SetLossFunction(B.loss.get());
            B.problem->AddEdge(e);
        }
    }
    if (prior && prior->dim > 0) {
        MatXX H(prior->dim, prior->dim);
        VecX b(prior->dim);
        for (int r = 0; r < prior->dim; ++r) {
            b[r] = prior->b[r];
            for (int c = 0; c < prior->dim; ++c) H(r, c) = prior->H[(size_t)r * prior->dim + c];
        }
        B.problem->SetHessianPrior(H);
        B.problem->SetbPrior(b);
        if (prior->err_dim > 0) {
            VecX e(prior->err_dim);
            MatXX J(prior->err_dim, prior->err_dim);
            for (int r = 0; r < prior->err_dim; ++r) {
                e[r] = prior->err[r];
                for (int c = 0; c < prior->err_dim; ++c) J(r, c) = prior->jt_inv[(size_t)r * prior->err_dim + c];
            }
            B.problem->SetErrPrior(e);
            B.problem->SetJtPrior(J);
        }
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

int ref17_hessian(const vio_graph *g, const ref_prior *prior, double *H, double *b, int32_t *P, int32_t *M) {
    Built B;
    if (!build(g, prior, B)) return VIO_ERR_UNSUPPORTED;
    CoutCapture cap;
    B.problem->SetOrdering();
    B.problem->MakeHessian();
    copy_out(B.problem->Hessian_, H);
    copy_out(B.problem->b_, b);
    if (P) *P = (int32_t)B.problem->ordering_poses_;
    if (M) *M = (int32_t)B.problem->ordering_landmarks_;
    return VIO_OK;
}

int ref17_init(const vio_graph *g, const ref_prior *prior, double *chi2, double *lambda) {
    Built B;
    if (!build(g, prior, B)) return VIO_ERR_UNSUPPORTED;
    CoutCapture cap;
    B.problem->SetOrdering();
    B.problem->MakeHessian();
    B.problem->ComputeLambdaInitLM();
    *chi2 = B.problem->currentChi_;
    *lambda = B.problem->currentLambda_;
    return VIO_OK;
}

int ref17_step(const vio_graph *g, const ref_prior *prior, double lambda, double *S, double *bS, double *dx) {
    Built B;
    if (!build(g, prior, B)) return VIO_ERR_UNSUPPORTED;
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

int ref17_solve_points(const vio_graph *g, const ref_prior *prior, int32_t iterations, double *pose, double *speedbias,
                       double *inv_depth, double *point_xyz, double *b_prior_out, double *err_prior_out, ref_result *res);
int ref17_solve(const vio_graph *g, const ref_prior *prior, int32_t iterations, double *pose, double *speedbias,
                double *inv_depth, double *b_prior_out, double *err_prior_out, ref_result *res) {
    return ref17_solve_points(g, prior, iterations, pose, speedbias, inv_depth, nullptr, b_prior_out, err_prior_out, res);
}
int ref17_solve_points(const vio_graph *g, const ref_prior *prior, int32_t iterations, double *pose, double *speedbias,
                       double *inv_depth, double *point_xyz, double *b_prior_out, double *err_prior_out, ref_result *res) {
    Built B;
    if (!build(g, prior, B)) return VIO_ERR_UNSUPPORTED;
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
    for (int i = 0; i < g->n_speedbias; ++i)
        for (int k = 0; k < 9; ++k) speedbias[9 * i + k] = B.sbs[i]->Parameters()[k];
    for (int i = 0; i < g->n_landmark; ++i) inv_depth[i] = B.landmarks[i]->Parameters()[0];
    if (point_xyz)
        for (int i = 0; i < g->n_point; ++i)
            for (int k = 0; k < 3; ++k) point_xyz[3 * i + k] = B.points[i]->Parameters()[k];
    if (b_prior_out) copy_out(B.problem->b_prior_, b_prior_out);
    if (err_prior_out) copy_out(B.problem->err_prior_, err_prior_out);
    return VIO_OK;
}

// ---- helpers used only to BUILD config-2 fixtures (not on the hot path) -----------------------

// IntegrationBase::push_back over n samples (A17/include/factor/integration_base.h:30-158): produces the
// EdgeImu constants.  acc/gyr are n x 3, sample 0 is (acc_0, gyr_0).
int ref17_preintegrate(int32_t n, const double *dt, const double *acc, const double *gyr, const double *ba,
                       const double *bg, const double *noise /* ACC_N ACC_W GYR_N GYR_W */, double *sum_dt,
                       double *delta_p, double *delta_q_xyzw, double *delta_v, double *jac225, double *cov225) {
    ACC_N = noise[0];
    ACC_W = noise[1];
    GYR_N = noise[2];
    GYR_W = noise[3];
    Eigen::Vector3d a0(acc[0], acc[1], acc[2]), g0(gyr[0], gyr[1], gyr[2]);
    IntegrationBase ib(a0, g0, Eigen::Vector3d(ba[0], ba[1], ba[2]), Eigen::Vector3d(bg[0], bg[1], bg[2]));
    for (int i = 1; i < n; ++i)
        ib.push_back(dt[i], Eigen::Vector3d(acc[3 * i], acc[3 * i + 1], acc[3 * i + 2]),
                     Eigen::Vector3d(gyr[3 * i], gyr[3 * i + 1], gyr[3 * i + 2]));
    *sum_dt = ib.sum_dt;
    for (int k = 0; k < 3; ++k) {
        delta_p[k] = ib.delta_p[k];
        delta_v[k] = ib.delta_v[k];
    }
    delta_q_xyzw[0] = ib.delta_q.x();
    delta_q_xyzw[1] = ib.delta_q.y();
    delta_q_xyzw[2] = ib.delta_q.z();
    delta_q_xyzw[3] = ib.delta_q.w();
    for (int r = 0; r < 15; ++r)
        for (int c = 0; c < 15; ++c) {
            jac225[15 * r + c] = ib.jacobian(r, c);
            cov225[15 * r + c] = ib.covariance(r, c);
        }
    return VIO_OK;
}

// Problem::Marginalize(margVertexs = {pose[marg_pose], speedbias[marg_sb]}, pose_dim)
// (A17/src/backend/problem.cc:617-795): produces the prior handed to the next window.
int ref17_marginalize(const vio_graph *g, const ref_prior *prior, int32_t marg_pose, int32_t marg_sb, int32_t pose_dim,
                      int32_t *out_dim, double *H_out, double *b_out, double *err_out, double *jt_inv_out) {
    Built B;
    if (!build(g, prior, B)) return VIO_ERR_UNSUPPORTED;
    CoutCapture cap;
    std::vector<std::shared_ptr<Vertex>> marg{B.poses[marg_pose], B.sbs[marg_sb]};
    B.problem->Marginalize(marg, pose_dim);
    *out_dim = (int32_t)B.problem->H_prior_.rows();
    copy_out(B.problem->H_prior_, H_out);
    copy_out(B.problem->b_prior_, b_out);
    copy_out(B.problem->err_prior_, err_out);
    copy_out(B.problem->Jt_prior_inv_, jt_inv_out);
    return VIO_OK;
}
}
