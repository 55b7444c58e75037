// ref_sparse17.cpp — SPARSE restatement of backend::Problem::Solve around the reference's OWN factor code.
// TEST INFRASTRUCTURE ONLY (see ref_shim.h): the at-scale CPU baseline / parity target of SURVEY.md 8(d).
//
// The unmodified v17 backend keeps Hessian_ as a dense (P+M)^2 matrix (A17/src/backend/problem.cc:306-307) and forms
// Hpm * Hmm^-1 * Hmp as a dense product (:412-419), which caps it at a few thousand landmarks.  Here every per-edge
// quantity still comes from the reference's compiled classes -
//     EdgeReprojection::ComputeResidual / ComputeJacobians   (A17/src/backend/edge_reprojection.cc:18-108)
//     EdgeSE3Prior::ComputeResidual / ComputeJacobians       (A17/src/backend/edge_prior.cpp)
//     Edge::RobustInfo / RobustChi2 / Chi2                   (A17/src/backend/edge.cc:32-74)
//     Vertex::Plus / BackUpParameters / RollBackParameters   (A17/src/backend/vertex.cc, vertex_pose.cc)
// and only the CONTAINERS are replaced: 6x6 block rows for H_pp, one scalar + <= K 6-vectors per landmark for H_mm / H_pm,
// a block-sparse Schur complement, Eigen::SimplicialLDLT instead of the dense ldlt (:439).  The LM control is the
// reference's, statement for statement (Solve :169-250, ComputeLambdaInitLM :497-522, IsGoodStepInLM :541-573,
// SolveLinearSystem :394-449 with lambda on the pose block only, UpdateStates / RollbackStates :452-494).
// Edges are built per landmark on the fly (the reference's Edge objects cost ~1 KB each; 10^7 of them do not fit).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <unordered_map>
#include <vector>
#include <Eigen/Dense>
#include <Eigen/Sparse>
#include <Eigen/SparseCholesky>

#include "backend/vertex_pose.h"
#include "backend/vertex_inverse_depth.h"
#include "backend/edge_reprojection.h"
#include "backend/edge_prior.h"
#include "backend/loss_function.h"

#define REF_FN(x) ref17_##x
#include "ref_shim.h"

using namespace myslam::backend;

namespace {
typedef Eigen::Matrix<double, 6, 6> Mat6;
typedef Eigen::Matrix<double, 6, 1> Vec6d;
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Sparse17 {
    const vio_graph *g;
    int C, L;
    std::vector<std::shared_ptr<VertexPose>> poses;
    std::vector<std::shared_ptr<VertexInverseDepth>> lms;
    std::vector<std::shared_ptr<EdgeSE3Prior>> priors;
    std::unique_ptr<LossFunction> loss;
    std::vector<int> eptr, eidx;  // edges by landmark
    // linear system
    std::vector<std::unordered_map<int, Mat6>> Hpp;  // block row a: column b >= a
    std::vector<Vec6d> bp;
    std::vector<double> Hll, bl;
    std::vector<int> lp_ptr;       // per landmark: its pose blocks
    std::vector<int> lp_pose;
    std::vector<Vec6d> lp_H;       // H_pl block (6 x 1) per (landmark, pose)
    Eigen::VectorXd dxp;
    std::vector<double> dxl;
    double chi = 0, lambda = 0, ni = 2;
    double t_lin = 0, t_solve = 0, t_chi = 0;

    std::shared_ptr<EdgeReprojection> make_edge(int e) const {
        Vec3 pi(g->rp_pts_i[3 * (size_t)e], g->rp_pts_i[3 * (size_t)e + 1], g->rp_pts_i[3 * (size_t)e + 2]);
        Vec3 pj(g->rp_pts_j[2 * (size_t)e], g->rp_pts_j[2 * (size_t)e + 1], 1.0);
        std::shared_ptr<EdgeReprojection> ed(new EdgeReprojection(pi, pj));
        std::vector<std::shared_ptr<Vertex>> vs{lms[g->rp_landmark[e]], poses[g->rp_pose_i[e]], poses[g->rp_pose_j[e]], poses[g->ext_pose]};
        ed->SetVertex(vs);
        ed->SetInformation(MatXX::Identity(2, 2) * g->rp_info);
        if (loss) ed->SetLossFunction(loss.get());
        return ed;
    }
    // Problem::MakeHessian (:303-389) + the Schur part of SolveLinearSystem (:406-431), block sparse
    void linearize() {
        const double t0 = now_s();
        for (auto &r : Hpp) r.clear();
        for (auto &b : bp) b.setZero();
        lp_ptr.assign(L + 1, 0); lp_pose.clear(); lp_H.clear();
        auto add_pp = [&](int a, int b, const Mat6 &h) {  // block (a, b) with a <= b kept; the mirror is implied
            if (a <= b) { auto it = Hpp[a].find(b); if (it == Hpp[a].end()) Hpp[a][b] = h; else it->second += h; }
            else { auto it = Hpp[b].find(a); if (it == Hpp[b].end()) Hpp[b][a] = h.transpose(); else it->second += h.transpose(); }
        };
        for (int l = 0; l < L; ++l) {
            double hll = 0.0, b_l = 0.0;
            const int p0 = (int)lp_pose.size();
            auto slot = [&](int pose) -> int {
                for (int k = p0; k < (int)lp_pose.size(); ++k) if (lp_pose[k] == pose) return k;
                lp_pose.push_back(pose); lp_H.push_back(Vec6d::Zero());
                return (int)lp_pose.size() - 1;
            };
            for (int q = eptr[l]; q < eptr[l + 1]; ++q) {
                const int e = eidx[q];
                auto ed = make_edge(e);
                ed->ComputeResidual();
                ed->ComputeJacobians();
                auto jac = ed->Jacobians();
                auto verts = ed->Verticies();
                double drho;
                MatXX robustInfo(ed->Information().rows(), ed->Information().cols());
                ed->RobustInfo(drho, robustInfo);
                const int vp[4] = {-1, g->rp_pose_i[e], g->rp_pose_j[e], g->ext_pose};
                for (size_t i = 0; i < verts.size(); ++i) {
                    if (verts[i]->IsFixed()) continue;
                    const MatXX JtW = jac[i].transpose() * robustInfo;
                    for (size_t j = i; j < verts.size(); ++j) {
                        if (verts[j]->IsFixed()) continue;
                        const MatXX h = JtW * jac[j];
                        if (i == 0 && j == 0) hll += h(0, 0);
                        else if (i == 0) lp_H[slot(vp[j])] += h.transpose();            // H_pl (6 x 1) = H_lp^T
                        else add_pp(vp[i], vp[j], h);
                    }
                    const VecX bi = -drho * jac[i].transpose() * ed->Information() * ed->Residual();
                    if (i == 0) b_l += bi[0];
                    else bp[vp[i]] += bi;
                }
            }
            Hll[l] = hll; bl[l] = b_l;
            lp_ptr[l + 1] = (int)lp_pose.size();
        }
        for (size_t k = 0; k < priors.size(); ++k) {
            auto &ed = priors[k];
            const int a = g->sp_pose[k];
            if (poses[a]->IsFixed()) continue;
            ed->ComputeResidual();
            ed->ComputeJacobians();
            double drho;
            MatXX robustInfo(6, 6);
            ed->RobustInfo(drho, robustInfo);
            const MatXX J = ed->Jacobians()[0];
            add_pp(a, a, J.transpose() * robustInfo * J);
            bp[a] += -drho * J.transpose() * ed->Information() * ed->Residual();
        }
        t_lin += now_s() - t0;
    }
    double chi2() {
        const double t0 = now_s();
        double c = 0.0;
        for (int l = 0; l < L; ++l)
            for (int q = eptr[l]; q < eptr[l + 1]; ++q) {
                auto ed = make_edge(eidx[q]);
                ed->ComputeResidual();
                c += ed->RobustChi2();
            }
        for (auto &ed : priors) { ed->ComputeResidual(); c += ed->RobustChi2(); }
        t_chi += now_s() - t0;
        return 0.5 * c;
    }
    // SolveLinearSystem: Schur complement over the landmarks, lambda on the pose block, LDL^T, back-substitution
    bool solve_step() {
        const double t0 = now_s();
        const int P = 6 * C;
        std::vector<std::unordered_map<int, Mat6>> S = Hpp;
        std::vector<Vec6d> bS = bp;
        for (int l = 0; l < L; ++l) {
            if (!(Hll[l] != 0.0)) continue;
            const double inv = 1.0 / Hll[l];
            for (int a = lp_ptr[l]; a < lp_ptr[l + 1]; ++a) {
                bS[lp_pose[a]] -= lp_H[a] * (inv * bl[l]);
                for (int b = lp_ptr[l]; b < lp_ptr[l + 1]; ++b) {
                    if (lp_pose[a] > lp_pose[b]) continue;
                    const Mat6 h = lp_H[a] * inv * lp_H[b].transpose();
                    auto it = S[lp_pose[a]].find(lp_pose[b]);
                    if (it == S[lp_pose[a]].end()) S[lp_pose[a]][lp_pose[b]] = -h; else it->second -= h;
                }
            }
        }
        std::vector<Eigen::Triplet<double>> trip;
        for (int a = 0; a < C; ++a) {
            bool have_diag = false;
            for (auto &kv : S[a]) {
                const int b = kv.first;
                have_diag |= b == a;
                for (int r = 0; r < 6; ++r)
                    for (int c = 0; c < 6; ++c) {
                        if (b == a && c > r) continue;  // lower triangle of the diagonal block
                        // SimplicialLDLT reads the lower triangle: block (a, b), a < b, is stored transposed at (b, a)
                        const double v = kv.second(r, c) + ((b == a && r == c) ? lambda : 0.0);
                        if (b == a) trip.emplace_back(6 * a + r, 6 * a + c, v);
                        else trip.emplace_back(6 * b + c, 6 * a + r, v);
                    }
            }
            if (!have_diag)
                for (int r = 0; r < 6; ++r) trip.emplace_back(6 * a + r, 6 * a + r, lambda);  // fixed vertex: zero rows + lambda
        }
        Eigen::SparseMatrix<double> A(P, P);
        A.setFromTriplets(trip.begin(), trip.end());
        Eigen::VectorXd rhs(P);
        for (int a = 0; a < C; ++a) rhs.segment<6>(6 * a) = bS[a];
        Eigen::SimplicialLDLT<Eigen::SparseMatrix<double>> ldlt(A);
        if (ldlt.info() != Eigen::Success) return false;
        dxp = ldlt.solve(rhs);
        for (int l = 0; l < L; ++l) {
            double t = bl[l];
            for (int a = lp_ptr[l]; a < lp_ptr[l + 1]; ++a) t -= lp_H[a].dot(dxp.segment<6>(6 * lp_pose[a]));
            dxl[l] = Hll[l] != 0.0 ? t / Hll[l] : 0.0;
        }
        t_solve += now_s() - t0;
        return true;
    }
    void update() {
        for (int a = 0; a < C; ++a) {
            if (poses[a]->IsFixed()) continue;
            poses[a]->BackUpParameters();
            poses[a]->Plus(dxp.segment<6>(6 * a));
        }
        for (int l = 0; l < L; ++l) {
            lms[l]->BackUpParameters();
            VecX d(1);
            d[0] = dxl[l];
            lms[l]->Plus(d);
        }
    }
    void rollback() {
        for (int a = 0; a < C; ++a) if (!poses[a]->IsFixed()) poses[a]->RollBackParameters();
        for (int l = 0; l < L; ++l) lms[l]->RollBackParameters();
    }
    double scale() const {  // 0.5 * dx^T (lambda dx + b) + 1e-6 over [poses | landmarks]  (IsGoodStepInLM :543-546)
        double s = 0.0;
        for (int a = 0; a < C; ++a) s += dxp.segment<6>(6 * a).dot(lambda * dxp.segment<6>(6 * a) + bp[a]);
        for (int l = 0; l < L; ++l) s += dxl[l] * (lambda * dxl[l] + bl[l]);
        return 0.5 * s + 1e-6;
    }
};
}  // namespace

extern "C" int ref17_sparse_solve(const vio_graph *g, int32_t iterations, int32_t fixed_iterations, double *pose_out, double *inv_depth_out,
                                  ref_result *res, double *timing /* [4] linearise, reduced solve + back-substitution, chi2, total (s) */) {
    if (g->n_speedbias != 0 || g->n_imu != 0 || g->n_point != 0 || g->ext_pose < 0) return VIO_ERR_UNSUPPORTED;
    const double t_begin = now_s();
    Sparse17 S;
    S.g = g; S.C = g->n_pose; S.L = g->n_landmark;
    S.poses.resize(S.C);
    for (int a = 0; a < S.C; ++a) {
        S.poses[a].reset(new VertexPose());
        VecX x(7);
        for (int c = 0; c < 7; ++c) x[c] = g->pose[7 * (size_t)a + c];
        S.poses[a]->SetParameters(x);
        if (g->pose_fixed && g->pose_fixed[a]) S.poses[a]->SetFixed();
        S.poses[a]->SetOrderingId(6 * a);
    }
    S.lms.resize(S.L);
    for (int l = 0; l < S.L; ++l) {
        S.lms[l].reset(new VertexInverseDepth());
        VecX x(1);
        x[0] = g->inv_depth[l];
        S.lms[l]->SetParameters(x);
    }
    switch (g->rp_loss) {
        case VIO_LOSS_HUBER: S.loss.reset(new HuberLoss(g->rp_loss_delta)); break;
        case VIO_LOSS_CAUCHY: S.loss.reset(new CauchyLoss(g->rp_loss_delta)); break;
        case VIO_LOSS_TUKEY: S.loss.reset(new TukeyLoss(g->rp_loss_delta)); break;
        default: break;
    }
    for (int k = 0; k < g->n_se3prior; ++k) {
        Vec3 p(g->sp_p[3 * k], g->sp_p[3 * k + 1], g->sp_p[3 * k + 2]);
        Qd q(g->sp_q[4 * k + 3], g->sp_q[4 * k], g->sp_q[4 * k + 1], g->sp_q[4 * k + 2]);
        std::shared_ptr<EdgeSE3Prior> e(new EdgeSE3Prior(p, q));
        std::vector<std::shared_ptr<Vertex>> vs{S.poses[g->sp_pose[k]]};
        e->SetVertex(vs);
        MatXX info(6, 6);
        for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 6; ++c) info(r, c) = g->sp_info[36 * k + 6 * r + c];
        e->SetInformation(info);
        S.priors.push_back(e);
    }
    S.eptr.assign(S.L + 1, 0);
    for (int64_t e = 0; e < g->n_reproj; ++e) S.eptr[g->rp_landmark[e] + 1]++;
    for (int l = 0; l < S.L; ++l) S.eptr[l + 1] += S.eptr[l];
    S.eidx.resize(g->n_reproj);
    {
        std::vector<int> cur(S.eptr.begin(), S.eptr.end() - 1);
        for (int64_t e = 0; e < g->n_reproj; ++e) S.eidx[cur[g->rp_landmark[e]]++] = (int)e;
    }
    S.Hpp.resize(S.C); S.bp.assign(S.C, Vec6d::Zero()); S.Hll.assign(S.L, 0.0); S.bl.assign(S.L, 0.0); S.dxl.assign(S.L, 0.0);
    S.dxp = Eigen::VectorXd::Zero(6 * S.C);
    // ---- Solve (:169-250)
    S.linearize();
    S.chi = S.chi2();
    {   // ComputeLambdaInitLM (:497-522): max |diag| over poses and landmarks, clamped, times 1e-5
        double mx = 0.0;
        for (int a = 0; a < S.C; ++a) {
            auto it = S.Hpp[a].find(a);
            if (it != S.Hpp[a].end()) for (int r = 0; r < 6; ++r) mx = std::max(mx, std::fabs(it->second(r, r)));
        }
        for (int l = 0; l < S.L; ++l) mx = std::max(mx, std::fabs(S.Hll[l]));
        mx = std::min(5e10, mx);
        S.lambda = 1e-5 * mx;
        S.ni = 2.0;
    }
    if (res) std::memset(res, 0, sizeof(*res));
    bool stop = false;
    int iter = 0;
    double last_chi = 1e20;
    while (!stop && iter < iterations) {
        if (res && iter < VIO_TRACE_MAX) { res->chi2_trace[iter] = S.chi; res->lambda_trace[iter] = S.lambda; }
        bool ok = false;
        int false_cnt = 0;
        while (!ok && false_cnt < 10) {
            if (!S.solve_step()) return VIO_ERR_INVALID;
            S.update();
            const double sc = S.scale();
            const double temp = S.chi2();
            const double rho = (S.chi - temp) / sc;
            if (rho > 0 && std::isfinite(temp)) {
                double alpha = 1.0 - std::pow(2 * rho - 1, 3);
                alpha = std::min(alpha, 2.0 / 3.0);
                S.lambda *= std::max(1.0 / 3.0, alpha);
                S.ni = 2;
                S.chi = temp;
                ok = true;
            } else {
                S.lambda *= S.ni;
                S.ni *= 2;
            }
            if (ok) { S.linearize(); false_cnt = 0; }
            else { false_cnt++; S.rollback(); }
        }
        iter++;
        if (!fixed_iterations && last_chi - S.chi < 1e-5) stop = true;
        last_chi = S.chi;
    }
    if (res) { res->iterations = iter; res->chi2_final = S.chi; res->lambda_final = S.lambda; }
    if (pose_out)
        for (int a = 0; a < S.C; ++a)
            for (int c = 0; c < 7; ++c) pose_out[7 * (size_t)a + c] = S.poses[a]->Parameters()[c];
    if (inv_depth_out)
        for (int l = 0; l < S.L; ++l) inv_depth_out[l] = S.lms[l]->Parameters()[0];
    if (timing) { timing[0] = S.t_lin; timing[1] = S.t_solve; timing[2] = S.t_chi; timing[3] = now_s() - t_begin; }
    return VIO_OK;
}
