// backend_b200.cc — host side of the drop-in `myslam::backend` mirror (include/backend/myslam_backend_b200.h).
// Problem::Solve packs the graph into a flat vio_graph, hands it to libvio_b200.so through the C-ABI and writes
// the optimised parameters back into the caller's vertex objects.  No numerics of the LM iteration run here.
//
// Reference behaviour mirrored (paths under /root/reference/workspace/assignments):
//   AddVertex/AddEdge/Remove*   15-vio-backend/backend/problem.cc:40-54,91-153  (bool returns, duplicate ids)
//   SetOrdering                 15-vio-backend/backend/problem.cc:224-262       (pose class first, id order)
//   Solve                       15-vio-backend/backend/problem.cc:155-222, vins-mono/src/backend/problem.cc:169-250
//   ExtendHessiansPriorSize     17-vins-initialization/vins-mono/src/backend/problem.cc:83-92
//   TestMarginalize             15-vio-backend/backend/problem.cc:571-657 (toy demo printed by TestMonoBA)
#include "backend/myslam_backend_b200.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <iostream>
#include <mutex>
#include <stdexcept>

#include <Eigen/Dense>

#include "vio_b200.h"

namespace myslam {
namespace backend {

unsigned long global_vertex_id = 0;
unsigned long global_edge_id = 0;

// ---- Vertex -----------------------------------------------------------------------------------------------------
Vertex::Vertex(int num_dimension, int local_dimension) {
    parameters_.resize(num_dimension, 1);
    local_dimension_ = local_dimension > 0 ? local_dimension : num_dimension;
    id_ = global_vertex_id++;
}
Vertex::~Vertex() {}
void Vertex::Plus(const VecX &delta) { parameters_ += delta; }

// t += dt ; q <- q * exp(dtheta)  (right multiplication).  Kept for callers that update vertices themselves;
// Problem::Solve performs the same update on the device (k_update_pose).
void VertexPose::Plus(const VecX &delta) {
    VecX &p = Parameters();
    p.head<3>() += delta.head<3>();
    const Vec3 w(delta[3], delta[4], delta[5]);
    const double th = w.norm();
    Qd dq;
    if (th < 1e-10) {
        const double th2 = th * th, th4 = th2 * th2;
        const double im = 0.5 - th2 / 48.0 + th4 / 3840.0;
        dq = Qd(1.0 - 0.5 * th2 + th4 / 384.0, im * w.x(), im * w.y(), im * w.z());
    } else {
        const double im = std::sin(0.5 * th) / th;
        dq = Qd(std::cos(0.5 * th), im * w.x(), im * w.y(), im * w.z());
    }
    dq.normalize();
    Qd q(p[6], p[3], p[4], p[5]);
    q = q * dq;
    p[3] = q.x(); p[4] = q.y(); p[5] = q.z(); p[6] = q.w();
}

// ---- loss functions (host copies for callers that evaluate RobustChi2 on their own edges) --------------------------
void TrivalLoss::Compute(double e2, Eigen::Vector3d &rho) const { rho[0] = e2; rho[1] = 1; rho[2] = 0; }
void HuberLoss::Compute(double e, Eigen::Vector3d &rho) const {
    const double d2 = delta_ * delta_;
    if (e <= d2) { rho[0] = e; rho[1] = 1.; rho[2] = 0.; }
    else { const double s = std::sqrt(e); rho[0] = 2 * s * delta_ - d2; rho[1] = delta_ / s; rho[2] = -0.5 * rho[1] / e; }
}
void CauchyLoss::Compute(double e2, Eigen::Vector3d &rho) const {
    const double d2 = delta_ * delta_, rec = 1. / d2, aux = rec * e2 + 1.0;
    rho[0] = d2 * std::log(aux); rho[1] = 1. / aux; rho[2] = -rec * rho[1] * rho[1];
}
void TukeyLoss::Compute(double e2, Eigen::Vector3d &rho) const {
    const double e = std::sqrt(e2), d2 = delta_ * delta_;
    if (e <= delta_) { const double om = 1. - e2 / d2; rho[0] = d2 * (1. - om * om * om) / 3.; rho[1] = om * om; rho[2] = -2. * om / d2; }
    else { rho[0] = d2 / 3.; rho[1] = 0; rho[2] = 0; }
}

// ---- Edge -------------------------------------------------------------------------------------------------------
Edge::Edge(int residual_dimension, int num_verticies, const std::vector<std::string> &verticies_types) {
    residual_.resize(residual_dimension, 1);
    if (!verticies_types.empty()) verticies_types_ = verticies_types;
    jacobians_.resize(num_verticies);
    id_ = global_edge_id++;
    information_ = MatXX::Identity(residual_dimension, residual_dimension);
    sqrt_information_ = information_;
}
Edge::~Edge() {}
void Edge::SetInformation(const MatXX &information) {
    information_ = information;
    sqrt_information_ = Eigen::LLT<MatXX>(information_).matrixL().transpose();
}
double Edge::Chi2() const { return residual_.transpose() * information_ * residual_; }
double Edge::RobustChi2() const {
    double e2 = Chi2();
    if (lossfunction_) { Eigen::Vector3d rho; lossfunction_->Compute(e2, rho); e2 = rho[0]; }
    return e2;
}
void Edge::RobustInfo(double &drho, MatXX &info) const {
    if (!lossfunction_) { drho = 1.0; info = information_; return; }
    const double e2 = Chi2();
    Eigen::Vector3d rho;
    lossfunction_->Compute(e2, rho);
    const VecX we = sqrt_information_ * residual_;
    MatXX ri = MatXX::Identity(information_.rows(), information_.cols()) * rho[1];
    if (rho[1] + 2 * rho[2] * e2 > 0.) ri += 2 * rho[2] * we * we.transpose();
    info = ri * information_;
    drho = rho[1];
}
bool Edge::CheckValid() {
    if (!verticies_types_.empty())
        for (size_t i = 0; i < verticies_.size(); ++i)
            if (verticies_types_[i] != verticies_[i]->TypeInfo()) {
                std::cout << "Vertex type does not match, should be " << verticies_types_[i] << ", but set to "
                          << verticies_[i]->TypeInfo() << std::endl;
                return false;
            }
    return true;
}

static void device_only(const char *what) {
    throw std::logic_error(std::string(what) +
                           ": built-in edges are evaluated on the GPU inside Problem::Solve (vio_b200 has no host evaluator)");
}
EdgeReprojection::EdgeReprojection(const Vec3 &pts_i, const Vec3 &pts_j)
    : Edge(2, 4, std::vector<std::string>()), pts_i_(pts_i), pts_j_(pts_j) {}
void EdgeReprojection::SetTranslationImuFromCamera(Eigen::Quaterniond &qic_, Vec3 &tic_) { qic = qic_; tic = tic_; }
void EdgeReprojection::ComputeResidual() { device_only("EdgeReprojection::ComputeResidual"); }
void EdgeReprojection::ComputeJacobians() { device_only("EdgeReprojection::ComputeJacobians"); }
EdgeReprojectionXYZ::EdgeReprojectionXYZ(const Vec3 &pts_i)
    : Edge(2, 2, std::vector<std::string>{"VertexXYZ", "VertexPose"}), obs_(pts_i) {}
void EdgeReprojectionXYZ::SetTranslationImuFromCamera(Eigen::Quaterniond &qic_, Vec3 &tic_) { qic = qic_; tic = tic_; }
void EdgeReprojectionXYZ::ComputeResidual() { device_only("EdgeReprojectionXYZ::ComputeResidual"); }
void EdgeReprojectionXYZ::ComputeJacobians() { device_only("EdgeReprojectionXYZ::ComputeJacobians"); }
EdgeSE3Prior::EdgeSE3Prior(const Vec3 &p, const Qd &q) : Edge(6, 1, std::vector<std::string>{"VertexPose"}), Pp_(p), Qp_(q) {}
void EdgeSE3Prior::ComputeResidual() { device_only("EdgeSE3Prior::ComputeResidual"); }
void EdgeSE3Prior::ComputeJacobians() { device_only("EdgeSE3Prior::ComputeJacobians"); }
EdgeImuB200::EdgeImuB200(const ImuPreintegrationB200 &pre)
    : Edge(15, 4, std::vector<std::string>{"VertexPose", "VertexSpeedBias", "VertexPose", "VertexSpeedBias"}), pre_(pre) {}
void EdgeImuB200::ComputeResidual() { device_only("EdgeImu::ComputeResidual"); }
void EdgeImuB200::ComputeJacobians() { device_only("EdgeImu::ComputeJacobians"); }

// ---- device handles outlive Problem objects ---------------------------------------------------------------------------
// The VINS estimator builds a NEW Problem for every frame (A17/src/estimator.cpp:902-1037 problemSolve, :693-829
// MargOldFrame), solves it once and drops it.  A C-ABI handle owns a stream, events, pinned staging and the device
// buffers of the packed graph; creating one costs more than solving a window.  Handles therefore go back to a small
// per-device pool when their Problem dies and the next Problem's vio_set_graph reuses the buffers (same window size:
// no allocation at all).  SURVEY 8(f-2): "reuse previous window's device buffers".
namespace {
struct HandlePool {
    std::mutex mu;
    std::vector<std::pair<int, vio_problem *>> idle;
};
HandlePool &handle_pool() {
    static HandlePool *p = new HandlePool();  // never destroyed: the CUDA context may be gone before static destructors run
    return *p;
}
vio_problem *acquire_handle(int device) {
    {
        HandlePool &hp = handle_pool();
        std::lock_guard<std::mutex> lk(hp.mu);
        for (size_t i = 0; i < hp.idle.size(); ++i)
            if (hp.idle[i].first == device) {
                vio_problem *h = hp.idle[i].second;
                hp.idle.erase(hp.idle.begin() + i);
                return h;
            }
    }
    vio_problem *h = nullptr;
    if (vio_create(device, nullptr, &h) != VIO_OK) return nullptr;
    return h;
}
void release_handle(int device, vio_problem *h) {
    HandlePool &hp = handle_pool();
    {
        std::lock_guard<std::mutex> lk(hp.mu);
        if (hp.idle.size() < 4) { hp.idle.emplace_back(device, h); return; }
    }
    vio_destroy(h);
}
}  // namespace

// ---- Problem ------------------------------------------------------------------------------------------------------
Problem::Problem(ProblemType problemType, bool v17_flavour) : problemType_(problemType), v17_(v17_flavour) {}
Problem::~Problem() {
    if (handle_) release_handle(device_, handle_);
    if (v17_) global_vertex_id = 0;  // the v17 destructor does this (vins-mono/src/backend/problem.cc:38-41)
}

bool Problem::IsPoseVertex(std::shared_ptr<Vertex> v) {
    const std::string t = v->TypeInfo();
    return t == "VertexPose" || (v17_ && t == "VertexSpeedBias");
}
bool Problem::IsLandmarkVertex(std::shared_ptr<Vertex> v) {
    const std::string t = v->TypeInfo();
    return t == "VertexPointXYZ" || t == "VertexInverseDepth";
}

bool Problem::AddVertex(std::shared_ptr<Vertex> vertex) {
    if (verticies_.find(vertex->Id()) != verticies_.end()) return false;
    verticies_.insert(std::make_pair(vertex->Id(), vertex));
    if (problemType_ == ProblemType::SLAM_PROBLEM && IsPoseVertex(vertex)) ExtendHessiansPriorSize(vertex->LocalDimension());
    return true;
}
void Problem::ExtendHessiansPriorSize(int dim) {
    const int old = (int)H_prior_.rows(), size = old + dim;
    H_prior_.conservativeResize(size, size);
    b_prior_.conservativeResize(size);
    b_prior_.tail(dim).setZero();
    H_prior_.rightCols(dim).setZero();
    H_prior_.bottomRows(dim).setZero();
}
bool Problem::AddEdge(std::shared_ptr<Edge> edge) {
    if (edges_.find(edge->Id()) != edges_.end()) return false;
    edges_.insert(std::make_pair(edge->Id(), edge));
    for (auto &v : edge->Verticies()) vertexToEdge_.insert(std::make_pair(v->Id(), edge));
    return true;
}
std::vector<std::shared_ptr<Edge>> Problem::GetConnectedEdges(std::shared_ptr<Vertex> vertex) {
    std::vector<std::shared_ptr<Edge>> out;
    auto range = vertexToEdge_.equal_range(vertex->Id());
    for (auto it = range.first; it != range.second; ++it)
        if (edges_.find(it->second->Id()) != edges_.end()) out.emplace_back(it->second);
    return out;
}
bool Problem::RemoveVertex(std::shared_ptr<Vertex> vertex) {
    if (verticies_.find(vertex->Id()) == verticies_.end()) return false;
    for (auto &e : GetConnectedEdges(vertex)) RemoveEdge(e);
    if (IsPoseVertex(vertex)) idx_pose_vertices_.erase(vertex->Id());
    else idx_landmark_vertices_.erase(vertex->Id());
    vertex->SetOrderingId(-1);
    verticies_.erase(vertex->Id());
    vertexToEdge_.erase(vertex->Id());
    return true;
}
bool Problem::RemoveEdge(std::shared_ptr<Edge> edge) {
    if (edges_.find(edge->Id()) == edges_.end()) return false;
    edges_.erase(edge->Id());
    return true;
}

void Problem::SetOrdering() {
    ordering_poses_ = ordering_generic_ = ordering_landmarks_ = 0;
    idx_pose_vertices_.clear();
    idx_landmark_vertices_.clear();
    for (auto &kv : verticies_) {
        auto &v = kv.second;
        ordering_generic_ += v->LocalDimension();
        if (problemType_ != ProblemType::SLAM_PROBLEM) continue;
        if (IsPoseVertex(v)) {
            v->SetOrderingId(ordering_poses_);
            idx_pose_vertices_.insert(std::make_pair(v->Id(), v));
            ordering_poses_ += v->LocalDimension();
        } else if (IsLandmarkVertex(v)) {
            v->SetOrderingId(ordering_landmarks_);
            ordering_landmarks_ += v->LocalDimension();
            idx_landmark_vertices_.insert(std::make_pair(v->Id(), v));
        }
    }
    for (auto &kv : idx_landmark_vertices_) kv.second->SetOrderingId(kv.second->OrderingId() + ordering_poses_);
}

// everything Problem::Solve / Marginalize hand to the C-ABI: flat arrays + the vertex objects to write back into
struct PackB200 {
    std::vector<double> pose, sb, invd;
    std::vector<uint8_t> pose_fixed, sb_fixed, lm_fixed, pt_fixed;
    std::vector<int32_t> pclass;
    std::unordered_map<unsigned long, int> pose_idx, sb_idx, lm_idx;
    std::vector<std::shared_ptr<Vertex>> pose_v, sb_v, lm_v;
    std::vector<int32_t> rp_lm, rp_i, rp_j, sp_pose, imu_pi, imu_si, imu_pj, imu_sj;
    std::vector<double> rp_pti, rp_ptj, sp_p, sp_q, sp_info, imu_dt, imu_dp, imu_dq, imu_dv, imu_ba, imu_bg, imu_jac, imu_cov;
    // VertexPointXYZ landmarks + EdgeReprojectionXYZ observations
    std::unordered_map<unsigned long, int> pt_idx;
    std::vector<std::shared_ptr<Vertex>> pt_v;
    std::vector<double> pt, rx_obs;
    std::vector<int32_t> rx_point, rx_pose;
    vio_graph g;
};

bool Problem::PackGraphB200(PackB200 &K) {
    auto &pose = K.pose; auto &sb = K.sb; auto &invd = K.invd; auto &pose_fixed = K.pose_fixed; auto &sb_fixed = K.sb_fixed;
    auto &pclass = K.pclass; auto &pose_idx = K.pose_idx; auto &sb_idx = K.sb_idx; auto &lm_idx = K.lm_idx;
    auto &pose_v = K.pose_v; auto &sb_v = K.sb_v; auto &lm_v = K.lm_v; auto &lm_fixed = K.lm_fixed;
    auto &rp_lm = K.rp_lm; auto &rp_i = K.rp_i; auto &rp_j = K.rp_j; auto &sp_pose = K.sp_pose;
    auto &imu_pi = K.imu_pi; auto &imu_si = K.imu_si; auto &imu_pj = K.imu_pj; auto &imu_sj = K.imu_sj;
    auto &rp_pti = K.rp_pti; auto &rp_ptj = K.rp_ptj; auto &sp_p = K.sp_p; auto &sp_q = K.sp_q; auto &sp_info = K.sp_info;
    auto &imu_dt = K.imu_dt; auto &imu_dp = K.imu_dp; auto &imu_dq = K.imu_dq; auto &imu_dv = K.imu_dv;
    auto &imu_ba = K.imu_ba; auto &imu_bg = K.imu_bg; auto &imu_jac = K.imu_jac; auto &imu_cov = K.imu_cov;
    vio_graph &g = K.g;
    SetOrdering();
    // ---- pack: pose-class vertices in id order, landmarks in id order ---------------------------------------------
    for (auto &kv : verticies_) {
        auto &v = kv.second;
        const std::string t = v->TypeInfo();
        const VecX &x = static_cast<Vertex &>(*v).Parameters();
        if (t == "VertexPose") {
            pose_idx[v->Id()] = (int)pose_v.size();
            pclass.push_back((int32_t)pose_v.size());
            pose_v.push_back(v);
            for (int k = 0; k < 7; ++k) pose.push_back(x[k]);
            pose_fixed.push_back(v->IsFixed());
        } else if (t == "VertexSpeedBias" && v17_) {
            sb_idx[v->Id()] = (int)sb_v.size();
            pclass.push_back(~(int32_t)sb_v.size());
            sb_v.push_back(v);
            for (int k = 0; k < 9; ++k) sb.push_back(x[k]);
            sb_fixed.push_back(v->IsFixed());
        } else if (t == "VertexInverseDepth") {
            lm_idx[v->Id()] = (int)lm_v.size();
            lm_v.push_back(v);
            invd.push_back(x[0]);
            lm_fixed.push_back(v->IsFixed());
        } else if (t == "VertexPointXYZ") {
            K.pt_idx[v->Id()] = (int)K.pt_v.size();
            K.pt_v.push_back(v);
            for (int k = 0; k < 3; ++k) K.pt.push_back(x[k]);
            K.pt_fixed.push_back(v->IsFixed());
        } else {
            std::cerr << "vio_b200: vertex type " << t << " is not on the device path" << std::endl;
            return false;
        }
    }
    std::vector<unsigned long> eids;
    eids.reserve(edges_.size());
    for (auto &kv : edges_) eids.push_back(kv.first);
    std::sort(eids.begin(), eids.end());  // deterministic packing order (the reference walks an unordered_map)
    std::memset(&g, 0, sizeof(g));
    g.ext_pose = -1;
    g.q_ic[3] = 1.0;
    g.rp_info = 1.0;
    bool have_rp = false;
    for (unsigned long id : eids) {
        auto &e = edges_[id];
        const std::string t = e->TypeInfo();
        auto vs = e->Verticies();
        if (t == "EdgeReprojection") {
            auto *er = dynamic_cast<EdgeReprojection *>(e.get());
            if (!er || vs.size() < 3) { std::cerr << "vio_b200: foreign EdgeReprojection type" << std::endl; return false; }
            const MatXX info = e->Information();
            const double c = info(0, 0);
            if (info.rows() != 2 || info(1, 1) != c || info(0, 1) != 0.0 || info(1, 0) != 0.0) {
                std::cerr << "vio_b200: reprojection information must be c*I2" << std::endl;
                return false;
            }
            LossFunction *lf = e->GetLossFunction();
            const int kind = lf ? lf->KindB200() : 0;
            const double delta = lf ? lf->DeltaB200() : 1.0;
            if (kind < 0) { std::cerr << "vio_b200: user-defined loss functions are not supported" << std::endl; return false; }
            int ext = -1;
            if (vs.size() >= 4) ext = pose_idx.at(vs[3]->Id());
            if (!have_rp) {
                g.rp_info = c; g.rp_loss = kind; g.rp_loss_delta = delta; g.ext_pose = ext;
                g.q_ic[0] = er->Qic().x(); g.q_ic[1] = er->Qic().y(); g.q_ic[2] = er->Qic().z(); g.q_ic[3] = er->Qic().w();
                g.t_ic[0] = er->Tic().x(); g.t_ic[1] = er->Tic().y(); g.t_ic[2] = er->Tic().z();
                have_rp = true;
            } else if (g.rp_info != c || g.rp_loss != kind || g.rp_loss_delta != delta || g.ext_pose != ext ||
                       (ext < 0 && (g.q_ic[3] != er->Qic().w() || g.q_ic[0] != er->Qic().x() || g.t_ic[0] != er->Tic().x()))) {
                std::cerr << "vio_b200: reprojection edges must share information, loss and extrinsics" << std::endl;
                return false;
            }
            rp_lm.push_back(lm_idx.at(vs[0]->Id()));
            rp_i.push_back(pose_idx.at(vs[1]->Id()));
            rp_j.push_back(pose_idx.at(vs[2]->Id()));
            for (int k = 0; k < 3; ++k) rp_pti.push_back(er->PtsI()[k]);
            rp_ptj.push_back(er->PtsJ()[0]);
            rp_ptj.push_back(er->PtsJ()[1]);
        } else if (t == "EdgeReprojectionXYZ") {
            auto *ex = dynamic_cast<EdgeReprojectionXYZ *>(e.get());
            if (!ex || vs.size() < 2) { std::cerr << "vio_b200: foreign EdgeReprojectionXYZ type" << std::endl; return false; }
            const MatXX info = e->Information();
            const double c = info(0, 0);
            if (info.rows() != 2 || info(1, 1) != c || info(0, 1) != 0.0 || info(1, 0) != 0.0) {
                std::cerr << "vio_b200: reprojection information must be c*I2" << std::endl;
                return false;
            }
            LossFunction *lf = e->GetLossFunction();
            const int kind = lf ? lf->KindB200() : 0;
            const double delta = lf ? lf->DeltaB200() : 1.0;
            if (kind < 0) { std::cerr << "vio_b200: user-defined loss functions are not supported" << std::endl; return false; }
            if (!have_rp) {
                g.rp_info = c; g.rp_loss = kind; g.rp_loss_delta = delta; g.ext_pose = -1;
                g.q_ic[0] = ex->Qic().x(); g.q_ic[1] = ex->Qic().y(); g.q_ic[2] = ex->Qic().z(); g.q_ic[3] = ex->Qic().w();
                g.t_ic[0] = ex->Tic().x(); g.t_ic[1] = ex->Tic().y(); g.t_ic[2] = ex->Tic().z();
                have_rp = true;
            } else if (g.rp_info != c || g.rp_loss != kind || g.rp_loss_delta != delta ||
                       (g.ext_pose < 0 && (g.q_ic[3] != ex->Qic().w() || g.q_ic[0] != ex->Qic().x() || g.t_ic[0] != ex->Tic().x()))) {
                std::cerr << "vio_b200: reprojection edges must share information, loss and extrinsics" << std::endl;
                return false;
            }
            K.rx_point.push_back(K.pt_idx.at(vs[0]->Id()));
            K.rx_pose.push_back(pose_idx.at(vs[1]->Id()));
            K.rx_obs.push_back(ex->Obs()[0]);
            K.rx_obs.push_back(ex->Obs()[1]);
        } else if (t == "EdgeSE3Prior") {
            auto *ep = dynamic_cast<EdgeSE3Prior *>(e.get());
            if (!ep) return false;
            sp_pose.push_back(pose_idx.at(vs[0]->Id()));
            for (int k = 0; k < 3; ++k) sp_p.push_back(ep->Pp()[k]);
            sp_q.push_back(ep->Qp().x()); sp_q.push_back(ep->Qp().y()); sp_q.push_back(ep->Qp().z()); sp_q.push_back(ep->Qp().w());
            const MatXX info = e->Information();
            for (int r = 0; r < 6; ++r)
                for (int c2 = 0; c2 < 6; ++c2) sp_info.push_back(info(r, c2));
        } else if (t == "EdgeImu") {
            auto *ei = dynamic_cast<EdgeImuB200 *>(e.get());
            if (!ei || !v17_) { std::cerr << "vio_b200: EdgeImu needs the v17 flavour" << std::endl; return false; }
            const ImuPreintegrationB200 &q = ei->Pre();
            imu_pi.push_back(pose_idx.at(vs[0]->Id())); imu_si.push_back(sb_idx.at(vs[1]->Id()));
            imu_pj.push_back(pose_idx.at(vs[2]->Id())); imu_sj.push_back(sb_idx.at(vs[3]->Id()));
            imu_dt.push_back(q.sum_dt);
            for (int k = 0; k < 3; ++k) { imu_dp.push_back(q.delta_p[k]); imu_dv.push_back(q.delta_v[k]); imu_ba.push_back(q.linearized_ba[k]); imu_bg.push_back(q.linearized_bg[k]); }
            imu_dq.push_back(q.delta_q.x()); imu_dq.push_back(q.delta_q.y()); imu_dq.push_back(q.delta_q.z()); imu_dq.push_back(q.delta_q.w());
            for (int r = 0; r < 15; ++r)
                for (int c2 = 0; c2 < 15; ++c2) { imu_jac.push_back(q.jacobian(r, c2)); imu_cov.push_back(q.covariance(r, c2)); }
        } else {
            std::cerr << "vio_b200: edge type " << t << " is not on the device path (user-defined edges: next round)" << std::endl;
            return false;
        }
    }
    g.n_pose = (int32_t)pose_v.size(); g.pose = pose.data(); g.pose_fixed = pose_fixed.data();
    g.n_speedbias = (int32_t)sb_v.size(); g.speedbias = sb.data(); g.speedbias_fixed = sb_fixed.data();
    g.pclass_order = pclass.data();
    g.n_landmark = (int32_t)lm_v.size(); g.inv_depth = invd.data();
    g.n_reproj = (int64_t)rp_lm.size(); g.rp_landmark = rp_lm.data(); g.rp_pose_i = rp_i.data(); g.rp_pose_j = rp_j.data();
    g.rp_pts_i = rp_pti.data(); g.rp_pts_j = rp_ptj.data();
    g.n_se3prior = (int32_t)sp_pose.size(); g.sp_pose = sp_pose.data(); g.sp_p = sp_p.data(); g.sp_q = sp_q.data(); g.sp_info = sp_info.data();
    g.n_imu = (int32_t)imu_pi.size(); g.imu_pose_i = imu_pi.data(); g.imu_sb_i = imu_si.data(); g.imu_pose_j = imu_pj.data(); g.imu_sb_j = imu_sj.data();
    g.imu_sum_dt = imu_dt.data(); g.imu_delta_p = imu_dp.data(); g.imu_delta_q = imu_dq.data(); g.imu_delta_v = imu_dv.data();
    g.imu_lin_ba = imu_ba.data(); g.imu_lin_bg = imu_bg.data(); g.imu_jacobian = imu_jac.data(); g.imu_covariance = imu_cov.data();
    g.gravity[0] = 0; g.gravity[1] = 0; g.gravity[2] = 9.81;
    g.storage = VIO_STORAGE_AUTO;
    g.n_point = (int32_t)K.pt_v.size(); g.point_xyz = K.pt.data();
    g.n_reproj_xyz = (int64_t)K.rx_point.size(); g.rx_point = K.rx_point.data(); g.rx_pose = K.rx_pose.data(); g.rx_obs = K.rx_obs.data();
    // fixed landmark-class vertices are constants on the device path (see vio_b200.h)
    g.landmark_fixed = lm_fixed.empty() ? nullptr : lm_fixed.data();
    g.point_fixed = K.pt_fixed.empty() ? nullptr : K.pt_fixed.data();

    return true;
}

bool Problem::Solve(int iterations) {
    if (edges_.size() == 0 || verticies_.size() == 0) {
        std::cerr << "\nCannot solve problem without edges or verticies" << std::endl;
        return false;
    }
    const auto t0 = std::chrono::steady_clock::now();
    if (problemType_ != ProblemType::SLAM_PROBLEM) return SolveGenericB200(iterations);
    PackB200 K;
    if (!PackGraphB200(K)) return false;
    vio_graph &g = K.g;
    auto &pose = K.pose; auto &sb = K.sb; auto &invd = K.invd;
    auto &pose_v = K.pose_v; auto &sb_v = K.sb_v; auto &lm_v = K.lm_v;
    if (!handle_) {
        handle_ = acquire_handle(device_);
        if (!handle_) throw std::runtime_error("vio_b200: no CUDA device (the GPU backend has no CPU fallback)");
    }
    int rc = vio_set_graph(handle_, &g);
    if (rc != VIO_OK) { std::cerr << "vio_b200: " << vio_last_error(handle_) << std::endl; return false; }
    const int P = (int)ordering_poses_;
    const bool have_prior = v17_ && H_prior_.rows() == P && P > 0 && (H_prior_.array() != 0.0).any();
    if (have_prior) {
        std::vector<double> Hp((size_t)P * P), Jt;
        for (int r = 0; r < P; ++r)
            for (int c2 = 0; c2 < P; ++c2) Hp[(size_t)r * P + c2] = H_prior_(r, c2);
        const int ed = (int)err_prior_.rows();
        if (ed > 0) {
            Jt.resize((size_t)ed * ed);
            for (int r = 0; r < ed; ++r)
                for (int c2 = 0; c2 < ed; ++c2) Jt[(size_t)r * ed + c2] = Jt_prior_inv_(r, c2);
        }
        rc = vio_set_prior(handle_, P, Hp.data(), b_prior_.data(), ed, ed ? err_prior_.data() : nullptr, ed ? Jt.data() : nullptr);
        if (rc != VIO_OK) { std::cerr << "vio_b200: " << vio_last_error(handle_) << std::endl; return false; }
    }
    vio_lm_opts o;
    std::memset(&o, 0, sizeof(o));
    o.flavour = v17_ ? VIO_LM_V17 : VIO_LM_V15;
    o.solver = VIO_SOLVER_AUTO;
    o.verbose = 1;  // the reference prints "iter: .. , chi= .. , Lambda= .." per iteration
    vio_stats st;
    rc = vio_solve(handle_, iterations, &o, &st);
    if (rc != VIO_OK) { std::cerr << "vio_b200: " << vio_last_error(handle_) << std::endl; return false; }
    rc = vio_get_vertices(handle_, pose.data(), sb.empty() ? nullptr : sb.data(), invd.empty() ? nullptr : invd.data());
    if (rc != VIO_OK) return false;
    for (size_t i = 0; i < pose_v.size(); ++i) {
        VecX &x = pose_v[i]->Parameters();
        for (int k = 0; k < 7; ++k) x[k] = pose[7 * i + k];
    }
    for (size_t i = 0; i < sb_v.size(); ++i) {
        VecX &x = sb_v[i]->Parameters();
        for (int k = 0; k < 9; ++k) x[k] = sb[9 * i + k];
    }
    for (size_t i = 0; i < lm_v.size(); ++i) lm_v[i]->Parameters()[0] = invd[i];
    if (!K.pt_v.empty()) {
        if (vio_get_points(handle_, K.pt.data()) != VIO_OK) return false;
        for (size_t i = 0; i < K.pt_v.size(); ++i)
            for (int k = 0; k < 3; ++k) K.pt_v[i]->Parameters()[k] = K.pt[3 * i + k];
    }
    if (have_prior && err_prior_.rows() > 0) vio_get_prior(handle_, b_prior_.data(), err_prior_.data());
    last_hessian_ms_ = st.ms_linearize;
    last_solve_ms_ = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::cout << "problem solve cost: " << last_solve_ms_ << " ms" << std::endl;
    std::cout << "   makeHessian cost: " << last_hessian_ms_ << " ms" << std::endl;
    return true;
}

// GENERIC_PROBLEM: vertices and edges are user subclasses whose ComputeResidual / ComputeJacobians / Plus are host
// virtuals (A15/app/CurveFitting.cpp:14-48).  The host evaluates the factors through those virtuals exactly like the
// reference's MakeHessian does (A17/src/backend/problem.cc:311-330) and hands them to the device, which builds H, b,
// solves (H + lambda I) dx = b and reduces chi2 / scale (vio_dense_* in vio_b200.h).  The LM control below is the
// reference's Solve loop (v15: 15-vio-backend/backend/problem.cc:155-222, v17: vins-mono/src/backend/problem.cc:169-250).
// Deviation, on purpose: v15's generic branch solves with the UNDAMPED Hessian_ (problem.cc:347-352), a defect that
// makes its own CurveFitting demo return 0 0 0; both flavours use the damped system here.
bool Problem::SolveGenericB200(int iterations) {
    const auto t0 = std::chrono::steady_clock::now();
    if (!handle_) {
        handle_ = acquire_handle(device_);
        if (!handle_) throw std::runtime_error("vio_b200: no CUDA device (the GPU backend has no CPU fallback)");
    }
    // ordering: vertices in id order at consecutive offsets (a single vertex in every reference driver)
    int n = 0;
    std::unordered_map<unsigned long, int> off;
    std::vector<std::shared_ptr<Vertex>> verts;
    for (auto &kv : verticies_) {
        off[kv.first] = n;
        kv.second->SetOrderingId(n);
        n += kv.second->LocalDimension();
        verts.push_back(kv.second);
    }
    ordering_generic_ = n;
    std::vector<unsigned long> eids;
    for (auto &kv : edges_) eids.push_back(kv.first);
    std::sort(eids.begin(), eids.end());
    int R = 0, dmax = 1;
    for (auto id : eids) { const int d = (int)edges_[id]->Residual().rows(); R += d; dmax = std::max(dmax, d); }
    std::vector<double> J((size_t)R * n), r(R), W((size_t)R * dmax), Wb((size_t)R * dmax), Om((size_t)R * dmax), delta(R);
    std::vector<int32_t> e0(R), dim(R), kind(R);
    bool user_loss = false;
    auto evaluate = [&](bool jac) {
        int row = 0;
        for (auto id : eids) {
            auto &e = edges_[id];
            e->ComputeResidual();
            const VecX res = e->Residual();
            const int d = (int)res.rows();
            const MatXX info = e->Information();
            LossFunction *lf = e->GetLossFunction();
            const int k = lf ? lf->KindB200() : 0;
            if (k < 0) user_loss = true;
            for (int a = 0; a < d; ++a) {
                r[row + a] = res[a]; e0[row + a] = row; dim[row + a] = d; kind[row + a] = k < 0 ? 0 : k;
                delta[row + a] = lf ? lf->DeltaB200() : 1.0;
                for (int c = 0; c < d; ++c) Om[(size_t)(row + a) * dmax + c] = info(a, c);
            }
            if (jac) {
                e->ComputeJacobians();
                const auto jacs = e->Jacobians();
                const auto vs = e->Verticies();
                double drho;
                MatXX rinfo(d, d);
                e->RobustInfo(drho, rinfo);
                for (int a = 0; a < d; ++a) {
                    for (int c = 0; c < n; ++c) J[(size_t)(row + a) * n + c] = 0.0;
                    for (int c = 0; c < d; ++c) { W[(size_t)(row + a) * dmax + c] = rinfo(a, c); Wb[(size_t)(row + a) * dmax + c] = drho * info(a, c); }
                }
                for (size_t vi = 0; vi < vs.size(); ++vi) {
                    if (vs[vi]->IsFixed()) continue;
                    const int o = off.at(vs[vi]->Id()), ld = vs[vi]->LocalDimension();
                    for (int a = 0; a < d; ++a)
                        for (int c = 0; c < ld; ++c) J[(size_t)(row + a) * n + o + c] += jacs[vi](a, c);
                }
            }
            row += d;
        }
    };
    auto chi2_now = [&](double &out) -> bool {
        if (user_loss) { std::cerr << "vio_b200: user-defined loss functions are not supported" << std::endl; return false; }
        int rc = vio_dense_chi2(handle_, R, dmax, r.data(), e0.data(), dim.data(), Om.data(), kind.data(), delta.data(), &out);
        if (rc != VIO_OK) { std::cerr << "vio_b200: " << vio_last_error(handle_) << std::endl; return false; }
        if (v17_) out *= 0.5;
        return true;
    };
    double hess_ms = 0, maxdiag = 0;
    auto make_hessian = [&]() -> bool {
        const auto th = std::chrono::steady_clock::now();
        evaluate(true);
        vio_dense_system sys;
        std::memset(&sys, 0, sizeof(sys));
        sys.n = n; sys.rows = R; sys.dmax = dmax; sys.J = J.data(); sys.r = r.data(); sys.row_edge0 = e0.data();
        sys.row_dim = dim.data(); sys.W = W.data(); sys.Wb = Wb.data();
        int rc = vio_dense_accumulate(handle_, &sys, &maxdiag);
        hess_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - th).count();
        if (rc != VIO_OK) { std::cerr << "vio_b200: " << vio_last_error(handle_) << std::endl; return false; }
        return true;
    };
    if (!make_hessian()) return false;
    double chi = 0;
    if (!chi2_now(chi)) return false;
    if (v17_) maxdiag = std::min(5e10, maxdiag);
    double lambda = 1e-5 * maxdiag, ni = 2.0, last_chi = 1e20;
    const double stop_thr = 1e-6 * chi;
    bool stop = false;
    int iter = 0;
    std::vector<double> dx(n);
    while (!stop && iter < iterations) {
        std::cout << "iter: " << iter << " , chi= " << chi << " , Lambda= " << lambda << std::endl;
        bool ok = false;
        int false_cnt = 0;
        while (!ok && (!v17_ || false_cnt < 10)) {
            double dot = 0, dx2 = 0;
            int rc = vio_dense_solve(handle_, lambda, dx.data(), &dot, &dx2);
            if (rc != VIO_OK) { std::cerr << "vio_b200: " << vio_last_error(handle_) << std::endl; return false; }
            if (!v17_ && (dx2 <= 1e-6 || false_cnt > 10)) { stop = true; break; }
            for (auto &v : verts) {
                v->BackUpParameters();
                VecX d = Eigen::Map<VecX>(dx.data() + v->OrderingId(), v->LocalDimension());
                v->Plus(d);
            }
            const double scale = v17_ ? 0.5 * dot + 1e-6 : dot + 1e-3;
            evaluate(false);
            double temp = 0;
            if (!chi2_now(temp)) return false;
            const double rho = (chi - temp) / scale;
            if (rho > 0 && std::isfinite(temp)) {
                double alpha = 1. - std::pow(2 * rho - 1, 3);
                alpha = std::min(alpha, 2. / 3.);
                lambda *= std::max(1. / 3., alpha);
                ni = 2; chi = temp; ok = true;
            } else { lambda *= ni; ni *= 2; ok = false; }
            if (ok) { if (!make_hessian()) return false; false_cnt = 0; }
            else {
                false_cnt++;
                for (auto &v : verts) {
                    if (v17_) v->RollBackParameters();
                    else { VecX d = Eigen::Map<VecX>(dx.data() + v->OrderingId(), v->LocalDimension()); v->Plus(-d); }
                }
            }
        }
        iter++;
        if (v17_) { if (last_chi - chi < 1e-5) { std::cout << "sqrt(currentChi_) <= stopThresholdLM_" << std::endl; stop = true; } }
        else if (std::sqrt(chi) <= stop_thr) stop = true;
        last_chi = chi;
    }
    last_hessian_ms_ = hess_ms;
    last_solve_ms_ = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::cout << "problem solve cost: " << last_solve_ms_ << " ms" << std::endl;
    std::cout << "   makeHessian cost: " << last_hessian_ms_ << " ms" << std::endl;
    return true;
}

// Problem::Marginalize(margVertexs, pose_dim): vins-mono/src/backend/problem.cc:617-795.  margVertexs[0] is the frame
// pose, margVertexs[1] (optional) its speed-bias.  The elimination runs on the device (vio_marginalize); afterwards the
// marginalised vertices and the landmarks that were eliminated with them are removed from the graph, like upstream.
bool Problem::Marginalize(const std::vector<std::shared_ptr<Vertex>> margVertexs, int pose_dim) {
    if (margVertexs.empty() || !v17_) { std::cerr << "vio_b200: Marginalize needs the v17 flavour and a frame vertex" << std::endl; return false; }
    PackB200 K;
    if (!PackGraphB200(K)) return false;
    if ((int)ordering_poses_ != pose_dim) { std::cerr << "vio_b200: pose_dim does not match the pose-class dimension" << std::endl; return false; }
    if (!handle_) {
        handle_ = acquire_handle(device_);
        if (!handle_) throw std::runtime_error("vio_b200: no CUDA device (the GPU backend has no CPU fallback)");
    }
    int rc = vio_set_graph(handle_, &K.g);
    if (rc != VIO_OK) { std::cerr << "vio_b200: " << vio_last_error(handle_) << std::endl; return false; }
    const int P = pose_dim;
    if (H_prior_.rows() == P) {
        std::vector<double> Hp((size_t)P * P);
        for (int r = 0; r < P; ++r)
            for (int c = 0; c < P; ++c) Hp[(size_t)r * P + c] = H_prior_(r, c);
        rc = vio_set_prior(handle_, P, Hp.data(), b_prior_.data(), 0, nullptr, nullptr);
        if (rc != VIO_OK) { std::cerr << "vio_b200: " << vio_last_error(handle_) << std::endl; return false; }
    }
    auto ip = K.pose_idx.find(margVertexs[0]->Id());
    if (ip == K.pose_idx.end()) { std::cerr << "vio_b200: margVertexs[0] must be a VertexPose of this problem" << std::endl; return false; }
    int msb = -1;
    if (margVertexs.size() > 1) {
        auto is = K.sb_idx.find(margVertexs[1]->Id());
        if (is == K.sb_idx.end()) { std::cerr << "vio_b200: margVertexs[1] must be a VertexSpeedBias of this problem" << std::endl; return false; }
        msb = is->second;
    }
    std::vector<double> H((size_t)P * P), b(P), e(P), J((size_t)P * P);
    int32_t n = 0;
    rc = vio_marginalize(handle_, ip->second, msb, &n, H.data(), b.data(), e.data(), J.data());
    if (rc != VIO_OK) { std::cerr << "vio_b200: " << vio_last_error(handle_) << std::endl; return false; }
    H_prior_ = Eigen::Map<Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor>>(H.data(), n, n);
    Jt_prior_inv_ = Eigen::Map<Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor>>(J.data(), n, n);
    b_prior_ = Eigen::Map<VecX>(b.data(), n);
    err_prior_ = Eigen::Map<VecX>(e.data(), n);
    // remove the marginalised vertices and the landmarks connected to the frame (problem.cc:786-793)
    std::vector<std::shared_ptr<Vertex>> lms;
    for (auto &ed : GetConnectedEdges(margVertexs[0]))
        for (auto &vv : ed->Verticies())
            if (IsLandmarkVertex(vv) && std::find(lms.begin(), lms.end(), vv) == lms.end()) lms.push_back(vv);
    for (auto &mvx : margVertexs) RemoveVertex(mvx);
    for (auto &lv : lms) RemoveVertex(lv);
    return true;
}
bool Problem::Marginalize(const std::shared_ptr<Vertex>) { return true; }  // the v15 stub returns true as well

// The toy Schur-complement demo TestMonoBA prints after Solve (3x3 information matrix, variable 1 marginalised).
void Problem::TestMarginalize() {
    const int idx = 1, D = 1, N = 3, M = N - D;
    const double d1 = 0.1 * 0.1, d2 = 0.2 * 0.2, d3 = 0.3 * 0.3;
    MatXX H(MatXX::Zero(N, N));
    H << 1. / d1, -1. / d1, 0, -1. / d1, 1. / d1 + 1. / d2 + 1. / d3, -1. / d3, 0., -1. / d3, 1 / d3;
    std::cout << "---------- TEST Marg: before marg------------" << std::endl << H << std::endl;
    std::vector<int> perm;
    for (int i = 0; i < N; ++i) if (i != idx) perm.push_back(i);
    perm.push_back(idx);
    // the reference swaps row/col idx with the last one; for N = 3, idx = 1 both give the order {0, 2, 1}
    MatXX Hp(N, N);
    for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c) Hp(r, c) = H(perm[r], perm[c]);
    std::cout << "---------- TEST Marg: target variable has been moved to bottom right------------" << std::endl << Hp << std::endl;
    const double eps = 1e-8;
    Eigen::MatrixXd Amm = 0.5 * (Hp.block(M, M, D, D) + Hp.block(M, M, D, D).transpose());
    Eigen::SelfAdjointEigenSolver<Eigen::MatrixXd> saes(Amm);
    Eigen::MatrixXd Amm_inv = saes.eigenvectors() *
                              Eigen::VectorXd((saes.eigenvalues().array() > eps).select(saes.eigenvalues().array().inverse(), 0)).asDiagonal() *
                              saes.eigenvectors().transpose();
    Eigen::MatrixXd prior = Hp.block(0, 0, M, M) - Hp.block(0, M, M, D) * Amm_inv * Hp.block(M, 0, D, M);
    std::cout << "---------- TEST Marg: after marg------------" << std::endl << prior << std::endl;
}

}  // namespace backend
}  // namespace myslam
