// myslam_backend_b200.h — source-compatible mirror of the reference's `myslam::backend` API whose
// Problem::Solve runs on the GPU through the C-ABI of include/vio_b200.h.
//
// What is mirrored (so the reference's drivers compile and link unchanged; paths under
// /root/reference/workspace/assignments):
//   Vertex / VertexPose / VertexInverseDepth / VertexPointXYZ / VertexMotion / VertexSpeedBias
//       15-vio-backend/backend/vertex.h:13-74, vertex_pose.h:17-21, vertex_inverse_depth.h:12-18,
//       17-vins-initialization/vins-mono/include/backend/vertex.h:45-46, vertex_speedbias.h:15-23
//   Edge / EdgeReprojection / EdgeSE3Prior / EdgeImu, loss functions
//       15-vio-backend/backend/edge.h:17-123, edge_reprojection.h:21-49, edge_prior.h:23-45,
//       17-vins-initialization/vins-mono/include/backend/edge.h:84-110, edge_imu.h:18-65, loss_function.h:23-91
//   Problem
//       15-vio-backend/backend/problem.h:17-190, 17-vins-initialization/vins-mono/include/backend/problem.h:68-90
//
// Ownership is the reference's: drivers own vertices/edges through shared_ptr, Problem stores the pointers, and
// Solve writes the optimised parameters back into the SAME vertex objects.  The numerics of Solve never run on the
// host: built-in edge types are packed into a flat vio_graph and evaluated by the CUDA kernels.
//
// Flavour: the LM constants follow the v15 backend unless MYSLAM_B200_V17 is defined (or SetFlavourV17(true) is
// called), which selects the v17 constants, exact reduced solve, robust kernels and the prior hand-off.
#ifndef MYSLAM_BACKEND_B200_H
#define MYSLAM_BACKEND_B200_H

#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include <Eigen/Core>
#include <Eigen/Geometry>

typedef Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic> MatXX;
typedef Eigen::Matrix<double, Eigen::Dynamic, 1> VecX;
typedef Eigen::Matrix<double, 15, 15> Mat1515;
typedef Eigen::Matrix<double, 6, 6> Mat66;
typedef Eigen::Matrix<double, 3, 3> Mat33;
typedef Eigen::Matrix<double, 2, 2> Mat22;
typedef Eigen::Matrix<double, 2, 3> Mat23;
typedef Eigen::Matrix<double, 15, 1> Vec15;
typedef Eigen::Matrix<double, 9, 1> Vec9;
typedef Eigen::Matrix<double, 7, 1> Vec7;
typedef Eigen::Matrix<double, 6, 1> Vec6;
typedef Eigen::Matrix<double, 3, 1> Vec3;
typedef Eigen::Matrix<double, 2, 1> Vec2;
typedef Eigen::Matrix<double, 1, 1> Vec1;
typedef Eigen::Quaterniond Qd;
typedef std::vector<Vec2, Eigen::aligned_allocator<Vec2>> VecVec2;
typedef std::vector<Vec3, Eigen::aligned_allocator<Vec3>> VecVec3;

typedef unsigned long ulong;
struct vio_problem;

namespace myslam {
namespace backend {

struct PackB200;
extern unsigned long global_vertex_id;
extern unsigned long global_edge_id;

// ---- vertices -------------------------------------------------------------------------------------------------
class Vertex {
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW;
    explicit Vertex(int num_dimension, int local_dimension = -1);
    virtual ~Vertex();
    int Dimension() const { return (int)parameters_.rows(); }
    int LocalDimension() const { return local_dimension_; }
    unsigned long Id() const { return id_; }
    VecX Parameters() const { return parameters_; }
    VecX &Parameters() { return parameters_; }
    void SetParameters(const VecX &params) { parameters_ = params; }
    void BackUpParameters() { parameters_backup_ = parameters_; }
    void RollBackParameters() { parameters_ = parameters_backup_; }
    virtual void Plus(const VecX &delta);
    virtual std::string TypeInfo() const = 0;
    int OrderingId() const { return (int)ordering_id_; }
    void SetOrderingId(unsigned long id) { ordering_id_ = id; }
    void SetFixed(bool fixed = true) { fixed_ = fixed; }
    bool IsFixed() const { return fixed_; }

protected:
    VecX parameters_, parameters_backup_;
    int local_dimension_;
    unsigned long id_;
    unsigned long ordering_id_ = 0;
    bool fixed_ = false;
};

class VertexPose : public Vertex {  // [tx ty tz qx qy qz qw], 6 local
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW;
    VertexPose() : Vertex(7, 6) {}
    void Plus(const VecX &delta) override;
    std::string TypeInfo() const override { return "VertexPose"; }
};
class VertexInverseDepth : public Vertex {
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW;
    VertexInverseDepth() : Vertex(1) {}
    std::string TypeInfo() const override { return "VertexInverseDepth"; }
};
class VertexPointXYZ : public Vertex {
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW;
    VertexPointXYZ() : Vertex(3) {}
    std::string TypeInfo() const override { return "VertexPointXYZ"; }
};
class VertexMotion : public Vertex {  // v15 name
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW;
    VertexMotion() : Vertex(9) {}
    std::string TypeInfo() const override { return "VertexMotion"; }
};
class VertexSpeedBias : public Vertex {  // v17 name: [v ba bg]
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW;
    VertexSpeedBias() : Vertex(9) {}
    std::string TypeInfo() const override { return "VertexSpeedBias"; }
};

// ---- robust kernels -------------------------------------------------------------------------------------------
class LossFunction {
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW;
    virtual ~LossFunction() {}
    virtual void Compute(double err2, Eigen::Vector3d &rho) const = 0;
    virtual int KindB200() const { return -1; }  // VIO_LOSS_* or -1 for a user-defined kernel
    virtual double DeltaB200() const { return 0.0; }
};
class TrivalLoss : public LossFunction {
public:
    void Compute(double err2, Eigen::Vector3d &rho) const override;
    int KindB200() const override { return 0; }
};
class HuberLoss : public LossFunction {
public:
    explicit HuberLoss(double delta) : delta_(delta) {}
    void Compute(double err2, Eigen::Vector3d &rho) const override;
    int KindB200() const override { return 1; }
    double DeltaB200() const override { return delta_; }
private:
    double delta_;
};
class CauchyLoss : public LossFunction {
public:
    explicit CauchyLoss(double delta) : delta_(delta) {}
    void Compute(double err2, Eigen::Vector3d &rho) const override;
    int KindB200() const override { return 2; }
    double DeltaB200() const override { return delta_; }
private:
    double delta_;
};
class TukeyLoss : public LossFunction {
public:
    explicit TukeyLoss(double delta) : delta_(delta) {}
    void Compute(double err2, Eigen::Vector3d &rho) const override;
    int KindB200() const override { return 3; }
    double DeltaB200() const override { return delta_; }
private:
    double delta_;
};

// ---- edges ------------------------------------------------------------------------------------------------------
class Edge {
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW;
    explicit Edge(int residual_dimension, int num_verticies,
                  const std::vector<std::string> &verticies_types = std::vector<std::string>());
    virtual ~Edge();
    unsigned long Id() const { return id_; }
    bool AddVertex(std::shared_ptr<Vertex> vertex) { verticies_.emplace_back(vertex); return true; }
    bool SetVertex(const std::vector<std::shared_ptr<Vertex>> &vertices) { verticies_ = vertices; return true; }
    std::shared_ptr<Vertex> GetVertex(int i) { return verticies_[i]; }
    std::vector<std::shared_ptr<Vertex>> Verticies() const { return verticies_; }
    size_t NumVertices() const { return verticies_.size(); }
    virtual std::string TypeInfo() const = 0;
    virtual void ComputeResidual() = 0;
    virtual void ComputeJacobians() = 0;
    double Chi2() const;
    double RobustChi2() const;
    VecX Residual() const { return residual_; }
    std::vector<MatXX> Jacobians() const { return jacobians_; }
    void SetInformation(const MatXX &information);
    MatXX Information() const { return information_; }
    MatXX SqrtInformation() const { return sqrt_information_; }
    void SetLossFunction(LossFunction *ptr) { lossfunction_ = ptr; }
    LossFunction *GetLossFunction() { return lossfunction_; }
    void RobustInfo(double &drho, MatXX &info) const;
    void SetObservation(const VecX &observation) { observation_ = observation; }
    VecX Observation() const { return observation_; }
    bool CheckValid();
    int OrderingId() const { return ordering_id_; }
    void SetOrderingId(int id) { ordering_id_ = id; }

protected:
    unsigned long id_;
    int ordering_id_ = 0;
    std::vector<std::string> verticies_types_;
    std::vector<std::shared_ptr<Vertex>> verticies_;
    VecX residual_;
    std::vector<MatXX> jacobians_;
    MatXX information_, sqrt_information_;
    VecX observation_;
    LossFunction *lossfunction_ = nullptr;
};

// Vertices: [VertexInverseDepth, host VertexPose, observing VertexPose] (v15) plus an extrinsic VertexPose (v17).
class EdgeReprojection : public Edge {
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW;
    EdgeReprojection(const Vec3 &pts_i, const Vec3 &pts_j);
    std::string TypeInfo() const override { return "EdgeReprojection"; }
    void ComputeResidual() override;   // built-in edges are evaluated on the device inside Problem::Solve
    void ComputeJacobians() override;
    void SetTranslationImuFromCamera(Eigen::Quaterniond &qic_, Vec3 &tic_);
    const Vec3 &PtsI() const { return pts_i_; }
    const Vec3 &PtsJ() const { return pts_j_; }
    const Qd &Qic() const { return qic; }
    const Vec3 &Tic() const { return tic; }
private:
    Qd qic = Qd::Identity();
    Vec3 tic = Vec3::Zero();
    Vec3 pts_i_, pts_j_;
};

// 2-vertex [VertexPointXYZ, VertexPose] factor (15-vio-backend/backend/edge_reprojection.h:56-83)
class EdgeReprojectionXYZ : public Edge {
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW;
    explicit EdgeReprojectionXYZ(const Vec3 &pts_i);
    std::string TypeInfo() const override { return "EdgeReprojectionXYZ"; }
    void ComputeResidual() override;
    void ComputeJacobians() override;
    void SetTranslationImuFromCamera(Eigen::Quaterniond &qic_, Vec3 &tic_);
    const Vec3 &Obs() const { return obs_; }
    const Qd &Qic() const { return qic; }
    const Vec3 &Tic() const { return tic; }
private:
    Qd qic = Qd::Identity();
    Vec3 tic = Vec3::Zero();
    Vec3 obs_;
};

class EdgeSE3Prior : public Edge {
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW;
    EdgeSE3Prior(const Vec3 &p, const Qd &q);
    std::string TypeInfo() const override { return "EdgeSE3Prior"; }
    void ComputeResidual() override;
    void ComputeJacobians() override;
    const Vec3 &Pp() const { return Pp_; }
    const Qd &Qp() const { return Qp_; }
private:
    Vec3 Pp_;
    Qd Qp_;
};

// Constants of one IMU pre-integration (what EdgeImu reads from the estimator's IntegrationBase:
// A17/include/factor/integration_base.h:160-186,200-206).  EdgeImu(IntegrationBase*) is provided by
// backend/edge_imu.h when the integrator's own integration_base.h is on the include path.
struct ImuPreintegrationB200 {
    double sum_dt = 0;
    Vec3 delta_p = Vec3::Zero(), delta_v = Vec3::Zero(), linearized_ba = Vec3::Zero(), linearized_bg = Vec3::Zero();
    Qd delta_q = Qd::Identity();
    Mat1515 jacobian = Mat1515::Identity(), covariance = Mat1515::Zero();
};
class EdgeImuB200 : public Edge {  // vertices: [VertexPose i, VertexSpeedBias i, VertexPose j, VertexSpeedBias j]
public:
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW;
    explicit EdgeImuB200(const ImuPreintegrationB200 &pre);
    std::string TypeInfo() const override { return "EdgeImu"; }
    void ComputeResidual() override;
    void ComputeJacobians() override;
    const ImuPreintegrationB200 &Pre() const { return pre_; }
private:
    ImuPreintegrationB200 pre_;
};

// ---- problem ------------------------------------------------------------------------------------------------------
typedef std::map<unsigned long, std::shared_ptr<Vertex>> HashVertex;
typedef std::unordered_map<unsigned long, std::shared_ptr<Edge>> HashEdge;
typedef std::unordered_multimap<unsigned long, std::shared_ptr<Edge>> HashVertexIdToEdge;

class Problem {
public:
    enum class ProblemType { SLAM_PROBLEM, GENERIC_PROBLEM };
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW;
    // the flavour default is fixed by the DRIVER's translation unit (MYSLAM_B200_V17), not by how the library was built
#ifdef MYSLAM_B200_V17
    explicit Problem(ProblemType problemType) : Problem(problemType, true) {}
#else
    explicit Problem(ProblemType problemType) : Problem(problemType, false) {}
#endif
    Problem(ProblemType problemType, bool v17_flavour);
    ~Problem();
    bool AddVertex(std::shared_ptr<Vertex> vertex);
    bool RemoveVertex(std::shared_ptr<Vertex> vertex);
    bool AddEdge(std::shared_ptr<Edge> edge);
    bool RemoveEdge(std::shared_ptr<Edge> edge);
    void GetOutlierEdges(std::vector<std::shared_ptr<Edge>> &outlier_edges);  // declared, never defined upstream
#ifdef MYSLAM_B200_V17
    bool Solve(int iterations = 10);
#else
    bool Solve(int iterations);
#endif
    bool Marginalize(std::shared_ptr<Vertex> frameVertex, const std::vector<std::shared_ptr<Vertex>> &landmarkVerticies);
    bool Marginalize(const std::shared_ptr<Vertex> frameVertex);
    bool Marginalize(const std::vector<std::shared_ptr<Vertex>> frameVertex, int pose_dim);
    void TestMarginalize();
    void TestComputePrior();
    // v17 prior hand-off
    MatXX GetHessianPrior() { return H_prior_; }
    VecX GetbPrior() { return b_prior_; }
    VecX GetErrPrior() { return err_prior_; }
    MatXX GetJtPrior() { return Jt_prior_inv_; }
    void SetHessianPrior(const MatXX &H) { H_prior_ = H; }
    void SetbPrior(const VecX &b) { b_prior_ = b; }
    void SetErrPrior(const VecX &b) { err_prior_ = b; }
    void SetJtPrior(const MatXX &J) { Jt_prior_inv_ = J; }
    void ExtendHessiansPriorSize(int dim);
    // B200 additions (not in the reference)
    void SetFlavourV17(bool v17) { v17_ = v17; }
    void SetDevice(int device) { device_ = device; }
    double LastSolveMs() const { return last_solve_ms_; }
    double LastMakeHessianMs() const { return last_hessian_ms_; }

private:
    bool SolveGenericB200(int iterations);
    bool PackGraphB200(PackB200 &K);
    bool IsPoseVertex(std::shared_ptr<Vertex> v);
    bool IsLandmarkVertex(std::shared_ptr<Vertex> v);
    void SetOrdering();
    std::vector<std::shared_ptr<Edge>> GetConnectedEdges(std::shared_ptr<Vertex> vertex);

    ProblemType problemType_;
    bool v17_;
    int device_ = 0;
    vio_problem *handle_ = nullptr;
    MatXX H_prior_, Jt_prior_inv_;
    VecX b_prior_, err_prior_;
    HashVertex verticies_;
    HashEdge edges_;
    HashVertexIdToEdge vertexToEdge_;
    ulong ordering_poses_ = 0, ordering_landmarks_ = 0, ordering_generic_ = 0;
    std::map<unsigned long, std::shared_ptr<Vertex>> idx_pose_vertices_, idx_landmark_vertices_;
    double last_solve_ms_ = 0, last_hessian_ms_ = 0;
};

}  // namespace backend
}  // namespace myslam
#endif
