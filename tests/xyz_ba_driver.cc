// xyz_ba_driver.cc — TEST DRIVER written for this repository (no reference driver uses VertexPointXYZ /
// EdgeReprojectionXYZ).  It only uses the public API that the reference backend (15-vio-backend/backend) and the
// B200 drop-in (include/backend) share, so the SAME source is compiled twice: against the unmodified reference
// (oracle/_ref/xyz_ba_ref15, oracle/Makefile) and against include/backend + libvio_backend.so (build/xyz_ba_b200).
// tests/test_gpu_dropin.py compares what the two binaries print.
#include <cmath>
#include <cstdio>
#include <iostream>
#include <memory>
#include <vector>

#include "backend/edge_prior.h"
#include "backend/edge_reprojection.h"
#include "backend/problem.h"
#include "backend/vertex_point_xyz.h"
#include "backend/vertex_pose.h"

using namespace myslam::backend;

static unsigned long long lcg_state = 88172645463325252ULL;
static double urand() {  // xorshift64*: identical draws in both builds
    lcg_state ^= lcg_state >> 12; lcg_state ^= lcg_state << 25; lcg_state ^= lcg_state >> 27;
    return (double)((lcg_state * 2685821657736338717ULL) >> 11) / 9007199254740992.0;
}
static double nrand() { return std::sqrt(-2.0 * std::log(urand() + 1e-300)) * std::cos(6.283185307179586 * urand()); }

int main() {
    const int n_cam = 8, n_pt = 120;
    const double radius = 8.0;
    std::vector<Eigen::Matrix3d> Rw(n_cam);
    std::vector<Eigen::Vector3d> tw(n_cam), pts(n_pt);
    for (int i = 0; i < n_cam; ++i) {  // quarter circle, optical axis towards the centre (TestMonoBA-like geometry)
        const double th = i * 2.0 * M_PI / (n_cam * 4);
        Rw[i] = Eigen::AngleAxisd(th, Eigen::Vector3d::UnitZ()).toRotationMatrix();
        tw[i] = Eigen::Vector3d(radius * std::cos(th) - radius, radius * std::sin(th), 1.0 * std::sin(2 * th));
    }
    for (int k = 0; k < n_pt; ++k) pts[k] = Eigen::Vector3d(-4.0 + 8.0 * urand(), -4.0 + 8.0 * urand(), 4.0 + 4.0 * urand());

    Eigen::Quaterniond qic(1, 0, 0, 0);
    Eigen::Vector3d tic(0.05, -0.02, 0.01);
    Problem problem(Problem::ProblemType::SLAM_PROBLEM);
    std::vector<std::shared_ptr<VertexPose>> cams;
    for (int i = 0; i < n_cam; ++i) {
        std::shared_ptr<VertexPose> v(new VertexPose());
        Eigen::VectorXd x(7);
        Eigen::Quaterniond q(Rw[i]);
        Eigen::Vector3d t = tw[i];
        if (i >= 2) t += 0.05 * Eigen::Vector3d(nrand(), nrand(), nrand());  // perturbed initial guess
        x << t, q.x(), q.y(), q.z(), q.w();
        v->SetParameters(x);
        if (i < 2) v->SetFixed();
        problem.AddVertex(v);
        cams.push_back(v);
    }
    std::vector<std::shared_ptr<VertexPointXYZ>> points;
    for (int k = 0; k < n_pt; ++k) {
        std::shared_ptr<VertexPointXYZ> v(new VertexPointXYZ());
        Eigen::VectorXd x(3);
        x << pts[k] + 0.1 * Eigen::Vector3d(nrand(), nrand(), nrand());
        v->SetParameters(x);
        problem.AddVertex(v);
        points.push_back(v);
        for (int i = 0; i < n_cam; ++i) {
            // normalised image observation of the true point in camera i, with pixel noise
            Eigen::Vector3d pb = Rw[i].transpose() * (pts[k] - tw[i]);
            Eigen::Vector3d pc = qic.inverse() * (pb - tic);
            Eigen::Vector3d obs(pc.x() / pc.z() + 1e-3 * nrand(), pc.y() / pc.z() + 1e-3 * nrand(), 1.0);
            std::shared_ptr<EdgeReprojectionXYZ> e(new EdgeReprojectionXYZ(obs));
            e->SetTranslationImuFromCamera(qic, tic);
            std::vector<std::shared_ptr<Vertex>> vs{v, cams[i]};
            e->SetVertex(vs);
            problem.AddEdge(e);
        }
    }
    problem.Solve(10);
    std::cout.setf(std::ios::fixed);
    std::cout.precision(6);
    for (int i = 0; i < n_cam; ++i) {
        Eigen::VectorXd x = cams[i]->Parameters();
        std::cout << "cam " << i << " : " << x[0] << " " << x[1] << " " << x[2] << " (gt " << tw[i].transpose() << ")" << std::endl;
    }
    for (int k = 0; k < n_pt; k += 10) {
        Eigen::VectorXd x = points[k]->Parameters();
        std::cout << "point " << k << " : " << x[0] << " " << x[1] << " " << x[2] << " (gt " << pts[k].transpose() << ")" << std::endl;
    }
    return 0;
}
