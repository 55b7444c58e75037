// frame_stream_driver.cc — TEST DRIVER written for this repository.  It mimics the way the VINS estimator uses the
// backend (17-vins-initialization/vins-mono/src/estimator.cpp:902-1037): a NEW Problem object is built for every frame
// from freshly created vertices and edges, solved once and dropped.  Frames alternate between two window shapes, and the
// third / fourth frame repeat the first / second exactly, so a backend that keeps anything between Problem objects
// (the B200 drop-in pools its device handles) must still print the same numbers for a repeated frame.
// Only the API shared by the reference backend (15-vio-backend/backend) and the drop-in (include/backend) is used: the
// SAME source is compiled against the unmodified reference (oracle/_ref/frame_stream_ref15) and against
// include/backend + libvio_backend.so (build/frame_stream_b200); tests/test_gpu_dropin.py compares the outputs.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <iostream>
#include <memory>
#include <vector>

#include "backend/edge_prior.h"
#include "backend/edge_reprojection.h"
#include "backend/problem.h"
#include "backend/vertex_inverse_depth.h"
#include "backend/vertex_pose.h"

using namespace myslam::backend;

static unsigned long long rng_state;
static double urand() {  // xorshift64*: identical draws in both builds
    rng_state ^= rng_state >> 12; rng_state ^= rng_state << 25; rng_state ^= rng_state >> 27;
    return (double)((rng_state * 2685821657736338717ULL) >> 11) / 9007199254740992.0;
}
static double nrand() { return std::sqrt(-2.0 * std::log(urand() + 1e-300)) * std::cos(6.283185307179586 * urand()); }

static void frame(int idx, int n_cam, int n_lm, unsigned long long seed) {
    rng_state = seed;
    const double radius = 8.0;
    std::vector<Eigen::Matrix3d> Rw(n_cam);
    std::vector<Eigen::Vector3d> tw(n_cam);
    for (int i = 0; i < n_cam; ++i) {
        const double th = i * 2.0 * M_PI / (n_cam * 4);
        Rw[i] = Eigen::AngleAxisd(th, Eigen::Vector3d::UnitZ()).toRotationMatrix();
        tw[i] = Eigen::Vector3d(radius * std::cos(th) - radius, radius * std::sin(th), 1.0 * std::sin(2 * th));
    }
    Eigen::Quaterniond qic(1, 0, 0, 0);
    Eigen::Vector3d tic(0.0, 0.0, 0.0);
    const auto t0 = std::chrono::steady_clock::now();
    {
        Problem problem(Problem::ProblemType::SLAM_PROBLEM);
        std::vector<std::shared_ptr<VertexPose>> cams;
        for (int i = 0; i < n_cam; ++i) {
            std::shared_ptr<VertexPose> v(new VertexPose());
            Eigen::VectorXd x(7);
            Eigen::Quaterniond q(Rw[i]);
            Eigen::Vector3d t = tw[i];
            if (i >= 2) t += 0.03 * Eigen::Vector3d(nrand(), nrand(), nrand());
            x << t, q.x(), q.y(), q.z(), q.w();
            v->SetParameters(x);
            if (i < 2) v->SetFixed();
            problem.AddVertex(v);
            cams.push_back(v);
        }
        std::vector<std::shared_ptr<VertexInverseDepth>> lms;
        for (int k = 0; k < n_lm; ++k) {
            const Eigen::Vector3d pw(-4.0 + 8.0 * urand(), -4.0 + 8.0 * urand(), 4.0 + 4.0 * urand());
            const int host = k % 3;  // hosts 0, 1, 2: several landmark groups per window
            Eigen::Vector3d pc_h = Rw[host].transpose() * (pw - tw[host]);
            std::shared_ptr<VertexInverseDepth> v(new VertexInverseDepth());
            Eigen::VectorXd x(1);
            x << 1.0 / (pc_h.z() + 0.3 * nrand());
            v->SetParameters(x);
            problem.AddVertex(v);
            lms.push_back(v);
            const Eigen::Vector3d pts_i(pc_h.x() / pc_h.z(), pc_h.y() / pc_h.z(), 1.0);
            for (int j = 0; j < n_cam; ++j) {
                if (j == host) continue;
                Eigen::Vector3d pc = Rw[j].transpose() * (pw - tw[j]);
                Eigen::Vector3d pts_j(pc.x() / pc.z() + 1e-3 * nrand(), pc.y() / pc.z() + 1e-3 * nrand(), 1.0);
                std::shared_ptr<EdgeReprojection> e(new EdgeReprojection(pts_i, pts_j));
                e->SetTranslationImuFromCamera(qic, tic);
                std::vector<std::shared_ptr<Vertex>> vs{v, cams[host], cams[j]};
                e->SetVertex(vs);
                problem.AddEdge(e);
            }
        }
        problem.Solve(6);
        std::cout.setf(std::ios::fixed);
        std::cout.precision(6);
        for (int i = 2; i < n_cam; i += 2) {
            Eigen::VectorXd x = cams[i]->Parameters();
            std::cout << "frame " << idx << " cam " << i << " : " << x[0] << " " << x[1] << " " << x[2] << std::endl;
        }
        for (int k = 0; k < n_lm; k += 17) std::cout << "frame " << idx << " lm " << k << " : " << lms[k]->Parameters()[0] << std::endl;
    }  // ~Problem
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::cerr << "frame " << idx << " (" << n_cam << " cams, " << n_lm << " landmarks): " << ms << " ms incl. graph construction and ~Problem" << std::endl;
}

int main() {
    frame(0, 8, 120, 88172645463325252ULL);
    frame(1, 5, 60, 1234567890123ULL);
    frame(2, 8, 120, 88172645463325252ULL);  // = frame 0
    frame(3, 5, 60, 1234567890123ULL);       // = frame 1
    frame(4, 11, 300, 99991ULL);
    return 0;
}
