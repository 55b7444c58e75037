// scene_gen.cc — deterministic synthetic scenes for BASELINE.json's configs, emitted as the flat
// arrays of a vio_graph.  Host only (libvio_scenes.so); used by tests, bench.py and the drop-in demo.
//
//   monoba : the reference's assignment test scene, draw for draw
//            (/root/reference/workspace/assignments/15-vio-backend/app/TestMonoBA.cpp:28-87,96-184):
//            poses on a quarter arc, every landmark hosted by camera 0 and seen by all cameras,
//            SE3 priors on cameras 0 and 1.  std::default_random_engine (libstdc++ minstd_rand0).
//   ring   : SURVEY.md §8(d) configs 4/5 - C cameras on a full circle, landmarks hosted by camera
//            h = floor(l*C/L) and observed by h+1..h+K-1 (mod C).  std::mt19937_64(seed).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>

namespace {

struct V3 {
    double x, y, z;
    V3(double a, double b, double c) : x(a), y(b), z(c) {}
    V3() : x(0), y(0), z(0) {}
};

// rotation about +z; row-major 3x3
void rotz(double th, double R[9]) {
    const double c = std::cos(th), s = std::sin(th);
    R[0] = c; R[1] = -s; R[2] = 0;
    R[3] = s; R[4] = c; R[5] = 0;
    R[6] = 0; R[7] = 0; R[8] = 1;
}

// rotation matrix -> quaternion xyzw, the trace-based conversion Eigen uses for Quaterniond(Matrix3d)
void mat_to_quat(const double R[9], double q[4]) {
    double t = R[0] + R[4] + R[8];
    if (t > 0.0) {
        t = std::sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (R[7] - R[5]) * t;
        q[1] = (R[2] - R[6]) * t;
        q[2] = (R[3] - R[1]) * t;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[4 * i]) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (R[3 * k + j] - R[3 * j + k]) * t;
        q[j] = (R[3 * j + i] + R[3 * i + j]) * t;
        q[k] = (R[3 * k + i] + R[3 * i + k]) * t;
    }
}

// R^T (p - t)
V3 to_cam(const double R[9], const V3 &t, const V3 &p) {
    const double dx = p.x - t.x, dy = p.y - t.y, dz = p.z - t.z;
    return V3(R[0] * dx + R[3] * dy + R[6] * dz, R[1] * dx + R[4] * dy + R[7] * dz, R[2] * dx + R[5] * dy + R[8] * dz);
}

}  // namespace

extern "C" {

int vio_scene_monoba_sizes(int pose_nums, int feature_nums, int with_ext, int32_t *n_pose, int32_t *n_landmark,
                           int64_t *n_reproj) {
    if (pose_nums < 2 || feature_nums < 1) return 1;
    *n_pose = pose_nums + (with_ext ? 1 : 0);
    *n_landmark = feature_nums;
    *n_reproj = (int64_t)feature_nums * (pose_nums - 1);
    return 0;
}

// Arrays sized by vio_scene_monoba_sizes.  With with_ext, pose 0 is the fixed identity extrinsic
// vertex of the v17 4-vertex EdgeReprojection and cameras are poses 1..pose_nums.
int vio_scene_monoba_fill(int pose_nums, int feature_nums, int with_ext, double prior_weight, double *pose,
                          uint8_t *pose_fixed, double *pose_gt, double *inv_depth, double *inv_depth_gt,
                          int32_t *rp_landmark, int32_t *rp_pose_i, int32_t *rp_pose_j, double *rp_pts_i,
                          double *rp_pts_j, int32_t *sp_pose, double *sp_p, double *sp_q, double *sp_info) {
    const int base = with_ext ? 1 : 0;
    const double radius = 8;
    std::vector<std::vector<double>> Rgt(pose_nums, std::vector<double>(9)), Robs(pose_nums, std::vector<double>(9));
    std::vector<V3> tgt(pose_nums), tobs(pose_nums);
    if (with_ext) {
        const double id[7] = {0, 0, 0, 0, 0, 0, 1};
        std::memcpy(pose, id, sizeof(id));
        std::memcpy(pose_gt, id, sizeof(id));
        pose_fixed[0] = 1;
    }
    std::default_random_engine generator;
    for (int n = 0; n < pose_nums; ++n) {
        std::uniform_real_distribution<double> xyz_rand(-0.05, +0.05);
        std::uniform_real_distribution<double> theta_rand(-0.105, +0.105);
        const double theta_gt = n * 2 * M_PI / (pose_nums * 4);
        const double theta_obs = theta_gt + theta_rand(generator);
        rotz(theta_gt, Rgt[n].data());
        rotz(theta_obs, Robs[n].data());
        const double x_gt = radius * std::cos(theta_gt) - radius;
        const double y_gt = radius * std::sin(theta_gt);
        const double z_gt = 1 * std::sin(2 * theta_gt);
        const double x_obs = x_gt + xyz_rand(generator);
        const double y_obs = y_gt + xyz_rand(generator);
        const double z_obs = z_gt + xyz_rand(generator);
        tgt[n] = V3(x_gt, y_gt, z_gt);
        tobs[n] = V3(x_obs, y_obs, z_obs);
        double q[4];
        double *po = pose + 7 * (base + n), *pg = pose_gt + 7 * (base + n);
        mat_to_quat(Robs[n].data(), q);
        po[0] = x_obs; po[1] = y_obs; po[2] = z_obs; po[3] = q[0]; po[4] = q[1]; po[5] = q[2]; po[6] = q[3];
        mat_to_quat(Rgt[n].data(), q);
        pg[0] = x_gt; pg[1] = y_gt; pg[2] = z_gt; pg[3] = q[0]; pg[4] = q[1]; pg[5] = q[2]; pg[6] = q[3];
        pose_fixed[base + n] = 0;
    }
    // observations: normalised image coordinates + N(0, 1/1000)
    std::vector<V3> points;
    std::vector<std::vector<V3>> obs(pose_nums, std::vector<V3>(feature_nums));
    std::normal_distribution<double> noise_pdf(0., 1. / 1000.);
    for (int j = 0; j < feature_nums; ++j) {
        std::uniform_real_distribution<double> xy_rand(-4, 4.0);
        std::uniform_real_distribution<double> z_rand(4., 8.);
        // same expression shape as the reference driver: three draws as constructor arguments
        V3 Pw(xy_rand(generator), xy_rand(generator), z_rand(generator));
        points.push_back(Pw);
        for (int i = 0; i < pose_nums; ++i) {
            V3 Pc = to_cam(Rgt[i].data(), tgt[i], Pw);
            const double z = Pc.z;
            Pc.x = Pc.x / z; Pc.y = Pc.y / z; Pc.z = Pc.z / z;
            Pc.x += noise_pdf(generator);
            Pc.y += noise_pdf(generator);
            obs[i][j] = Pc;
        }
    }
    // gauge priors on cameras 0 and 1 (ground-truth pose, information = prior_weight * I6)
    for (int i = 0; i < 2; ++i) {
        sp_pose[i] = base + i;
        sp_p[3 * i] = tgt[i].x; sp_p[3 * i + 1] = tgt[i].y; sp_p[3 * i + 2] = tgt[i].z;
        mat_to_quat(Rgt[i].data(), sp_q + 4 * i);
        for (int k = 0; k < 36; ++k) sp_info[36 * i + k] = (k % 7 == 0) ? prior_weight : 0.0;
    }
    // landmarks: inverse depth in camera 0 (its noisy pose) with N(0,1) depth noise from a fresh engine
    std::default_random_engine generator2;
    std::normal_distribution<double> depth_noise(0, 1.);
    int64_t e = 0;
    for (int i = 0; i < feature_nums; ++i) {
        const V3 Pc = to_cam(Robs[0].data(), tobs[0], points[i]);
        const double noise = depth_noise(generator2);
        inv_depth[i] = 1. / (Pc.z + noise);
        inv_depth_gt[i] = 1. / points[i].z;
        for (int j = 1; j < pose_nums; ++j, ++e) {
            rp_landmark[e] = i;
            rp_pose_i[e] = base + 0;
            rp_pose_j[e] = base + j;
            rp_pts_i[3 * e] = obs[0][i].x; rp_pts_i[3 * e + 1] = obs[0][i].y; rp_pts_i[3 * e + 2] = obs[0][i].z;
            rp_pts_j[2 * e] = obs[j][i].x; rp_pts_j[2 * e + 1] = obs[j][i].y;
        }
    }
    return 0;
}

int vio_scene_ring_sizes(int n_cam, int n_landmark, int k_obs, int with_ext, int32_t *n_pose, int64_t *n_reproj) {
    if (n_cam < 3 || n_landmark < 1 || k_obs < 2 || k_obs > n_cam) return 1;
    *n_pose = n_cam + (with_ext ? 1 : 0);
    *n_reproj = (int64_t)n_landmark * (k_obs - 1);
    return 0;
}

// SURVEY.md §8(d) config 4/5 generator.  Camera c: theta = 2 pi c / C, radius 0.1 C, position
// (r cos - r, r sin, sin 2theta), R = Rz(theta).  Landmark l: host h = floor(l C / L), world point =
// host position + (U(-4,4), U(-4,4), U(4,8)); observed by h..h+K-1 (mod C); observation noise N(0,1/1000);
// pose initial noise as monoba; inverse depth init 1/(z_h + N(0, 0.25)); priors on cameras 0, 1.
int vio_scene_ring_fill(int n_cam, int n_landmark, int k_obs, int with_ext, uint64_t seed, double prior_weight,
                        double *pose, uint8_t *pose_fixed, double *pose_gt, double *inv_depth, double *inv_depth_gt,
                        int32_t *rp_landmark, int32_t *rp_pose_i, int32_t *rp_pose_j, double *rp_pts_i,
                        double *rp_pts_j, int32_t *sp_pose, double *sp_p, double *sp_q, double *sp_info) {
    const int base = with_ext ? 1 : 0;
    const double radius = 0.1 * n_cam;
    std::mt19937_64 gen(seed);
    std::uniform_real_distribution<double> xyz_rand(-0.05, +0.05), theta_rand(-0.105, +0.105);
    std::uniform_real_distribution<double> xy_rand(-4.0, 4.0), z_rand(4.0, 8.0);
    std::normal_distribution<double> obs_noise(0., 1. / 1000.), depth_noise(0., 0.25);
    std::vector<double> Rgt(9 * (size_t)n_cam), Robs(9 * (size_t)n_cam);
    std::vector<V3> tgt(n_cam), tobs(n_cam);
    if (with_ext) {
        const double id[7] = {0, 0, 0, 0, 0, 0, 1};
        std::memcpy(pose, id, sizeof(id));
        std::memcpy(pose_gt, id, sizeof(id));
        pose_fixed[0] = 1;
    }
    for (int c = 0; c < n_cam; ++c) {
        const double th = 2 * M_PI * c / n_cam;
        const double th_obs = th + theta_rand(gen);
        rotz(th, &Rgt[9 * (size_t)c]);
        rotz(th_obs, &Robs[9 * (size_t)c]);
        tgt[c] = V3(radius * std::cos(th) - radius, radius * std::sin(th), std::sin(2 * th));
        const double nx = xyz_rand(gen), ny = xyz_rand(gen), nz = xyz_rand(gen);
        tobs[c] = V3(tgt[c].x + nx, tgt[c].y + ny, tgt[c].z + nz);
        double q[4];
        double *po = pose + 7 * (size_t)(base + c), *pg = pose_gt + 7 * (size_t)(base + c);
        mat_to_quat(&Robs[9 * (size_t)c], q);
        po[0] = tobs[c].x; po[1] = tobs[c].y; po[2] = tobs[c].z; po[3] = q[0]; po[4] = q[1]; po[5] = q[2]; po[6] = q[3];
        mat_to_quat(&Rgt[9 * (size_t)c], q);
        pg[0] = tgt[c].x; pg[1] = tgt[c].y; pg[2] = tgt[c].z; pg[3] = q[0]; pg[4] = q[1]; pg[5] = q[2]; pg[6] = q[3];
        pose_fixed[base + c] = 0;
    }
    for (int i = 0; i < 2; ++i) {
        sp_pose[i] = base + i;
        sp_p[3 * i] = tgt[i].x; sp_p[3 * i + 1] = tgt[i].y; sp_p[3 * i + 2] = tgt[i].z;
        mat_to_quat(&Rgt[9 * (size_t)i], sp_q + 4 * i);
        for (int k = 0; k < 36; ++k) sp_info[36 * i + k] = (k % 7 == 0) ? prior_weight : 0.0;
    }
    int64_t e = 0;
    for (int l = 0; l < n_landmark; ++l) {
        const int h = (int)(((int64_t)l * n_cam) / n_landmark);
        // point in the host camera frame, then to world through the ground-truth host pose
        const double lx = xy_rand(gen), ly = xy_rand(gen), lz = z_rand(gen);
        const double *Rh = &Rgt[9 * (size_t)h];
        const V3 Pw(Rh[0] * lx + Rh[1] * ly + Rh[2] * lz + tgt[h].x, Rh[3] * lx + Rh[4] * ly + Rh[5] * lz + tgt[h].y,
                    Rh[6] * lx + Rh[7] * ly + Rh[8] * lz + tgt[h].z);
        inv_depth_gt[l] = 1.0 / lz;
        const V3 Pc0 = to_cam(&Robs[9 * (size_t)h], tobs[h], Pw);
        inv_depth[l] = 1.0 / (Pc0.z + depth_noise(gen));
        const double hx = lx / lz + obs_noise(gen), hy = ly / lz + obs_noise(gen);
        for (int k = 1; k < k_obs; ++k, ++e) {
            const int j = (h + k) % n_cam;
            V3 Pc = to_cam(&Rgt[9 * (size_t)j], tgt[j], Pw);
            rp_landmark[e] = l;
            rp_pose_i[e] = base + h;
            rp_pose_j[e] = base + j;
            rp_pts_i[3 * e] = hx; rp_pts_i[3 * e + 1] = hy; rp_pts_i[3 * e + 2] = 1.0;
            rp_pts_j[2 * e] = Pc.x / Pc.z + obs_noise(gen);
            rp_pts_j[2 * e + 1] = Pc.y / Pc.z + obs_noise(gen);
        }
    }
    return 0;
}
// hessian_nullspace_test scene (/root/reference/workspace/assignments/14-sliding-window/src/hessian_nullspace_test.cpp:
// 45-93): N = 10 cameras on the quarter arc, M = 20 world points drawn x, y ~ U(-4,4), z ~ U(8,10) from ONE
// default-seeded std::default_random_engine in the order x, y, z per point; every camera sees every point.
int vio_scene_nullspace_fill(int draw_order, double *pose /* 10 x 7 */, double *points /* 20 x 3 */) {
    const int N = 10, M = 20;
    const double radius = 8;
    for (int n = 0; n < N; ++n) {
        const double theta = n * 2 * M_PI / (N * 4);
        double R[9], q[4];
        rotz(theta, R);
        mat_to_quat(R, q);
        double *o = pose + 7 * n;
        o[0] = radius * std::cos(theta) - radius; o[1] = radius * std::sin(theta); o[2] = 1 * std::sin(2 * theta);
        o[3] = q[0]; o[4] = q[1]; o[5] = q[2]; o[6] = q[3];
    }
    std::default_random_engine generator;
    std::uniform_real_distribution<double> xy_uniform(-4.0, 4.0), z_uniform(8.0, 10.0);
    for (int m = 0; m < M; ++m) {
        // The reference draws inside a constructor call, Eigen::Vector3d(xy(gen), xy(gen), z(gen)), whose argument
        // evaluation order is unspecified: draw_order 0 = left to right (x, y, z), 1 = right to left (z, y, x; what g++
        // does).  tests/test_oracle.py pins the order against the singular values the reference publishes.
        double x, y, z;
        if (draw_order == 0) { x = xy_uniform(generator); y = xy_uniform(generator); z = z_uniform(generator); }
        else { z = z_uniform(generator); y = xy_uniform(generator); x = xy_uniform(generator); }
        points[3 * m] = x; points[3 * m + 1] = y; points[3 * m + 2] = z;
    }
    return 0;
}
}
