// ref_bal.cpp — TEST INFRASTRUCTURE (checker), not product code.  Reads a BAL file with the reference's OWN reader
// (07-backend-optimization/01-bal-g2o/src/bal.cpp, class BALProblem, include/bal.hpp:4-91; compiled unmodified next to this
// file by oracle/Makefile) and prints, in full precision, what it parsed plus the pixel the reference's camera model
// predicts for every observation.  The model is restated from VertexPoseAndIntrinsics::project and PoseAndIntrinsics
// (src/bal_g2o.cpp:25-42, 94-109; that file itself needs g2o, which is not in the image): R = SO3d::exp(r).matrix() with
// the reference's vendored Sophus, P = R X + t (g2o::SE3Quat::map), p = -P.xy / P.z, pixel = f (1 + k1 |p|^2 + k2 |p|^4) p.
// Used by tests/golden/make_golden_bal.py to pin visual-inertial-odometry_b200/bal.py.
#include <cstdio>
#include <string>

#include <Eigen/Core>
#include <Eigen/Dense>
#include "sophus/so3.hpp"

#include "bal.hpp"

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    BALProblem bal(argv[1]);
    printf("\nSIZES %d %d %d\n", bal.num_cameras(), bal.num_points(), bal.num_observations());
    for (int i = 0; i < bal.num_observations(); ++i) {
        const double *cam = bal.camera_for_observation(i);
        const double *pt = bal.point_for_observation(i);
        const Eigen::Matrix3d R = Sophus::SO3d::exp(Eigen::Vector3d(cam[0], cam[1], cam[2])).matrix();
        const Eigen::Vector3d t(cam[3], cam[4], cam[5]);
        const Eigen::Vector3d P = R * Eigen::Vector3d(pt[0], pt[1], pt[2]) + t;
        const Eigen::Vector2d p(-P(0) / P(2), -P(1) / P(2));
        const double r2 = p.squaredNorm();
        const double distortion = 1.0 + r2 * (cam[7] + cam[8] * r2);
        printf("OBS %d %d %.17g %.17g %.17g %.17g\n", bal.camera_index()[i], bal.point_index()[i], bal.observations()[2 * i],
               bal.observations()[2 * i + 1], cam[6] * distortion * p(0), cam[6] * distortion * p(1));
    }
    for (int i = 0; i < bal.num_cameras(); ++i) {
        const double *c = bal.cameras() + 9 * i;
        printf("CAM");
        for (int k = 0; k < 9; ++k) printf(" %.17g", c[k]);
        printf("\n");
    }
    for (int i = 0; i < bal.num_points(); ++i) {
        const double *x = bal.points() + 3 * i;
        printf("PT %.17g %.17g %.17g\n", x[0], x[1], x[2]);
    }
    return 0;
}
