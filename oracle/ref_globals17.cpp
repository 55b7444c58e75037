// Definitions of the globals the reference declares in include/parameters.h (their real definitions
// live in parameters.cpp, which needs OpenCV).  TEST INFRASTRUCTURE ONLY.
#include <Eigen/Dense>
double ACC_N = 0.2687, ACC_W = 7.07e-6;
double GYR_N = 0.2121, GYR_W = 7.07e-7;
Eigen::Vector3d G(0.0, 0.0, 9.81);
