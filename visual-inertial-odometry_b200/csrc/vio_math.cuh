// vio_math.cuh — FP64 device math for the backend::Problem hot path (sm_100a).
//
// Formulas follow the reference's arithmetic (paths relative to
// /root/reference/workspace/assignments; A15 = 15-vio-backend, A17 = 17-vins-initialization/vins-mono,
// EIG = 02-kinematics-in-3D-space/workspace/Eigen):
//   quat_to_R        EIG/Eigen/src/Geometry/Quaternion.h:531-563  (toRotationMatrix, no normalisation)
//   quat_mul         EIG/Eigen/src/Geometry/Quaternion.h:430-446  (operator*)
//   so3_exp          A15/thirdparty/Sophus/sophus/so3.hpp:393-420, 682-685 (normalised by the ctor)
//   so3_log          A15/thirdparty/Sophus/sophus/so3.hpp:541-580
//   so3_jr_inv       A15/thirdparty/Sophus/sophus/so3.hpp:130-145
//   loss_*           A17/src/backend/loss_function.cc:10-47, A17/include/backend/loss_function.h:36-44
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define VIO_HD __host__ __device__ __forceinline__

// 1 / z to ~1 ulp without the slow-path branch of the IEEE division: 20-bit hardware estimate + two Newton steps.
// (z is a depth or an inverse depth here: never denormal; 0 / inf / NaN still come out as inf / 0 / NaN.)
VIO_HD double vio_rcp(double z) {
#ifdef __CUDA_ARCH__
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(z));
    double e = fma(-z, y, 1.0);
    y = fma(y, e, y);
    e = fma(-z, y, 1.0);
    y = fma(y, e, y);
    return y;
#else
    return 1.0 / z;
#endif
}

struct Mat3 {
    double m[9];  // row-major
};

VIO_HD void quat_to_R(const double q[4] /*xyzw*/, double R[9]) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1.0 - (tyy + tzz);
    R[1] = txy - twz;
    R[2] = txz + twy;
    R[3] = txy + twz;
    R[4] = 1.0 - (txx + tzz);
    R[5] = tyz - twx;
    R[6] = txz - twy;
    R[7] = tyz + twx;
    R[8] = 1.0 - (txx + tyy);
}

// c = a * b  (Hamilton product, xyzw storage)
VIO_HD void quat_mul(const double a[4], const double b[4], double c[4]) {
    const double ax = a[0], ay = a[1], az = a[2], aw = a[3];
    const double bx = b[0], by = b[1], bz = b[2], bw = b[3];
    c[3] = aw * bw - ax * bx - ay * by - az * bz;
    c[0] = aw * bx + ax * bw + ay * bz - az * by;
    c[1] = aw * by + ay * bw + az * bx - ax * bz;
    c[2] = aw * bz + az * bw + ax * by - ay * bx;
}

// Eigen's Quaternion::inverse(): conjugate / squaredNorm (EIG/Eigen/src/Geometry/Quaternion.h:659-670)
VIO_HD void quat_inv(const double a[4], double c[4]) {
    const double n2 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3];
    c[0] = -a[0] / n2;
    c[1] = -a[1] / n2;
    c[2] = -a[2] / n2;
    c[3] = a[3] / n2;
}

VIO_HD void so3_exp(const double w[3], double q[4] /*xyzw, unit*/) {
    const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    const double th = sqrt(th2);
    double imag, real;
    if (th < 1e-10) {
        const double th4 = th2 * th2;
        imag = 0.5 - (1.0 / 48.0) * th2 + (1.0 / 3840.0) * th4;
        real = 1.0 - 0.5 * th2 + (1.0 / 384.0) * th4;
    } else {
        imag = sin(0.5 * th) / th;
        real = cos(0.5 * th);
    }
    double x = imag * w[0], y = imag * w[1], z = imag * w[2];
    const double len = sqrt(x * x + y * y + z * z + real * real);
    q[0] = x / len;
    q[1] = y / len;
    q[2] = z / len;
    q[3] = real / len;
}

VIO_HD void so3_log(const double q[4] /*xyzw unit*/, double w[3]) {
    const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
    const double n = sqrt(n2);
    const double qw = q[3];
    double f;
    if (n < 1e-10) {
        f = 2.0 / qw - 2.0 * n2 / (qw * qw * qw);
    } else if (fabs(qw) < 1e-10) {
        f = (qw > 0.0 ? 3.14159265358979323846 : -3.14159265358979323846) / n;
    } else {
        f = 2.0 * atan(n / qw) / n;
    }
    w[0] = f * q[0];
    w[1] = f * q[1];
    w[2] = f * q[2];
}

// I + 0.5*hat(k) + (1 - (1+cos t) t / (2 sin t)) hat(k)^2,  k = w/|w|  (as written in the reference)
VIO_HD void so3_jr_inv(const double w[3], double J[9]) {
    const double th = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    J[0] = 1; J[1] = 0; J[2] = 0; J[3] = 0; J[4] = 1; J[5] = 0; J[6] = 0; J[7] = 0; J[8] = 1;
    if (th < 1e-10) return;
    const double kx = w[0] / th, ky = w[1] / th, kz = w[2] / th;
    const double K[9] = {0, -kz, ky, kz, 0, -kx, -ky, kx, 0};
    const double c = 1.0 - (1.0 + cos(th)) * th / (2.0 * sin(th));
    for (int r = 0; r < 3; ++r)
        for (int cc = 0; cc < 3; ++cc) {
            double kk = 0;
            for (int k = 0; k < 3; ++k) kk += K[3 * r + k] * K[3 * k + cc];
            J[3 * r + cc] += 0.5 * K[3 * r + cc] + c * kk;
        }
}

// rho[0..2] = rho(e2), rho'(e2), rho''(e2)
VIO_HD void loss_compute(int kind, double delta, double e2, double rho[3]) {
    if (kind == 1) {  // Huber
        const double dsqr = delta * delta;
        if (e2 <= dsqr) {
            rho[0] = e2; rho[1] = 1.0; rho[2] = 0.0;
        } else {
            const double sq = sqrt(e2);
            rho[0] = 2.0 * sq * delta - dsqr;
            rho[1] = delta / sq;
            rho[2] = -0.5 * rho[1] / e2;
        }
    } else if (kind == 2) {  // Cauchy
        const double dsqr = delta * delta;
        const double rec = 1.0 / dsqr;
        const double aux = rec * e2 + 1.0;
        rho[0] = dsqr * log(aux);
        rho[1] = 1.0 / aux;
        rho[2] = -rec * (rho[1] * rho[1]);
    } else if (kind == 3) {  // Tukey
        const double e = sqrt(e2);
        const double d2 = delta * delta;
        if (e <= delta) {
            const double aux = e2 / d2;
            const double om = 1.0 - aux;
            rho[0] = d2 * (1.0 - om * om * om) / 3.0;
            rho[1] = om * om;
            rho[2] = -2.0 * om / d2;
        } else {
            rho[0] = d2 / 3.0; rho[1] = 0.0; rho[2] = 0.0;
        }
    } else {
        rho[0] = e2; rho[1] = 1.0; rho[2] = 0.0;
    }
}

// y = R x, y = R^T x
VIO_HD void mat3_mul_vec(const double R[9], const double x[3], double y[3]) {
    y[0] = R[0] * x[0] + R[1] * x[1] + R[2] * x[2];
    y[1] = R[3] * x[0] + R[4] * x[1] + R[5] * x[2];
    y[2] = R[6] * x[0] + R[7] * x[1] + R[8] * x[2];
}
VIO_HD void mat3t_mul_vec(const double R[9], const double x[3], double y[3]) {
    y[0] = R[0] * x[0] + R[3] * x[1] + R[6] * x[2];
    y[1] = R[1] * x[0] + R[4] * x[1] + R[7] * x[2];
    y[2] = R[2] * x[0] + R[5] * x[1] + R[8] * x[2];
}
// C = A B
VIO_HD void mat3_mul(const double A[9], const double B[9], double C[9]) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) C[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
}
// C = A^T B
VIO_HD void mat3t_mul(const double A[9], const double B[9], double C[9]) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) C[3 * r + c] = A[r] * B[c] + A[3 + r] * B[3 + c] + A[6 + r] * B[6 + c];
}
// R * hat(v)
VIO_HD void mat3_mul_hat(const double R[9], const double v[3], double C[9]) {
    // hat(v) = [0 -vz vy; vz 0 -vx; -vy vx 0];  column c of R*hat(v) = R * hat(v)[:,c]
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double a = R[3 * r], b = R[3 * r + 1], c = R[3 * r + 2];
        C[3 * r + 0] = b * v[2] - c * v[1];
        C[3 * r + 1] = c * v[0] - a * v[2];
        C[3 * r + 2] = a * v[1] - b * v[0];
    }
}
